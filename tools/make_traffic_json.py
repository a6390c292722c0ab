#!/usr/bin/env python
"""profiles/r02_ncu_traffic.json from `ncu --set full` reports of one V-cycle's big SpMV launches.

    python tools/make_traffic_json.py gpurun_out/rXX_ops_4096.csv  rep1.ncu-rep skip1  [rep2.ncu-rep skip2 ...]

Every report was captured with `-k regex:spmv_ -s <skip> -c <count>` on `bench.py --profile-one-cycle`: its launches are the
SpMV launches number skip, skip+1, ... of the cycle, i.e. the SpMV rows (positive algorithmic bytes, not elementwise /
dense tail) of the per-op table `bench.py --dump-ops` wrote in the same call, in order.  For every captured launch:
dram__bytes_read.sum + dram__bytes_write.sum against the algorithmic bytes of the same op.  The file is stamped with the
kernel family name and the sha256 of pflare_b200/csrc/kernels.cuh: bench.py reports `roofline.traffic` only when the
stamp matches the source that runs.
"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOT_SPMV = ("elementwise/permute", "dense tail", "exchange", "coarse-child")


def read_report(rep):
    if rep.endswith(".csv"):   # `ncu -i rep --page raw --csv` saved on the GPU box (the report itself can be too big to travel)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def val(d, name):
        i = hdr.index(name)
        v = float(d[i].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                    "msecond": 1e3}.get(units[i], 1)
    out = []
    for d in data:
        name = re.sub(r"\(.*", "", d[hdr.index("Kernel Name")]).replace("void ", "").replace("pfb::", "")
        out.append((name, val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"), val(d, "gpu__time_duration.sum")))
    return out


def main():
    ops_csv = sys.argv[1]
    pairs = [(sys.argv[i], int(sys.argv[i + 1])) for i in range(2, len(sys.argv) - 1, 2)]
    ops = [r for r in csv.DictReader(open(ops_csv))]
    spmv = [o for o in ops if o["op"] not in NOT_SPMV and float(o["alg_bytes"]) > 0]
    out = []
    for rep, skip in pairs:
        for k, (name, dram, us) in enumerate(read_report(rep)):
            o = spmv[skip + k]
            out.append({"level": int(o["level"]), "op": o["op"], "kernel": name, "algorithmic_bytes": int(float(o["alg_bytes"])),
                        "dram_bytes": dram, "ncu_us": us, "report": os.path.basename(rep)})
    alg = sum(x["algorithmic_bytes"] for x in out)
    dram = sum(x["dram_bytes"] for x in out)
    sha = hashlib.sha256(open(os.path.join(ROOT, "pflare_b200", "csrc", "kernels.cuh"), "rb").read()).hexdigest()
    js = {"kernel": "spmv_wc_kernel + spmv_sv_kernel", "kernels_cuh_sha256": sha,
          "source": "ncu --set full --clock-control none (%s): %d SpMV launches of one V-cycle, 4096^2" % (", ".join(os.path.basename(r) for r, _ in pairs), len(out)),
          "launches": out, "mean_dram_bytes_per_launch": dram / len(out), "mean_algorithmic_bytes_per_launch": alg / len(out),
          "dram_over_algorithmic": dram / alg}
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    json.dump(js, open(path, "w"), indent=1)
    print(path, "dram/alg = %.3f over %d launches" % (dram / alg, len(out)))
    for x in out:
        print("L%-2d %-14s %-44s alg %7.1f MB dram %7.1f MB (%.2f) %6.1f us" % (x["level"], x["op"], x["kernel"][:44], x["algorithmic_bytes"] / 1e6,
                                                                             x["dram_bytes"] / 1e6, x["dram_bytes"] / x["algorithmic_bytes"], x["ncu_us"]))


if __name__ == "__main__":
    main()
