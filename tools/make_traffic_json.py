#!/usr/bin/env python
"""profiles/r02_ncu_traffic.json from an `ncu --set full` report of one V-cycle's big SpMV launches.

    python tools/make_traffic_json.py gpurun_out/rXX_prof.ncu-rep gpurun_out/rXX_ops_4096.csv <first launch index in the ops table>

For every captured launch: dram__bytes_read.sum + dram__bytes_write.sum against the algorithmic bytes of the same op
(the per-op table bench.py --dump-ops wrote in the same call).  The file is stamped with the kernel family name and the
sha256 of pflare_b200/csrc/kernels.cuh: bench.py reports `roofline.traffic` only when the stamp matches the source that runs.
"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, ops_csv, first = sys.argv[1], sys.argv[2], int(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(name):
        return hdr.index(name)

    def val(d, name):
        v = float(d[col(name)].replace(",", ""))
        u = units[col(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
    ops = [r for r in csv.DictReader(open(ops_csv))]
    out = []
    for k, d in enumerate(data):
        o = ops[first + k]
        name = re.sub(r"\(.*", "", d[col("Kernel Name")]).replace("void ", "").replace("pfb::", "")
        out.append({"level": int(o["level"]), "op": o["op"], "kernel": name, "algorithmic_bytes": int(float(o["alg_bytes"])),
                    "dram_bytes": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
                    "ncu_us": val(d, "gpu__time_duration.sum")})
    alg = sum(x["algorithmic_bytes"] for x in out)
    dram = sum(x["dram_bytes"] for x in out)
    sha = hashlib.sha256(open(os.path.join(ROOT, "pflare_b200", "csrc", "kernels.cuh"), "rb").read()).hexdigest()
    js = {"kernel": "spmv_wc_kernel + spmv_sv_kernel", "kernels_cuh_sha256": sha,
          "source": "ncu --set full --clock-control none (%s): %d SpMV launches of one V-cycle, 4096^2" % (os.path.basename(rep), len(out)),
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
