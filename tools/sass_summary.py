#!/usr/bin/env python
"""profiles/r02_sass_spmv.txt: per kernel family, the SASS instructions that identify the design (cuobjdump -sass of the built library).

    python tools/sass_summary.py            # needs pflare_b200/libpflare_b200.so (python -c "import __graft_entry__ as g; g.build()")
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pflare_b200", "libpflare_b200.so")
COLS = ["UBLKCP", "SYNCS", "LDGSTS", "SHFL", "DFMA", "DMUL", "DADD", "BAR", "LDS", "STS", "LDG", "STG", "ATOMG", "SYS", "UTMALDG", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    sample = {}
    cur = None
    names = []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            names.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        ins = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not ins:
            continue
        sample.setdefault(cur, []).append(line.rstrip())
        fam.setdefault(cur, collections.Counter())
        txt = ins.group(1)
        op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
        for c in COLS:
            if c == "SYS":
                if ".SYS" in txt:
                    fam[cur][c] += 1
            elif op.startswith(c):
                fam[cur][c] += 1
    dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
    agg = collections.OrderedDict()
    for k, cnt in fam.items():
        d = dem[k].replace("void ", "").replace("pfb::", "").replace("(anonymous namespace)::", "")
        base = re.sub(r"[<(].*", "", d)
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(cnt)
    out = ["# SASS evidence (cuobjdump -sass pflare_b200/libpflare_b200.so, sm_100a), round 2 final; regenerate with tools/sass_summary.py", "",
           "Per kernel family: number of instantiations and, summed over them, the instructions that identify the design:",
           "`UBLKCP` = 1-D TMA bulk copy (cp.async.bulk), `SYNCS` = mbarrier ops, `LDGSTS` = cp.async, `SHFL` = warp shuffles,",
           "`DFMA/DMUL/DADD` = fp64 pipe, `BAR` = CTA barrier, `LDS/STS` = shared memory, `LDG/STG` = global memory, `SYS` = system-scope",
           "memory operations and fences (the flags / peer stores of the fused ghost exchange: `ST.E.STRONG.SYS`, `LD.E.STRONG.SYS`, `MEMBAR.*.SYS`).",
           "No `UTMALDG`, `UTC*MMA`, `HMMA`: 1-D sparse fp64 streams, tensor cores deliberately unused.", "",
           "| kernel | inst. | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for base, (n, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append("| `%s` | %d | %s |" % (base, n, " | ".join(str(cnt[c]) for c in COLS)))
    for want, pats, title in (("spmv_wc_kernel<1, 8, false, 8, 2>", ("UBLKCP", "SYNCS", "FENCE"), "TMA issue, mbarrier arm / wait, proxy fence"),
                              ("spmv_sv_kernel<1, 4, true, 0>", (".SYS", "ATOMG"), "fused ghost exchange: peer stores, system-scope fence, done counter, flag release / acquire")):
        for k in fam:
            if want in dem[k]:
                out += ["", "Representative lines of `%s` (%s):" % (dem[k].replace("pfb::", ""), title), "```"]
                out += [l for l in sample[k] if any(p in l for p in pats)][:14]
                out.append("```")
                break
    open(os.path.join(ROOT, "profiles", "r02_sass_spmv.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:34]))


if __name__ == "__main__":
    main()
