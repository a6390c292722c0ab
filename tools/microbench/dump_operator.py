"""Dump real operators of an AIRG hierarchy for `spmv_pipe r <files...>`:
    python tools/microbench/dump_operator.py 2048 /tmp/ops     (writes /tmp/ops/L<l>_Aff.bin, L<l>_Minv.bin)
File format: int64 n, int64 nnz, int32 rp[n+1], int32 col[nnz], float64 val[nnz] (square, F-local = device ordering)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hiergen  # noqa: E402


def dump(path, m):
    m = m.tocsr()
    m.sort_indices()
    with open(path, "wb") as f:
        np.array([m.shape[0], m.nnz], dtype=np.int64).tofile(f)
        m.indptr.astype(np.int32).tofile(f)
        m.indices.astype(np.int32).tofile(f)
        m.data.astype(np.float64).tofile(f)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/pflare_ops"
    os.makedirs(out, exist_ok=True)
    H = hiergen.build_hierarchy(hiergen.adv_diff_fd(n, n), hiergen.AirOptions())
    for l, lv in enumerate(H.levels[:10], start=1):
        if lv.A_ff.shape[0] < 20000:
            break
        dump(os.path.join(out, "L%d_Aff.bin" % l), lv.A_ff)
        if lv.inv_A_ff.kind == "csr":
            dump(os.path.join(out, "L%d_Minv.bin" % l), lv.inv_A_ff.mat)
        print("level", l, lv.A_ff.shape[0], "rows", lv.A_ff.nnz / lv.A_ff.shape[0], "nnz/row")


if __name__ == "__main__":
    main()
