"""Run the PRODUCT SpMV kernel (through pflare_b200_inv_apply, assembled inverse, device pointers, no copies)
on the same synthetic banded fixed-row-length matrices as tools/microbench/spmv_pipe.cu, to separate
"kernel generality overhead" from "structure of the real operators"."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hiergen  # noqa: E402
import pflare_b200  # noqa: E402


def banded(n, L, spread=3000, seed=1):
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.int64)
    cols = np.empty((n, L), dtype=np.int64)
    for k in range(L):
        cols[:, k] = np.maximum(i - spread * (L - 1 - k) // max(L - 1, 1) - rng.integers(0, 3, n) - k, 0)
    cols.sort(axis=1)
    vals = 1.0 / (1 + rng.integers(0, 7, (n, L)))
    indptr = np.arange(0, n * L + 1, L, dtype=np.int64)
    m = sp.csr_matrix((vals.ravel(), cols.ravel().astype(np.int32), indptr.astype(np.int32)), shape=(n, n))
    m.sum_duplicates()
    m.sort_indices()
    return m


def main():
    torch.cuda.set_device(0)
    for L, n in ((2, 1 << 24), (8, 1 << 23), (32, 1 << 21)):
        M = banded(n, L)
        H = hiergen.Hierarchy(A=M, levels=[], coarse_matrix=M, inv_coarse=hiergen.Inverse("csr", mat=M))
        for opts in ({}, {"kernel": 25}, {"kernel": 0}):
            dev = pflare_b200.DeviceAIR(1)
            for k, v in opts.items():
                dev.set_option(k, v)
            dev.set_level(1, n, [], [], [])
            dev.set_csr(1, pflare_b200.INV_AFF, M)
            dev.finalize()
            x = torch.rand(n, dtype=torch.float64, device="cuda")
            y = torch.empty_like(x)
            stream = torch.cuda.ExternalStream(dev.stream_ptr())
            torch.cuda.synchronize()
            for _ in range(3):
                dev.inv_apply_ptr(1, pflare_b200.INV_AFF, x.data_ptr(), y.data_ptr(), 1)
            dev.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            reps = 10
            for _ in range(reps):
                dev.inv_apply_ptr(1, pflare_b200.INV_AFF, x.data_ptr(), y.data_ptr(), 1)
            e1.record(stream)
            dev.synchronize()
            ms = e0.elapsed_time(e1) / reps
            by = 12.0 * M.nnz + 4.0 * n + 8.0 * n + 8.0 * n
            ref = (M[:1000] @ x.cpu().numpy())
            err = np.abs(y[:1000].cpu().numpy() - ref).max()
            print("L=%-3d nnz/row %.2f %-14s %8.3f ms  %6.0f GB/s  (err %.1e)" % (L, M.nnz / n, str(opts), ms, by / (ms * 1e-3) / 1e9, err), flush=True)
            dev.close()


if __name__ == "__main__":
    main()
