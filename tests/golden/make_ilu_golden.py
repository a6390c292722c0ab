#!/usr/bin/env python
"""Golden fixture for BASELINE.json configs[4] (tests/ilu_factors.c): the L and U factors of the reference's own data fixture
`tests/data/mat_stream_2364`, produced by a scipy restatement of the ParILU(0) sweep of /root/reference/tests/ilu_factors.c:355-575
(natural ordering, -parilu_tol 1e-4, -parilu_max_sweeps 100; U left-scaled to a unit diagonal, :565-570), plus what the CPU oracle
gets for the two solves the reference test runs on them with PCPFLAREINV (PFLAREINV_NEWTON, matrix-free, default order 6;
:118-122, Richardson rtol 1e-6 unpreconditioned norm, max_it 2000, :183-186).

Run in the build container (needs /root/reference):   python tests/golden/make_ilu_golden.py
Writes tests/golden/ilu_mat_stream_2364.npz.  The original matrix file is NOT copied; only the derived factors are stored.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pflare_b200.petsc_io import Reader  # noqa: E402


def pattern_restrict(M, P):
    """values of M on the sparsity pattern of P (remove_from_sparse_match(M, R, ...) of the reference test)"""
    P = P.tocsr()
    M = M.tocsr()
    rows = np.repeat(np.arange(P.shape[0]), np.diff(P.indptr))
    vals = np.asarray(M[rows, P.indices]).ravel()
    return sp.csr_matrix((vals, P.indices.copy(), P.indptr.copy()), shape=P.shape)


def parilu(A, tol=1e-4, max_sweeps=100):
    A = sp.csr_matrix(A)
    A.sort_indices()
    n = A.shape[0]
    A_L = sp.tril(A, k=-1, format="csr")          # strict lower pattern + values of A
    A_U = sp.triu(A, k=0, format="csr")           # upper incl. diagonal
    # explicit zeros must stay in the patterns: build them from index arrays, not from arithmetic
    L_strict = sp.csr_matrix((np.zeros(A_L.nnz), A_L.indices.copy(), A_L.indptr.copy()), shape=A.shape)
    U = A_U.copy()
    thr = tol * np.sqrt((A.data ** 2).sum())
    res0 = None
    for sweep in range(max_sweeps):
        L = L_strict + sp.identity(n, format="csr")
        M = (L @ U).tocsr()
        R_L = sp.csr_matrix((A_L.data - pattern_restrict(M, A_L).data, A_L.indices.copy(), A_L.indptr.copy()), shape=A.shape)
        R_U = sp.csr_matrix((A_U.data - pattern_restrict(M, A_U).data, A_U.indices.copy(), A_U.indptr.copy()), shape=A.shape)
        res = np.sqrt((R_L.data ** 2).sum() + (R_U.data ** 2).sum())
        if res0 is None:
            res0 = res
        print("  ParILU sweep %3d  stencil residual = %.6e" % (sweep, res))
        if res < thr:
            break
        if not np.isfinite(res) or res > 1000.0 * res0:
            raise RuntimeError("ParILU diverged")
        inv_dU = 1.0 / U.diagonal()
        R_L = sp.csr_matrix((R_L.data * inv_dU[R_L.indices], R_L.indices, R_L.indptr), shape=A.shape)   # right (column) scaling
        L_strict = sp.csr_matrix((L_strict.data + R_L.data, L_strict.indices, L_strict.indptr), shape=A.shape)
        U = sp.csr_matrix((U.data + R_U.data, U.indices, U.indptr), shape=A.shape)
    L = (L_strict + sp.identity(n, format="csr")).tocsr()
    inv_d = 1.0 / U.diagonal()
    U = sp.csr_matrix((U.data * np.repeat(inv_d, np.diff(U.indptr)), U.indices, U.indptr), shape=A.shape)   # left scaling -> unit diagonal
    L.sort_indices(); U.sort_indices()
    return L, U


def main():
    import cases  # noqa: F401  (sets up the import path of the oracle / hiergen packages)
    import hiergen
    import oracle
    from hiergen import poly
    from krylov import richardson
    r = Reader(open("/root/reference/tests/data/mat_stream_2364", "rb").read())
    A = r.mat()
    print("mat_stream_2364: %d rows, %d nonzeros" % (A.shape[0], A.nnz))
    L, U = parilu(A)
    out = {}
    n = A.shape[0]
    b = np.random.default_rng(1234).random(n)
    for name, F in (("L", L), ("U", U)):
        H = hiergen.build_pflareinv(F, inverse_type=poly.NEWTON, poly_order=6, matrix_free=True)
        O = hiergen.feed(H, oracle.OracleAIR(H.no_levels))
        y = O.inv_apply(1, oracle.INV_AFF, b)
        x, its, conv = richardson(F, b, np.zeros(n), lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-6, max_it=2000)
        print("%s solve (richardson + GMRES poly): %d iterations, converged %s, ||b - F x|| / ||b|| = %.3e" % (
            name, its, conv, np.linalg.norm(b - F @ x) / np.linalg.norm(b)))
        co = np.atleast_2d(np.asarray(H.inv_coarse.coeffs, dtype=np.float64))
        out.update({name + "_indptr": F.indptr.astype(np.int32), name + "_indices": F.indices.astype(np.int32), name + "_data": F.data,
                    name + "_coeffs": co, name + "_apply": y, name + "_its": np.int64(its), name + "_conv": np.bool_(conv)})
    out["b"] = b
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ilu_mat_stream_2364.npz"), **out)
    print("wrote tests/golden/ilu_mat_stream_2364.npz")


if __name__ == "__main__":
    main()
