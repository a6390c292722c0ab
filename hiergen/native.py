"""ctypes loader for hiergen's OpenMP helpers, with scipy fallbacks."""
import ctypes
import os
import subprocess
import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hg_kernels.so")
_SRC = os.path.join(_HERE, "csrc", "hg_kernels.cpp")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["g++", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            try:
                build()
            except Exception:
                return None
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _csr64(a):
    a = a.tocsr()
    return (np.ascontiguousarray(a.indptr, dtype=np.int64), np.ascontiguousarray(a.indices, dtype=np.int32),
            np.ascontiguousarray(a.data, dtype=np.float64))


def _mk(data, indices, indptr, shape):
    ip = indptr.astype(np.int32) if indptr[-1] < 2**31 - 1 else indptr
    m = sp.csr_matrix((data, indices, ip), shape=shape)
    m.has_sorted_indices = True
    return m


def spgemm(a, b, use_native=True):
    """C = A @ B with sorted column indices (explicit zeros kept)."""
    L = lib() if use_native else None
    if L is None:
        c = (a @ b).tocsr()
        c.sort_indices()
        return c
    ai, aj, av = _csr64(a)
    bi, bj, bv = _csr64(b)
    m = a.shape[0]
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    L.hg_spgemm_run.restype = ctypes.c_void_p
    tot = ctypes.c_int64(0)
    h = L.hg_spgemm_run(i64(m), i64(b.shape[1]), _p(ai, i64), _p(aj, i32), _p(av, f64), _p(bi, i64), _p(bj, i32), _p(bv, f64), ctypes.byref(tot))
    ci = np.zeros(m + 1, dtype=np.int64)
    cj = np.empty(max(tot.value, 1), dtype=np.int32)[:tot.value]
    cv = np.empty(max(tot.value, 1), dtype=np.float64)[:tot.value]
    L.hg_spgemm_fetch(ctypes.c_void_p(h), _p(ci, i64), _p(cj, i32), _p(cv, f64))
    return _mk(cv, cj, ci, (a.shape[0], b.shape[1]))


def drop_small(a, tol, relative=1, lump=False, drop_diagonal=0):
    """Native remove_small_from_sparse; returns None when the helper library is unavailable."""
    L = lib()
    if L is None or a.nnz >= 2**31 - 1:
        return None
    a = a.tocsr()
    ai = np.ascontiguousarray(a.indptr, dtype=np.int32)
    aj = np.ascontiguousarray(a.indices, dtype=np.int32)
    av = np.ascontiguousarray(a.data, dtype=np.float64)
    m = a.shape[0]
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    cnt = np.zeros(m, dtype=np.int64)
    args = (i64(m), _p(ai, i32), _p(aj, i32), _p(av, f64), ctypes.c_double(tol), ctypes.c_int(relative), ctypes.c_int(int(lump)),
            ctypes.c_int(drop_diagonal))
    L.hg_drop_small(*args, _p(cnt, i64), None, None, None)
    oi = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(cnt, out=oi[1:])
    oj = np.empty(max(oi[-1], 1), dtype=np.int32)[:oi[-1]]
    ov = np.empty(max(oi[-1], 1), dtype=np.float64)[:oi[-1]]
    bad = L.hg_drop_small(*args, None, _p(oi, i64), _p(oj, i32), _p(ov, f64))
    if bad:
        raise ValueError("lumping onto a missing diagonal")
    return _mk(ov, oj, oi, a.shape)


def masked_powers(S, A, coeff, sparsity_order, use_native=True):
    """acc = sum_{term>=s+2} coeff[term-1] * T_{term-1} on the sparsity of S = A^s."""
    coeff = np.ascontiguousarray(coeff, dtype=np.float64)
    ncoef = coeff.size
    L = lib() if use_native else None
    if L is None:
        mask = S.copy()
        mask.data[:] = 1.0
        T = S.copy()
        acc = S.copy()
        acc.data[:] = 0.0
        for term in range(sparsity_order + 2, ncoef + 1):
            T = (T @ A).multiply(mask).tocsr()
            T = (T + 0.0 * mask).tocsr()  # restore full pattern (explicit zeros)
            T.sort_indices()
            acc = (acc + coeff[term - 1] * T + 0.0 * mask).tocsr()
            acc.sort_indices()
        return acc.data.copy() if acc.nnz == S.nnz else _align(acc, S)
    si, sj, sv = _csr64(S)
    ai, aj, av = _csr64(A)
    out = np.zeros(S.nnz, dtype=np.float64)
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    L.hg_masked_powers(i64(S.shape[0]), _p(si, i64), _p(sj, i32), _p(sv, f64), _p(ai, i64), _p(aj, i32),
                       _p(av, f64), ctypes.c_int(ncoef), ctypes.c_int(sparsity_order), _p(coeff, f64), _p(out, f64))
    return out


def _align(acc, S):
    # values of acc at the pattern of S (acc pattern is a subset of S's)
    out = np.zeros(S.nnz)
    accd = acc.todok()
    rows = np.repeat(np.arange(S.shape[0]), np.diff(S.indptr))
    for t, (r, c) in enumerate(zip(rows, S.indices)):
        out[t] = accd.get((r, c), 0.0)
    return out


def extract(a, rows, col_map, ncols):
    """a[rows][:, col_map >= 0] with relabelled columns (native); None if the helper is unavailable."""
    L = lib()
    if L is None:
        return None
    ai = np.ascontiguousarray(a.indptr, dtype=np.int32)
    aj = np.ascontiguousarray(a.indices, dtype=np.int32)
    av = np.ascontiguousarray(a.data, dtype=np.float64)
    cm = np.ascontiguousarray(col_map, dtype=np.int64)
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    if rows is None:
        nrows, rp = a.shape[0], None
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        nrows, rp = rows.size, _p(rows, i32)
    cnt = np.zeros(nrows, dtype=np.int64)
    L.hg_extract(i64(nrows), rp, _p(ai, i32), _p(aj, i32), _p(av, f64), _p(cm, i64), _p(cnt, i64), None, None, None)
    oi = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(cnt, out=oi[1:])
    oj = np.empty(max(oi[-1], 1), dtype=np.int32)[:oi[-1]]
    ov = np.empty(max(oi[-1], 1), dtype=np.float64)[:oi[-1]]
    L.hg_extract(i64(nrows), rp, _p(ai, i32), _p(aj, i32), _p(av, f64), _p(cm, i64), None, _p(oi, i64), _p(oj, i32), _p(ov, f64))
    return _mk(ov, oj, oi, (nrows, ncols))


def pmisr(S, measure, cf):
    L = lib()
    if L is None:
        return None
    si = np.ascontiguousarray(S.indptr, dtype=np.int32)
    sj = np.ascontiguousarray(S.indices, dtype=np.int32)
    me = np.ascontiguousarray(measure, dtype=np.float64)
    cf = np.ascontiguousarray(cf, dtype=np.int8)
    L.hg_pmisr(ctypes.c_int64(S.shape[0]), _p(si, ctypes.c_int32), _p(sj, ctypes.c_int32), _p(me, ctypes.c_double),
               cf.ctypes.data_as(ctypes.POINTER(ctypes.c_byte)))
    return cf


def diag_dom_ratio(a, cf):
    L = lib()
    if L is None:
        return None
    ai = np.ascontiguousarray(a.indptr, dtype=np.int32)
    aj = np.ascontiguousarray(a.indices, dtype=np.int32)
    av = np.ascontiguousarray(a.data, dtype=np.float64)
    cf8 = np.ascontiguousarray(cf, dtype=np.int8)
    out = np.zeros(a.shape[0])
    L.hg_diag_dom_ratio(ctypes.c_int64(a.shape[0]), _p(ai, ctypes.c_int32), _p(aj, ctypes.c_int32), _p(av, ctypes.c_double),
                        cf8.ctypes.data_as(ctypes.POINTER(ctypes.c_byte)), _p(out, ctypes.c_double))
    return out


def lair_z(A_ff, A_cf, sparsity):
    """lAIR Z (incomplete SAI): Z(i, J) A_ff(J, J) = -A_cf(i, J) on the F neighbourhood J = sparsity row i
    (/root/reference/src/SAI_Z.F90:24-640).  Returns Z with the pattern of ``sparsity``."""
    L = lib()
    S = sparsity.tocsr()
    S.sort_indices()
    ff = A_ff.tocsr(); ff.sort_indices()
    cf = A_cf.tocsr(); cf.sort_indices()
    si = np.ascontiguousarray(S.indptr, dtype=np.int32)
    sj = np.ascontiguousarray(S.indices, dtype=np.int32)
    zv = np.zeros(sj.size, dtype=np.float64)
    if L is None:   # scipy / numpy fallback (small problems only)
        ffc = ff.tocsc()
        for i in range(S.shape[0]):
            J = sj[si[i]:si[i + 1]]
            if J.size == 0:
                continue
            M = ff[J][:, J].toarray()
            rhs = -np.asarray(cf[i, J].todense()).ravel()
            zv[si[i]:si[i + 1]] = np.linalg.solve(M.T, rhs)
    else:
        i32, f64 = ctypes.c_int32, ctypes.c_double
        ffi = np.ascontiguousarray(ff.indptr, dtype=np.int32); ffj = np.ascontiguousarray(ff.indices, dtype=np.int32)
        ffv = np.ascontiguousarray(ff.data, dtype=np.float64)
        cfi = np.ascontiguousarray(cf.indptr, dtype=np.int32); cfj = np.ascontiguousarray(cf.indices, dtype=np.int32)
        cfv = np.ascontiguousarray(cf.data, dtype=np.float64)
        L.hg_lair_z.restype = None
        L.hg_lair_z(ctypes.c_int64(S.shape[0]), _p(si, i32), _p(sj, i32), _p(ffi, i32), _p(ffj, i32), _p(ffv, f64),
                    _p(cfi, i32), _p(cfj, i32), _p(cfv, f64), _p(zv, f64))
    return _mk(zv, sj, si, S.shape)
