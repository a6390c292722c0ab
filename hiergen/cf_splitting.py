"""CF splitting restatement: strength matrix, PMISR (Luby), DDC cleanup.

Follows ``/root/reference/src/SAbs.F90:16-144`` (strength), ``src/PMISR_Module.F90:121-671``
(PMISR: measure = rand + #strong neighbours, smallest measure wins, measure<1 -> F,
leftovers -> C) and ``src/DDC_Module.F90:221-481`` + ``src/MatDiagDom.F90:98-273``
(swap the least diagonally dominant ``ddc_fraction`` of F rows to C, 1000 bins).
Vectorised numpy; serial (one rank) semantics.  RNG = numpy, so CF markers are NOT
bit-identical to a gfortran build of the reference.
"""
import numpy as np
import scipy.sparse as sp

F_POINT = -1
C_POINT = 1


def _rows_of(a):
    return np.repeat(np.arange(a.shape[0], dtype=np.int64), np.diff(a.indptr))


def _row_reduce(ufunc, vals, indptr, empty):
    """ufunc.reduceat over CSR rows with a value for empty rows."""
    n = indptr.size - 1
    out = np.full(n, empty, dtype=vals.dtype if vals.size else np.float64)
    nz = np.diff(indptr) > 0
    if vals.size:
        starts = indptr[:-1][nz]
        out[nz] = ufunc.reduceat(vals, starts)
    return out


def drop_small(a, tol, relative=1, lump=False, drop_diagonal=0, use_native=True):
    """remove_small_from_sparse (``src/PETSc_Helper.F90:207-412``).

    relative: 1 = tol * max|row| incl. diagonal, -1 = excl. diagonal, 0 = absolute.
    drop_diagonal: 0 never, -1 always, 1 allowed.  Entries with |v| >= row tol are kept.
    """
    a = a.tocsr()
    if use_native:
        from . import native
        out = native.drop_small(a, tol, relative, lump, drop_diagonal)
        if out is not None:
            return out
    n = a.shape[0]
    rows = _rows_of(a)
    cols = a.indices
    av = np.abs(a.data)
    isdiag = cols == rows
    if relative == 1:
        rowtol = tol * _row_reduce(np.maximum, av, a.indptr, 0.0)
    elif relative == -1:
        off = np.where(isdiag, -np.inf, av)
        mx = _row_reduce(np.maximum, off, a.indptr, -np.inf)
        mx = np.where(np.isfinite(mx), mx, -np.finfo(np.float64).max)
        rowtol = tol * mx
    else:
        rowtol = np.full(n, tol)
    keep = av >= rowtol[rows]
    if drop_diagonal == -1:
        keep &= ~isdiag
    elif drop_diagonal == 0:
        keep |= isdiag
    data = a.data.copy()
    if lump:
        dropped = np.where(keep, 0.0, data)
        lumpsum = np.bincount(rows, weights=dropped, minlength=n)
        has_diag = np.bincount(rows[isdiag], minlength=n) > 0
        if not np.all(has_diag | (lumpsum == 0.0)):
            raise ValueError("lumping onto a missing diagonal")
        data = data + np.where(isdiag, lumpsum[rows], 0.0)
    out = sp.csr_matrix((data[keep], cols[keep], np.concatenate(([0], np.cumsum(np.bincount(rows[keep], minlength=n))))),
                        shape=a.shape)
    out.has_sorted_indices = True
    return out


def generate_sabs(a, strong_threshold, symmetrize=True):
    s = drop_small(a, strong_threshold, relative=-1, lump=False, drop_diagonal=-1)
    if symmetrize:
        s = (s + s.T).tocsr()
    s.data[:] = 1.0
    s.sort_indices()
    return s


def pmisr(S, rng, measure=None, cf=None, use_native=True):
    """PMISR Luby loop on a symmetric strength matrix; returns int8 markers (F=-1, C=+1)."""
    n = S.shape[0]
    indptr, cols = S.indptr, S.indices
    rows = _rows_of(S)
    if measure is None:
        measure = rng.random(n) + np.diff(indptr)
    if cf is None:
        cf = np.zeros(n, dtype=np.int8)
    if use_native:
        from . import native
        out = native.pmisr(S, measure, cf)
        if out is not None:
            return out
    assigned = cf != 0
    zero = (~assigned) & (np.abs(measure) < 1)
    cf[zero] = F_POINT
    assigned |= zero
    idx = np.arange(n)
    while not assigned.all():
        nb = np.where(assigned[cols], np.inf, measure[cols])
        rowmin = _row_reduce(np.minimum, nb, indptr, np.inf)
        in_set = (~assigned) & (measure < rowmin)
        if not in_set.any():  # exact ties: larger index loses (PMISR_Module.F90:519-521)
            key = np.where(assigned[cols], np.inf, measure[cols] + 0.0)
            tie = (~assigned[rows]) & (key == measure[rows]) & (cols < rows)
            loses = np.zeros(n, dtype=bool)
            loses[rows[tie]] = True
            in_set = (~assigned) & (measure <= rowmin) & ~loses
        cf[in_set] = F_POINT
        assigned |= in_set
        sel = in_set[rows]
        assigned[cols[sel]] = True
    cf[cf == 0] = C_POINT
    return cf


def diag_dom_ratio(a, cf, use_native=True):
    """Per-F-row sum|a_ij, j in F, j!=i| / |a_ii| (MatDiagDom.F90:98-273)."""
    if use_native:
        from . import native
        out = native.diag_dom_ratio(a, cf)
        if out is not None:
            return out[cf == F_POINT]
    rows = _rows_of(a)
    cols = a.indices
    isF = cf == F_POINT
    m = isF[rows] & isF[cols]
    av = np.abs(a.data)
    n = a.shape[0]
    diag = np.bincount(rows[m & (rows == cols)], weights=av[m & (rows == cols)], minlength=n)
    offs = np.bincount(rows[m & (rows != cols)], weights=av[m & (rows != cols)], minlength=n)
    ratio = np.zeros(n)
    nzd = diag != 0
    ratio[nzd] = offs[nzd] / diag[nzd]
    return ratio[isF]


def ddc(a, cf, fraction_swap):
    """One DDC pass, fixed-fraction path (DDC_Module.F90:416-477)."""
    if fraction_swap == 0.0:
        return cf
    fidx = np.flatnonzero(cf == F_POINT)
    nf = fidx.size
    ratio = diag_dom_ratio(a, cf)
    if fraction_swap < 0:
        search = nf
        swap_val = -fraction_swap
    else:
        search = int(float(nf) * fraction_swap)
        if search <= 0:
            return cf
        nb = 1000
        bins = np.minimum(np.floor(ratio * nb).astype(np.int64) + 1, nb)
        bins[bins < 0] = nb
        hist = np.bincount(bins, minlength=nb + 1)[1:]
        csum = np.cumsum(hist[::-1])
        k = int(np.argmax(csum >= search))
        bin_boundary = nb - k
        swap_val = (bin_boundary - 1) / nb
    swap = ~((ratio == 0) | (ratio < swap_val))
    cf = cf.copy()
    cf[fidx[swap]] *= -1
    return cf


def compute_cf_splitting(a, strong_threshold=0.5, ddc_its=1, ddc_fraction=0.1, symmetric=False, rng=None):
    """CF_PMISR_DDC (CF_Splitting.F90:235-460). Returns sorted int32 (is_fine, is_coarse)."""
    if rng is None:
        rng = np.random.default_rng(1)
    S = generate_sabs(a, strong_threshold, symmetrize=not symmetric)
    cf = pmisr(S, rng)
    if strong_threshold != 0.0:
        for _ in range(ddc_its):
            cf = ddc(a, cf, ddc_fraction)
    is_f = np.flatnonzero(cf == F_POINT).astype(np.int32)
    is_c = np.flatnonzero(cf != F_POINT).astype(np.int32)
    return is_f, is_c
