"""Split a (serially built) hierarchy into the per-rank pieces a PETSc MPIAIJ run would hold.

Row ownership is contiguous on every level (PETSc's layout): level 1 is split evenly (PETSc's default
``PetscSplitOwnership``: n//P rows, the first n%P ranks get one more); level l+1 is owned by the rank
that owns the corresponding C point of level l, so the coarse ranges are the per-rank C counts -- just
as ``MatCreateSubMatrix`` / the grid-transfer builders of the reference lay them out
(/root/reference/src/Grid_Transfer.F90:329-461,588-815).  Every operator is handed over as PETSc's
``MatMPIAIJGetSeqAIJ`` triple: diag block (local column indices), off-diag block over compressed ghost
columns, and ``garray`` (sorted global column ids) -- /root/reference/src/Grid_Transferk.kokkos.cxx:30-42.

The partitioned hierarchy is mathematically the serial one (same CF splitting and operators), so
the distributed V-cycle must reproduce the serial oracle to round-off.
"""
import numpy as np
import scipy.sparse as sp

AFF, AFC, ACF, ACC, INV_AFF, INV_ACC, R, P, COARSE = range(9)


def split_ownership(n, nranks):
    base, rem = divmod(n, nranks)
    counts = np.array([base + (1 if r < rem else 0) for r in range(nranks)], dtype=np.int64)
    return np.concatenate(([0], np.cumsum(counts)))


def _split_cols(sub, c0, c1):
    """rows already restricted; returns (diag CSR with local cols, offdiag CSR over compressed ghosts, garray)."""
    sub = sub.tocsr()
    sub.sort_indices()
    m = sub.shape[0]
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(sub.indptr))
    cols = sub.indices.astype(np.int64)
    isd = (cols >= c0) & (cols < c1)
    d = sp.csr_matrix((sub.data[isd], (rows[isd], cols[isd] - c0)), shape=(m, c1 - c0))
    d.sort_indices()
    garray = np.unique(cols[~isd])
    if garray.size:
        oc = np.searchsorted(garray, cols[~isd])
        o = sp.csr_matrix((sub.data[~isd], (rows[~isd], oc)), shape=(m, garray.size))
        o.sort_indices()
    else:
        o = None
    # keep explicitly stored zeros out of trouble: csr_matrix((data,(i,j))) sums duplicates only
    return d, o, garray.astype(np.int64)


class LocalOperator:
    def __init__(self, diag, offdiag, garray, cstart):
        self.diag, self.offdiag, self.garray, self.cstart = diag, offdiag, garray, int(cstart)


class LocalHierarchy:
    """What ONE rank uploads."""

    def __init__(self, rank, nranks, no_levels):
        self.rank, self.nranks, self.no_levels = rank, nranks, no_levels
        self.levels = []      # dicts: n, rstart, is_fine, is_coarse, smooth, ops{which: LocalOperator}, inv_ff, inv_cc
        self.rangesV = []     # per level: global ownership offsets (nranks + 1)
        self.rangesF = []
        self.full_smoothing = False   # -pc_air_full_smoothing_up_and_down
        self.coarse_its = 1           # -mg_coarse_ksp_max_it

    def local_rows(self):
        return self.levels[0]["n"]

    def sizes(self):
        return [int(r[-1]) for r in self.rangesV]

    def feed(self, sink):
        if self.full_smoothing:
            sink.set_option("full_smoothing_up_and_down", 1)
        if self.coarse_its > 1:
            sink.set_option("mg_coarse_ksp_max_it", self.coarse_its)
        for l, lv in enumerate(self.levels, start=1):
            sink.set_level(l, lv["n"], lv["is_fine"], lv["is_coarse"], lv["smooth"], rstart=lv["rstart"])
            for which, op in lv["ops"].items():
                sink.set_csr(l, which, op.diag, op.offdiag, op.garray, op.cstart)
            for which, inv in ((INV_AFF, lv.get("inv_ff")), (INV_ACC, lv.get("inv_cc"))):
                if inv is None:
                    continue
                kind, payload = inv
                if kind == "csr":
                    sink.set_csr(l, which, payload.diag, payload.offdiag, payload.garray, payload.cstart)
                elif kind == "diag":
                    sink.set_diag(l, which, payload)
                else:
                    sink.set_poly(l, which, payload["type"], payload["coeffs"], payload["diag_scale"])
        sink.finalize()   # collective over the ranks (a no-op for the ranks of an in-process group)
        return sink


def partition(H, nranks, only=None):
    """Returns [LocalHierarchy for rank 0 .. nranks-1]; with ``only=r`` just rank r's piece is
    materialised (the other entries carry the ownership ranges only)."""
    NL = H.no_levels
    out = [LocalHierarchy(r, nranks, NL) for r in range(nranks)]
    full = bool(getattr(H.options, "full_smoothing_up_and_down", False))
    for lh in out:
        lh.full_smoothing = full
        lh.coarse_its = int(getattr(H.options, "mg_coarse_ksp_max_it", 1) or 1)
    todo = range(nranks) if only is None else [only]
    rv = split_ownership(H.levels[0].n if H.levels else H.coarse_matrix.shape[0], nranks)

    def local_inv(inv, rrange, crange, r):
        if inv is None:
            return None
        if inv.kind == "csr":
            d, o, g = _split_cols(inv.mat[rrange[r]:rrange[r + 1]], crange[r], crange[r + 1])
            return ("csr", LocalOperator(d, o, g, crange[r]))
        if inv.kind == "diag":
            return ("diag", np.asarray(inv.diag)[rrange[r]:rrange[r + 1]].copy())
        return ("poly", {"type": inv.inverse_type, "coeffs": inv.coeffs, "diag_scale": inv.diag_scale})

    for l, lv in enumerate(H.levels):
        isf, isc = np.asarray(lv.is_fine, dtype=np.int64), np.asarray(lv.is_coarse, dtype=np.int64)
        rf = np.searchsorted(isf, rv)          # F ownership offsets
        rc = np.searchsorted(isc, rv)          # C ownership offsets == next level's row ownership
        for r in todo:
            a, b = rv[r], rv[r + 1]
            d = {"n": int(b - a), "rstart": int(a), "smooth": list(lv.smooth_order),
                 "is_fine": (isf[rf[r]:rf[r + 1]] - a).astype(np.int32),
                 "is_coarse": (isc[rc[r]:rc[r + 1]] - a).astype(np.int32), "ops": {}}

            def put(which, mat, rrange, crange):
                dd, oo, gg = _split_cols(mat[rrange[r]:rrange[r + 1]], crange[r], crange[r + 1])
                d["ops"][which] = LocalOperator(dd, oo, gg, crange[r])
            put(R, lv.R, rc, rv)
            put(P, lv.P, rv, rc)
            if full:
                # the smoother acts on all unknowns of the level: coarse_matrix(level) and its approximate inverse
                put(COARSE, lv.A, rv, rv)
                d["inv_ff"] = local_inv(lv.inv_A_ff, rv, rv, r)
            else:
                put(AFF, lv.A_ff, rf, rf)
                put(AFC, lv.A_fc, rf, rc)
                if lv.A_cf is not None and lv.A_cc is not None:
                    put(ACF, lv.A_cf, rc, rf)
                    put(ACC, lv.A_cc, rc, rc)
                    d["inv_cc"] = local_inv(lv.inv_A_cc, rc, rc, r)
                d["inv_ff"] = local_inv(lv.inv_A_ff, rf, rf, r)
            out[r].levels.append(d)
            out[r].rangesV.append(rv.copy())
            out[r].rangesF.append(rf.copy())
        rv = rc
    for r in todo:
        a, b = rv[r], rv[r + 1]
        d = {"n": int(b - a), "rstart": int(a), "smooth": [], "is_fine": np.zeros(0, np.int32),
             "is_coarse": np.zeros(0, np.int32), "ops": {}}
        dd, oo, gg = _split_cols(H.coarse_matrix[a:b], a, b)
        d["ops"][COARSE] = LocalOperator(dd, oo, gg, a)
        d["inv_ff"] = local_inv(H.inv_coarse, rv, rv, r)
        out[r].levels.append(d)
        out[r].rangesV.append(rv.copy())
        out[r].rangesF.append(rv.copy())
    return out


def scatter_vector(v, ranges):
    return [np.ascontiguousarray(v[ranges[r]:ranges[r + 1]]) for r in range(len(ranges) - 1)]
