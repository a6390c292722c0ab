"""AIR hierarchy construction (stand-in for the reference's PCSetUp).

Level loop restating ``/root/reference/src/AIR_MG_Setup.F90:44-1231`` with the operator
construction of ``src/AIR_Operators_Setup.F90:36-1085``:

  CF split -> A_ff (diagonal detection) -> approximate inverse M_ff -> one-point W
  -> Z = -A_cf M_ff, drop r_drop -> R=[Z I], P=[W;I] in natural numbering
  -> A_c = R (A P), drop a_drop (optionally lump) -> next level; stop when
  n_c <= coarse_eq_limit or no F points or max_levels; then the coarsest inverse.

Supported options: the subset of ``-pc_air_*`` that changes what the *apply* sees
(inverse types power/arnoldi/newton/neumann/jacobi/wjacobi, assembled or matrix-free,
diag scaling, smooth_order incl. C smooths, z_type product only, one-point or ideal W,
truncation via max_levels, coarsest inverse options).  lAIR/SAI Z, constraints, improve
iterations, CR/aggregation splittings, reuse and processor agglomeration are not
restated here (out of scope: SURVEY.md section 2).
"""
from dataclasses import dataclass, field
from typing import List, Optional
import numpy as np
import scipy.sparse as sp

from . import native, poly
from .cf_splitting import compute_cf_splitting, drop_small


@dataclass
class AirOptions:
    """Defaults = ``air_options`` in /root/reference/src/AIR_Data_Type.F90:34-264."""
    max_levels: int = 300
    coarse_eq_limit: int = 6
    strong_threshold: float = 0.5
    ddc_its: int = 1
    ddc_fraction: float = 0.1
    smooth_order: tuple = (2,)
    diag_scale_polys: bool = False
    matrix_free_polys: bool = False
    one_point_classical_prolong: bool = True
    symmetric: bool = False
    strong_r_threshold: float = 0.0
    inverse_type: int = poly.ARNOLDI
    poly_order: int = 6
    inverse_sparsity_order: int = 1
    c_inverse_type: int = poly.ARNOLDI
    c_poly_order: int = 6
    c_inverse_sparsity_order: int = 1
    coarsest_inverse_type: int = poly.ARNOLDI
    coarsest_poly_order: int = 6
    coarsest_inverse_sparsity_order: int = 1
    coarsest_matrix_free_polys: bool = False
    coarsest_diag_scale_polys: bool = False
    r_drop: float = 0.01
    a_drop: float = 1e-4
    a_lump: bool = False
    # -pc_air_full_smoothing_up_and_down (AIR_Data_Type.F90:126-127): one Richardson sweep with an approximate
    # inverse of the WHOLE level matrix down and up (PCMG multiplicative V-cycle) instead of F/C smoothing on the way up
    full_smoothing_up_and_down: bool = False
    # -pc_air_z_type product | lair (AIR_Data_Type.F90: z_type, lair_distance = 2): lAIR solves, per C point, the local
    # system Z(i,J) A_ff(J,J) = -A_cf(i,J) on the distance-d F neighbourhood J (pattern of A_cf A_ff^(d-1))
    z_type: str = "product"
    lair_distance: int = 2
    # -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N (apply-time only: N Richardson sweeps of the coarse solve, AIR_MG_Setup.F90:1094-1102)
    mg_coarse_ksp_max_it: int = 1
    seed: int = 1

    @property
    def any_c_smooths(self):
        return any(s < 0 for s in self.smooth_order)


@dataclass
class Inverse:
    """Approximate inverse as the reference stores it in inv_A_ff / inv_A_cc.

    kind 'csr'  : assembled AIJ (mat)
    kind 'diag' : MATDIAGONAL (diag)
    kind 'poly' : matrix-free MatShell; inverse_type in PCPFLAREINVType; ``coeffs`` has
                  shape (ncoef, 1) (power/arnoldi/neumann) or (nroots, 2) (newton real,imag)
                  exactly like ``poly_data%coefficients`` (Gmres_Poly.F90:77-83)
    """
    kind: str
    mat: Optional[sp.csr_matrix] = None
    diag: Optional[np.ndarray] = None
    inverse_type: int = poly.ARNOLDI
    coeffs: Optional[np.ndarray] = None
    diag_scale: bool = False

    def nnz(self):
        return self.mat.nnz if self.kind == "csr" else (self.diag.size if self.kind == "diag" else 0)


@dataclass
class Level:
    n: int
    is_fine: np.ndarray
    is_coarse: np.ndarray
    A_ff: sp.csr_matrix
    A_fc: sp.csr_matrix
    inv_A_ff: Inverse
    R: sp.csr_matrix
    P: sp.csr_matrix
    smooth_order: List[int]
    aff_diag: bool = False
    A_cf: Optional[sp.csr_matrix] = None
    A_cc: Optional[sp.csr_matrix] = None
    inv_A_cc: Optional[Inverse] = None
    A: Optional[sp.csr_matrix] = None     # coarse_matrix(our_level): only kept with full_smoothing_up_and_down


@dataclass
class Hierarchy:
    A: sp.csr_matrix
    levels: List[Level]
    coarse_matrix: sp.csr_matrix
    inv_coarse: Inverse
    options: AirOptions = field(default_factory=AirOptions)

    @property
    def no_levels(self):
        return len(self.levels) + 1

    def sizes(self):
        return [lv.n for lv in self.levels] + [self.coarse_matrix.shape[0]]


def _i32(m):
    m = m.tocsr()
    m.sort_indices()
    return sp.csr_matrix((m.data.astype(np.float64), m.indices.astype(np.int32), m.indptr.astype(np.int32)), shape=m.shape)


def _submatrix(a, row_idx, col_map, ncols):
    """MatCreateSubMatrix(a, rows, cols): col_map[j] = new col index or -1."""
    sub = a[row_idx] if row_idx is not None else a
    sub = sub.tocsr()
    rows = np.repeat(np.arange(sub.shape[0], dtype=np.int64), np.diff(sub.indptr))
    newc = col_map[sub.indices]
    keep = newc >= 0
    indptr = np.concatenate(([0], np.cumsum(np.bincount(rows[keep], minlength=sub.shape[0]))))
    out = sp.csr_matrix((sub.data[keep], newc[keep].astype(np.int32), indptr.astype(np.int32)), shape=(sub.shape[0], ncols))
    out.has_sorted_indices = True  # col_map is monotone on the kept set
    return out


def _extract(a, rows, col_map, ncols):
    """a[rows][:, kept cols] through the native helper (falls back to scipy)."""
    out = native.extract(a, rows, col_map, ncols)
    if out is None:
        out = _submatrix(a[rows].tocsr(), None, col_map, ncols)
    return out


def _clamp_orders(n_rows, poly_order, sparsity_order):
    """setup_gmres_poly_data (Gmres_Poly.F90:40-86)."""
    po = poly_order
    if po + 1 > n_rows:
        po = int(n_rows - 1)
    so = min(sparsity_order, po)
    return po, so


def make_inverse(A, inverse_type, poly_order, sparsity_order, matrix_free, diag_scale, rng, want_assembled=False):
    """start/finish_approximate_inverse (Approx_Inverse_Setup.F90:394-500).

    Returns (inverse_for_apply, assembled_or_None).  ``assembled`` is the AIJ version the
    grid-transfer construction needs when the smoother itself is matrix-free.
    """
    n = A.shape[0]
    po, so = _clamp_orders(n, poly_order, sparsity_order)
    if inverse_type in (poly.JACOBI, poly.WJACOBI):
        d = poly.jacobi_diag(A, inverse_type == poly.WJACOBI)
        inv = Inverse("diag", diag=d, inverse_type=inverse_type)
        return inv, sp.diags(d).tocsr()
    if inverse_type == poly.NEUMANN:
        coeffs = np.ones((po + 1, 1))
        asm = None
        if (not matrix_free) or want_assembled:
            asm = _i32(poly.neumann_assembled(A, po, so))
        if matrix_free:
            return Inverse("poly", inverse_type=poly.NEUMANN, coeffs=coeffs, diag_scale=True), asm
        return Inverse("csr", mat=asm, inverse_type=poly.NEUMANN), asm
    if inverse_type in (poly.NEWTON, poly.NEWTON_NO_EXTRA):
        As = A
        if diag_scale:
            As = (sp.diags(1.0 / A.diagonal()) @ A).tocsr()
        re, im = poly.roots_newton(As, po, rng, add_roots=(inverse_type == poly.NEWTON))
        coeffs = np.stack((re, im), axis=1)
        if not matrix_free:
            raise NotImplementedError("assembled Newton-basis inverse is not restated; use matrix_free")
        asm = None
        if want_assembled:
            # STAND-IN: the reference assembles the Newton polynomial itself for the grid transfers
            # (src/Gmres_Poly_Newton.F90:1094-1929); here the Arnoldi-basis polynomial of the same order
            # is assembled instead.  Only Z / W (inputs of the apply) are affected.
            asm = poly.assembled_poly_inverse(A, poly.coefficients_arnoldi(As, po, rng), po, so, diag_scale)
        return Inverse("poly", inverse_type=inverse_type, coeffs=coeffs, diag_scale=diag_scale), asm
    if inverse_type in (poly.POWER, poly.ARNOLDI):
        As = A
        if diag_scale:
            As = (sp.diags(1.0 / A.diagonal()) @ A).tocsr()
        if inverse_type == poly.ARNOLDI:
            c = poly.coefficients_arnoldi(As, po, rng)
        else:
            c = poly.coefficients_power(As, po, rng)
        asm = None
        if (not matrix_free) or want_assembled:
            asm = poly.assembled_poly_inverse(A, c, po, so, diag_scale)
        if matrix_free:
            return Inverse("poly", inverse_type=inverse_type, coeffs=c.reshape(-1, 1).copy(), diag_scale=diag_scale), asm
        if so == 0 and po > 0:
            return Inverse("diag", diag=asm.diagonal().copy(), inverse_type=inverse_type), asm
        return Inverse("csr", mat=asm, inverse_type=inverse_type), asm
    raise NotImplementedError("inverse type %d" % inverse_type)


def one_point_W(A_fc):
    """generate_one_point_with_one_entry_from_sparse (Grid_Transfer.F90:94-220)."""
    nf, nc = A_fc.shape
    indptr = A_fc.indptr
    has = np.diff(indptr) > 0
    av = np.abs(A_fc.data)
    rows = np.repeat(np.arange(nf, dtype=np.int64), np.diff(indptr))
    mx = np.zeros(nf)
    if av.size:
        mx[has] = np.maximum.reduceat(av, indptr[:-1][has])
    ismax = av == mx[rows]
    pos = np.flatnonzero(ismax)
    first = np.full(nf, -1, dtype=np.int64)
    # first maximal entry per row (maxloc semantics)
    first[rows[pos][::-1]] = pos[::-1]
    sel = first[has]
    wptr = np.concatenate(([0], np.cumsum(has.astype(np.int64))))
    W = sp.csr_matrix((np.ones(sel.size), A_fc.indices[sel].astype(np.int32), wptr.astype(np.int32)), shape=(nf, nc))
    return W


def compute_R_from_Z(Z, is_f, is_c, n):
    """R = [Z I] in the level's natural numbering (Grid_Transfer.F90:588-815)."""
    nc = is_c.size
    Zc = Z.tocoo()
    rows = np.concatenate((Zc.row, np.arange(nc)))
    cols = np.concatenate((is_f[Zc.col], is_c))
    vals = np.concatenate((Zc.data, np.ones(nc)))
    return _i32(sp.coo_matrix((vals, (rows, cols)), shape=(nc, n)))


def compute_P_from_W(W, is_f, is_c, n):
    """P = [W; I] in the level's natural numbering (Grid_Transfer.F90:329-461)."""
    nc = is_c.size
    Wc = W.tocoo()
    rows = np.concatenate((is_f[Wc.row], is_c))
    cols = np.concatenate((Wc.col, np.arange(nc)))
    vals = np.concatenate((Wc.data, np.ones(nc)))
    return _i32(sp.coo_matrix((vals, (rows, cols)), shape=(n, nc)))


def _is_diag_only(a):
    rows = np.repeat(np.arange(a.shape[0], dtype=np.int64), np.diff(a.indptr))
    return bool(np.all(a.indices == rows))


def build_hierarchy(A, opts: AirOptions = None, verbose=False):
    opts = opts or AirOptions()
    rng = np.random.default_rng(opts.seed)
    A = _i32(A)
    levels = []
    cur = A
    for our_level in range(1, opts.max_levels):
        n = cur.shape[0]
        is_f, is_c = compute_cf_splitting(cur, opts.strong_threshold, opts.ddc_its, opts.ddc_fraction,
                                          opts.symmetric, rng)
        nf, nc = is_f.size, is_c.size
        if not (nc > opts.coarse_eq_limit and nf != 0):
            break
        fmap = np.full(n, -1, dtype=np.int64); fmap[is_f] = np.arange(nf)
        cmap = np.full(n, -1, dtype=np.int64); cmap[is_c] = np.arange(nc)
        A_ff = _extract(cur, is_f, fmap, nf)
        A_fc = _extract(cur, is_f, cmap, nc)
        A_cf = _extract(cur, is_c, fmap, nf)
        smooth = list(opts.smooth_order)
        inv_type = opts.inverse_type
        sparsity = opts.inverse_sparsity_order
        aff_diag = (opts.strong_threshold == 0.0) or _is_diag_only(A_ff)
        if aff_diag and inv_type not in (poly.SAI, poly.ISAI):
            sparsity = 0
            if inv_type != poly.WJACOBI and opts.poly_order > 2:
                smooth = [1 if s > 0 else s for s in smooth]
        # strong R threshold: dropped copies used only for the grid transfers
        if opts.strong_r_threshold != 0.0:
            Adrop = drop_small(cur, opts.strong_r_threshold, relative=1, drop_diagonal=0)
            Afd = Adrop[is_f].tocsr()
            A_ff_drop = _submatrix(Afd, None, fmap, nf)
            A_cf_drop = _submatrix(Adrop[is_c].tocsr(), None, fmap, nf)
            A_fc_drop = _submatrix(Afd, None, cmap, nc)
        else:
            A_ff_drop, A_cf_drop, A_fc_drop = A_ff, A_cf, A_fc
        if opts.full_smoothing_up_and_down:
            # the smoother inverts the whole level matrix (AIR_Operators_Setup.F90:115-119, 373-377); the grid
            # transfers get their own assembled inverse of the (dropped) A_ff (:139-151, 413-427)
            inv_ff, _ = make_inverse(cur, opts.inverse_type, opts.poly_order, opts.inverse_sparsity_order,
                                     opts.matrix_free_polys, opts.diag_scale_polys, rng)
            newton = inv_type in (poly.NEWTON, poly.NEWTON_NO_EXTRA)
            _, asm = make_inverse(A_ff_drop, inv_type, opts.poly_order, sparsity, newton, opts.diag_scale_polys, rng,
                                  want_assembled=True)
        else:
            inv_ff, asm = make_inverse(A_ff, inv_type, opts.poly_order, sparsity, opts.matrix_free_polys,
                                       opts.diag_scale_polys, rng, want_assembled=(opts.strong_r_threshold == 0.0))
            if opts.strong_r_threshold != 0.0:
                _, asm = make_inverse(A_ff_drop, inv_type, opts.poly_order, sparsity, False, opts.diag_scale_polys, rng)
        if asm is None:
            raise NotImplementedError("grid transfers need an assembled inverse (z_type product)")
        # W
        if opts.one_point_classical_prolong and not opts.symmetric:
            W = one_point_W(A_fc)
        else:
            W = native.spgemm(asm, A_fc_drop)
            W.data *= -1.0
            W = drop_small(W, opts.r_drop, relative=1)
        if opts.z_type == "lair":
            # lAIR (src/AIR_Operators_Setup.F90:699-779, src/SAI_Z.F90): neighbourhood from the dropped A_cf, A_ff,
            # local solves with the undropped ones
            pat = A_cf_drop
            for _ in range(2, opts.lair_distance + 1):
                pat = native.spgemm(pat, A_ff_drop)
            Z = native.lair_z(A_ff, A_cf, pat)
        else:
            # Z = -A_cf * inv(A_ff), drop
            Z = native.spgemm(A_cf_drop, asm)
            Z.data *= -1.0
        Z = drop_small(Z, opts.r_drop, relative=1)
        R = compute_R_from_Z(Z, is_f, is_c, n)
        P = compute_P_from_W(W, is_f, is_c, n) if not opts.symmetric else _i32(R.T)
        # coarse matrix
        AP = native.spgemm(cur, P)
        RAP = native.spgemm(R, AP)
        del AP
        coarse = _i32(drop_small(RAP, opts.a_drop, relative=1, lump=opts.a_lump))
        del RAP
        lv = Level(n=n, is_fine=is_f, is_coarse=is_c, A_ff=_i32(A_ff), A_fc=_i32(A_fc), inv_A_ff=inv_ff,
                   R=R, P=P, smooth_order=smooth, aff_diag=aff_diag)
        if opts.full_smoothing_up_and_down:
            lv.A = cur
        if opts.any_c_smooths and not opts.full_smoothing_up_and_down:
            lv.A_cf = _i32(A_cf)
            lv.A_cc = _i32(_extract(cur, is_c, cmap, nc))
            lv.inv_A_cc, _ = make_inverse(lv.A_cc, opts.c_inverse_type, opts.c_poly_order,
                                          opts.c_inverse_sparsity_order, opts.matrix_free_polys,
                                          opts.diag_scale_polys, rng)
        levels.append(lv)
        if verbose:
            print("level %2d rows %10d F %10d C %10d nnz(A) %11d nnz(Aff) %10d nnz(M) %10d nnz(Z) %10d nnz(Ac) %11d" % (
                our_level, n, nf, nc, cur.nnz, A_ff.nnz, inv_ff.nnz(), Z.nnz, coarse.nnz), flush=True)
        cur = coarse
    inv_c, _ = make_inverse(cur, opts.coarsest_inverse_type, opts.coarsest_poly_order,
                            opts.coarsest_inverse_sparsity_order, opts.coarsest_matrix_free_polys,
                            opts.coarsest_diag_scale_polys, rng)
    return Hierarchy(A=A, levels=levels, coarse_matrix=_i32(cur), inv_coarse=inv_c, options=opts)


def build_pflareinv(A, inverse_type=poly.ARNOLDI, poly_order=6, sparsity_order=1, matrix_free=False, seed=1):
    """PCPFLAREINV setup (src/PCPFLAREINV.c:689-783): a single approximate inverse of A.

    Defaults follow PCCreate_PFLAREINV; diag scaling is forced off (PCPFLAREINV.c:715-721).
    Returned as a 1-level Hierarchy whose apply is ``MatMult(mat_inverse)``.
    """
    rng = np.random.default_rng(seed)
    A = _i32(A)
    inv, _ = make_inverse(A, inverse_type, poly_order, sparsity_order, matrix_free, False, rng)
    return Hierarchy(A=A, levels=[], coarse_matrix=A, inv_coarse=inv, options=AirOptions())
