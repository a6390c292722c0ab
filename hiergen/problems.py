"""Problem generators restating the reference's test drivers.

* ``adv_1d``       -- ``/root/reference/tests/adv_1d.c:79-105``
* ``adv_diff_fd``  -- ``/root/reference/tests/adv_diff_fd.c:366-586`` (``ComputeMat``)
* ``read_petsc_binary`` -- PETSc binary Mat/Vec files (``tests/data/*``)
* ``dg_upwind_surrogate`` -- block-structured stand-in for ``tests/adv_dg_upwind.c``
  (the real driver needs DMPlex + gmsh meshes; documented as a surrogate).

All matrices are scipy CSR with int32 indices, float64 values, sorted columns,
rows in the natural (i fastest, then j, then k) DMDA ordering of a 1-rank run.
"""
import numpy as np
import scipy.sparse as sp


def _csr(rows, cols, vals, n):
    a = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    a = sp.csr_matrix((a.data.astype(np.float64), a.indices.astype(np.int32),
                       a.indptr.astype(np.int32)), shape=(n, n))
    return a


def adv_1d(n=1000):
    """1D upwind advection: row 0 = (1); row i = (-1 @ i-1, +1 @ i)."""
    i = np.arange(1, n)
    rows = np.concatenate(([0], i, i))
    cols = np.concatenate(([0], i - 1, i))
    vals = np.concatenate(([1.0], -np.ones(n - 1), np.ones(n - 1)))
    return _csr(rows, cols, vals, n)


def adv_diff_fd(nx, ny, nz=None, theta=None, alpha=0.0, u=None, v=None, w=None,
                adv_nondim=None, L=(1.0, 1.0, 1.0)):
    """Upwind finite-difference advection(-diffusion), 2D (nz=None) or 3D.

    Defaults follow the driver: velocity (1,1[,1]) normalised to unit length
    (theta=pi/4 in 2D), alpha=0 -> nondimensional pure advection with outflow
    rows keeping the upwind stencil; alpha!=0 -> Dirichlet on every face and the
    equation scaled by the cell volume.  Zero entries are not stored
    (MAT_IGNORE_ZERO_ENTRIES in the driver).
    """
    dim = 2 if nz is None else 3
    if theta is not None:
        uu, vv, ww = np.cos(theta), np.sin(theta), 0.0
    else:
        uu, vv, ww = 1.0, 1.0, 1.0
    unit = True
    if u is not None or v is not None or w is not None:
        unit = False
        uu = uu if u is None else u
        vv = vv if v is None else v
        ww = ww if w is None else w
    if unit:
        mag = np.sqrt(uu * uu + vv * vv + (ww * ww if dim == 3 else 0.0))
        if mag > 1e-12:
            uu, vv, ww = uu / mag, vv / mag, ww / mag
    if adv_nondim is None:
        adv_nondim = (alpha == 0.0)
    M, N = nx, ny
    P = nz if dim == 3 else 1
    Hx = L[0] / (M - 1)
    Hy = L[1] / (N - 1)
    Hz = L[2] / (P - 1) if dim == 3 else 1.0
    n = M * N * P
    k, j, i = np.meshgrid(np.arange(P), np.arange(N), np.arange(M), indexing="ij")
    i = i.ravel(); j = j.ravel(); k = k.ravel()
    idx = (k * N + j) * M + i
    R, C, V = [], [], []

    def add(mask, di, dj, dk, val):
        if np.isscalar(val):
            if val == 0.0:
                return
            val = np.full(int(mask.sum()), val)
        r = idx[mask]
        c = ((k[mask] + dk) * N + (j[mask] + dj)) * M + (i[mask] + di)
        R.append(r); C.append(c); V.append(val)

    if dim == 2:
        inflow = (i == 0) | (j == 0)
        outflow = ((i == M - 1) | (j == N - 1)) & ~inflow
        interior = ~(inflow | outflow)
        ax = 1.0 if adv_nondim else Hx          # goes with v (south)
        ay = Hy / Hx if adv_nondim else Hy      # goes with u (west)
        add(inflow, 0, 0, 0, 1.0)
        if alpha == 0.0:
            advmask = interior | outflow
        else:
            add(outflow, 0, 0, 0, 1.0)
            advmask = interior
            add(interior, 0, -1, 0, -alpha * Hx / Hy)
            add(interior, -1, 0, 0, -alpha * Hy / Hx)
            add(interior, 0, 0, 0, alpha * 2.0 * (Hx / Hy + Hy / Hx))
            add(interior, 1, 0, 0, -alpha * Hy / Hx)
            add(interior, 0, 1, 0, -alpha * Hx / Hy)
        if uu != 0.0 or vv != 0.0:
            add(advmask, 0, -1, 0, -vv * ax)
            add(advmask, -1, 0, 0, -uu * ay)
            add(advmask, 0, 0, 0, uu * ay + vv * ax)
    else:
        inflow = (i == 0) | (j == 0) | (k == 0)
        outflow = ((i == M - 1) | (j == N - 1) | (k == P - 1)) & ~inflow
        interior = ~(inflow | outflow)
        if adv_nondim:
            ayz, axz, axy = (Hy * Hz / Hx) / Hx, Hz / Hx, Hy / Hx
        else:
            ayz, axz, axy = Hy * Hz, Hx * Hz, Hx * Hy
        add(inflow, 0, 0, 0, 1.0)
        if alpha == 0.0:
            advmask = interior | outflow
        else:
            add(outflow, 0, 0, 0, 1.0)
            advmask = interior
            dxy, dxz, dyz = Hx * Hy / Hz, Hx * Hz / Hy, Hy * Hz / Hx
            add(interior, 0, 0, -1, -alpha * dxy)
            add(interior, 0, -1, 0, -alpha * dxz)
            add(interior, -1, 0, 0, -alpha * dyz)
            add(interior, 0, 0, 0, alpha * 2.0 * (dyz + dxz + dxy))
            add(interior, 1, 0, 0, -alpha * dyz)
            add(interior, 0, 1, 0, -alpha * dxz)
            add(interior, 0, 0, 1, -alpha * dxy)
        if uu != 0.0 or vv != 0.0 or ww != 0.0:
            add(advmask, 0, 0, -1, -ww * axy)
            add(advmask, 0, -1, 0, -vv * axz)
            add(advmask, -1, 0, 0, -uu * ayz)
            add(advmask, 0, 0, 0, uu * ayz + vv * axz + ww * axy)
    return _csr(np.concatenate(R), np.concatenate(C), np.concatenate(V), n)


def dg_upwind_surrogate(ncx, ncy, nb=3, seed=7):
    """Block-structured stand-in for DG-P1 upwind advection (tests/adv_dg_upwind.c).

    ncx*ncy cells with nb unknowns each; every cell couples to itself (dense
    nb x nb block, diagonally dominant) and to its west and south upwind
    neighbours (dense blocks with non-positive row sums cancelling the volume
    term).  SURROGATE: the real driver assembles on DMPlex meshes.
    """
    rng = np.random.default_rng(seed)
    nc = ncx * ncy
    cj, ci = np.meshgrid(np.arange(ncy), np.arange(ncx), indexing="ij")
    ci = ci.ravel(); cj = cj.ravel()
    cell = cj * ncx + ci
    R, C, V = [], [], []
    loc = np.arange(nb)
    lr, lc = np.meshgrid(loc, loc, indexing="ij")
    lr = lr.ravel(); lc = lc.ravel()
    base_mass = (np.eye(nb) * 2.0 + 0.25 * rng.random((nb, nb))).ravel()
    base_up = (-(0.5 + 0.5 * rng.random((nb, nb))) / nb).ravel()

    def add_block(rcell, ccell, blk):
        m = rcell.size
        R.append((rcell[:, None] * nb + lr[None, :]).ravel())
        C.append((ccell[:, None] * nb + lc[None, :]).ravel())
        V.append(np.tile(blk, m))

    add_block(cell, cell, base_mass)
    west = ci > 0
    add_block(cell[west], cell[west] - 1, base_up)
    south = cj > 0
    add_block(cell[south], cell[south] - ncx, base_up)
    return _csr(np.concatenate(R), np.concatenate(C), np.concatenate(V), nc * nb)


def read_petsc_binary(path):
    """Read a PETSc binary file: one AIJ Mat followed by any number of Vecs.

    Format (big endian): Mat = int32 [1211216, M, N, nnz], M row lengths,
    nnz int32 columns, nnz float64 values; Vec = int32 [1211214, n], n float64.
    """
    buf = open(path, "rb").read()
    off = 0
    mats, vecs = [], []
    while off < len(buf):
        cid = int(np.frombuffer(buf, ">i4", 1, off)[0])
        if cid == 1211216:
            _, M, N, nnz = np.frombuffer(buf, ">i4", 4, off).astype(np.int64); off += 16
            rl = np.frombuffer(buf, ">i4", M, off).astype(np.int64); off += 4 * M
            cols = np.frombuffer(buf, ">i4", nnz, off).astype(np.int32); off += 4 * nnz
            vals = np.frombuffer(buf, ">f8", nnz, off).astype(np.float64); off += 8 * nnz
            indptr = np.concatenate(([0], np.cumsum(rl))).astype(np.int32)
            a = sp.csr_matrix((vals, cols, indptr), shape=(M, N))
            a.sort_indices()
            mats.append(a)
        elif cid == 1211214:
            n = int(np.frombuffer(buf, ">i4", 2, off)[1]); off += 8
            vecs.append(np.frombuffer(buf, ">f8", n, off).astype(np.float64)); off += 8 * n
        else:
            raise ValueError("unknown PETSc class id %d at offset %d" % (cid, off))
    return mats, vecs


def parilu_factors(A, tol=1e-4, max_sweeps=100):
    """Matrix-form Chow ParILU(0) sweep of ``tests/ilu_factors.c:484-535`` followed by the left
    scaling of U by 1/diag(U) (``:563-567``).  Returns (L, U_scaled, inv_diag_U_raw, sweeps).

    L = unit lower factor on the strict-lower pattern of A (+ diagonal), U on the upper pattern.
    Each sweep: M = L U; R_L = A_Lstrict - M|pat(Lstrict); R_U = A_U - M|pat(U); stop when
    sqrt(|R_L|_F^2 + |R_U|_F^2) < tol*|A|_F; else L += R_L diag(U)^-1, U += R_U.
    """
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    lower = A.indices < rows
    A_Ls = sp.csr_matrix((A.data[lower], (rows[lower], A.indices[lower])), shape=A.shape)
    A_U = sp.csr_matrix((A.data[~lower], (rows[~lower], A.indices[~lower])), shape=A.shape)
    mask_L = A_Ls.copy(); mask_L.data[:] = 1.0
    mask_U = A_U.copy(); mask_U.data[:] = 1.0
    L = (sp.identity(n, format="csr") + 0.0 * A_Ls).tocsr()
    U = A_U.copy()
    thr = tol * np.sqrt((A.data ** 2).sum())
    sweeps = 0
    for sweep in range(max_sweeps):
        M = (L @ U).tocsr()
        R_L = (A_Ls - M.multiply(mask_L)).tocsr()
        R_U = (A_U - M.multiply(mask_U)).tocsr()
        res = np.sqrt((R_L.data ** 2).sum() + (R_U.data ** 2).sum())
        sweeps = sweep + 1
        if res < thr:
            break
        inv_dU = 1.0 / U.diagonal()
        L = (L + R_L @ sp.diags(inv_dU)).tocsr()
        U = (U + R_U).tocsr()
    inv_diag_U_raw = 1.0 / U.diagonal()
    U = (sp.diags(inv_diag_U_raw) @ U).tocsr()

    def fix(m):
        m = m.tocsr(); m.sort_indices()
        return sp.csr_matrix((m.data.astype(np.float64), m.indices.astype(np.int32), m.indptr.astype(np.int32)), shape=m.shape)
    return fix(L), fix(U), inv_diag_U_raw, sweeps
