"""GMRES polynomial coefficients / roots and assembled approximate inverses.

Restates (setup, input generation only):
* Box-Muller rhs            ``/root/reference/src/Gmres_Poly.F90:139-241``
* Arnoldi + coefficients    ``src/Gmres_Poly.F90:308-548``
* power basis + QR + gelsd  ``src/Gmres_Poly.F90:552-773``
* harmonic Ritz roots, clustering, extra roots, modified Leja ``src/Gmres_Poly_Newton.F90:21-712``
* assembled fixed-sparsity inverse ``src/Gmres_Poly.F90:920-1337,1522-1813``
* Neumann ``src/Neumann_Poly.F90:108-175``; (weighted) Jacobi ``src/Weighted_Jacobi.F90:15-88``
"""
import numpy as np
import scipy.sparse as sp
from . import native

TOL_ZERO = float(np.float32(1e-12))     # PFLARE_TOL_ZERO: single literal widened (Pflare_Parameters.F90:206)
TOL_RCOND = float(np.float32(1e-12))
TOL_ARNOLDI = float(np.float32(1e-14))
TOL_LUCKY = 1e-30
TOL_LEJA_PERTURB = float(np.float32(5e-8))
EPS = np.finfo(np.float64).eps

POWER, ARNOLDI, NEWTON, NEWTON_NO_EXTRA, NEUMANN, SAI, ISAI, WJACOBI, JACOBI = range(9)


def box_muller(n, rng):
    u = rng.random((n, 2))
    u1 = np.maximum(u[:, 0], np.finfo(np.float64).tiny)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u[:, 1])


def arnoldi(A, poly_order, v0, want_C=False, rel_tol=-1.0, lucky_tol=TOL_LUCKY):
    sub = poly_order + 1
    H = np.zeros((poly_order + 2, poly_order + 1))
    C = np.zeros((poly_order + 2, poly_order + 2)) if want_C else None
    n = v0.size
    V = np.zeros((sub + 1, n))
    beta = np.linalg.norm(v0)
    V[0] = v0 / beta
    if want_C:
        C[0, 0] = 1.0 / beta
    y = np.zeros(poly_order + 1)
    m_out = sub
    for m in range(1, sub + 1):
        w = A @ V[m - 1]
        if want_C:
            c_j = np.zeros(poly_order + 2)
            c_j[1:m + 1] = C[0:m, m - 1]
        h = V[:m] @ w
        H[:m, m - 1] = h
        w = w - h @ V[:m]
        if want_C:
            for i in range(m):
                c_j[:i + 1] -= C[:i + 1, i] * H[i, m - 1]
        H[m, m - 1] = np.linalg.norm(w)
        if H[m, m - 1] < lucky_tol:
            if rel_tol > 0:
                y = _ls_arnoldi(beta, m, H, poly_order)
            m_out = m
            break
        V[m] = w / H[m, m - 1]
        if want_C:
            C[:m + 1, m] = c_j[:m + 1] / H[m, m - 1]
        if rel_tol > 0:
            y = _ls_arnoldi(beta, m, H, poly_order)
            g0 = H[:m + 1, :m] @ y[:m]
            g0[0] -= beta
            if np.linalg.norm(g0[:m]) / beta < rel_tol:
                m_out = m
                break
    return H, C, y, m_out, beta


def _ls_arnoldi(beta, m, H, poly_order):
    g0 = np.zeros(m + 1)
    g0[0] = beta
    sol = np.linalg.lstsq(H[:m + 1, :m], g0, rcond=None)[0]
    y = np.zeros(poly_order + 1)
    y[:m] = sol
    return y


def coefficients_arnoldi(A, poly_order, rng):
    v0 = box_muller(A.shape[0], rng)
    H, C, y, m, beta = arnoldi(A, poly_order, v0, want_C=True, rel_tol=TOL_ARNOLDI)
    coeff = np.zeros(poly_order + 1)
    coeff[:m] = C[:m, :m] @ y[:m]
    return coeff


def coefficients_power(A, poly_order, rng):
    sub = poly_order + 1
    K = np.zeros((A.shape[0], sub + 1))
    K[:, 0] = box_muller(A.shape[0], rng)
    for i in range(sub):
        K[:, i + 1] = A @ K[:, i]
    R = np.linalg.qr(K, mode="r")
    g0 = np.zeros(R.shape[0])   # fewer rows than columns on tiny levels (n == poly_order + 1)
    g0[0] = R[0, 0]
    return np.linalg.lstsq(R[:, 1:], g0, rcond=None)[0]


def _modified_leja(re, im):
    n = re.size
    mag = np.sqrt(re ** 2 + im ** 2)
    order = []
    done = np.zeros(n, dtype=bool)

    def take(i):
        order.append(i); done[i] = True
        if im[i] != 0.0:
            j = i + 1 if im[i] > 0 else i - 1
            order.append(j); done[j] = True

    take(int(np.argmax(mag)))
    while len(order) < n:
        best, best_i = -np.finfo(np.float64).max, -1
        for i in range(n):
            if done[i]:
                continue
            with np.errstate(divide="ignore"):
                d = np.sqrt((re[i] - re[order]) ** 2 + (im[i] - im[order]) ** 2)
                val = 1.0 + np.sum(np.log10(d))
            if val > best:
                best, best_i = val, i
        if best < 0:
            best_i = int(np.flatnonzero(~done)[0])
        take(best_i)
    return np.array(order)


def _cluster(re, im, rel_tol, abs_tol):
    n = re.size
    used = np.zeros(n, dtype=bool)
    outr, outi = [], []
    for i in range(n):
        if used[i]:
            continue
        if re[i] == 0.0 and im[i] == 0.0:
            used[i] = True
            continue
        sr, si, cnt = re[i], im[i], 1
        used[i] = True
        mi = np.hypot(re[i], im[i])
        for j in range(i + 1, n):
            if used[j]:
                continue
            if re[j] == 0.0 and im[j] == 0.0:
                used[j] = True
                continue
            mj = np.hypot(re[j], im[j])
            dist = np.hypot(re[j] - re[i], im[j] - im[i])
            if dist <= abs_tol + rel_tol * max(mi, mj, 1.0):
                sr += re[j]; si += im[j]; cnt += 1; used[j] = True
        outr.append(sr / cnt); outi.append(si / cnt)
    r = np.zeros(n); q = np.zeros(n)
    r[:len(outr)] = outr; q[:len(outi)] = outi
    return r, q


def _extra_roots(re, im):
    n = re.size
    pof = np.ones(n)
    extra = np.zeros(n, dtype=np.int64)
    overflow = np.zeros(n, dtype=np.int64)
    for k in range(n):
        a, b = re[k], im[k]
        if b < 0:
            continue
        if abs(a) < TOL_ZERO or a * a + b * b < TOL_ZERO:
            continue
        for i in range(n):
            if i == k:
                continue
            c, d = re[i], im[i]
            if abs(c) < TOL_ZERO or c * c + d * d < TOL_ZERO:
                continue
            dr = (a * c + b * d) / (c * c + d * d)
            di = (b * c - a * d) / (c * c + d * d)
            dm = np.sqrt((1 - dr) ** 2 + di ** 2)
            with np.errstate(divide="ignore"):
                if np.log10(pof[k]) + np.log10(dm) > 307:
                    overflow[k] += int(np.log10(pof[k]))
                    pof[k] = 1.0
            pof[k] *= dm
        with np.errstate(divide="ignore"):
            lp = np.log10(pof[k])
        if lp > 4 or overflow[k] != 0:
            extra[k] = int(np.ceil((lp + overflow[k] - 4.0) / 14.0))
    outr, outi = list(re), list(im)
    for i in range(n):
        for _ in range(int(extra[i])):
            outr.append(re[i]); outi.append(im[i])
            if im[i] > 0:
                outr.append(re[i]); outi.append(-im[i])
    return np.array(outr), np.array(outi)


def roots_newton(A, poly_order, rng, add_roots=True):
    """Harmonic Ritz roots in modified Leja order; returns (real, imag), zeros at the end."""
    v0 = box_muller(A.shape[0], rng)
    H, _, _, m, beta = arnoldi(A, poly_order, v0)
    p1 = poly_order + 1
    e_d = np.zeros(p1); e_d[poly_order] = 1.0
    sol = np.linalg.lstsq(H[:p1, :p1].T.copy(), e_d, rcond=TOL_RCOND)[0]
    Hs = H[:p1, :p1].copy()
    Hs[:, poly_order] += sol * H[poly_order + 1, poly_order] ** 2
    ev = np.linalg.eigvals(Hs)
    # LAPACK geev order: conjugate pairs adjacent, positive imaginary first
    re, im = _pair_order(ev)
    H_norm = np.linalg.norm(H[:m, :m])
    rel_tol = np.sqrt(EPS)
    abs_tol = EPS * max(H_norm, beta)
    small = re ** 2 + im ** 2 < (abs_tol + rel_tol * H_norm) ** 2
    re[small] = 0.0; im[small] = 0.0
    re, im = _cluster(re, im, rel_tol, abs_tol)
    nz = (re != 0.0) | (im != 0.0)
    num = int(nz.sum())
    if num == 0:
        return np.zeros(p1), np.zeros(p1)
    r0, i0 = re[nz], im[nz]
    if add_roots:
        ra, ia = _extra_roots(r0, i0)
        pr = ra.copy()
        for i in range(num):
            k = 0
            for j in range(num, ra.size):
                if ra[j] == r0[i] and abs(ia[j]) == abs(i0[i]):
                    k += 1
                    pr[j] = ra[j] + k * TOL_LEJA_PERTURB
        idx = _modified_leja(pr, ia)
        outr = np.zeros(ra.size + (p1 - num)); outi = np.zeros_like(outr)
        outr[:ra.size] = ra[idx]; outi[:ra.size] = ia[idx]
        return outr, outi
    idx = _modified_leja(r0, i0)
    outr = np.zeros(p1); outi = np.zeros(p1)
    outr[:num] = r0[idx]; outi[:num] = i0[idx]
    return outr, outi


def _pair_order(ev):
    """Order eigenvalues like LAPACK dgeev: complex pairs adjacent, +imag first."""
    ev = list(ev)
    re, im = [], []
    used = [False] * len(ev)
    for i, z in enumerate(ev):
        if used[i]:
            continue
        used[i] = True
        if abs(z.imag) == 0.0:
            re.append(z.real); im.append(0.0)
            continue
        # find conjugate
        best, bj = None, -1
        for j in range(len(ev)):
            if used[j]:
                continue
            d = abs(ev[j] - z.conjugate())
            if best is None or d < best:
                best, bj = d, j
        used[bj] = True
        a = 0.5 * (z.real + ev[bj].real)
        b = 0.5 * (abs(z.imag) + abs(ev[bj].imag))
        re += [a, a]; im += [b, -b]
    return np.array(re), np.array(im)


# ---------------------------------------------------------------- assembled inverses

def _plus_diag(a):
    """mat_duplicate_copy_plus_diag: same values, diagonal entries present (explicit zeros)."""
    n = a.shape[0]
    if _diag_positions(a) is not None:
        return a.copy()
    d = sp.csr_matrix((np.zeros(n), np.arange(n, dtype=np.int32), np.arange(n + 1, dtype=np.int32)), shape=(n, n))
    return _add_keep_pattern(a, d)


def _diag_positions(a):
    """Index of each row's diagonal entry in a.data, or None if any row lacks one."""
    n = a.shape[0]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(a.indptr))
    pos = np.flatnonzero(a.indices == rows)
    if pos.size != n:
        return None
    return pos


def _add_diag(a, c):
    pos = _diag_positions(a)
    if pos is None:
        return _add_keep_pattern(a, sp.eye(a.shape[0], format="csr"), c)
    a = a.copy()
    a.data[pos] += c
    return a


def _add_keep_pattern(a, b, alpha=1.0):
    """a + alpha*b keeping the union pattern including explicit zeros."""
    a = a.tocoo(); b = b.tocoo()
    rows = np.concatenate((a.row, b.row)); cols = np.concatenate((a.col, b.col))
    vals = np.concatenate((a.data, alpha * b.data))
    out = sp.coo_matrix((vals, (rows, cols)), shape=a.shape).tocsr()  # sums duplicates, keeps zeros
    out.sort_indices()
    return out


def assembled_poly_inverse(A, coeff, poly_order, sparsity_order, diag_scale=False):
    """build_gmres_polynomial_inverse, assembled branch (Gmres_Poly.F90:1661-1811)."""
    n = A.shape[0]
    A = A.tocsr(); A.sort_indices()
    if diag_scale:
        dinv = 1.0 / A.diagonal()
        As = sp.diags(dinv) @ A
        As = As.tocsr(); As.sort_indices()
    else:
        As = A
    I = sp.eye(n, format="csr")
    if poly_order == 0:
        inv = (I * coeff[0]).tocsr()
    elif poly_order == 1 and sparsity_order == 1:
        inv = _add_diag(_plus_diag(As) * coeff[1], coeff[0])
    elif sparsity_order < poly_order:
        powers = [As]
        for _ in range(2, sparsity_order + 1):
            powers.append(native.spgemm(As, powers[-1]))
        S = powers[sparsity_order - 1] if sparsity_order >= 1 else None
        if sparsity_order == 0:
            # diagonal sparsity: powers of the diagonal only
            d = As.diagonal()
            vals = np.zeros(n)
            p = np.ones(n)
            for c in coeff:
                vals += c * p
                p = p * d
            inv = sp.diags(vals).tocsr()
        else:
            S = _plus_diag(S)
            acc = native.masked_powers(S, As, coeff, sparsity_order)
            inv = sp.csr_matrix((coeff[sparsity_order] * S.data + acc, S.indices.copy(), S.indptr.copy()), shape=S.shape)
            for order in range(sparsity_order - 1, 0, -1):
                inv = _add_keep_pattern(inv, powers[order - 1], coeff[order])
            inv = _add_diag(inv, coeff[0])
    else:
        inv = _plus_diag(As) * coeff[1]
        power = As
        for order in range(2, poly_order + 1):
            power = native.spgemm(As, power)
            if coeff[order] == 0.0:
                continue
            inv = _add_keep_pattern(inv, power, coeff[order])
        inv = _add_diag(inv, coeff[0])
    inv = inv.tocsr()
    if diag_scale:
        inv = (inv @ sp.diags(dinv)).tocsr()
    inv.sort_indices()
    return sp.csr_matrix((inv.data.astype(np.float64), inv.indices.astype(np.int32), inv.indptr.astype(np.int32)),
                         shape=inv.shape)


def neumann_assembled(A, poly_order, sparsity_order):
    n = A.shape[0]
    dinv = 1.0 / A.diagonal()
    T = (sp.eye(n, format="csr") - sp.diags(dinv) @ A).tocsr()
    T = _add_keep_pattern(T, A * 0.0)
    inv = assembled_poly_inverse(T, np.ones(poly_order + 1), poly_order, sparsity_order, False)
    inv = (inv @ sp.diags(dinv)).tocsr()
    inv.sort_indices()
    return inv


def jacobi_diag(A, weighted):
    d = A.diagonal().astype(np.float64)
    w = 1.0
    if weighted:
        s = 1.0 / np.sqrt(np.abs(d))
        T = sp.diags(s) @ A @ sp.diags(s)
        norm_inf = np.abs(T).sum(axis=1).max()
        w = 3.0 / (4.0 * norm_inf)
    return w / d
