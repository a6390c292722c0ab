"""Independent (scipy, "textbook form") evaluation of the AIR cycle, used to cross-check the C oracle.

Kaskade cycle: b_{l+1} = R_l b_l; x_L = M_L b_L; going up x_l = P_l x_{l+1} followed by
the F/C point smoothing x_f <- x_f + M_ff (b_f - A_fc x_c - A_ff x_f).  Polynomial inverses are expanded
explicitly as sums of matrix powers / products of Newton factors, i.e. NOT in the reference's Horner /
Newton recurrence order -- agreement is therefore to round-off (1e-10), not bit-wise.
"""
import numpy as np
import scipy.sparse as sp

TOL_ZERO = float(np.float32(1e-12))


def poly_apply_explicit(inv, A, x):
    """y = q(A') D^-1 x written out term by term."""
    n = A.shape[0]
    d = A.diagonal()
    neumann = inv.inverse_type == 4
    scaled = neumann or inv.diag_scale
    Ap = A
    rhs = x
    if scaled:
        Ap = sp.diags(1.0 / d) @ A
        rhs = x / d
    if neumann:
        Ap = sp.identity(n) - Ap
    co = np.asarray(inv.coeffs)
    if inv.inverse_type in (2, 3):
        re, im = co[:, 0], co[:, 1]
        # q(A) = sum_k (1/theta_k) prod_{j<k} (I - A/theta_j); complex pairs combined into real quadratics
        y = np.zeros(n)
        t = rhs.copy()
        i = 0
        nr = re.size
        while i < nr:
            if im[i] == 0.0:
                if abs(re[i]) < TOL_ZERO:
                    i += 1
                    continue
                y = y + t / re[i]
                if i < nr - 1:
                    t = t - (Ap @ t) / re[i]
                i += 1
            else:
                a, b = re[i], im[i]
                sq = a * a + b * b
                if sq < TOL_ZERO:
                    i += 2
                    continue
                u = 2 * a * t - Ap @ t
                y = y + u / sq
                if i < nr - 2:
                    t = t - (Ap @ u) / sq
                i += 2
        return y
    y = np.zeros(n)
    pw = rhs.copy()
    for k in range(co.shape[0]):
        y = y + co[k, 0] * pw
        if k + 1 < co.shape[0]:
            pw = Ap @ pw
    return y


def inv_apply(inv, A, x):
    if inv.kind == "csr":
        return inv.mat @ x
    if inv.kind == "diag":
        return inv.diag * x
    return poly_apply_explicit(inv, A, x)


def vcycle(H, b):
    bs = [np.asarray(b, dtype=np.float64)]
    for lv in H.levels:
        bs.append(lv.R @ bs[-1])
    x = inv_apply(H.inv_coarse, H.coarse_matrix, bs[-1])
    for l in range(len(H.levels) - 1, -1, -1):
        lv = H.levels[l]
        x = lv.P @ x
        bl = bs[l]
        for s in lv.smooth_order:
            if s == 0:
                break
            F, C = lv.is_fine, lv.is_coarse
            if s > 0:
                rhs = bl[F] - lv.A_fc @ x[C]
                xf = x[F]
                for _ in range(s):
                    xf = xf + inv_apply(lv.inv_A_ff, lv.A_ff, rhs - lv.A_ff @ xf)
                x[F] = xf
            else:
                rhs = bl[C] - lv.A_cf @ x[F]
                xc = x[C]
                for _ in range(-s):
                    xc = xc + inv_apply(lv.inv_A_cc, lv.A_cc, rhs - lv.A_cc @ xc)
                x[C] = xc
    return x


def vcycle_full(H, b):
    """-pc_air_full_smoothing_up_and_down: PCMG multiplicative V(1,1) with inv_A_ff(l) ~ A_l^-1 on all unknowns,
    residual restriction R (b - A x), x += P e (src/AIR_MG_Setup.F90:978-1074)."""
    As = [lv.A for lv in H.levels]

    def rec(l, bl):
        if l == len(H.levels):
            return inv_apply(H.inv_coarse, H.coarse_matrix, bl)
        lv = H.levels[l]
        x = inv_apply(lv.inv_A_ff, As[l], bl)
        e = rec(l + 1, lv.R @ (bl - As[l] @ x))
        x = x + lv.P @ e
        return x + inv_apply(lv.inv_A_ff, As[l], bl - As[l] @ x)
    return rec(0, np.asarray(b, dtype=np.float64))
