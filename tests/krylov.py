"""Outer Krylov loops restating what the reference's drivers ask PETSc for (test helper).

GMRES(30) with left or right preconditioning and PETSc's default convergence test
(KSPConvergedDefault: ||r|| <= max(rtol * ||b||, atol); with a zero rhs and a nonzero initial
guess PETSc falls back to rtol * ||r0||), plus preconditioned Richardson.  Returns the iteration
count PETSc would report (number of Krylov steps until convergence).
"""
import numpy as np


def gmres(A, b, x0, M, rtol=1e-5, atol=1e-50, max_it=10000, restart=30, side="left"):
    x = x0.copy()
    apply_A = (lambda v: A @ v)
    its = 0
    ref = None
    while True:
        r = b - apply_A(x)
        if side == "left":
            r = M(r)
        beta = np.linalg.norm(r)
        if ref is None:
            bn = np.linalg.norm(M(b) if side == "left" else b)
            ref = bn if bn > 0 else beta
            if beta <= max(rtol * ref, atol):
                return x, 0, True
        V = np.zeros((restart + 1, b.size))
        H = np.zeros((restart + 1, restart))
        V[0] = r / beta
        g = np.zeros(restart + 1); g[0] = beta
        cs = np.zeros(restart); sn = np.zeros(restart)
        k_done = 0
        converged = False
        for k in range(restart):
            if side == "left":
                w = M(apply_A(V[k]))
            else:
                w = apply_A(M(V[k]))
            for i in range(k + 1):
                H[i, k] = np.dot(w, V[i]); w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] != 0:
                V[k + 1] = w / H[k + 1, k]
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k] = d; H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k]
            its += 1; k_done = k + 1
            if abs(g[k + 1]) <= max(rtol * ref, atol):
                converged = True
                break
            if its >= max_it:
                break
        y = np.linalg.solve(np.triu(H[:k_done, :k_done]), g[:k_done])
        upd = y @ V[:k_done]
        x = x + (upd if side == "left" else M(upd))
        if converged or its >= max_it:
            return x, its, converged


def richardson(A, b, x0, M, rtol=1e-5, atol=1e-50, max_it=10000):
    """KSPRICHARDSON with -ksp_norm_type unpreconditioned."""
    x = x0.copy()
    r = b - A @ x
    bn = np.linalg.norm(b)
    ref = bn if bn > 0 else np.linalg.norm(r)
    for it in range(max_it + 1):
        if np.linalg.norm(r) <= max(rtol * ref, atol):
            return x, it, True
        if it == max_it:
            break
        x = x + M(r)
        r = b - A @ x
    return x, max_it, False
