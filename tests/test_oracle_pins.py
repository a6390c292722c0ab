"""CPU suite, part 1: pin the oracle (oracle/air_oracle.c).

The reference holds no vector-level golden outputs for PCApply; what its tests DO pin
are iteration-count bounds (`-ksp_max_it N` in tests/Makefile, exit code 1 when not converged).  The
oracle -- fed by hiergen, the restated setup -- must meet every one of those bounds for the configs
BASELINE.json names; it is also cross-checked against an independent scipy evaluation of the cycle
and frozen by the fixtures of tests/golden/.
"""
import os

import numpy as np
import pytest

import cases
import hiergen
import oracle
import pycycle
from hiergen import io as hio, poly
from krylov import gmres, richardson

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle(H):
    return hiergen.feed(H, oracle.OracleAIR(H.no_levels))


def _its(A, M, rtol, side):
    n = A.shape[0]
    _, its, conv = gmres(A, np.zeros(n), np.ones(n), M, rtol=rtol, side=side)   # b = 0, x0 = 1 like the drivers
    return its, conv


# ------------------------------------------------------------------ the reference's known answers
def test_pin_adv1d_makefile_537(built_libs):
    A, H = cases.build("adv1d_makefile")
    its, conv = _its(A, _oracle(H).apply, 1e-10, "right")
    assert conv and its <= 2          # tests/Makefile:537-540 -ksp_max_it 2


def test_pin_fd2d_run_check(built_libs):
    A, H = cases.build("fd2d_25")
    its, conv = _its(A, _oracle(H).apply, 1e-5, "left")
    assert conv and its <= 5          # tests/Makefile:1322-1323 -ksp_max_it 5 (PETSc default rtol, left PC)


def test_pin_fd3d_makefile_543(built_libs):
    A, H = cases.build("fd3d_10_lump")
    its, conv = _its(A, _oracle(H).apply, 1e-10, "right")
    assert conv and its <= 4          # tests/Makefile:543-546 -ksp_max_it 4


@pytest.mark.parametrize("n", [100, 200])
def test_pin_fd2d_scaling_study(built_libs, n):
    A = hiergen.adv_diff_fd(n, n)
    H = hiergen.build_hierarchy(A, hiergen.AirOptions(a_lump=True, a_drop=1e-5, strong_threshold=0.99))
    its, conv = _its(A, _oracle(H).apply, 1e-10, "right")
    assert conv and its <= 6          # tests/Makefile:1128-1134 -ksp_max_it 6, grid independent


@pytest.mark.parametrize("name", ["inv_newton_5_o16", "inv_newton_10_o50"])
def test_pin_pflareinv_newton(built_libs, name):
    A, H = cases.build_inv(name)
    O = _oracle(H)
    its, conv = _its(A, lambda v: O.inv_apply(1, oracle.INV_AFF, v), 1e-5, "left")
    assert conv and its <= 1          # tests/Makefile:548-553 -ksp_max_it 1


def test_pin_ilu_factors_newton(built_libs):
    # configs[4]: Richardson + PCPFLAREINV Newton matrix-free on the ParILU factors, rtol 1e-6
    # (tests/ilu_factors.c:122-126, run at tests/Makefile:105-109; the driver fails if not converged)
    z = np.load(os.path.join(GOLD, "ilu_mat_stream.npz"))
    import scipy.sparse as sp
    n = z["b"].size
    for nm in ("L", "U"):
        T = sp.csr_matrix((z[nm + "_data"], z[nm + "_indices"], z[nm + "_indptr"]), shape=(n, n))
        H = hiergen.build_pflareinv(T, poly.NEWTON, 6, 1, True)
        H.inv_coarse.coeffs = z[nm + "_roots"]
        O = _oracle(H)
        y = O.inv_apply(1, oracle.INV_AFF, z["b"])
        assert cases.rel_l2(y, z[nm + "_y_oracle"]) < 1e-13
        _, its, conv = richardson(T, z["b"], np.zeros(n), lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-6, max_it=100)
        assert conv and its <= 5


@pytest.mark.parametrize("name", cases.MAT_STREAM_CASES)
def test_pin_mat_stream_2364(built_libs, name):
    """tests/Makefile:89-95,113,208: `ex12f -f data/mat_stream_2364 [...] -ksp_max_it 5` on the reference's own data fixture
    (matrix AND rhs from the file, zero initial guess, left-preconditioned GMRES, rtol 1e-5; the run fails unless it converges within
    5 iterations): default PCAIR, power basis + a_drop 1e-3 with fcf smoothing / matrix-free smoothing / lAIR Z, Newton matrix-free."""
    A, H = cases.build(name)
    b = cases._mat_stream("b")
    O = _oracle(H)
    _, its, conv = gmres(A, b, np.zeros(A.shape[0]), O.apply, rtol=1e-5, side="left")
    assert conv and its <= 5, (name, its)


@pytest.mark.parametrize("name,coarse_its", [("ms2364_exact_arnoldi18", 5), ("ms2364_exact_newton60", 1)])
def test_pin_mat_stream_exact_solver(built_libs, name, coarse_its):
    """tests/Makefile:132-145: AIRG as an exact solver on data/mat_stream_2364 -- outer `-ksp_type richardson -ksp_norm_type
    unpreconditioned -ksp_max_it 1` must converge (rtol 1e-5) in ONE iteration.  The Arnoldi-18 variant runs
    `-mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it 5`: with a single coarse sweep the residual stalls at 1.3e-5 (> rtol), with
    the five Richardson sweeps of the reference's run it drops to 1e-15 -- this pins the coarse-iteration semantics."""
    A, H = cases.build(name)
    b = cases._mat_stream("b")
    n = A.shape[0]
    O = _oracle(H)
    O.set_option("mg_coarse_ksp_max_it", coarse_its)
    x, its, conv = richardson(A, b, np.zeros(n), O.apply, rtol=1e-5, max_it=1)
    assert conv and its == 1
    if coarse_its > 1:
        O1 = _oracle(H)
        _, _, conv1 = richardson(A, b, np.zeros(n), O1.apply, rtol=1e-5, max_it=1)
        assert not conv1
        assert np.linalg.norm(b - A @ x) <= 1e-12 * np.linalg.norm(b)


@pytest.mark.parametrize("label,kw,bound", [("default arnoldi", dict(inverse_type=poly.ARNOLDI, poly_order=6, matrix_free=False), 21),
                                            ("power", dict(inverse_type=poly.POWER, poly_order=6, matrix_free=False), 21),
                                            ("newton matrix-free", dict(inverse_type=poly.NEWTON, poly_order=6, matrix_free=True), 13)])
def test_pin_mat_stream_pflareinv(built_libs, label, kw, bound):
    """tests/Makefile:119-127: `ex6 -f data/mat_stream_2364 -pc_type pflareinv [-pc_pflareinv_type power | newton
    -pc_pflareinv_matrix_free] -ksp_max_it 21 | 13` (matrix and rhs from the reference's file, GMRES, rtol 1e-5): single-level
    polynomial preconditioning; measured here 20, 20 and 12 iterations."""
    A, b = cases._mat_stream(), cases._mat_stream("b")
    H = hiergen.build_pflareinv(A, **kw)
    O = _oracle(H)
    _, its, conv = gmres(A, b, np.zeros(A.shape[0]), lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-5, side="left")
    assert conv and its <= bound, (label, its)


@pytest.mark.parametrize("label,kw,bound,coarse_its", [
    ("default", dict(), 5, 1),
    ("arnoldi + a_drop", dict(inverse_type=poly.ARNOLDI, coarsest_inverse_type=poly.ARNOLDI, a_drop=1e-3), 5, 1),
    ("neumann", dict(inverse_type=poly.NEUMANN, a_drop=1e-3), 5, 1),
    ("neumann matrix-free", dict(inverse_type=poly.NEUMANN, a_drop=1e-3, matrix_free_polys=True), 5, 1),
    ("wjacobi", dict(inverse_type=poly.WJACOBI, a_drop=1e-3), 8, 1),
    ("jacobi", dict(inverse_type=poly.JACOBI, a_drop=1e-3), 5, 1),
    ("exact solver, 10 coarse sweeps", dict(strong_threshold=0.0, a_drop=0.0, r_drop=0.0, inverse_type=poly.JACOBI), 1, 10)])
def test_pin_diffusion_8x8(built_libs, label, kw, bound, coarse_its):
    """tests/Makefile:386-424: `adv_diff_fd -u 0 -v 0 -alpha 1.0 -da_grid_x 8 -da_grid_y 8 -pc_type air [...] -ksp_max_it N` (b = 0,
    x0 = 1, GMRES rtol 1e-5; the last one `-ksp_type richardson -ksp_norm_type unpreconditioned -mg_coarse_ksp_type richardson
    -mg_coarse_ksp_max_it 10 -ksp_max_it 1`): the polynomial / Neumann / Jacobi inverse families on a diffusion stencil."""
    from hiergen import AirOptions
    A = hiergen.adv_diff_fd(8, 8, alpha=1.0, u=0.0, v=0.0)
    n = A.shape[0]
    H = hiergen.build_hierarchy(A, AirOptions(**kw))
    O = _oracle(H)
    O.set_option("mg_coarse_ksp_max_it", coarse_its)
    if coarse_its > 1:
        _, its, conv = richardson(A, np.zeros(n), np.ones(n), O.apply, rtol=1e-5, max_it=bound)
    else:
        _, its, conv = gmres(A, np.zeros(n), np.ones(n), O.apply, rtol=1e-5, side="left")
    assert conv and its <= bound, (label, its)


@pytest.mark.parametrize("kw", [dict(smooth_order=(1, -1)), dict(smooth_order=(1, -1), c_inverse_sparsity_order=0)], ids=["fc", "fc_c_sparsity0"])
def test_pin_advection_8x8_fc_smoothing(built_libs, kw):
    """tests/Makefile:299-303: `adv_diff_fd -da_grid_x 8 -da_grid_y 8 -pc_type air -ksp_max_it 3 -pc_air_smooth_type fc
    [-pc_air_c_inverse_sparsity_order 0]`: one F and one C smooth per level (c_smooths, src/FC_Smooth.F90:572-640)."""
    from hiergen import AirOptions
    A = hiergen.adv_diff_fd(8, 8)
    n = A.shape[0]
    O = _oracle(hiergen.build_hierarchy(A, AirOptions(**kw)))
    _, its, conv = gmres(A, np.zeros(n), np.ones(n), O.apply, rtol=1e-5, side="left")
    assert conv and its <= 3, its


def test_pin_e05r0100_power(built_libs):
    """tests/Makefile:157: `ex6 -f data/e05r0100_petsc -b_in_f 0 -pc_air_a_drop 1e-3 -pc_air_inverse_type power -ksp_max_it 26` on the
    reference's data fixture (b = 0, random initial guess, GMRES rtol 1e-5): a hard non-symmetric problem where AIRG needs ~23 iterations."""
    import scipy.sparse as sp
    from hiergen import AirOptions
    z = np.load(os.path.join(GOLD, "e05r0100_system.npz"))
    n = z["indptr"].size - 1
    A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    H = hiergen.build_hierarchy(A, AirOptions(a_drop=1e-3, inverse_type=poly.POWER))
    O = _oracle(H)
    _, its, conv = gmres(A, np.zeros(n), np.random.default_rng(0).random(n), O.apply, rtol=1e-5, side="left")
    assert conv and its <= 26, its


def _bus1138(order):
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLD, "bus1138_newton.npz"))
    n = z["x0"].size
    A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    H = hiergen.build_pflareinv(A, poly.NEWTON, order, 1, True)
    H.inv_coarse.coeffs = z["roots_%d" % order]
    return A, H, z


@pytest.mark.parametrize("order,bound", [(60, 6), (120, 5)])
def test_pin_1138_bus_high_order_newton(built_libs, order, bound):
    """tests/Makefile:199-205: `ex6 -b_in_f 0 -f data/1138_bus -pc_type pflareinv -pc_pflareinv_type newton -pc_pflareinv_poly_order
    60|120 -pc_pflareinv_matrix_free -ksp_norm_type unpreconditioned -ksp_max_it 6|5` on the reference's own data fixture (GMRES,
    b = 0, random initial guess, rtol 1e-5; the unpreconditioned norm makes PETSc's GMRES right-preconditioned).  The Newton
    polynomial with the added roots (86 / 239 stored roots, src/Gmres_Poly_Newton.F90:630-700) is the longest product chain the
    apply path runs: the oracle must stay within the reference's iteration bound and reproduce the frozen apply."""
    A, H, z = _bus1138(order)
    O = _oracle(H)
    assert cases.rel_l2(O.inv_apply(1, oracle.INV_AFF, z["v"]), z["y_oracle_%d" % order]) < 1e-13
    n = A.shape[0]
    _, its, conv = gmres(A, np.zeros(n), z["x0"], lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-5, side="right")
    assert conv and its <= bound


# ------------------------------------------------------------------ oracle vs independent evaluation
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_textbook_cycle(built_libs, name):
    A, H = cases.build(name)
    b = cases.rhs(A.shape[0])
    x = _oracle(H).apply(b)
    xr = pycycle.vcycle_full(H, b) if H.options.full_smoothing_up_and_down else pycycle.vcycle(H, b)
    assert cases.rel_l2(x, xr) < 1e-9, name


@pytest.mark.parametrize("name", sorted(cases.PFLAREINV_CASES))
def test_oracle_inverse_matches_explicit_polynomial(built_libs, name):
    A, H = cases.build_inv(name)
    x = cases.rhs(A.shape[0])
    y = _oracle(H).inv_apply(1, oracle.INV_AFF, x)
    yr = pycycle.inv_apply(H.inv_coarse, H.coarse_matrix, x)
    assert cases.rel_l2(y, yr) < 1e-8, name


def test_horner_skips_zero_coefficients(built_libs):
    # src/Gmres_Poly.F90:1468: a zero coefficient skips the matvec AND the shift (not Horner-exact)
    A, H = cases.build_inv("inv_arnoldi_mf")
    H.inv_coarse.coeffs = np.array([[1.0], [0.0], [0.5], [0.25]])
    x = cases.rhs(A.shape[0])
    y = _oracle(H).inv_apply(1, oracle.INV_AFF, x)
    # reference order: y = c3 x; (order 2) y = A y + c2 x; (order 1 skipped); (order 0) y = A y + c0 x
    t = 0.25 * x
    t = A @ t + 0.5 * x
    t = A @ t + 1.0 * x
    assert cases.rel_l2(y, t) < 1e-14


def test_newton_zero_roots_and_final_pair(built_libs):
    A, H = cases.build_inv("inv_newton_noextra_mf")
    # real zero root skipped, complex pair last => second matvec of the pair skipped (:845)
    H.inv_coarse.coeffs = np.array([[2.0, 0.0], [0.0, 0.0], [1.5, 0.5], [1.5, -0.5]])
    x = cases.rhs(A.shape[0])
    y = _oracle(H).inv_apply(1, oracle.INV_AFF, x)
    t = x.copy()
    yy = t / 2.0
    t = t - (A @ t) / 2.0
    sq = 1.5 * 1.5 + 0.25
    u = 3.0 * t - A @ t
    yy = yy + u / sq
    assert cases.rel_l2(y, yy) < 1e-14


# ------------------------------------------------------------------ fixtures
@pytest.mark.parametrize("name", cases.GOLDEN)
def test_golden_fixture_matches_oracle(built_libs, name):
    H, d = hio.load(os.path.join(GOLD, name + ".npz"))
    x = _oracle(H).apply(d["b"])
    assert cases.rel_l2(x, d["x_oracle"]) < 1e-13


@pytest.mark.parametrize("name", cases.GOLDEN)
def test_golden_fixture_is_what_hiergen_builds(built_libs, name):
    """The integer data of the fixture (CF lists) must be reproduced bit-exactly by the seeded generator."""
    H, _ = hio.load(os.path.join(GOLD, name + ".npz"))
    _, H2 = cases.build(name)
    assert H.no_levels == H2.no_levels
    for a, b in zip(H.levels, H2.levels):
        assert np.array_equal(a.is_fine, b.is_fine) and np.array_equal(a.is_coarse, b.is_coarse)
        assert np.array_equal(a.R.indices, b.R.indices) and np.array_equal(a.P.indices, b.P.indices)


def test_cf_splitting_is_a_disjoint_cover(built_libs):
    # structural assertion of tests/ex6_cf_splitting.c:34-64
    A, H = cases.build("fd2d_64")
    for lv in H.levels:
        both = np.concatenate((lv.is_fine, lv.is_coarse))
        assert both.size == lv.n and np.array_equal(np.sort(both), np.arange(lv.n))
        assert np.all(np.diff(lv.is_fine) > 0) and np.all(np.diff(lv.is_coarse) > 0)

