"""GPU suite: parity of the CUDA path (through the C-ABI) against the CPU oracle on identical operators.

Bar (BASELINE.json north_star): integer/index data bit-exact; fp64 V-cycle output relative L2
difference <= 1e-12; outer GMRES iteration count equal (+-1) to the oracle's / within the reference's
pinned bound.
"""
import os

import numpy as np
import pytest

import cases
import hiergen
import oracle
import pflare_b200
from hiergen import io as hio, poly
from krylov import gmres, richardson

pytestmark = pytest.mark.gpu

TOL = 1e-12          # north_star: relative L2 difference of the fp64 V-cycle output
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle(H):
    return hiergen.feed(H, oracle.OracleAIR(H.no_levels))


def _device(H, **opts):
    d = pflare_b200.DeviceAIR(H.no_levels)
    for k, v in opts.items():
        d.set_option(k, v)
    hiergen.feed(H, d)
    return d


@pytest.mark.parametrize("name", sorted(cases.CASES))
@pytest.mark.parametrize("mode", [dict(dense_rows=0), dict(dense_rows=4096), dict(fuse_perm=2), dict(fuse_perm=2, dense_rows=0, engine=0)],
                         ids=["sparse_all_levels", "dense_tail_default", "fused_entry_exit_permutation", "fused_permutation_tma_ring_engine"])
def test_vcycle_parity(built_libs, name, mode):
    A, H = cases.build(name)
    b = cases.rhs(A.shape[0])
    xo = _oracle(H).apply(b)
    d = _device(H, **mode)
    x = d.apply(b)
    assert cases.rel_l2(x, xo) <= TOL, name
    # repeated applies are deterministic and do not depend on leftover state
    x2 = d.apply(b)
    assert np.array_equal(x, x2)
    d.close()


@pytest.mark.parametrize("opts", [dict(graph=0), dict(dense_rows=0), dict(fuse=0, dense_rows=0), dict(pdl=0), dict(pdl=0, dense_rows=0),
                                  dict(graph=0, fuse=0, dense_rows=0),
                                  # SpMV kernels: 0 = smem-staged CSR stream kernel, 1 = round-1 TMA kernel (CTA tiles), 2 = warp tiles (default)
                                  dict(kernel=0, dense_rows=0), dict(kernel=1, dense_rows=0), dict(kernel=2, dense_rows=0),
                                  dict(kernel=1, ctas_per_sm=1),
                                  # kernel 2 engines: 1 = direct (no shared memory, default), 0 = TMA ring; entry / exit permutation fused into level 1
                                  dict(engine=0, dense_rows=0), dict(engine=0), dict(engine=0, max_ctas=1, dense_rows=0), dict(engine=1, max_ctas=1, dense_rows=0),
                                  dict(engine=1, dense_rows=0), dict(engine=1), dict(wt_format=1, dense_rows=0), dict(wt_format=2, dense_rows=0), dict(wt_format=1, max_ctas=1, dense_rows=0), dict(engine=1, fuse_perm=2),
                                  dict(fuse_perm=2), dict(fuse_perm=2, graph=0, pdl=0), dict(fuse_perm=2, epi_classes=0, dense_rows=0), dict(fuse_perm=2, kernel=0), dict(fuse_perm=2, kernel=1, dense_rows=0),
                                  dict(fuse_perm=0),
                                  # warp-tile kernel: ring depth, generic (run-time branched) epilogue instead of the compiled classes,
                                  # ONE persistent CTA per SM / in total (many tiles per warp: the mbarrier ring wraps many times)
                                  dict(epi_classes=0, dense_rows=0), dict(epi_classes=0, fuse=0, dense_rows=0),
                                  dict(ctas_per_sm=1, dense_rows=0), dict(max_ctas=1, dense_rows=0), dict(max_ctas=1, dense_rows=0, graph=0, pdl=0),
                                  dict(kernel=1, max_ctas=2, dense_rows=0),
                                  # coarse levels collapsed into one dense operator (built from the same kernels at setup)
                                  dict(dense_rows=600), dict(dense_rows=16384), dict(dense_rows=300, graph=0)],
                         ids=lambda o: ",".join("%s=%g" % kv for kv in o.items()))
@pytest.mark.parametrize("name", ["fd2d_64", "fd2d_mf_newton", "fd2d_fcf", "fd2d_diagAff", "dg_mf", "fd2d_idealW",
                                  "fd2d_mf_neumann", "fd2d_mf_newton_noextra_dscale"])
def test_vcycle_parity_execution_modes(built_libs, name, opts):
    A, H = cases.build(name)
    b = cases.rhs(A.shape[0], seed=7)
    xo = _oracle(H).apply(b)
    d = _device(H, **opts)
    assert cases.rel_l2(d.apply(b), xo) <= TOL
    d.close()


@pytest.mark.parametrize("name", cases.GOLDEN)
def test_golden_fixtures(built_libs, name):
    H, g = hio.load(os.path.join(GOLD, name + ".npz"))
    d = _device(H)
    assert cases.rel_l2(d.apply(g["b"]), g["x_oracle"]) <= TOL
    d.close()


@pytest.mark.parametrize("name", ["fd2d_25", "fd2d_full_mf"])
def test_golden_petsc_binary_container(built_libs, name):
    """A hierarchy handed over as PETSc binary objects (the container a PFLARE build dumps, pflare_b200/petsc_io.py)."""
    from pflare_b200 import petsc_io
    H = petsc_io.load_hierarchy(os.path.join(GOLD, name + ".petsc"))
    pc = pflare_b200.PC().setType("air").setHierarchy(H)
    assert cases.rel_l2(pc.apply(H.b), H.x) <= TOL
    pc.destroy()


def test_golden_ilu_factors_pflareinv_newton(built_libs):
    """configs[4]: PCPFLAREINV Newton-basis polynomial (matrix-free) on the ParILU factors of the reference's
    own fixture tests/data/mat_stream_2364."""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLD, "ilu_mat_stream.npz"))
    n = z["b"].size
    for nm in ("L", "U"):
        T = sp.csr_matrix((z[nm + "_data"], z[nm + "_indices"], z[nm + "_indptr"]), shape=(n, n))
        H = hiergen.build_pflareinv(T, poly.NEWTON, 6, 1, True)
        H.inv_coarse.coeffs = z[nm + "_roots"]
        pc = pflare_b200.PC().setType("pflareinv").setHierarchy(H)
        y = pc.apply(z["b"])
        assert cases.rel_l2(y, z[nm + "_y_oracle"]) <= TOL
        _, its, conv = richardson(T, z["b"], np.zeros(n), pc.apply, rtol=1e-6, max_it=100)
        assert conv and its <= 5
        pc.destroy()


@pytest.mark.parametrize("name", sorted(cases.PFLAREINV_CASES))
def test_pflareinv_parity(built_libs, name):
    A, H = cases.build_inv(name)
    x = cases.rhs(A.shape[0])
    yo = _oracle(H).inv_apply(1, oracle.INV_AFF, x)
    pc = pflare_b200.PC().setType("pflareinv").setHierarchy(H)
    assert cases.rel_l2(pc.apply(x), yo) <= TOL
    pc.destroy()


@pytest.mark.parametrize("name", ["fd2d_64", "fd2d_fcf", "fd2d_cf", "fd2d_mf_neumann", "fd2d_ffcc_mf"])
def test_level_smoother_and_inverse_parity(built_libs, name):
    """Seams 2 and 3 of SURVEY.md section 8b: one mg_FC_point_richardson on a level, one inverse apply on a level."""
    A, H = cases.build(name)
    O = _oracle(H)
    d = _device(H)
    for l in (1, 2, H.no_levels - 1):
        lv = H.levels[l - 1]
        b, x0 = cases.rhs(lv.n, seed=l), cases.rhs(lv.n, seed=100 + l)
        assert cases.rel_l2(d.fc_smooth(l, b, x0), O.fc_smooth(l, b, x0)) <= TOL
        v = cases.rhs(lv.is_fine.size, seed=200 + l)
        assert cases.rel_l2(d.inv_apply(l, pflare_b200.INV_AFF, v), O.inv_apply(l, oracle.INV_AFF, v)) <= TOL
        if lv.inv_A_cc is not None:
            v = cases.rhs(lv.is_coarse.size, seed=300 + l)
            assert cases.rel_l2(d.inv_apply(l, pflare_b200.INV_ACC, v), O.inv_apply(l, oracle.INV_ACC, v)) <= TOL
    d.close()


def test_64bit_petscint_upload(built_libs):
    """The *_i64 upload entry points (PetscInt = 64-bit builds) give the same V-cycle."""
    A, H = cases.build("fd2d_fcf")
    b = cases.rhs(A.shape[0])
    d = pflare_b200.DeviceAIR(H.no_levels, idx64=True)
    hiergen.feed(H, d)
    assert cases.rel_l2(d.apply(b), _oracle(H).apply(b)) <= TOL
    d.close()


def test_integer_data_round_trip_bit_exact(built_libs):
    A, H = cases.build("fd2d_64")
    d = _device(H)
    for l, lv in enumerate(H.levels, start=1):
        assert np.array_equal(d.get_is(l, 0), lv.is_fine)
        assert np.array_equal(d.get_is(l, 1), lv.is_coarse)
    d.close()


@pytest.mark.parametrize("name,rtol,side,bound", [("adv1d_makefile", 1e-10, "right", 2), ("fd2d_25", 1e-5, "left", 5),
                                                  ("fd3d_10_lump", 1e-10, "right", 4), ("fd2d_100_study", 1e-10, "right", 6)])
def test_outer_gmres_iteration_counts(built_libs, name, rtol, side, bound):
    A, H = cases.build(name)
    n = A.shape[0]
    pc = pflare_b200.PC().setType("air").setHierarchy(H)
    _, its_gpu, conv = gmres(A, np.zeros(n), np.ones(n), pc.apply, rtol=rtol, side=side)
    _, its_cpu, _ = gmres(A, np.zeros(n), np.ones(n), _oracle(H).apply, rtol=rtol, side=side)
    assert conv and its_gpu <= bound and abs(its_gpu - its_cpu) <= 1
    pc.destroy()


@pytest.mark.parametrize("name,rtol,side,bound", [("adv1d_makefile", 1e-10, "right", 2), ("fd2d_25", 1e-5, "left", 5),
                                                  ("fd3d_10_lump", 1e-10, "right", 4), ("fd2d_100_study", 1e-10, "right", 6),
                                                  ("fd2d_full_mf", 1e-8, "right", 40)])
def test_outer_krylov_on_device(built_libs, name, rtol, side, bound):
    """KSPSolve on the device (pflare_b200_ksp_solve: GMRES(30) around the V-cycle, b = 0, x0 = 1 like the reference's drivers):
    iteration count equal (+-1) to the host loop around the oracle and within the reference's -ksp_max_it bound, and the
    solution solves the system."""
    A, H = cases.build(name)
    n = A.shape[0]
    d = pflare_b200.DeviceAIR(H.no_levels)
    d.ksp_set_operator(A)
    hiergen.feed(H, d)
    x, its_dev, conv, rn = d.ksp_solve(np.zeros(n), np.ones(n), ksp_type="gmres", side=side, rtol=rtol)
    _, its_cpu, _ = gmres(A, np.zeros(n), np.ones(n), _oracle(H).apply, rtol=rtol, side=side)
    assert conv and its_dev <= bound and abs(its_dev - its_cpu) <= 1, (its_dev, its_cpu)
    if side == "right":
        assert np.linalg.norm(A @ x) <= 10 * rtol * np.linalg.norm(A @ np.ones(n))
    # a restart shorter than the iteration count exercises the restart path
    x2, its2, conv2, _ = d.ksp_solve(np.zeros(n), np.ones(n), ksp_type="gmres", side=side, rtol=rtol, restart=2)
    _, its2_cpu, _ = gmres(A, np.zeros(n), np.ones(n), _oracle(H).apply, rtol=rtol, side=side, restart=2)
    assert conv2 and abs(its2 - its2_cpu) <= 1
    d.close()


def test_outer_richardson_on_device_pflareinv(built_libs):
    """KSPRICHARDSON on the device around a PCPFLAREINV handle (tests/Makefile:548-553 style)."""
    A, H = cases.build_inv("inv_arnoldi_asm")
    n = A.shape[0]
    b = cases.rhs(n)
    d = pflare_b200.DeviceAIR(1)
    d.ksp_set_operator(A)
    hiergen.feed(H, d)
    x, its, conv, rn = d.ksp_solve(b, np.zeros(n), ksp_type="richardson", rtol=1e-6, max_it=200)
    O = _oracle(H)
    _, its_cpu, conv_cpu = richardson(A, b, np.zeros(n), lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-6, max_it=200)
    assert conv == conv_cpu and abs(its - its_cpu) <= 1
    if conv:
        assert np.linalg.norm(b - A @ x) <= 1e-5 * np.linalg.norm(b)
    d.close()


def test_edge_cases(built_libs):
    # zero rhs -> exactly zero; a single-level PCAIR is refused; wrong vector length is an error
    A, H = cases.build("fd2d_25")
    pc = pflare_b200.PC().setType("air").setHierarchy(H)
    assert not np.any(pc.apply(np.zeros(A.shape[0])))
    with pytest.raises(ValueError):
        pc.apply(np.zeros(A.shape[0] + 1))
    pc.destroy()
    H1 = hiergen.build_pflareinv(A)
    with pytest.raises(pflare_b200.PflareB200Error):
        pflare_b200.PC().setType("air").setHierarchy(H1).setUp()
    # two PCs alive at once (tests/ex6_two_airg.c: per-instance handles)
    A2, H2 = cases.build("fd2d_64")
    p1 = pflare_b200.PC().setType("air").setHierarchy(H)
    p2 = pflare_b200.PC().setType("air").setHierarchy(H2)
    b1, b2 = cases.rhs(A.shape[0]), cases.rhs(A2.shape[0])
    x1, x2 = p1.apply(b1), p2.apply(b2)
    assert cases.rel_l2(x1, _oracle(H).apply(b1)) <= TOL and cases.rel_l2(x2, _oracle(H2).apply(b2)) <= TOL
    p1.destroy()
    assert cases.rel_l2(p2.apply(b2), x2) == 0.0
    p2.destroy()


@pytest.mark.parametrize("opts", [dict(), dict(wt_format=1), dict(wt_format=2), dict(wt_format=2, engine=0),
                                  dict(kernel=0), dict(kernel=1)], ids=str)
def test_rows_longer_than_a_tile(built_libs, opts):
    """A dense-ish operator: the rows longer than a warp tile (> 256 nonzeros) are left out of the tiles and run on the
    CSR stream kernel (whole-CTA reduction), the others stay on the warp-tile kernels."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n = 3000
    D = sp.random(n, n, density=0.002, random_state=5, format="lil")
    D[7, :] = rng.random(n)            # one 3000-nnz row
    D[11, :2500] = rng.random(2500)
    D[12, :300] = rng.random(300)      # just above a tile
    D[2999, 100:700] = rng.random(600)
    A = (D.tocsr() + sp.identity(n) * 50.0).tocsr()
    H = hiergen.build_pflareinv(A, poly.ARNOLDI, 4, 1, True)
    x = cases.rhs(n)
    yo = _oracle(H).inv_apply(1, oracle.INV_AFF, x)
    pc = pflare_b200.PC().setType("pflareinv").setHierarchy(H)
    for k, v in opts.items():
        pc.setOption(k, v)
    assert cases.rel_l2(pc.apply(x), yo) <= TOL
    pc.destroy()


def test_vcycle_with_long_rows_on_the_coarse_levels(built_libs):
    """3D diffusion-dominated hierarchy without the dense tail: the coarse operators have rows of several hundred nonzeros."""
    A = hiergen.adv_diff_fd(14, 14, 14, alpha=1.0)
    H = hiergen.build_hierarchy(A, hiergen.AirOptions(a_drop=0.0, r_drop=0.0))
    longest = max(int(np.diff(m.indptr).max()) for lv in H.levels for m in (lv.A_ff, lv.A_fc, lv.R))
    assert longest > 256, longest
    b = cases.rhs(A.shape[0])
    xo = _oracle(H).apply(b)
    for opts in (dict(dense_rows=0), dict(dense_rows=0, wt_format=1), dict(dense_rows=0, wt_format=2, engine=0)):
        d = _device(H, **opts)
        assert cases.rel_l2(d.apply(b), xo) <= TOL, opts
        d.close()


def test_medium_size_parity_and_linearity(built_libs):
    """512^2 (262k rows, ~30 levels): parity against the oracle plus linearity of PCApply."""
    A = hiergen.adv_diff_fd(512, 512)
    H = hiergen.build_hierarchy(A, hiergen.AirOptions())
    n = A.shape[0]
    d = _device(H)
    b1, b2 = cases.rhs(n, 1), cases.rhs(n, 2)
    x1, x2 = d.apply(b1), d.apply(b2)
    assert cases.rel_l2(x1, _oracle(H).apply(b1)) <= TOL
    x12 = d.apply(2.5 * b1 - b2)
    assert cases.rel_l2(x12, 2.5 * x1 - x2) <= 1e-11
    d.set_option("dense_rows", 4096)        # collapsed coarse tail: same answer to round-off
    assert cases.rel_l2(d.apply(b1), x1) <= TOL
    st = d.stats()
    assert st["kernel_launches"] > 0 and st["algorithmic_bytes"] > 12 * st["nnz_per_cycle"]
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["L", "U"])
def test_ilu_factor_richardson_on_device(built_libs, which):
    """BASELINE.json configs[4] (tests/ilu_factors.c:175-196): the Richardson solve (rtol 1e-6, unpreconditioned norm) the reference
    test runs with PCPFLAREINV (Newton basis, matrix-free, order 6) on the ParILU(0) factors of its fixture mat_stream_2364, here
    entirely on the device (pflare_b200_ksp_solve): converges, same iteration count as the host loop around the oracle."""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLD, "ilu_mat_stream.npz"))
    n = z["b"].size
    T = sp.csr_matrix((z[which + "_data"], z[which + "_indices"], z[which + "_indptr"]), shape=(n, n))
    H = hiergen.build_pflareinv(T, poly.NEWTON, 6, 1, True)
    H.inv_coarse.coeffs = z[which + "_roots"]
    O = _oracle(H)
    _, its_cpu, conv_cpu = richardson(T, z["b"], np.zeros(n), lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-6, max_it=2000)
    d = pflare_b200.DeviceAIR(1)
    d.ksp_set_operator(T)
    hiergen.feed(H, d)
    x, its, conv, rn = d.ksp_solve(z["b"], np.zeros(n), ksp_type="richardson", rtol=1e-6, max_it=2000)
    assert conv and conv_cpu and abs(its - its_cpu) <= 1 and its <= 5
    assert np.linalg.norm(z["b"] - T @ x) <= 1e-6 * np.linalg.norm(z["b"])
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("order,bound", [(60, 6), (120, 5)])
def test_1138_bus_high_order_newton_on_device(built_libs, order, bound):
    """tests/Makefile:199-205 on the reference's fixture data/1138_bus: PCPFLAREINV Newton basis, matrix-free, order 60 / 120 with
    the added roots (86 / 239 roots = as many fused SpMV launches per apply).  The device apply matches the frozen oracle vector (1e-12
    at order 60; at order 120 within the chain's measured rounding sensitivity, see below) and the right-preconditioned GMRES of the reference's run (b = 0, random x0, rtol 1e-5) stays within its -ksp_max_it."""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLD, "bus1138_newton.npz"))
    n = z["x0"].size
    A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    H = hiergen.build_pflareinv(A, poly.NEWTON, order, 1, True)
    H.inv_coarse.coeffs = z["roots_%d" % order]
    d = pflare_b200.DeviceAIR(1)
    d.ksp_set_operator(A)
    hiergen.feed(H, d)
    O = _oracle(H)
    # Order 60 meets the 1e-12 bar.  The order-120 chain (239 products) amplifies rounding: the ORACLE evaluated on the symmetrically
    # permuted matrix (same polynomial, same arithmetic, only the order of the row sums changes) already differs from itself by
    # ~4e-11, so the bar for that case is 20 x this measured rounding sensitivity (the device sums rows in tile order).
    p = np.random.default_rng(5).permutation(n)
    Ap = A[p][:, p].tocsr()
    Ap.sort_indices()
    Hp = hiergen.build_pflareinv(Ap, poly.NEWTON, order, 1, True)
    Hp.inv_coarse.coeffs = z["roots_%d" % order]
    yp = np.empty(n)
    yp[p] = _oracle(Hp).inv_apply(1, oracle.INV_AFF, z["v"][p])
    noise = cases.rel_l2(yp, z["y_oracle_%d" % order])
    assert noise <= (TOL if order == 60 else 1e-9)
    assert cases.rel_l2(d.inv_apply(1, pflare_b200.INV_AFF, z["v"]), z["y_oracle_%d" % order]) <= max(TOL, 20 * noise)
    _, its_cpu, _ = gmres(A, np.zeros(n), z["x0"], lambda v: O.inv_apply(1, oracle.INV_AFF, v), rtol=1e-5, side="right")
    x, its, conv, rn = d.ksp_solve(np.zeros(n), z["x0"], ksp_type="gmres", side="right", rtol=1e-5)
    assert conv and its <= bound and abs(its - its_cpu) <= 1
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [dict(dense_rows=0), dict()], ids=["sparse_all_levels", "dense_tail_default"])
def test_coarse_richardson_iterations(built_libs, mode):
    """-mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it 5 (option mg_coarse_ksp_max_it; tests/Makefile:132-136 on the reference's
    fixture mat_stream_2364): parity with the oracle, set before and after finalize_setup, and the outer Richardson of the reference's
    run converges in one iteration only with the five coarse sweeps."""
    A, H = cases.build("ms2364_exact_arnoldi18")
    b = cases._mat_stream("b")
    n = A.shape[0]
    O5 = _oracle(H)
    O5.set_option("mg_coarse_ksp_max_it", 5)
    x5, x1 = O5.apply(b), _oracle(H).apply(b)
    d = _device(H, mg_coarse_ksp_max_it=5, **mode)
    x = d.apply(b)
    assert cases.rel_l2(x, x5) <= TOL
    assert np.linalg.norm(b - A @ x) <= 1e-5 * np.linalg.norm(b) < np.linalg.norm(b - A @ x1)
    # back to one sweep on the live handle (the program and the collapsed tail are rebuilt).  A single application of the order-18
    # coarse polynomial is rounding-sensitive at the 5e-9 level (the oracle differs from itself by that much when only the order of
    # its row sums changes), so this leg is compared at 1e-6; the five sweeps above contract that error away, hence TOL there.
    d.set_option("mg_coarse_ksp_max_it", 1)
    y1 = d.apply(b)
    assert cases.rel_l2(y1, x1) <= 1e-6 and np.linalg.norm(b - A @ y1) > 1e-5 * np.linalg.norm(b)
    d.set_option("mg_coarse_ksp_max_it", 5)
    assert cases.rel_l2(d.apply(b), x5) <= TOL
    d.close()


@pytest.mark.gpu
def test_resetup_on_a_live_handle(built_libs):
    """The reference's re-setup (src/PCAIR_Shell.F90:148-162): the upload hook runs again on the SAME handle with new operators and
    finalize_setup releases everything the previous setup put on the device (operators, plans, dense tail, graph) before it rebuilds --
    old and new state are never mixed (the collapsed coarse tail in particular is rebuilt from the new values)."""
    A, H = cases.build("fd2d_64")
    H2 = hiergen.build_hierarchy(1.7 * A, cases.CASES["fd2d_64"]()[1])
    b = cases.rhs(A.shape[0])
    x1, x2 = _oracle(H).apply(b), _oracle(H2).apply(b)
    assert cases.rel_l2(x1, x2) > 0.1
    d = _device(H)
    assert cases.rel_l2(d.apply(b), x1) <= TOL
    before = d.stats()["device_bytes"]
    hiergen.feed(H2, d)
    assert cases.rel_l2(d.apply(b), x2) <= TOL
    hiergen.feed(H, d)
    assert cases.rel_l2(d.apply(b), x1) <= TOL
    assert abs(d.stats()["device_bytes"] - before) <= 0.01 * before          # nothing of the two earlier setups is still allocated
    d.close()
