"""Named test hierarchies (inputs for both the oracle and the CUDA path).

Each case restates a configuration the reference's own tests run (tests/Makefile line cited) or
an option of the apply path.  Built by hiergen (setup stand-in), seeded.
"""
import functools
import numpy as np

import hiergen
from hiergen import poly
from hiergen import AirOptions as O


def _adv2(n):
    return hiergen.adv_diff_fd(n, n)


def _mat_stream(part="A"):
    """The reference's own data fixture tests/data/mat_stream_2364 (unstructured streaming problem, 2364 rows, 15948 nonzeros) and
    its rhs, as converted by tests/golden/make_golden.py (the GPU box has no /root/reference)."""
    import os
    import scipy.sparse as sp
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mat_stream_2364_system.npz"))
    if part == "b":
        return z["b"]
    n = z["b"].size
    return sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))


CASES = {
    # tests/Makefile:537-540 (adv_1d, Newton matrix-free coarse solver, power-basis smoother)
    "adv1d_makefile": lambda: (hiergen.adv_1d(1000), O(coarsest_inverse_type=poly.NEWTON, coarsest_poly_order=10,
                                                       coarsest_matrix_free_polys=True, a_drop=1e-3, inverse_type=poly.POWER)),
    "adv1d_default": lambda: (hiergen.adv_1d(1000), O()),
    "adv1d_order4": lambda: (hiergen.adv_1d(1000), O(poly_order=4)),          # BASELINE.json configs[0]
    # tests/Makefile:1322-1323 (run_check)
    "fd2d_25": lambda: (_adv2(25), O()),
    "fd2d_64": lambda: (_adv2(64), O()),
    # tests/Makefile:1128-1134 scaling-study options
    "fd2d_100_study": lambda: (_adv2(100), O(a_lump=True, a_drop=1e-5, strong_threshold=0.99)),
    # tests/Makefile:543-546
    "fd3d_10_lump": lambda: (hiergen.adv_diff_fd(10, 10, 10), O(a_lump=True)),
    "fd3d_12_diffusion": lambda: (hiergen.adv_diff_fd(12, 12, 12, alpha=0.1), O()),
    "fd2d_diffusion": lambda: (hiergen.adv_diff_fd(40, 40, alpha=1.0), O()),
    # -pc_air_matrix_free_polys (Horner), +diag scale
    "fd2d_mf_arnoldi": lambda: (_adv2(48), O(matrix_free_polys=True)),
    "fd2d_mf_power_dscale": lambda: (_adv2(48), O(matrix_free_polys=True, inverse_type=poly.POWER, diag_scale_polys=True)),
    "fd2d_mf_newton": lambda: (_adv2(48), O(matrix_free_polys=True, inverse_type=poly.NEWTON)),
    "fd2d_mf_newton_noextra_dscale": lambda: (_adv2(40), O(matrix_free_polys=True, inverse_type=poly.NEWTON_NO_EXTRA,
                                                            diag_scale_polys=True)),
    "fd2d_mf_neumann": lambda: (_adv2(40), O(matrix_free_polys=True, inverse_type=poly.NEUMANN)),
    "fd2d_neumann": lambda: (_adv2(40), O(inverse_type=poly.NEUMANN)),
    "fd2d_jacobi": lambda: (_adv2(40), O(inverse_type=poly.JACOBI)),
    "fd2d_wjacobi": lambda: (hiergen.adv_diff_fd(30, 30, alpha=1.0), O(inverse_type=poly.WJACOBI)),
    "fd2d_sparsity0": lambda: (_adv2(40), O(inverse_sparsity_order=0)),
    "fd2d_sparsity2": lambda: (_adv2(32), O(inverse_sparsity_order=2)),
    # smoothing orders (-pc_air_smooth_type): fc, cf, fcf, ffc
    "fd2d_fc": lambda: (_adv2(32), O(smooth_order=(1, -1))),
    "fd2d_cf": lambda: (_adv2(32), O(smooth_order=(-1, 1))),
    "fd2d_fcf": lambda: (_adv2(32), O(smooth_order=(1, -1, 1))),
    "fd2d_ffcc_mf": lambda: (_adv2(32), O(smooth_order=(2, -2), matrix_free_polys=True)),
    "fd2d_f1": lambda: (_adv2(32), O(smooth_order=(1,))),
    # ideal (non one-point) prolongator, symmetric
    "fd2d_idealW": lambda: (hiergen.adv_diff_fd(32, 32, alpha=1.0), O(one_point_classical_prolong=False)),
    # strong threshold 0 -> diagonal A_ff on every level, 1 F smooth
    "fd2d_diagAff": lambda: (_adv2(32), O(strong_threshold=0.0)),
    # truncated hierarchy + high-order matrix-free Newton coarse solve (docs/gpus.md:22-40)
    "fd2d_trunc_newton": lambda: (_adv2(64), O(max_levels=4, coarsest_inverse_type=poly.NEWTON, coarsest_poly_order=10,
                                               coarsest_matrix_free_polys=True)),
    "fd2d_two_level": lambda: (_adv2(20), O(max_levels=2)),
    "fd2d_coarse_mf_dscale": lambda: (_adv2(32), O(coarsest_matrix_free_polys=True, coarsest_diag_scale_polys=True)),
    # DG upwind surrogate (configs[2]), matrix-free smoothing
    "dg_mf": lambda: (hiergen.dg_upwind_surrogate(24, 24, 3), O(matrix_free_polys=True)),
    "dg_assembled": lambda: (hiergen.dg_upwind_surrogate(16, 16, 4), O()),
    # -pc_air_full_smoothing_up_and_down (src/AIR_MG_Setup.F90:978-1074): PCMG multiplicative V(1,1), residual restriction
    # R (b - A x), the smoother inverts the whole level matrix; assembled, matrix-free Horner / Newton, Jacobi
    "fd2d_full": lambda: (hiergen.adv_diff_fd(40, 40, alpha=0.5), O(full_smoothing_up_and_down=True)),
    "fd2d_full_mf": lambda: (hiergen.adv_diff_fd(40, 40, alpha=0.5), O(full_smoothing_up_and_down=True, matrix_free_polys=True)),
    "fd2d_full_mf_newton": lambda: (_adv2(40), O(full_smoothing_up_and_down=True, matrix_free_polys=True, inverse_type=poly.NEWTON)),
    "fd2d_full_jacobi": lambda: (hiergen.adv_diff_fd(32, 32, alpha=1.0), O(full_smoothing_up_and_down=True, inverse_type=poly.JACOBI)),
    # the reference's runs on its data fixture mat_stream_2364 (tests/Makefile:89-95,113,208; all with -ksp_max_it 5)
    "ms2364_default": lambda: (_mat_stream(), O()),
    "ms2364_power_fcf": lambda: (_mat_stream(), O(a_drop=1e-3, inverse_type=poly.POWER, smooth_order=(1, -1, 1))),
    "ms2364_power_mf": lambda: (_mat_stream(), O(a_drop=1e-3, inverse_type=poly.POWER, matrix_free_polys=True)),
    "ms2364_power_lair": lambda: (_mat_stream(), O(a_drop=1e-3, inverse_type=poly.POWER, z_type="lair")),
    "ms2364_newton_mf": lambda: (_mat_stream(), O(a_drop=1e-3, inverse_type=poly.NEWTON, matrix_free_polys=True)),
}

MAT_STREAM_CASES = ["ms2364_default", "ms2364_power_fcf", "ms2364_power_mf", "ms2364_power_lair", "ms2364_newton_mf"]

# Not part of CASES (the generic per-case tests compare against an explicitly evaluated polynomial, which these order-18 / order-60
# coarse polynomials are too ill-conditioned for): used by the exact-solver pins and the coarse-iteration tests only.
EXACT_CASES = {
    # AIRG as an exact solver on the same fixture (tests/Makefile:132-145): no dropping, Jacobi smoothing, truncated hierarchy, high-order
    # matrix-free coarse polynomial; the first one needs -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it 5 (option mg_coarse_ksp_max_it)
    "ms2364_exact_arnoldi18": lambda: (_mat_stream(), O(strong_threshold=0.0, a_drop=0.0, r_drop=0.0, inverse_type=poly.JACOBI, max_levels=30,
                                                         coarsest_poly_order=18, coarsest_matrix_free_polys=True, coarsest_inverse_type=poly.ARNOLDI)),
    "ms2364_exact_newton60": lambda: (_mat_stream(), O(strong_threshold=0.0, a_drop=0.0, r_drop=0.0, inverse_type=poly.JACOBI, max_levels=10,
                                                        coarsest_poly_order=60, coarsest_matrix_free_polys=True, coarsest_inverse_type=poly.NEWTON)),
}

FULL_CASES = ["fd2d_full", "fd2d_full_mf", "fd2d_full_mf_newton", "fd2d_full_jacobi"]

# cases small enough for the CPU-only suite and the golden fixtures
GOLDEN = ["adv1d_makefile", "fd2d_25", "fd3d_10_lump", "fd2d_mf_newton", "fd2d_fcf", "dg_mf"]


@functools.lru_cache(maxsize=None)
def build(name):
    A, opts = (CASES[name] if name in CASES else EXACT_CASES[name])()
    H = hiergen.build_hierarchy(A, opts)
    return A, H


def rhs(n, seed=1234):
    """Seeded uniform(0,1) rhs for apply parity."""
    return np.random.default_rng(seed).random(n)


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


PFLAREINV_CASES = {
    # tests/Makefile:548-553
    "inv_newton_5_o16": lambda: (_adv2(5), dict(inverse_type=poly.NEWTON, poly_order=16, matrix_free=True)),
    "inv_newton_10_o50": lambda: (_adv2(10), dict(inverse_type=poly.NEWTON, poly_order=50, matrix_free=True)),
    "inv_arnoldi_asm": lambda: (_adv2(30), dict(inverse_type=poly.ARNOLDI, poly_order=6, matrix_free=False)),
    "inv_arnoldi_mf": lambda: (_adv2(30), dict(inverse_type=poly.ARNOLDI, poly_order=6, matrix_free=True)),
    "inv_power_mf": lambda: (_adv2(30), dict(inverse_type=poly.POWER, poly_order=6, matrix_free=True)),
    "inv_neumann_mf": lambda: (_adv2(30), dict(inverse_type=poly.NEUMANN, poly_order=6, matrix_free=True)),
    "inv_newton_noextra_mf": lambda: (_adv2(30), dict(inverse_type=poly.NEWTON_NO_EXTRA, poly_order=6, matrix_free=True)),
    "inv_jacobi": lambda: (_adv2(30), dict(inverse_type=poly.JACOBI)),
}


@functools.lru_cache(maxsize=None)
def build_inv(name):
    A, kw = PFLAREINV_CASES[name]()
    return A, hiergen.build_pflareinv(A, **kw)

