"""CPU suite, part 2: the drop-in boundary.  The C-ABI library loads and exports every symbol that
include/pflare_b200.h declares; the host mirror behaves like the reference's PC interface; without a
GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
import hiergen
import pflare_b200
from pflare_b200 import _capi
from hiergen import io as hio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "pflare_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pflare_b200_\w+)\s*\(", txt)))


def test_library_exports_every_header_symbol(built_libs):
    L = ctypes.CDLL(pflare_b200.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s
    # and the ctypes table covers exactly the header
    assert sorted(_capi.SIGNATURES) == syms


def test_no_oracle_or_cpu_path_in_product():
    """The product package must never import/link the oracle (it is test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pflare_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "air_oracle" not in src, f


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(built_libs):
    with pytest.raises(pflare_b200.PflareB200Error) as e:
        pflare_b200.DeviceAIR(2)
    assert e.value.code == 10 and "no CPU fallback" in str(e.value)
    A, H = cases.build("fd2d_25")
    pc = pflare_b200.PC().setType("air").setHierarchy(H)
    with pytest.raises(pflare_b200.PflareB200Error):
        pc.apply(np.ones(A.shape[0]))


def test_pc_mirror_error_behaviour():
    pc = pflare_b200.PC()
    with pytest.raises(ValueError):
        pc.setType("gamg")                      # only 'air' and 'pflareinv' are registered (src/PCAIR.c:3649-3663)
    with pytest.raises(RuntimeError):
        pc.setUp()                              # PCSetUp before PCSetType
    pc.setType("air")
    with pytest.raises(RuntimeError):
        pc.setUp()                              # no operators
    assert pc.getType() == "air"


class _Recorder:
    def __init__(self):
        self.calls = []

    def set_level(self, l, n, f, c, s):
        self.calls.append(("level", l, n, len(f), len(c), tuple(s)))

    def set_csr(self, l, which, m):
        self.calls.append(("csr", l, which, m.shape, m.nnz))

    def set_diag(self, l, which, d):
        self.calls.append(("diag", l, which, len(d)))

    def set_poly(self, l, which, t, c, ds):
        self.calls.append(("poly", l, which, t, np.asarray(c).shape, bool(ds)))

    def finalize(self):
        self.calls.append(("finalize",))


def test_upload_walk_hands_over_every_operator():
    A, H = cases.build("fd2d_fcf")
    r = pflare_b200.feed(H, _Recorder())
    NL = H.no_levels
    levels = [c for c in r.calls if c[0] == "level"]
    assert [c[1] for c in levels] == list(range(1, NL + 1))
    for l, lv in enumerate(H.levels, start=1):
        which = sorted(c[2] for c in r.calls if c[0] in ("csr", "diag", "poly") and c[1] == l)
        assert which == sorted([pflare_b200.AFF, pflare_b200.AFC, pflare_b200.ACF, pflare_b200.ACC,
                                pflare_b200.INV_AFF, pflare_b200.INV_ACC, pflare_b200.R, pflare_b200.P])
        assert levels[l - 1][2:5] == (lv.n, lv.is_fine.size, lv.is_coarse.size)
    assert r.calls[-1] == ("finalize",)
    assert sorted(c[2] for c in r.calls if c[0] != "level" and len(c) > 1 and c[1] == NL) == [pflare_b200.INV_AFF, pflare_b200.COARSE]


def test_hierarchy_container_round_trip(tmp_path):
    """Integer data (CF lists, CSR structure) must survive the container bit-exactly."""
    A, H = cases.build("fd2d_ffcc_mf")
    p = str(tmp_path / "h.npz")
    hio.save(p, H)
    H2, _ = hio.load(p)
    assert H2.no_levels == H.no_levels
    for a, b in zip(H.levels, H2.levels):
        assert np.array_equal(a.is_fine, b.is_fine) and np.array_equal(a.is_coarse, b.is_coarse)
        for nm in ("A_ff", "A_fc", "R", "P", "A_cf", "A_cc"):
            ma, mb = getattr(a, nm), getattr(b, nm)
            assert np.array_equal(ma.indptr, mb.indptr) and np.array_equal(ma.indices, mb.indices)
            assert np.array_equal(ma.data, mb.data)
        assert a.smooth_order == b.smooth_order
        assert np.array_equal(a.inv_A_ff.coeffs, b.inv_A_ff.coeffs)


def test_problem_generators_match_reference_stencils():
    # tests/adv_1d.c:79-105
    A = hiergen.adv_1d(5).toarray()
    assert np.array_equal(A, np.eye(5) - np.eye(5, k=-1))
    # tests/adv_diff_fd.c:434-491: interior row of the 2D upwind stencil, theta = pi/4, nondimensional
    n = 8
    A = hiergen.adv_diff_fd(n, n).tocsr()
    i, j = 3, 4
    row = A[j * n + i].toarray().ravel()
    u = v = np.sqrt(0.5)
    assert np.isclose(row[(j - 1) * n + i], -v) and np.isclose(row[j * n + i - 1], -u) and np.isclose(row[j * n + i], u + v)
    assert A[0, 0] == 1.0 and A[0].nnz == 1                       # inflow row = identity
    assert A.nnz == (n - 1) * (n - 1) * 3 + (2 * n - 1)


def test_c_abi_from_plain_c(built_libs, tmp_path):
    """include/pflare_b200.h compiles as C99 and the library can be driven from C (host-only planning context):
    the language and calling convention of the reference's FFI layer (src/C_PETSc_Routines.c)."""
    import subprocess
    exe = str(tmp_path / "capi_host_check")
    libdir = os.path.dirname(pflare_b200.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c", "capi_host_check.c"), "-L", libdir, "-lpflare_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "CAPI_HOST_OK" in out.stdout, out.stdout + out.stderr


def test_warp_tile_format_and_row_sum_algorithm(tmp_path):
    """The operator storage of spmv_wt_kernel (pflare_b200/csrc/wt_format.h) and the kernel's row-sum algorithm
    (lane-local walk + segmented warp scan), emulated lane by lane on the CPU against a plain CSR product:
    random row lengths 1..256, empty rows, the merged A_fc|W (last entry = W) mode, ghost columns."""
    import subprocess
    for name, ok in (("wt_format_check", "WT_FORMAT_OK"), ("wc_format_check", "WC_FORMAT_OK")):   # row-aligned lanes / chunk format
        exe = str(tmp_path / name)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-o", exe, os.path.join(ROOT, "tests", "c", name + ".cpp")])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and ok in out.stdout, out.stdout + out.stderr


def test_petsc_binary_hierarchy_container_round_trip(built_libs, tmp_path):
    """The PETSc-binary hierarchy container (pflare_b200/petsc_io.py): write -> read is bit exact, the
    committed fixtures reproduce their stored PCApply output on the oracle, and the reader's primitives parse the
    reference's own PETSc fixture tests/data/mat_stream_2364 (as committed under tests/golden/ilu_mat_stream.npz)."""
    import numpy as np
    import cases
    import hiergen
    import oracle
    from pflare_b200 import petsc_io
    from pflare_b200.upload import feed
    for name in ("fd2d_fcf", "fd2d_mf_newton", "fd2d_full", "fd2d_jacobi"):
        A, H = cases.build(name)
        b = cases.rhs(A.shape[0])
        xo = hiergen.feed(H, oracle.OracleAIR(H.no_levels)).apply(b)
        path = str(tmp_path / (name + ".petsc"))
        petsc_io.save_hierarchy(path, H, b, xo)
        H2 = petsc_io.load_hierarchy(path)
        assert H2.no_levels == H.no_levels and H2.options.full_smoothing_up_and_down == H.options.full_smoothing_up_and_down
        for l1, l2 in zip(H.levels, H2.levels):
            assert np.array_equal(l1.is_fine, l2.is_fine) and np.array_equal(l1.is_coarse, l2.is_coarse)
            assert (l1.R != l2.R).nnz == 0 and (l1.P != l2.P).nnz == 0
        x2 = feed(H2, oracle.OracleAIR(H2.no_levels)).apply(H2.b)
        assert np.array_equal(x2, H2.x) and np.array_equal(x2, xo)
    gold = os.path.join(ROOT, "tests", "golden")
    for name in ("fd2d_25", "fd2d_full_mf"):
        Hg = petsc_io.load_hierarchy(os.path.join(gold, name + ".petsc"))
        xg = feed(Hg, oracle.OracleAIR(Hg.no_levels)).apply(Hg.b)
        assert np.linalg.norm(xg - Hg.x) <= 1e-13 * np.linalg.norm(Hg.x)
    with pytest.raises(ValueError):
        petsc_io.load_hierarchy(os.path.join(gold, "fd2d_25.npz"))


def test_64bit_petscint_upload_entry_points(built_libs):
    """PetscInt = 64-bit builds: the *_i64 upload calls narrow a rank's local block to 32 bits and refuse anything that does
    not fit (host-only planning context: no GPU needed)."""
    import numpy as np
    A, H = cases.build("fd2d_fcf")
    d32 = pflare_b200.DeviceAIR(H.no_levels, device=-1)
    d64 = pflare_b200.DeviceAIR(H.no_levels, device=-1, idx64=True)
    hiergen.feed(H, d32)
    hiergen.feed(H, d64)
    for l, lv in enumerate(H.levels, start=1):
        assert np.array_equal(d64.get_is(l, 0), lv.is_fine) and np.array_equal(d64.get_is(l, 1), lv.is_coarse)
    assert d32.stats()["nnz_per_cycle"] == d64.stats()["nnz_per_cycle"] and d32.stats()["algorithmic_bytes"] == d64.stats()["algorithmic_bytes"]
    # an index beyond 2^31 is refused loudly
    bad = np.array([0, 1, 2 ** 31 + 5], dtype=np.int64)
    L = pflare_b200.lib()
    rc = L.pflare_b200_set_level_i64(d64.h, 1, 0, 3, 3, ctypes.c_void_p(bad.ctypes.data), 0, None, None, 0)
    assert rc == 3 and b"32-bit" in L.pflare_b200_last_error()
    d32.close(); d64.close()
