"""Python binding of the CPU oracle (oracle/air_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (pflare_b200) must never import this module.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libair_oracle.so")
_lib = None

AFF, AFC, ACF, ACC, INV_AFF, INV_ACC, R, P, COARSE = range(9)


def build(force=False):
    src = os.path.join(_HERE, "air_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libair_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.c_int]
        for name in ("oracle_set_level", "oracle_set_csr", "oracle_set_diag", "oracle_set_poly", "oracle_pcapply",
                     "oracle_inv_apply", "oracle_matmult", "oracle_fc_smooth"):
            getattr(L, name).restype = ctypes.c_int
        L.oracle_destroy.restype = None
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class OracleAIR:
    """Sink for ``hiergen.feed`` + apply entry points of the oracle."""

    def __init__(self, no_levels):
        self.L = lib()
        self.no_levels = no_levels
        self.h = ctypes.c_void_p(self.L.oracle_create(no_levels))
        self.n = {}
        self._keep = []

    def set_level(self, our_level, n, is_f, is_c, smooth_order):
        is_f = np.ascontiguousarray(is_f, dtype=np.int32)
        is_c = np.ascontiguousarray(is_c, dtype=np.int32)
        sm = np.ascontiguousarray(smooth_order, dtype=np.int32)
        self.n[our_level] = n
        self.L.oracle_set_level(self.h, our_level, n, is_f.size, _ip(is_f), is_c.size, _ip(is_c), _ip(sm), sm.size)

    def set_csr(self, our_level, which, mat):
        ia = np.ascontiguousarray(mat.indptr, dtype=np.int32)
        ja = np.ascontiguousarray(mat.indices, dtype=np.int32)
        a = np.ascontiguousarray(mat.data, dtype=np.float64)
        self.L.oracle_set_csr(self.h, our_level, which, mat.shape[0], mat.shape[1], _ip(ia), _ip(ja), _dp(a))

    def set_diag(self, our_level, which, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        self.L.oracle_set_diag(self.h, our_level, which, d.size, _dp(d))

    def set_poly(self, our_level, which, inverse_type, coeffs, diag_scale):
        c = np.asarray(coeffs, dtype=np.float64)
        re = np.ascontiguousarray(c[:, 0])
        im = np.ascontiguousarray(c[:, 1]) if c.shape[1] > 1 else np.zeros_like(re)
        self.L.oracle_set_poly(self.h, our_level, which, inverse_type, re.size, _dp(re), _dp(im), int(diag_scale))

    def set_option(self, key, value):
        """Options that change the apply (mirrors pflare_b200_set_option); unknown keys are execution-only and ignored."""
        if key == "full_smoothing_up_and_down":
            self.L.oracle_set_full_smoothing(self.h, int(bool(value)))
        elif key == "mg_coarse_ksp_max_it":
            self.L.oracle_set_coarse_its(self.h, int(value))

    def finalize(self):
        pass

    def apply(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        rc = self.L.oracle_pcapply(self.h, _dp(b), _dp(x))
        if rc:
            raise RuntimeError("oracle_pcapply rc=%d" % rc)
        return x

    def inv_apply(self, our_level, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        rc = self.L.oracle_inv_apply(self.h, our_level, which, _dp(x), _dp(y))
        if rc:
            raise RuntimeError("oracle_inv_apply rc=%d" % rc)
        return y

    def fc_smooth(self, our_level, b, x):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x, dtype=np.float64, copy=True)
        self.L.oracle_fc_smooth(self.h, our_level, _dp(b), _dp(x))
        return x

    def threads(self):
        return int(self.L.oracle_omp_threads())

    def set_threads(self, n):
        self.L.oracle_set_omp_threads(int(n))
        return self.threads()

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
