"""ctypes declarations of every symbol of include/pflare_b200.h (the C-ABI boundary).

The shared library is built in-tree by ``pflare_b200/csrc/Makefile`` (``__graft_entry__.build()``)
as ``pflare_b200/libpflare_b200.so``.  There is no CPU fallback: if the library is missing the
import of any compute entry point raises, and every compute call fails when no CUDA device is
usable (``pflare_b200_create`` returns error 10).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpflare_b200.so")

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p
P = ctypes.POINTER

# name -> (restype, argtypes); kept in one table so tests can check the export list against
# the header.
SIGNATURES = {
    "pflare_b200_get_unique_id": (c_int, [c_vp]),
    "pflare_b200_create": (c_int, [P(c_vp), c_int, c_int, c_vp, c_int, c_int]),
    "pflare_b200_set_level": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_int]),
    "pflare_b200_set_csr": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "pflare_b200_set_level_i64": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64]),
    "pflare_b200_set_csr_i64": (c_int, [c_vp, c_int, c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "pflare_b200_set_diag": (c_int, [c_vp, c_int, c_int, c_int, c_vp]),
    "pflare_b200_set_poly": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int]),
    "pflare_b200_finalize_setup": (c_int, [c_vp]),
    "pflare_b200_apply": (c_int, [c_vp, c_vp, c_vp, c_int]),
    "pflare_b200_inv_apply": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int]),
    "pflare_b200_fc_smooth": (c_int, [c_vp, c_int, c_vp, c_vp, c_int]),
    "pflare_b200_ksp_set_operator": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "pflare_b200_ksp_solve": (c_int, [c_vp, c_int, c_int, c_dbl, c_dbl, c_int, c_int, c_vp, c_vp, c_int, P(c_int), P(c_int), P(c_dbl)]),
    "pflare_b200_get_stream": (c_int, [c_vp, P(c_vp)]),
    "pflare_b200_synchronize": (c_int, [c_vp]),
    "pflare_b200_get_is": (c_int, [c_vp, c_int, c_int, c_vp]),
    "pflare_b200_get_garray": (c_int, [c_vp, c_int, c_int, c_vp, P(c_int)]),
    "pflare_b200_get_stats": (c_int, [c_vp, c_vp, c_int]),
    "pflare_b200_profile_apply": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, P(c_int)]),
    "pflare_b200_set_option": (c_int, [c_vp, ctypes.c_char_p, c_dbl]),
    "pflare_b200_last_error": (ctypes.c_char_p, []),
    "pflare_b200_destroy": (c_int, [P(c_vp)]),
    "pflare_b200_set_host_exchange": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "pflare_b200_get_ghost_plan": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, P(c_int)]),
    "pflare_b200_get_layout": (c_int, [c_vp, P(c_int), c_vp, c_int]),
    "pflare_b200_cluster_create": (c_int, [P(c_vp), c_int, c_int, c_int]),
    "pflare_b200_cluster_rank": (c_int, [c_vp, c_int, P(c_vp)]),
    "pflare_b200_cluster_finalize": (c_int, [c_vp]),
    "pflare_b200_cluster_apply": (c_int, [c_vp, c_vp, c_vp, c_int]),
    "pflare_b200_cluster_inv_apply": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int]),
    "pflare_b200_cluster_set_option": (c_int, [c_vp, ctypes.c_char_p, c_dbl]),
    "pflare_b200_cluster_get_stream": (c_int, [c_vp, P(c_vp)]),
    "pflare_b200_cluster_destroy": (c_int, [P(c_vp)]),
}

_lib = None


class PflareB200Error(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "pflare_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libpflare_b200.so (fails loudly when it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C pflare_b200/csrc`). pflare_b200 has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().pflare_b200_last_error()
        raise PflareB200Error(rc, msg.decode() if msg else "")
