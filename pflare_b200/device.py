"""`DeviceAIR`: one per-PC device context of libpflare_b200.so, driven through the C-ABI only.

It implements the upload-hook "sink" protocol (set_level / set_csr / set_diag / set_poly /
finalize) that the reference-side shim calls at the end of ``setup_air_pcmg``
(/root/reference/src/AIR_MG_Setup.F90:1178-1211) and that ``hiergen.feed``
replays in this repository, plus the apply entry points.  All arithmetic happens in the CUDA
library; nothing here computes on the CPU.
"""
import ctypes
import numpy as np

from . import _capi
from ._capi import check

AFF, AFC, ACF, ACC, INV_AFF, INV_ACC, R, P, COARSE = range(9)

STAT_NAMES = ("kernel_launches", "algorithmic_bytes", "nnz_per_cycle", "device_bytes", "ghost_bytes_sent",
              "largest_kernel_bytes", "tail_levels", "exchange_groups")


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class DeviceAIR:
    """Handle to one uploaded hierarchy (``void *handle`` of include/pflare_b200.h)."""

    def __init__(self, no_levels, rank=0, nranks=1, unique_id=None, device=0):
        self.L = _capi.lib()
        self.no_levels = int(no_levels)
        self.rank, self.nranks, self.device = rank, nranks, device
        self.h = ctypes.c_void_p()
        uid = None
        if unique_id is not None:
            self._uid = ctypes.create_string_buffer(bytes(unique_id), 128)
            uid = ctypes.cast(self._uid, ctypes.c_void_p)
        check(self.L.pflare_b200_create(ctypes.byref(self.h), rank, nranks, uid, device, self.no_levels))
        self.n = {}
        self.nf = {}
        self.nc = {}
        self._stream = None

    # ------------------------------------------------------------------ upload hook
    def set_level(self, our_level, n, is_f, is_c, smooth_order, rstart=0):
        is_f, is_c, sm = _i32(is_f), _i32(is_c), _i32(smooth_order)
        self.n[our_level], self.nf[our_level], self.nc[our_level] = int(n), is_f.size, is_c.size
        check(self.L.pflare_b200_set_level(self.h, our_level, int(rstart), int(n), is_f.size, _ptr(is_f), is_c.size,
                                           _ptr(is_c), _ptr(sm), sm.size))

    def set_csr(self, our_level, which, mat, offdiag=None, garray=None, cstart=0):
        """mat = diag block (scipy CSR, local columns); offdiag = CSR over compressed ghost columns."""
        ia, ja, a = _i32(mat.indptr), _i32(mat.indices), _f64(mat.data)
        if offdiag is not None and garray is not None and len(garray) > 0:
            oi, oj, oa = _i32(offdiag.indptr), _i32(offdiag.indices), _f64(offdiag.data)
            ga = np.ascontiguousarray(garray, dtype=np.int64)
            check(self.L.pflare_b200_set_csr(self.h, our_level, which, mat.shape[0], mat.shape[1], int(cstart), _ptr(ia),
                                             _ptr(ja), _ptr(a), ga.size, _ptr(oi), _ptr(oj), _ptr(oa), _ptr(ga)))
        else:
            check(self.L.pflare_b200_set_csr(self.h, our_level, which, mat.shape[0], mat.shape[1], int(cstart), _ptr(ia),
                                             _ptr(ja), _ptr(a), 0, None, None, None, None))

    def set_diag(self, our_level, which, d):
        d = _f64(d)
        check(self.L.pflare_b200_set_diag(self.h, our_level, which, d.size, _ptr(d)))

    def set_poly(self, our_level, which, inverse_type, coeffs, diag_scale):
        c = np.asarray(coeffs, dtype=np.float64)
        if c.ndim == 1:
            c = c.reshape(-1, 1)
        re = _f64(c[:, 0])
        im = _f64(c[:, 1]) if c.shape[1] > 1 else None
        check(self.L.pflare_b200_set_poly(self.h, our_level, which, int(inverse_type), re.size, _ptr(re), _ptr(im),
                                          int(bool(diag_scale))))

    def finalize(self):
        check(self.L.pflare_b200_finalize_setup(self.h))

    def set_option(self, key, value):
        check(self.L.pflare_b200_set_option(self.h, key.encode(), float(value)))

    # ------------------------------------------------------------------ apply (host buffers)
    def apply(self, b):
        """x = PCApply(b) with HOST buffers (copies in, runs the V-cycle on the GPU, copies out)."""
        b = _f64(b)
        x = np.empty_like(b)
        check(self.L.pflare_b200_apply(self.h, _ptr(b), _ptr(x), 0))
        return x

    def inv_apply(self, our_level, which, x):
        x = _f64(x)
        y = np.empty_like(x)
        check(self.L.pflare_b200_inv_apply(self.h, our_level, which, _ptr(x), _ptr(y), 0))
        return y

    def fc_smooth(self, our_level, b, x):
        b = _f64(b)
        x = np.array(x, dtype=np.float64, copy=True)
        check(self.L.pflare_b200_fc_smooth(self.h, our_level, _ptr(b), _ptr(x), 0))
        return x

    # ------------------------------------------------------------------ apply (device pointers)
    def apply_ptr(self, b_ptr, x_ptr, on_device=1):
        """Raw-pointer apply; asynchronous on the handle's stream when on_device != 0."""
        check(self.L.pflare_b200_apply(self.h, ctypes.c_void_p(b_ptr), ctypes.c_void_p(x_ptr), on_device))

    def inv_apply_ptr(self, our_level, which, x_ptr, y_ptr, on_device=1):
        check(self.L.pflare_b200_inv_apply(self.h, our_level, which, ctypes.c_void_p(x_ptr), ctypes.c_void_p(y_ptr), on_device))

    def stream_ptr(self):
        if self._stream is None:
            s = ctypes.c_void_p()
            check(self.L.pflare_b200_get_stream(self.h, ctypes.byref(s)))
            self._stream = s.value or 0
        return self._stream

    def synchronize(self):
        check(self.L.pflare_b200_synchronize(self.h))

    # ------------------------------------------------------------------ introspection
    def get_is(self, our_level, which_is):
        n = self.nf[our_level] if which_is == 0 else self.nc[our_level]
        out = np.zeros(max(n, 1), dtype=np.int32)
        check(self.L.pflare_b200_get_is(self.h, our_level, which_is, _ptr(out)))
        return out[:n]

    def get_garray(self, our_level, which):
        ng = ctypes.c_int(0)
        check(self.L.pflare_b200_get_garray(self.h, our_level, which, None, ctypes.byref(ng)))
        out = np.zeros(max(ng.value, 1), dtype=np.int64)
        check(self.L.pflare_b200_get_garray(self.h, our_level, which, _ptr(out), ctypes.byref(ng)))
        return out[:ng.value]

    def stats(self):
        v = np.zeros(8)
        check(self.L.pflare_b200_get_stats(self.h, _ptr(v), 8))
        return dict(zip(STAT_NAMES, v.tolist()))

    def profile_apply(self, b_ptr, x_ptr, max_ops=4096):
        ms = np.zeros(max_ops, dtype=np.float32)
        by = np.zeros(max_ops)
        lev = np.zeros(max_ops, dtype=np.int32)
        kind = np.zeros(max_ops, dtype=np.int32)
        n = ctypes.c_int(0)
        check(self.L.pflare_b200_profile_apply(self.h, ctypes.c_void_p(b_ptr), ctypes.c_void_p(x_ptr), max_ops, _ptr(ms),
                                               _ptr(by), _ptr(lev), _ptr(kind), ctypes.byref(n)))
        k = n.value
        return ms[:k], by[:k], lev[:k], kind[:k]

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.pflare_b200_destroy(ctypes.byref(self.h))
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
