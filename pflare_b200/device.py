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

    def __init__(self, no_levels, rank=0, nranks=1, unique_id=None, device=0, _handle=None, idx64=False):
        self.L = _capi.lib()
        self.idx64 = bool(idx64)     # hand the index data over as 64-bit PetscInt (the *_i64 entry points)
        self.no_levels = int(no_levels)
        self.rank, self.nranks, self.device = rank, nranks, device
        self._owned = _handle is None
        if _handle is not None:      # a rank of an in-process group (owned by the ClusterAIR)
            self.h = _handle
        else:
            self.h = ctypes.c_void_p()
            uid = None
            if unique_id is not None:
                self._uid = ctypes.create_string_buffer(bytes(unique_id), 128)
                uid = ctypes.cast(self._uid, ctypes.c_void_p)
            check(self.L.pflare_b200_create(ctypes.byref(self.h), rank, nranks, uid, device, self.no_levels))
        self._cb = None
        self.n = {}
        self.nf = {}
        self.nc = {}
        self._stream = None

    # ------------------------------------------------------------------ upload hook
    def set_level(self, our_level, n, is_f, is_c, smooth_order, rstart=0):
        is_f, is_c, sm = _i32(is_f), _i32(is_c), _i32(smooth_order)
        self.n[our_level], self.nf[our_level], self.nc[our_level] = int(n), is_f.size, is_c.size
        if self.idx64:
            f8, c8, s8 = (np.ascontiguousarray(a, dtype=np.int64) for a in (is_f, is_c, sm))
            check(self.L.pflare_b200_set_level_i64(self.h, our_level, int(rstart), int(n), f8.size, _ptr(f8), c8.size, _ptr(c8), _ptr(s8), s8.size))
            return
        check(self.L.pflare_b200_set_level(self.h, our_level, int(rstart), int(n), is_f.size, _ptr(is_f), is_c.size,
                                           _ptr(is_c), _ptr(sm), sm.size))

    def set_csr(self, our_level, which, mat, offdiag=None, garray=None, cstart=0):
        """mat = diag block (scipy CSR, local columns); offdiag = CSR over compressed ghost columns."""
        ia, ja, a = _i32(mat.indptr), _i32(mat.indices), _f64(mat.data)
        if self.idx64:
            i8 = lambda v: np.ascontiguousarray(v, dtype=np.int64)
            ia8, ja8 = i8(ia), i8(ja)
            if offdiag is not None and garray is not None and len(garray) > 0:
                oi8, oj8, oa, ga = i8(offdiag.indptr), i8(offdiag.indices), _f64(offdiag.data), i8(garray)
                check(self.L.pflare_b200_set_csr_i64(self.h, our_level, which, mat.shape[0], mat.shape[1], int(cstart), _ptr(ia8), _ptr(ja8),
                                                     _ptr(a), ga.size, _ptr(oi8), _ptr(oj8), _ptr(oa), _ptr(ga)))
            else:
                check(self.L.pflare_b200_set_csr_i64(self.h, our_level, which, mat.shape[0], mat.shape[1], int(cstart), _ptr(ia8), _ptr(ja8),
                                                     _ptr(a), 0, None, None, None, None))
            return
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
        if self._owned:
            check(self.L.pflare_b200_finalize_setup(self.h))
        # ranks of an in-process group are finalized together by ClusterAIR.finalize()

    def set_host_exchange(self, alltoall, alltoallv):
        """Setup-time host communicator as two Python callables with MPI_Alltoall / MPI_Alltoallv semantics:
        alltoall(send: int64[P]) -> int64[P];  alltoallv(sendbuf: bytes, scnt, sdsp, rcnt, rdsp) -> bytes."""
        P_ = self.nranks
        A2A = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64))
        A2AV = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                                ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                                ctypes.POINTER(ctypes.c_int64))

        def _a2a(ctx, send, recv):
            try:
                out = alltoall(np.array([send[p] for p in range(P_)], dtype=np.int64))
                for p in range(P_):
                    recv[p] = int(out[p])
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1

        def _a2av(ctx, sbuf, scnt, sdsp, rbuf, rcnt, rdsp):
            try:
                sc = [scnt[p] for p in range(P_)]
                sd = [sdsp[p] for p in range(P_)]
                rc_ = [rcnt[p] for p in range(P_)]
                rd = [rdsp[p] for p in range(P_)]
                tot = max([sd[p] + sc[p] for p in range(P_)] + [0])
                data = ctypes.string_at(sbuf, tot) if tot else b""
                out = alltoallv(data, sc, sd, rc_, rd)
                if len(out):
                    ctypes.memmove(rbuf, out, len(out))
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1

        self._cb = (A2A(_a2a), A2AV(_a2av))
        check(self.L.pflare_b200_set_host_exchange(self.h, ctypes.cast(self._cb[0], ctypes.c_void_p),
                                                   ctypes.cast(self._cb[1], ctypes.c_void_p), None))

    def ghost_plan(self, our_level, which):
        P_ = self.nranks
        sc = np.zeros(P_, dtype=np.int32)
        rc_ = np.zeros(P_, dtype=np.int32)
        ro = np.zeros(P_, dtype=np.int32)
        n = ctypes.c_int(0)
        check(self.L.pflare_b200_get_ghost_plan(self.h, our_level, which, _ptr(sc), _ptr(rc_), _ptr(ro), None, ctypes.byref(n)))
        idx = np.zeros(max(n.value, 1), dtype=np.int32)
        check(self.L.pflare_b200_get_ghost_plan(self.h, our_level, which, None, None, None, _ptr(idx), ctypes.byref(n)))
        return {"send_count": sc, "recv_count": rc_, "recv_off": ro, "send_idx": idx[:n.value]}

    def layout(self):
        la = ctypes.c_int(0)
        rows = np.zeros(self.no_levels, dtype=np.int64)
        check(self.L.pflare_b200_get_layout(self.h, ctypes.byref(la), _ptr(rows), self.no_levels))
        return la.value, rows

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

    # ------------------------------------------------------------------ outer Krylov method (KSPSolve)
    def ksp_set_operator(self, mat, offdiag=None, garray=None, cstart=0):
        """System matrix of the outer solve (level-1 rows, natural numbering); call before finalize."""
        ia, ja, a = _i32(mat.indptr), _i32(mat.indices), _f64(mat.data)
        if offdiag is not None and garray is not None and len(garray) > 0:
            oi, oj, oa = _i32(offdiag.indptr), _i32(offdiag.indices), _f64(offdiag.data)
            ga = np.ascontiguousarray(garray, dtype=np.int64)
            check(self.L.pflare_b200_ksp_set_operator(self.h, mat.shape[0], mat.shape[1], int(cstart), _ptr(ia), _ptr(ja), _ptr(a),
                                                      ga.size, _ptr(oi), _ptr(oj), _ptr(oa), _ptr(ga)))
        else:
            check(self.L.pflare_b200_ksp_set_operator(self.h, mat.shape[0], mat.shape[1], int(cstart), _ptr(ia), _ptr(ja), _ptr(a),
                                                      0, None, None, None, None))

    def ksp_solve(self, b, x0, ksp_type="gmres", side="right", rtol=1e-5, atol=1e-50, max_it=10000, restart=30):
        """KSPSolve on the device with host buffers: returns (x, its, converged, rnorm)."""
        b = _f64(b)
        x = np.array(x0, dtype=np.float64, copy=True)
        its, why, rn = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
        check(self.L.pflare_b200_ksp_solve(self.h, 0 if ksp_type == "gmres" else 1, 0 if side == "left" else 1, float(rtol), float(atol),
                                           int(max_it), int(restart), _ptr(b), _ptr(x), 0, ctypes.byref(its), ctypes.byref(why),
                                           ctypes.byref(rn)))
        return x, its.value, why.value > 0, rn.value

    def ksp_solve_ptr(self, b_ptr, x_ptr, ksp_type=0, side=1, rtol=1e-5, atol=1e-50, max_it=10000, restart=30):
        its, why, rn = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
        check(self.L.pflare_b200_ksp_solve(self.h, ksp_type, side, float(rtol), float(atol), int(max_it), int(restart),
                                           ctypes.c_void_p(b_ptr), ctypes.c_void_p(x_ptr), 1, ctypes.byref(its), ctypes.byref(why),
                                           ctypes.byref(rn)))
        return its.value, why.value, rn.value

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
        if getattr(self, "h", None) is not None and self.h and self._owned:
            self.L.pflare_b200_destroy(ctypes.byref(self.h))
        self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ClusterAIR:
    """In-process rank group (``pflare_b200_cluster_*``): a partitioned hierarchy driven by one process."""

    def __init__(self, no_levels, nranks, device=0):
        self.L = _capi.lib()
        self.no_levels, self.nranks, self.device = int(no_levels), int(nranks), device
        self.c = ctypes.c_void_p()
        check(self.L.pflare_b200_cluster_create(ctypes.byref(self.c), self.nranks, device, self.no_levels))
        self.ranks = []
        for r in range(self.nranks):
            h = ctypes.c_void_p()
            check(self.L.pflare_b200_cluster_rank(self.c, r, ctypes.byref(h)))
            self.ranks.append(DeviceAIR(no_levels, rank=r, nranks=nranks, device=device, _handle=h))

    def set_option(self, key, value):
        check(self.L.pflare_b200_cluster_set_option(self.c, key.encode(), float(value)))

    def upload(self, local_hierarchies):
        for r, lh in enumerate(local_hierarchies):
            lh.feed(self.ranks[r])
        self.finalize()
        return self

    def finalize(self):
        check(self.L.pflare_b200_cluster_finalize(self.c))

    def _ptr_array(self, arrs):
        return (ctypes.c_void_p * self.nranks)(*[a.ctypes.data for a in arrs])

    def apply(self, b_parts):
        b_parts = [_f64(b) for b in b_parts]
        x_parts = [np.empty_like(b) for b in b_parts]
        check(self.L.pflare_b200_cluster_apply(self.c, self._ptr_array(b_parts), self._ptr_array(x_parts), 0))
        return x_parts

    def apply_ptrs(self, b_ptrs, x_ptrs, on_device=1):
        bp = (ctypes.c_void_p * self.nranks)(*b_ptrs)
        xp = (ctypes.c_void_p * self.nranks)(*x_ptrs)
        check(self.L.pflare_b200_cluster_apply(self.c, bp, xp, on_device))

    def inv_apply(self, our_level, which, x_parts):
        x_parts = [_f64(x) for x in x_parts]
        y_parts = [np.empty_like(x) for x in x_parts]
        check(self.L.pflare_b200_cluster_inv_apply(self.c, our_level, which, self._ptr_array(x_parts), self._ptr_array(y_parts), 0))
        return y_parts

    def stream_ptr(self):
        s = ctypes.c_void_p()
        check(self.L.pflare_b200_cluster_get_stream(self.c, ctypes.byref(s)))
        return s.value or 0

    def stats(self):
        return [r.stats() for r in self.ranks]

    def close(self):
        if getattr(self, "c", None) is not None and self.c:
            for r in self.ranks:
                r.h = ctypes.c_void_p()
            self.L.pflare_b200_cluster_destroy(ctypes.byref(self.c))
            self.c = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
