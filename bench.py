#!/usr/bin/env python
"""bench.py -- AIRG V-cycle apply (PCApply of PCAIR) on B200, BASELINE.json's metric.

A "step" is ONE V-cycle: PCApply on one right-hand side, on the hierarchy of the workload
(default: BASELINE.json configs[1], tests/adv_diff_fd.c 2D upwind advection 4096^2, default PCAIR
options).  The hierarchy is an input (the reference's setup builds it; here hiergen restates that
setup on the host before the timed region) and is uploaded once as device CSR.

  value    : DOF/s = level-1 rows / V-cycle time, rhs and result resident in HBM, CUDA events on
             the library's own stream, max over ranks.
  e2e      : the same through the reference-facing call with HOST buffers (pinned): H2D of the rhs,
             V-cycle, D2H of the result inside the timed region.
  roofline : HBM; algorithmic bytes (SURVEY.md section 8d model, computed by the library per op) of the
             SpMV launches of one V-cycle / the graph-mode cycle time (the SpMV kernel is all but a
             handful of launches, so the whole cycle time is charged to it: a conservative figure).
  parity   : relative L2 difference between the GPU result and the CPU oracle's result on the same
             seeded rhs, at the benchmarked size and transport (every rank checks its rows); the
             run FAILS above 1e-12.
  cpu_baseline / --impl reference : the CPU oracle (OpenMP restatement of the reference's PETSc
             path, kind "port": PETSc/gfortran are absent so the reference cannot be built) on the
             same hierarchy, all host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the host-side setup (hierarchy load / partition, warp-tile packing at upload) is
# OpenMP code, so give each rank its share of the host cores instead -- before any OpenMP runtime is loaded.  (The oracle's thread
# count is set explicitly, host_threads().)
if "LOCAL_RANK" in os.environ and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    try:
        _cores = len(os.sched_getaffinity(0))
    except Exception:
        _cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(max(1, _cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL's own banner / debug lines go to stderr.  NCCL honours NCCL_DEBUG_FILE only above
# the VERSION level (at NCCL_DEBUG=VERSION the "NCCL version ..." banner is written to stdout regardless), so VERSION becomes WARN:
# same banner, no extra output, but routed to the file.
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------- workload
def make_problem(args):
    import hiergen
    from hiergen import poly
    w = args.workload
    if w == "adv_diff_fd_2d":
        n = args.n
        A = hiergen.adv_diff_fd(n, n)
        opts = hiergen.AirOptions()
        name = "tests/adv_diff_fd.c 2D upwind advection %dx%d, default PCAIR (AIRG, assembled Arnoldi order 6, ff smoothing)" % (n, n)
    elif w == "adv_diff_fd_2d_mf":
        n = args.n
        A = hiergen.adv_diff_fd(n, n)
        opts = hiergen.AirOptions(matrix_free_polys=True)
        name = "tests/adv_diff_fd.c 2D upwind advection %dx%d, PCAIR -pc_air_matrix_free_polys" % (n, n)
    elif w == "adv_diff_fd_3d":
        n = args.n
        A = hiergen.adv_diff_fd(n, n, n)
        opts = hiergen.AirOptions(a_lump=True)
        name = "tests/adv_diff_fd.c 3D upwind advection %d^3, PCAIR -pc_air_a_lump" % n
    elif w == "adv_diff_fd_3d_lair":
        # BASELINE.json configs[3]: 3D FD, AIRG with lAIR Z (-pc_air_z_type lair, lair_distance 2), at the size that fits the budget
        n = args.n
        A = hiergen.adv_diff_fd(n, n, n, alpha=args.alpha)
        opts = hiergen.AirOptions(a_lump=True, z_type="lair")
        name = "tests/adv_diff_fd.c 3D advection-diffusion FD %d^3 (alpha %g), PCAIR -pc_air_z_type lair -pc_air_a_lump" % (n, args.alpha)
    elif w == "dg_upwind":
        n = args.n
        A = hiergen.dg_upwind_surrogate(n, n, 3)
        opts = hiergen.AirOptions(matrix_free_polys=True)
        name = "DG-P1 upwind surrogate %dx%d cells x3, PCAIR -pc_air_matrix_free_polys" % (n, n)
    else:
        raise SystemExit("unknown workload " + w)
    return A, opts, name


DEFAULT_P2P = "2"   # the library's default ghost exchange (pflare_b200_set_option "p2p")


def cache_dir():
    for d in (os.environ.get("PFLARE_BENCH_CACHE"), "/dev/shm", "/tmp"):
        if d and os.path.isdir(d) and os.access(d, os.W_OK):
            p = os.path.join(d, "pflare_b200_cache")
            os.makedirs(p, exist_ok=True)
            return p
    return None


def build_hierarchy(args):
    """Build (or load from the per-box cache) the workload's hierarchy.  The cache only saves host
    time between back-to-back bench invocations on one box (reference arm, b200 arm, N = 1/2/4/8);
    it holds inputs, never results."""
    import hiergen
    from hiergen import io as hio
    t = time.time()
    A, opts, name = make_problem(args)
    cd = None if args.no_cache else cache_dir()
    path = os.path.join(cd, "%s_%d_%g.npz" % (args.workload, args.n, args.alpha)) if cd else None
    if path and os.path.exists(path):
        try:
            H, _ = hio.load(path)
            H.A = A
            log("[bench] hierarchy: %d rows, %d levels, loaded from %s in %.1f s" % (A.shape[0], H.no_levels, path, time.time() - t))
            return A, H, name
        except Exception as e:
            log("[bench] cache load failed (%s); rebuilding" % e)
    H = hiergen.build_hierarchy(A, opts, verbose=args.verbose)
    log("[bench] hierarchy: %d rows, %d levels, built on the host in %.1f s" % (A.shape[0], H.no_levels, time.time() - t))
    if path and int(os.environ.get("RANK", "0")) == 0:
        try:
            t = time.time()
            tmp = path + ".tmp.%d.npz" % os.getpid()
            d = hio.to_dict(H, with_A=False)
            np.savez(tmp, **d)
            os.replace(tmp, path)
            log("[bench] hierarchy cached at %s (%.1f s)" % (path, time.time() - t))
        except Exception as e:
            log("[bench] could not cache the hierarchy: %s" % e)
    return A, H, name


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log("[bench] NVML unavailable:", e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max), "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def make_oracle(H):
    """The CPU oracle on this hierarchy, with ALL host threads (torchrun exports OMP_NUM_THREADS=1)."""
    import hiergen
    import oracle
    O = hiergen.feed(H, oracle.OracleAIR(H.no_levels))
    O.set_threads(host_threads())
    return O


def workload_config(name, A, H):
    """`config` of the JSON line: identical in the b200 and the reference arm."""
    return {"workload": name, "rows": int(A.shape[0]), "levels": int(H.no_levels), "nnz_level1": int(A.nnz),
            "rhs": "seeded uniform(0,1), seed 1234",
            "l2": "inputs larger than L2 (every V-cycle streams the whole hierarchy, > 126 MB), no flush"}


def time_oracle(O, b, cycles, warm=1):
    for _ in range(warm):
        O.apply(b)
    ts = []
    for _ in range(cycles):
        t = time.perf_counter()
        x = O.apply(b)
        ts.append(time.perf_counter() - t)
    return float(np.median(ts)), x


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    A, H, name = build_hierarchy(args)
    n = A.shape[0]
    # each step = one V-cycle of the CPU path on the full workload (a cycle is O(0.1-1 s): bounded)
    O = make_oracle(H)
    b = np.random.default_rng(1234).random(n)
    for _ in range(args.warmup):
        O.apply(b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.apply(b)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sample = "%d full V-cycles of the CPU oracle on the whole workload (%d rows)" % (args.steps, n)
    out = {
        "impl": "reference", "metric": "AIRG V-cycle DOF/s", "value": val, "unit": "DOF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(name, A, H),
        "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": O.threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import pflare_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- pflare_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        # rank 0 builds (or loads) the hierarchy and caches it; the others load it from the cache
        if rank == 0:
            A, H, name = build_hierarchy(args)
        dist.barrier()
        if rank != 0:
            A, H, name = build_hierarchy(args)
    else:
        A, H, name = build_hierarchy(args)
    n = A.shape[0]
    t = time.time()
    if world > 1:
        import ctypes
        import hiergen
        from pflare_b200 import _capi
        uid = [None]
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _capi.check(_capi.lib().pflare_b200_get_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
            uid[0] = buf.raw
        dist.broadcast_object_list(uid, src=0)
        part = hiergen.partition(H, world, only=rank)[rank]
        rows0 = part.rangesV[0]
        n_local = int(rows0[rank + 1] - rows0[rank])
        log("[bench] rank %d: rows [%d, %d), partitioned in %.1f s" % (rank, rows0[rank], rows0[rank + 1], time.time() - t))
        pc = pflare_b200.PC(rank=rank, nranks=world, unique_id=uid[0], device=local).setType("air").setHierarchy(part)
    else:
        n_local = n
        pc = pflare_b200.PC(device=local).setType("air").setHierarchy(H)
    for kv in args.opt:
        k, v = kv.split("=")
        pc.setOption(k, float(v))
    t = time.time()
    pc.setUp()
    dev = pc.device()
    log("[bench] rank %d upload + finalize: %.1f s" % (rank, time.time() - t))
    st = dev.stats()
    stream = torch.cuda.ExternalStream(dev.stream_ptr(), device=torch.device("cuda", local))

    b_full = np.random.default_rng(1234).random(n)
    lo = int(rows0[rank]) if world > 1 else 0
    b_host = torch.from_numpy(np.ascontiguousarray(b_full[lo:lo + n_local])).pin_memory()
    x_host = torch.empty(n_local, dtype=torch.float64).pin_memory()
    b = b_host.cuda()
    x = torch.empty_like(b)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / steps

    if args.profile_one_cycle:
        for _ in range(args.warmup):
            dev.apply_ptr(b.data_ptr(), x.data_ptr(), 1)
        dev.synchronize()
        torch.cuda.profiler.start()
        dev.apply_ptr(b.data_ptr(), x.data_ptr(), 1)
        dev.synchronize()
        torch.cuda.profiler.stop()
        log("[bench] profiled one V-cycle")
        pc.destroy()
        return
    # device-resident leg
    with ClockSampler(local) as clk:
        ms_dev = timed(lambda: dev.apply_ptr(b.data_ptr(), x.data_ptr(), 1), args.steps, args.warmup)
    clocks = clk.summary()
    extra = {}
    for kv in args.compare_opt:
        for one in kv.split(","):
            k, v = one.split("=")
            pc.setOption(k, float(v))
        extra[kv] = timed(lambda: dev.apply_ptr(b.data_ptr(), x.data_ptr(), 1), args.steps, args.warmup)
        log("[bench] with %s: %.4f ms per V-cycle (main run %.4f)" % (kv, extra[kv], ms_dev))
    # end-to-end leg: host (pinned) buffers through the reference-facing call
    ms_e2e = timed(lambda: dev.apply_ptr(b_host.data_ptr(), x_host.data_ptr(), 0), args.steps, args.warmup)

    # device-resident and host-buffer applies must agree bit for bit
    dev.synchronize()
    xs = x.cpu().numpy()
    assert np.all(np.isfinite(xs)) and np.array_equal(xs, x_host.numpy()), "device and host-buffer applies differ"

    # parity against the CPU oracle on the same rhs, at THIS size and over THIS transport: rank 0 runs the
    # oracle (all host threads), every rank checks its own rows
    cpu = None
    parity = None
    if not args.no_parity:
        xo_full = torch.empty(n, dtype=torch.float64, device="cuda")
        if rank == 0:
            O = make_oracle(H)
            t_cpu, xo = time_oracle(O, b_full, cycles=args.cpu_cycles if (world == 1 and not args.no_cpu_baseline) else 1,
                                    warm=1 if world == 1 else 0)
            xo_full.copy_(torch.from_numpy(xo))
            if world == 1 and not args.no_cpu_baseline:
                cpu = {"value": n / t_cpu, "unit": "DOF/s", "cores": O.threads(), "kind": "port",
                       "sample": "median of %d full V-cycles of the CPU oracle (OpenMP restatement of the PETSc path) on the whole workload" % args.cpu_cycles,
                       "ms_per_cycle": t_cpu * 1e3}
            O.close()
        if world > 1:
            dist.broadcast(xo_full, src=0)
        xo_loc = xo_full[lo:lo + n_local]
        num = torch.sum((x - xo_loc) ** 2)
        den = torch.sum(xo_loc ** 2)
        nd = torch.stack([num, den])
        if world > 1:
            dist.all_reduce(nd)
        rel = float(torch.sqrt(nd[0] / nd[1]).item())
        parity = {"rel_l2": rel, "n": int(n), "tol": 1e-12, "against": "CPU oracle (oracle/air_oracle.c) on the same seeded rhs",
                  "ranks_checked": world}
        log("[bench] rank %d parity vs oracle: rel L2 = %.3e" % (rank, rel))
        if not (rel <= 1e-12):
            raise SystemExit("bench.py: PARITY FAILURE: relative L2 difference %.3e > 1e-12 against the CPU oracle" % rel)
        del xo_full

    # roofline of the dominant kernel: per-launch CUDA events (graph off for this pass)
    reps = 3
    tot_ms = tot_by = 0.0
    allms = allby = 0.0
    biggest = (0.0, 0.0, 0)
    for r in range(reps + 1):
        ms, by, lev, kind = dev.profile_apply(b.data_ptr(), x.data_ptr())
        if r == 0:
            continue
        is_spmv = np.isin(kind, [1, 2, 3, 4, 5, 7, 8, 9])     # SpMV mega-op launches (6 = elementwise / permutation, 10 = exchange, 11 = dense tail)
        tot_ms += float(ms[is_spmv].sum())
        tot_by += float(by[is_spmv].sum())
        allms += float(ms.sum())
        allby += float(by.sum())
        k = int(np.argmax(by))
        biggest = (float(by[k]), float(ms[k]), int(lev[k]))
        if r == reps and args.dump_ops and rank == 0:
            tags = {1: "restrict Z", 2: "coarse", 3: "A_fc(+W)", 4: "A_ff resid", 5: "inverse", 6: "elementwise/permute", 7: "fused local F", 8: "A_cf", 9: "A_cc", 10: "exchange", 11: "dense tail"}
            nspmv = int(np.sum(is_spmv & (by > 0)))        # SpMV kernel launches of one cycle (used to aim ncu -s/-c)
            with open(args.dump_ops + ".nspmv", "w") as f:
                f.write("%d\n" % nspmv)
            with open(args.dump_ops, "w") as f:
                f.write("idx,level,op,alg_bytes,ms,GBps\n")
                for i in range(len(ms)):
                    f.write("%d,%d,%s,%d,%.5f,%.1f\n" % (i, lev[i], tags.get(int(kind[i]), str(kind[i])), by[i], ms[i],
                                                        by[i] / (ms[i] * 1e-3) / 1e9 if ms[i] > 0 else 0.0))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # The SpMV mega-op kernel is all but a handful of launches of the cycle.  achieved = algorithmic bytes of ALL SpMV
    # launches of one V-cycle / the graph-mode cycle time (the non-SpMV launches' time is charged to the kernel too).
    spmv_by = tot_by / reps
    achieved = spmv_by / (ms_dev * 1e-3) / 1e9
    per_launch = tot_by / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    kernel_id = {0: "spmv_stream_kernel", 1: "spmv_tma_kernel<256,1024,2>", 2: "spmv_wc_kernel + spmv_sv_kernel"}[int(float(dict(kv.split("=") for kv in args.opt).get("kernel", 2)))]
    roof = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
        "kernel": kernel_id + " (SpMV mega-op on warp-tile storage: chunk-format TMA-ring kernel for short rows, direct kernel on row-aligned lanes otherwise)",
        "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6650",
        "how": "algorithmic bytes (SURVEY.md 8d) of all SpMV launches of one V-cycle / graph-mode cycle time (CUDA events on the library stream)",
        "achieved_per_launch_events": per_launch,
        "per_launch_events_note": "same bytes / sum of per-launch CUDA-event durations with the graph off (adds event + launch gaps)",
        "frac_of_8TBs_nominal": achieved / 8000.0,
        "spmv_share_of_cycle_events": tot_ms / allms if allms > 0 else None,
        "largest_launch": {"bytes": biggest[0], "ms": biggest[1], "level": biggest[2],
                           "GBps": biggest[0] / (biggest[1] * 1e-3) / 1e9 if biggest[1] > 0 else None},
        "algorithmic_bytes_per_cycle": st["algorithmic_bytes"], "spmv_algorithmic_bytes_per_cycle": spmv_by,
        "nnz_per_cycle": st["nnz_per_cycle"],
    }
    # DRAM traffic of the dominant kernel: from the committed ncu --set full capture of this same workload, valid
    # only for the kernel source it was taken with (the file is stamped with the kernel name and the sha256 of
    # kernels.cuh; a mismatch means the capture is stale and is NOT reported)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        sha = hashlib.sha256(open(os.path.join(ROOT, "pflare_b200", "csrc", "kernels.cuh"), "rb").read()).hexdigest()
        if tr.get("kernel") == kernel_id and tr.get("kernels_cuh_sha256") == sha:
            roof["traffic"] = tr["mean_dram_bytes_per_launch"]
            roof["traffic_detail"] = {"launches": len(tr["launches"]), "mean_algorithmic_bytes_per_launch": tr["mean_algorithmic_bytes_per_launch"],
                                      "dram_over_algorithmic": tr["dram_over_algorithmic"], "source": tr["source"]}
        else:
            roof["traffic_detail"] = {"stale": "profiles/r02_ncu_traffic.json was captured with another kernel source (%s)" % tr.get("kernel")}
    except Exception:
        pass

    # whole-job counters (bytes / launches summed over the ranks)
    tot = torch.tensor([st["algorithmic_bytes"], st["kernel_launches"], st["ghost_bytes_sent"], st["device_bytes"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    job_bytes, job_launches, job_ghost, job_dev = [float(v) for v in tot.tolist()]
    l_agg, glob_rows = dev.layout() if world > 1 else (H.no_levels + 1, None)

    if rank == 0:
        out = {
            "metric": "AIRG V-cycle DOF/s", "value": n / (ms_dev * 1e-3),
            "unit": "DOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, A, H),
            "details": {"streamed_GB_per_cycle": job_bytes / 1e9, "device_bytes": job_dev,
                        "partition": "1 GPU" if world == 1 else "%d ranks, contiguous row blocks (PETSc MPIAIJ ownership), ghost exchange = %s, levels >= %d agglomerated on rank 0" % (world, {"1": "peer-memory push kernel + acks (CUDA IPC)", "2": "peer-memory push fused into the consuming SpMV kernel (CUDA IPC)"}.get(dict(o.split("=") for o in args.opt if "=" in o).get("p2p", DEFAULT_P2P), "NCCL send/recv"), l_agg),
                        "ghost_bytes_per_cycle": job_ghost, "exchange_groups_per_cycle_rank0": int(st["exchange_groups"]),
                        "library_options": args.opt},
            "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "DOF/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "note": "summed over ranks"},
            "gpu_launches": int(job_launches) * args.steps,
            "launches_per_cycle": int(st["kernel_launches"]), "tail_levels": int(st["tail_levels"]),
            "clocks": clocks, "compare_opt_ms": extra,
        }
        print(json.dumps(out), flush=True)
    pc.destroy()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PFLARE_BENCH_WORKLOAD", "adv_diff_fd_2d"))
    ap.add_argument("--size", dest="n", type=int, default=int(os.environ.get("PFLARE_BENCH_N", "4096")))
    ap.add_argument("--alpha", type=float, default=0.0, help="diffusion coefficient of the 3D workloads (0 = pure upwind advection)")
    ap.add_argument("--cpu-cycles", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check (development runs only)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (repeatable)")
    ap.add_argument("--compare-opt", action="append", default=[], help="after the main run, re-time the device-resident leg with this option changed (key=value)")
    ap.add_argument("--profile-one-cycle", action="store_true",
                    help="bracket exactly one V-cycle with cudaProfilerStart/Stop (run under `ncu --profile-from-start off`) and exit")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch table of one V-cycle (CUDA events) to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
