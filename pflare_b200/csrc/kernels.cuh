// kernels.cuh -- sm_100a device code of the AIRG V-cycle apply.
//
// One SpMV "mega-op" covers every MatMult + Vec-AXPY chain of the reference's apply path
// (SURVEY.md section 2c): the CSR product of a row is reduced once and the row epilogue applies
// the fused vector updates.  The path is HBM-bound fp64/int32 work (no tensor cores by design).
//
//   spmv_wt_kernel      the kernel of the cycle (round 2).  The operator is stored as WARP TILES: one
//                       contiguous blob per tile (<= 8 slots per lane x 32 lanes of consecutive rows,
//                       row-aligned lanes: wt_format.h) that ONE 1-D TMA bulk copy brings into a
//                       per-warp shared-memory ring.  Warps are autonomous (no CTA barrier anywhere):
//                       while a warp reduces tile t its x gathers and epilogue operands of tile t+1 are
//                       already in flight in registers and tiles t+2.. are in flight in the ring.  Rows
//                       are summed by lane-local adds plus a segmented shuffle reduction; the row
//                       epilogue is specialised at compile time per op class.
//   spmv_tma_kernel     round-1 kernel (CTA tiles of a CSR stream, 3 bulk copies per tile, CTA barriers);
//                       kept as option kernel=1 for A/B measurements.
//   spmv_stream_kernel  first-generation smem-staged kernel (option kernel=0); also the fallback for
//                       operators with rows longer than a warp tile.
//   ew_kernel           diagonal inverses / scalings / permutations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wt_format.h"

namespace pfb {

constexpr int kThreads = 256;       // CTA size of the streaming kernels
constexpr int kTile = 2048;         // nnz per CTA tile of the stream kernel (16 KB of fp64 products)
constexpr int kMaxRowsPerBlk = 1024;

struct TileDesc { int r0, nrows, s, n; };  // first row, #rows, first nnz, #nnz (n > tile size: one long row)

// epilogue classes of the warp-tile kernel (compile-time specialisations of the row epilogue)
enum { EPI_GENERIC = 0,     // everything below, decided at run time
       EPI_ADD = 1,         // out[i] += s                         restriction b_c += Z b_f ; x_f += M r
       EPI_SET = 2,         // out[i]  = s                         coarse solve, PCPFLAREINV, x_f = W x_c
       EPI_AXPBY = 3,       // out[i]  = alpha aux[i] + beta s     residuals, Horner steps
       EPI_AFCW = 4,        // merged A_fc|W: out = aux - s ; wout = w x_c
       EPI_AFCW_LOCAL = 5,  // merged A_fc|W + row-local F smooth (diagonal A_ff and M_ff)
       EPI_AXPBY_ACC = 6,   // Newton-basis steps: out = alpha aux + beta s ; acc (=|+=) gamma (acc_src | out)
       EPI_NCLASS = 7 };

struct SpmvOp {
  // CSR block + its row-block partition (kernel 0 / 1 and the long-row fallback)
  const int *rp, *col;
  const double *val;
  int m, nblk;
  const int *blk;
  const int *rowmap;           // stream kernel on a row subset (the rows longer than a warp tile): CSR row r is operator row rowmap[r]
  const TileDesc *tiles;       // tile list of the round-1 TMA kernel
  int ntiles;
  // warp-tile storage (kernel 2)
  const unsigned char *blob;
  const WtDesc *wdesc;
  int nwt;
  int kp;                      // row-aligned format: slots per lane of a sub-tile (1, 2, 4, 8); chunk format: rows per lane of a tile
  int fmt;                     // 1 = chunk format (spmv_wc_kernel), 2 = row-aligned format (spmv_wt / spmv_sv / spmv_thin kernels)
  int epi;                     // epilogue class
  // gather sources: column c < nloc reads x[c], otherwise xg[c - nloc] (ghost buffer)
  const double *x, *xg;
  int nloc;
  // s = sum_j a_ij x_j ; optional s /= D[i] ; optional s = x[i] - s (Neumann I - D^-1 A)
  const double *D;
  int neumann;
  // v = alpha * aux[i] + beta * s
  const double *aux;
  double alpha, beta;
  double *out; int out_mode;   // 0 none, 1 out[i] = v, 2 out[i] += v
  double *out2; double delta;  // out2[i] = delta * v
  double *acc; double gamma; const double *acc_src; int acc_mode;  // acc[i] (=|+=) gamma * (acc_src ? acc_src[i] : v)
  // one-point prolongation companion (A_fc|W merged CSR): the LAST stored entry of every row is
  // the W entry (wval, wcol); its product is not part of the row sum but gives wout[i] = W x_c
  int wlast; double *wout;
  // fully local F smooth (diagonal A_ff and diagonal inverse): x = wout value; repeat fd_its:
  // x += fd_m[i] * (v - fd_a[i] * x); wout[i] = x
  const double *fd_a, *fd_m; int fd_its;
  // entry / exit permutation of the cycle fused into the level-1 ops (natural <-> nested ordering):
  //   aux_idx  : aux is read as aux[aux_idx[i]]          (b in PETSc's natural ordering)
  //   wout_idx : wout / out is ALSO stored to xnat[wout_idx[i]]  (x in natural ordering)
  const int *aux_idx; const int *wout_idx; double *xnat;
  // peer-memory ghost exchange: before the first ghost read, wait until every source rank has pushed
  // its chunk of THIS exchange instance (ready[q] >= *epoch for the ranks q in srcmask)
  const unsigned *gw_ready; const unsigned *gw_epoch; unsigned gw_srcmask;
  // fused peer-memory exchange (option p2p=2): the kernel itself first pushes THIS rank's boundary entries of x into
  // the peers' ghost buffers of this exchange instance (xpush), runs its interior tiles, and only the warp that reaches
  // a tile >= gw_first (the first tile with ghost columns) waits for the peers' chunks
  const struct XPush *xpush; int gw_first;
};

// out[i] (=|+=) alpha * a[i] * (b ? b[i] : 1) / (dv ? dv[i] : 1)
struct EwOp {
  int n;
  const double *a, *b, *dv;
  double alpha;
  double *out; int mode;  // 1 set, 2 add
  const int *gather;      // optional: read a[gather[i]]
  const int *scatter;     // optional: write out[scatter[i]]
};

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization
// attribute may start while its predecessor drains; everything before pdl_wait() must touch only
// data no kernel of the cycle writes (matrix arrays, tile lists), everything after sees the
// predecessor's results.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- system-scope flags of the peer-memory ghost exchange (written by one GPU, polled by another)
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded spin: a peer that is legitimately late (still in its setup, or running the agglomerated levels) is waited for;
// after kSpinTimeoutNs the wait gives up and sets *err, so that a lost peer can never hang the GPU.
const unsigned long long kSpinTimeoutNs = 120ull * 1000000000ull;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void spin_until(const unsigned *flag, unsigned e, unsigned *err) {
  unsigned long long t0 = 0;
  int it = 0;
  while ((int)(ld_acquire_sys(flag) - e) < 0) {
    if (err && *reinterpret_cast<volatile unsigned *>(err)) break;   // an earlier wait already gave up: fall through, do not wait again
    __nanosleep(20);
    if ((++it & 1023) == 0) {
      const unsigned long long t = global_ns();
      if (!t0) t0 = t;
      else if (t - t0 > kSpinTimeoutNs) { if (err) *err = 1u; break; }
    }
  }
}
// consumer side: called by ONE thread of a CTA (or warp) before its first ghost read (word 1 of the flag block records a timeout)
__device__ __forceinline__ void ghost_wait(const unsigned *ready, const unsigned *epoch, unsigned srcmask) {
  if (!ready) return;
  const unsigned e = *epoch;
  for (unsigned m = srcmask; m; m &= m - 1) spin_until(ready + (__ffs(m) - 1), e, const_cast<unsigned *>(epoch) + 1);
}

// ---- fused push (p2p=2).  Flag block of a rank: word 0 = epoch (number of the running cycle), words 32..63 =
// started[q] (rank q has entered cycle e, i.e. finished reading every ghost buffer of cycle e-1), then ready[inst][q].
// Every exchange instance owns its own ghost buffer, so inside one cycle nothing is ever overwritten and the only
// back-pressure needed is the once-per-cycle started[] flag (epoch2_kernel).  Word 1 is set when a bounded spin gave up
// (a peer never arrived): the result is then wrong, nothing hangs, and the next host-buffer apply reports it.
struct XPush {
  int n, nranks, me, inst;
  const int *idx;                        // [n] positions in x of the entries the peers need
  const int *send_off;                   // [nranks + 1]
  const unsigned long long *dst;         // [nranks] my chunk inside peer p's ghost buffer of this instance (my address space)
  const unsigned long long *peer_flags;  // [nranks] peer p's flag block
  const unsigned *epoch;                 // my flag block
  unsigned *done;                        // warp-unit counter of this instance
  unsigned dstmask;
};
// called by every warp of the consuming kernel after pdl_wait(): warp-unit u pushes entries [32u, 32u + 32)
__device__ __forceinline__ void ghost_push(const XPush *__restrict__ px, const double *__restrict__ x, int gwarp, int nwarps, int lane) {
  if (!px) return;
  const int n = px->n;
  const int units = (n + 31) >> 5;
  if (gwarp >= units) return;
  const unsigned e = *px->epoch;
  unsigned cnt = 0;
  for (int u = gwarp; u < units; u += nwarps) {
    const int j = u * 32 + lane;
    if (j < n) {
      int p = 0;
      while (j >= px->send_off[p + 1]) ++p;
      reinterpret_cast<double *>(px->dst[p])[j - px->send_off[p]] = x[px->idx[j]];
    }
    ++cnt;
  }
  __threadfence_system();
  __syncwarp();
  if (lane == 0) {
    const unsigned prev = atomicAdd(px->done, cnt);
    if (prev + cnt == (unsigned)units) {   // last unit: every chunk is written and fenced -> raise the consumers' flags
      *px->done = 0;
      __threadfence_system();
      for (unsigned m = px->dstmask; m; m &= m - 1) {
        const int p = __ffs(m) - 1;
        st_release_sys(reinterpret_cast<unsigned *>(px->peer_flags[p]) + 64 + (size_t)px->inst * 32 + px->me, e);
      }
    }
  }
}
// lazy consumer gate: armed while the warp has not yet waited for the peers' chunks
__device__ __forceinline__ void ghost_gate(const SpmvOp &op, bool &armed, int tile, int lane) {
  if (armed && tile >= op.gw_first) {
    if (lane == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
    __syncwarp();
    armed = false;
  }
}

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

__device__ __forceinline__ double gather_x(const SpmvOp &op, int c) {
  if (op.xg != nullptr && c >= op.nloc) return __ldcg(op.xg + (c - op.nloc));  // ghosts: written by peers, read through L2
  return op.x[c];
}

// ---- generic (run-time branched) row epilogue, split in two so that the loads that depend only on
// the row index can be issued long before the row sum exists
struct EpiPre { double aux, D, xi, fa, fm, out, accsrc, acc; };

__device__ __forceinline__ EpiPre epi_prefetch(const SpmvOp &op, int i) {
  EpiPre p;
  p.aux = op.aux ? op.aux[op.aux_idx ? op.aux_idx[i] : i] : 0.0;
  p.D = op.D ? op.D[i] : 1.0;
  p.xi = op.neumann ? op.x[i] : 0.0;
  p.fa = op.fd_its > 0 ? op.fd_a[i] : 0.0;
  p.fm = op.fd_its > 0 ? op.fd_m[i] : 0.0;
  p.out = op.out_mode == 2 ? op.out[i] : 0.0;
  p.accsrc = (op.acc_mode && op.acc_src) ? op.acc_src[i] : 0.0;
  p.acc = op.acc_mode == 2 ? op.acc[i] : 0.0;
  return p;
}

__device__ __forceinline__ void epi_finish(const SpmvOp &op, int i, double s, double xw, const EpiPre &p) {
  if (op.D) s = s / p.D;
  if (op.neumann) s = p.xi - s;
  double v = op.beta * s;
  if (op.aux) v = op.alpha * p.aux + v;
  if (op.wout) {
    for (int it = 0; it < op.fd_its; ++it) xw = xw + p.fm * (v - p.fa * xw);
    op.wout[i] = xw;
    if (op.wout_idx) op.xnat[op.wout_idx[i]] = xw;
  }
  if (op.out_mode == 1) op.out[i] = v;
  else if (op.out_mode == 2) {
    op.out[i] = p.out + v;
    if (op.wout_idx && !op.wout) op.xnat[op.wout_idx[i]] = p.out + v;
  }
  if (op.out2) op.out2[i] = op.delta * v;
  if (op.acc_mode) {
    const double t = op.gamma * (op.acc_src ? p.accsrc : v);
    op.acc[i] = op.acc_mode == 1 ? t : p.acc + t;
  }
}

__device__ __forceinline__ void row_epilogue(const SpmvOp &op, int i, double s, double xw) {
  const EpiPre p = epi_prefetch(op, i);
  epi_finish(op, i, s, xw, p);
}

// ---- compile-time epilogue classes: Pre = the operands prefetched per row (one tile ahead)
template <int EPI> struct EpiT;
template <> struct EpiT<EPI_GENERIC> {
  static constexpr int kPre = 0; static constexpr bool kXw = true;
  struct Pre {};
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &, int) { return Pre(); }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double xw, const Pre &) { row_epilogue(op, i, s, xw); }
};
template <> struct EpiT<EPI_ADD> {
  static constexpr int kPre = 1; static constexpr bool kXw = false;
  struct Pre { double o; };
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &op, int i) { Pre p; p.o = op.out[i]; return p; }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double, const Pre &p) {
    const double v = p.o + s;
    if (op.wout_idx) op.xnat[op.wout_idx[i]] = v;   // last update of the level-1 F points: straight into the caller's x
    else op.out[i] = v;
  }
};
template <> struct EpiT<EPI_SET> {
  static constexpr int kPre = 0; static constexpr bool kXw = false;
  struct Pre {};
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &, int) { return Pre(); }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double, const Pre &) { op.out[i] = s; }
};
template <> struct EpiT<EPI_AXPBY> {
  static constexpr int kPre = 1; static constexpr bool kXw = false;
  struct Pre { double a; };
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &op, int i) { Pre p; p.a = op.aux[op.aux_idx ? op.aux_idx[i] : i]; return p; }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double, const Pre &p) {
    const double v = op.beta * s;
    op.out[i] = op.alpha * p.a + v;
  }
};
template <> struct EpiT<EPI_AFCW> {
  static constexpr int kPre = 1; static constexpr bool kXw = true;
  struct Pre { double a; };
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &op, int i) { Pre p; p.a = op.aux[op.aux_idx ? op.aux_idx[i] : i]; return p; }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double xw, const Pre &p) {
    const double v = op.beta * s;
    op.wout[i] = xw;
    op.out[i] = op.alpha * p.a + v;
  }
};
template <> struct EpiT<EPI_AFCW_LOCAL> {
  static constexpr int kPre = 3; static constexpr bool kXw = true;
  struct Pre { double a, fa, fm; };
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &op, int i) {
    Pre p; p.a = op.aux[op.aux_idx ? op.aux_idx[i] : i]; p.fa = op.fd_a[i]; p.fm = op.fd_m[i]; return p;
  }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double xw, const Pre &p) {
    double v = op.beta * s;
    v = op.alpha * p.a + v;
    for (int it = 0; it < op.fd_its; ++it) xw = xw + p.fm * (v - p.fa * xw);
    if (op.wout_idx) op.xnat[op.wout_idx[i]] = xw;   // level 1: straight into the caller's x (natural ordering)
    else op.wout[i] = xw;
  }
};
template <> struct EpiT<EPI_AXPBY_ACC> {
  static constexpr int kPre = 3; static constexpr bool kXw = false;
  struct Pre { double a, as, ac; };
  static __device__ __forceinline__ Pre prefetch(const SpmvOp &op, int i) {
    Pre p; p.a = op.aux[i]; p.as = op.acc_src ? op.acc_src[i] : 0.0; p.ac = op.acc_mode == 2 ? op.acc[i] : 0.0; return p;
  }
  static __device__ __forceinline__ void finish(const SpmvOp &op, int i, double s, double, const Pre &p) {
    double v = op.beta * s;
    v = op.alpha * p.a + v;
    op.out[i] = v;
    const double t = op.gamma * (op.acc_src ? p.as : v);
    op.acc[i] = op.acc_mode == 1 ? t : p.ac + t;
  }
};

// Process one row block with all threads of the CTA.  `prod` holds kTile doubles.
template <int NT>
__device__ __forceinline__ void process_block(const SpmvOp &op, int b, double *prod, double *red) {
  const int tid = threadIdx.x;
  const int r0 = __ldg(op.blk + b), r1 = __ldg(op.blk + b + 1);
  const int s = __ldg(op.rp + r0), e = __ldg(op.rp + r1);
  const int n = e - s;
  if (n <= kTile) {
    // stage products: coalesced streaming loads of val/col, gathered x
    constexpr int kIter = kTile / NT;
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int k = tid + it * NT;
      if (k < n) {
        const double a = ld_stream(op.val + s + k);
        const int c = ld_stream(op.col + s + k);
        prod[k] = a * gather_x(op, c);
      }
    }
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += NT) {
      int p = __ldg(op.rp + r) - s;
      int q = __ldg(op.rp + r + 1) - s;
      double xw = 0.0;
      if (op.wlast) { --q; xw = prod[q]; }
      double sum = 0.0;
      for (; p < q; ++p) sum += prod[p];
      row_epilogue(op, op.rowmap ? op.rowmap[r] : r, sum, xw);
    }
    __syncthreads();
  } else {
    // a single long row: whole-CTA reduction
    const int last = op.wlast ? e - 1 : e;
    double part = 0.0;
    for (int k = s + tid; k < last; k += NT) part += ld_stream(op.val + k) * gather_x(op, ld_stream(op.col + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int w = 0; w < NT / 32; ++w) sum += red[w];
      const double xw = op.wlast ? op.val[last] * gather_x(op, op.col[last]) : 0.0;
      row_epilogue(op, op.rowmap ? op.rowmap[r0] : r0, sum, xw);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) spmv_stream_kernel(const SpmvOp op) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double prod[kTile];
  __shared__ double red[kThreads / 32];
  ghost_push(op.xpush, op.x, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, (gridDim.x * blockDim.x) >> 5, threadIdx.x & 31);
  if (threadIdx.x == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
  __syncthreads();
  for (int b = blockIdx.x; b < op.nblk; b += gridDim.x) process_block<kThreads>(op, b, prod, red);
}

// Rows longer than a warp tile (compact CSR of just those rows, rowmap = their row numbers): one warp per row, the lanes
// stride the row with four independent 128-entry groups in flight, shuffle reduction, generic epilogue on lane 0.
__global__ void __launch_bounds__(256) spmv_longrow_kernel(const SpmvOp op) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int *__restrict__ rp = op.rp;
  const int *__restrict__ col = op.col;
  const double *__restrict__ val = op.val;
  int r = gw;
  int p0 = 0, p1 = 0;
  if (r < op.m) { p0 = rp[r]; p1 = rp[r + 1]; }
  pdl_wait();
  if (lane == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
  __syncwarp();
  for (; r < op.m; r += nw) {
    const int q0 = p0, q1 = p1 - (op.wlast ? 1 : 0), last = p1 - 1;
    if (r + nw < op.m) { p0 = rp[r + nw]; p1 = rp[r + nw + 1]; }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = q0 + lane;
    for (; p + 96 < q1; p += 128) {
      const int c0 = __ldcs(col + p), c1 = __ldcs(col + p + 32), c2 = __ldcs(col + p + 64), c3 = __ldcs(col + p + 96);
      const double v0 = __ldcs(val + p), v1 = __ldcs(val + p + 32), v2 = __ldcs(val + p + 64), v3 = __ldcs(val + p + 96);
      s0 += v0 * gather_x(op, c0); s1 += v1 * gather_x(op, c1); s2 += v2 * gather_x(op, c2); s3 += v3 * gather_x(op, c3);
    }
    for (; p < q1; p += 32) s0 += __ldcs(val + p) * gather_x(op, __ldcs(col + p));
    double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      const double xw = op.wlast ? val[last] * gather_x(op, col[last]) : 0.0;
      row_epilogue(op, op.rowmap ? op.rowmap[r] : r, sum, xw);
    }
  }
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA (1-D bulk copy) helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// The matrix stream is read exactly once per op: mark it evict-first in L2 so that it does not
// push the gathered vectors (which ARE re-read, by neighbouring rows and by the next op) out of
// the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

// ------------------------------------------------------------------------------------------
// Warp-tile SpMV (the kernel of the cycle).  Operator storage: wt_format.h (row-aligned lanes).
//
// Persistent CTAs of NW autonomous warps.  Global warp g owns the tiles g, g + #warps, ...; lane 0
// keeps STAGES bulk copies (one per tile) in flight into the warp's private ring, each signalled on
// its own mbarrier.  Per tile the warp runs two software-pipelined stages:
//   A (one tile ahead)  read the tile's column indices from shared memory and issue the x gathers
//                       and the loads of the rows' epilogue operands into registers;
//   B                   products into registers, after which the ring slot is refilled at once (the next
//                       bulk copy overlaps the rest of the tile's work); per sub-tile every lane sums
//                       its KP slots (pairwise) and a segmented shuffle reduction over the lanes of a
//                       row leaves the row sum in the row's head lane; row sums go through a small
//                       shared-memory buffer so that the epilogue runs with one row per lane
//                       (coalesced vector traffic), 8 / KP rows per lane.
// No __syncthreads: a warp never waits for another warp, so one warp's gather latency is hidden by
// the other warps of the SM and the bulk stream never drains.
template <int EPI, int KP, bool GHOST, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32, 2) spmv_wt_kernel(const SpmvOp op) {
  typedef EpiT<EPI> E;
  typedef typename E::Pre Pre;
  constexpr bool XW = E::kXw;
  constexpr int NSLOT = kWtSlots;
  constexpr int NS = kWtSlots / KP;            // sub-tiles per tile == rows per lane in the epilogue
  constexpr bool AHEAD = E::kPre * NS <= 16;   // epilogue operands prefetched one tile ahead (else at the start of stage B)
  constexpr int RS_BYTES = NS * 32 * 8;        // row sums of the tile being reduced (the ring slot itself is refilled early)
  constexpr int XW_BYTES = XW ? NS * 32 * 8 : 0;
  constexpr int WARP_BYTES = STAGES * kWtStageBytes + RS_BYTES + XW_BYTES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[NW][STAGES];
  __shared__ WtDesc sdesc[NW][STAGES];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + (size_t)w * WARP_BYTES;
  double *rs_s = reinterpret_cast<double *>(wbase + STAGES * kWtStageBytes);
  double *xw_s = rs_s + NS * 32;
  const int nwarps = gridDim.x * NW;
  const int first = blockIdx.x * NW + w;
  const int my = first < op.nwt ? (op.nwt - first + nwarps - 1) / nwarps : 0;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[w][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t pol = l2_policy_evict_first();
  const WtDesc *__restrict__ wdesc = op.wdesc;
  // producer (lane 0): one bulk copy per tile; the descriptor of the NEXT tile is fetched one issue ahead
  WtDesc dn = {0u, 0, 0, 0};
  if (lane == 0 && my > 0) dn = wdesc[first];
  auto issue = [&](int j) {
    const WtDesc d = dn;
    if (j + 1 < my) dn = wdesc[(size_t)first + (size_t)(j + 1) * nwarps];
    const int slot = j % STAGES;
    sdesc[w][slot] = d;
    const uint32_t bytes = (uint32_t)((d.geom & 0xff) * KP * 384 + 32);
    mbar_expect_tx(&full[w][slot], bytes);
    tma_load_1d(wbase + slot * kWtStageBytes, op.blob + (size_t)d.off16 * 16, bytes, &full[w][slot], pol);
  };
  pdl_launch_dependents();
  if (lane == 0)
    for (int j = 0; j < STAGES && j < my; ++j) issue(j);   // matrix data only: legal before pdl_wait
  pdl_wait();   // from here on the vectors written by the previous kernels are read
  if (GHOST) ghost_push(op.xpush, op.x, first, nwarps, lane);
  bool garmed = GHOST && op.gw_ready != nullptr;
  __syncwarp();

  const double *__restrict__ xv = op.x;
  const double *__restrict__ xg = op.xg;
  const int nloc = op.nloc;
  WtDesc dnx = {0u, 0, 0, 0};
  double xn[NSLOT];
  Pre pn[NS];
  // stage A of tile j
  auto stage_a = [&](int j) {
    const int slot = j % STAGES;
    if (GHOST) ghost_gate(op, garmed, first + j * nwarps, lane);
    mbar_wait(&full[w][slot], (uint32_t)((j / STAGES) & 1));
    dnx = sdesc[w][slot];
    const int nslots = (dnx.geom & 0xff) * KP;
    const int *col_s = reinterpret_cast<const int *>(wbase + slot * kWtStageBytes + nslots * 256);
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      xn[k] = 0.0;
      if (k < nslots) {
        const int c = col_s[k * 32 + lane];
        if (GHOST) xn[k] = (c >= nloc) ? __ldcg(xg + (c - nloc)) : xv[c];
        else xn[k] = xv[c];
      }
    }
    if (AHEAD) {
#pragma unroll
      for (int q = 0; q < NS; ++q)
        if (lane + 32 * q < dnx.nrows) pn[q] = E::prefetch(op, dnx.r0 + lane + 32 * q);
    }
  };
  if (my > 0) stage_a(0);
  for (int it = 0; it < my; ++it) {
    const WtDesc d = dnx;
    double p[NSLOT];
    Pre pc[NS];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) p[k] = xn[k];
#pragma unroll
    for (int q = 0; q < NS; ++q) pc[q] = pn[q];
    // ---- stage B, part 1: products and head masks into registers -> the ring slot is consumed
    const int slot = it % STAGES;
    unsigned char *st = wbase + slot * kWtStageBytes;
    const double *val_s = reinterpret_cast<const double *>(st);
    const int ns = d.geom & 0xff, gmax = d.geom >> 8;
    const int nslots = ns * KP;
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) p[k] = (k < nslots) ? val_s[k * 32 + lane] * p[k] : 0.0;
    const unsigned *heads = reinterpret_cast<const unsigned *>(st + nslots * 384);
    unsigned hd[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) hd[s] = (s < ns) ? heads[s] : 0u;
    // refill the slot NOW (not at the end of the iteration): the bulk copy of tile it + STAGES then has the
    // whole reduction / epilogue of this tile and stage A of the next one to land
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy reads of the slot before the next bulk copy into it
    __syncwarp();
    if (lane == 0 && it + STAGES < my) issue(it + STAGES);
    if (it + 1 < my) stage_a(it + 1);
    if (!AHEAD) {
#pragma unroll
      for (int q = 0; q < NS; ++q)
        if (lane + 32 * q < d.nrows) pc[q] = E::prefetch(op, d.r0 + lane + 32 * q);
    }
    // ---- stage B, part 2: row sums
    const bool wf = (EPI == EPI_AFCW || EPI == EPI_AFCW_LOCAL) || (EPI == EPI_GENERIC && op.wlast);
    int rowbase = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < ns) {
        const unsigned H = hd[s];
        const bool head = (H >> lane) & 1u;
        double xw = 0.0;
        double q0 = p[s * KP];
        if (XW && wf && head) { xw = q0; q0 = 0.0; }   // merged A_fc|W: the row's first entry is the W entry
        double acc;
        if (KP == 1) acc = q0;
        else if (KP == 2) acc = q0 + p[s * KP + (KP > 1 ? 1 : 0)];
        else if (KP == 4) acc = (q0 + p[s * KP + (KP > 1 ? 1 : 0)]) + (p[s * KP + (KP > 2 ? 2 : 0)] + p[s * KP + (KP > 2 ? 3 : 0)]);
        else acc = ((q0 + p[(KP > 1 ? 1 : 0)]) + (p[(KP > 2 ? 2 : 0)] + p[(KP > 2 ? 3 : 0)])) +
                   ((p[(KP > 4 ? 4 : 0)] + p[(KP > 4 ? 5 : 0)]) + (p[(KP > 4 ? 6 : 0)] + p[(KP > 4 ? 7 : 0)]));
        // segmented reduction over the lanes of a row (toward its head lane)
        if (gmax > 1) {
          const unsigned above = lane < 31 ? (H >> (lane + 1)) : 0u;
          const int dist = above ? __ffs((int)above) - 1 : 31 - lane;   // lanes of my row after me
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            if (o < gmax) {
              const double t = __shfl_down_sync(0xffffffffu, acc, o);
              if (o <= dist) acc += t;
            }
          }
        }
        if (head) {
          const int r = rowbase + __popc(H & ((1u << lane) - 1u));
          rs_s[r] = acc;
          if (XW && wf) xw_s[r] = xw;
        }
        rowbase += __popc(H);
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int r = lane + 32 * q;
      if (r < d.nrows) E::finish(op, d.r0 + r, rs_s[r], (XW && wf) ? xw_s[r] : 0.0, pc[q]);
    }
    __syncwarp();   // the row-sum buffer is reused by the next tile
  }
}

// ------------------------------------------------------------------------------------------
// Chunk-format engine (wt_format.h, "CHUNK format"; operators with short rows): the same per-warp TMA ring as
// spmv_wt_kernel, but lanes own consecutive runs of the CSR stream: every lane walks its kpl nonzeros, row ends
// are flagged in per-lane masks, and ONE segmented warp scan carries the partial sum of a row that spans lanes.
// Row sums go to shared memory (the consumed value area of the stage), the epilogue runs with RQ rows per lane.
template <int EPI, int RQ, bool GHOST, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32, 2) spmv_wc_kernel(const SpmvOp op) {
  typedef EpiT<EPI> E;
  typedef typename E::Pre Pre;
  constexpr bool XW = E::kXw;
  constexpr int KPL = kWcKpl;
  constexpr bool AHEAD = E::kPre * RQ <= 16;   // epilogue operands prefetched one tile ahead (else at the start of stage B)
  constexpr int XW_BYTES = XW ? RQ * 32 * 8 : 0;
  constexpr int WARP_BYTES = STAGES * kWcStageBytes + XW_BYTES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[NW][STAGES];
  __shared__ WtDesc sdesc[NW][STAGES];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + (size_t)w * WARP_BYTES;
  double *xw_s = reinterpret_cast<double *>(wbase + STAGES * kWcStageBytes);
  const int nwarps = gridDim.x * NW;
  const int first = blockIdx.x * NW + w;
  const int my = first < op.nwt ? (op.nwt - first + nwarps - 1) / nwarps : 0;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[w][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t pol = l2_policy_evict_first();
  const WtDesc *__restrict__ wdesc = op.wdesc;
  // producer (lane 0): one bulk copy per tile; the descriptor of the NEXT tile is fetched one issue ahead
  WtDesc dn = {0u, 0, 0, 0};
  if (lane == 0 && my > 0) dn = wdesc[first];
  auto issue = [&](int j) {
    const WtDesc d = dn;
    if (j + 1 < my) dn = wdesc[(size_t)first + (size_t)(j + 1) * nwarps];
    const int slot = j % STAGES;
    sdesc[w][slot] = d;
    const uint32_t bytes = (uint32_t)(d.geom * 384 + 64);
    mbar_expect_tx(&full[w][slot], bytes);
    tma_load_1d(wbase + slot * kWcStageBytes, op.blob + (size_t)d.off16 * 16, bytes, &full[w][slot], pol);
  };
  pdl_launch_dependents();
  if (lane == 0)
    for (int j = 0; j < STAGES && j < my; ++j) issue(j);   // matrix data only: legal before pdl_wait
  pdl_wait();   // from here on the vectors written by the previous kernels are read
  if (GHOST) ghost_push(op.xpush, op.x, first, nwarps, lane);
  bool garmed = GHOST && op.gw_ready != nullptr;
  __syncwarp();

  const double *__restrict__ xv = op.x;
  const double *__restrict__ xg = op.xg;
  const int nloc = op.nloc;
  WtDesc dnx = {0u, 0, 0, 0};
  double xn[KPL];
  Pre pn[RQ];
  // stage A of tile j
  auto stage_a = [&](int j) {
    const int slot = j % STAGES;
    if (GHOST) ghost_gate(op, garmed, first + j * nwarps, lane);
    mbar_wait(&full[w][slot], (uint32_t)((j / STAGES) & 1));
    dnx = sdesc[w][slot];
    const int *col_s = reinterpret_cast<const int *>(wbase + slot * kWcStageBytes + dnx.geom * 256);
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      xn[k] = 0.0;
      if (k < dnx.geom) {
        const int c = col_s[k * 32 + lane];
        if (GHOST) xn[k] = (c >= nloc) ? __ldcg(xg + (c - nloc)) : xv[c];
        else xn[k] = xv[c];
      }
    }
    if (AHEAD) {
#pragma unroll
      for (int q = 0; q < RQ; ++q)
        if (lane + 32 * q < dnx.nrows) pn[q] = E::prefetch(op, dnx.r0 + lane + 32 * q);
    }
  };
  if (my > 0) stage_a(0);
  for (int it = 0; it < my; ++it) {
    const WtDesc d = dnx;
    double p[KPL];
    Pre pc[RQ];
#pragma unroll
    for (int k = 0; k < KPL; ++k) p[k] = xn[k];
#pragma unroll
    for (int q = 0; q < RQ; ++q) pc[q] = pn[q];
    if (it + 1 < my) stage_a(it + 1);
    // ---- stage B
    const int slot = it % STAGES;
    unsigned char *st = wbase + slot * kWcStageBytes;
    double *val_s = reinterpret_cast<double *>(st);
    if (!AHEAD) {
#pragma unroll
      for (int q = 0; q < RQ; ++q)
        if (lane + 32 * q < d.nrows) pc[q] = E::prefetch(op, d.r0 + lane + 32 * q);
    }
#pragma unroll
    for (int k = 0; k < KPL; ++k) p[k] = (k < d.geom) ? val_s[k * 32 + lane] * p[k] : 0.0;
    const unsigned e = reinterpret_cast<const unsigned short *>(st + d.geom * 384)[lane];
    __syncwarp();   // every lane holds its values: the value area now receives the row sums
    // partial sum after this lane's last row end (the whole lane if it ends no row)
    const int last = 31 - __clz((int)e);
    double tail = 0.0;
#pragma unroll
    for (int k = 0; k < KPL; ++k)
      if (k > last) tail += p[k];
    // segmented inclusive scan of the tails over the lanes (a lane that ends a row starts a new segment)
    const unsigned has = __ballot_sync(0xffffffffu, e != 0u);
    const unsigned upto = has & (0xffffffffu >> (31 - lane));
    const int dist = lane - (upto ? 31 - __clz((int)upto) : 0);
    double sc = tail;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, sc, o);
      if (o <= dist) sc += t;
    }
    double acc = __shfl_up_sync(0xffffffffu, sc, 1);   // the open row's partial sum entering this lane
    if (lane == 0) acc = 0.0;
    // first tile-local row index this lane finishes
    const int cnt = __popc(e);
    int rb = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, rb, o);
      if (lane >= o) rb += t;
    }
    rb -= cnt;
    const bool wl = (EPI == EPI_AFCW || EPI == EPI_AFCW_LOCAL) || (EPI == EPI_GENERIC && op.wlast);
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      const bool end = (e >> k) & 1u;
      if (XW && wl) {
        if (end) { val_s[rb] = acc; xw_s[rb] = p[k]; ++rb; acc = 0.0; }
        else acc += p[k];
      } else {
        acc += p[k];
        if (end) { val_s[rb] = acc; ++rb; acc = 0.0; }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const int r = lane + 32 * q;
      if (r < d.nrows) E::finish(op, d.r0 + r, val_s[r], (XW && wl) ? xw_s[r] : 0.0, pc[q]);
    }
    // generic-proxy accesses to this slot are done; order them before the next bulk copy into it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0 && it + STAGES < my) issue(it + STAGES);
  }
}

// ------------------------------------------------------------------------------------------
// Direct (register-staged) engine on the same warp-tile storage: no shared memory, no barrier of any
// kind.  The lane-interleaved blob makes the tile's column / value loads perfectly coalesced
// (32 lanes x 4 / 8 bytes per slot, streaming cache policy), so a warp simply loads its slots, gathers
// x, multiplies and reduces; the head lane of every row runs the row epilogue (consecutive rows ->
// consecutive head lanes -> the vector accesses still fill whole sectors).  With no shared-memory
// ring and 64 registers, 32 warps per SM are resident (the TMA-ring engine: 16) and the whole L1
// serves the gathers: latency is hidden by occupancy instead of by a software pipeline.  Measured on
// the 4096^2 cycle: 3 / 4 / 5 CTAs per SM (80 / 64 / 48 registers) -> 2.62 / 2.56 / 3.01 ms; a 40-register
// "thin warp" variant with 48 warps per SM (tools/microbench/r02_thin_engine.cuh) -> 2.76 ms.
__device__ __forceinline__ int ld_stream_nc(const int *p) { return __ldcs(p); }
template <int EPI, int KP, bool GHOST, int PF>   // PF 1: the next tile's column indices are loaded one tile ahead (3 CTAs per SM instead of 4)
__global__ void __launch_bounds__(256, PF ? 3 : 4) spmv_sv_kernel(const SpmvOp op) {
  typedef EpiT<EPI> E;
  typedef typename E::Pre Pre;
  constexpr bool XW = E::kXw;
  constexpr int NSLOT = kWtSlots;
  constexpr int NS = kWtSlots / KP;
  constexpr bool AHEAD = E::kPre * NS <= 8;   // epilogue operands loaded before the gathers (else right before the row is finished)
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const WtDesc *__restrict__ wdesc = op.wdesc;
  const double *__restrict__ xv = op.x;
  const double *__restrict__ xg = op.xg;
  const int nloc = op.nloc;
  const bool wf = (EPI == EPI_AFCW || EPI == EPI_AFCW_LOCAL) || (EPI == EPI_GENERIC && op.wlast);
  pdl_launch_dependents();
  int t = gw;
  WtDesc dn = {0u, 0, 0, 0}, dn2 = {0u, 0, 0, 0};
  if (t < op.nwt) dn = wdesc[t];
  if (PF && t + nwarps < op.nwt) dn2 = wdesc[t + nwarps];
  int cn[NSLOT];
  if (PF && t < op.nwt) {
    const int nsl = (dn.geom & 0xff) * KP;
    const int *cg = reinterpret_cast<const int *>(op.blob + (size_t)dn.off16 * 16 + nsl * 256);
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) cn[k] = (k < nsl) ? __ldcs(cg + k * 32 + lane) : 0;
  }
  bool waited = false, garmed = GHOST && op.gw_ready != nullptr;
  for (; t < op.nwt; t += nwarps) {
    const WtDesc d = dn;
    if (PF) {
      dn = dn2;
      if (t + 2 * nwarps < op.nwt) dn2 = wdesc[t + 2 * nwarps];
    } else if (t + nwarps < op.nwt) {
      dn = wdesc[t + nwarps];
    }
    const int ns = d.geom & 0xff, gmax = d.geom >> 8;
    const int nslots = ns * KP;
    const unsigned char *b = op.blob + (size_t)d.off16 * 16;
    const double *val_g = reinterpret_cast<const double *>(b);
    const int *col_g = reinterpret_cast<const int *>(b + nslots * 256);
    const unsigned *heads = reinterpret_cast<const unsigned *>(b + nslots * 384);
    // matrix stream: coalesced, read once
    int c[NSLOT];
    double p[NSLOT];
    if (PF) {
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) c[k] = cn[k];
      if (t + nwarps < op.nwt) {   // the next tile's columns: in flight while this tile is gathered, multiplied and reduced
        const int nsl = (dn.geom & 0xff) * KP;
        const int *cg = reinterpret_cast<const int *>(op.blob + (size_t)dn.off16 * 16 + nsl * 256);
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) cn[k] = (k < nsl) ? __ldcs(cg + k * 32 + lane) : 0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) c[k] = (k < nslots) ? __ldcs(col_g + k * 32 + lane) : 0;
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) p[k] = (k < nslots) ? __ldcs(val_g + k * 32 + lane) : 0.0;
    unsigned hd[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) hd[s] = (s < ns) ? __ldg(heads + s) : 0u;
    if (!waited) {   // from here on the vectors written by the previous kernels are read
      pdl_wait(); waited = true;
      if (GHOST) ghost_push(op.xpush, op.x, gw, nwarps, lane);
    }
    if (GHOST) ghost_gate(op, garmed, t, lane);
    // gathers
    double xr[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      xr[k] = 0.0;
      if (k < nslots) {
        if (GHOST) xr[k] = (c[k] >= nloc) ? __ldcg(xg + (c[k] - nloc)) : xv[c[k]];
        else xr[k] = xv[c[k]];
      }
    }
    // epilogue operands of the rows whose head lane I am
    Pre pc[NS];
    int row[NS];
    {
      int rowbase = d.r0;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const unsigned H = hd[s];
        row[s] = ((H >> lane) & 1u) ? rowbase + __popc(H & ((1u << lane) - 1u)) : -1;
        if (AHEAD && row[s] >= 0) pc[s] = E::prefetch(op, row[s]);
        rowbase += __popc(H);
      }
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) p[k] *= xr[k];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < ns) {
        const unsigned H = hd[s];
        double xw = 0.0;
        double q0 = p[s * KP];
        if (XW && wf && row[s] >= 0) { xw = q0; q0 = 0.0; }   // merged A_fc|W: the row's first entry is the W entry
        double acc;
        if (KP == 1) acc = q0;
        else if (KP == 2) acc = q0 + p[s * KP + (KP > 1 ? 1 : 0)];
        else if (KP == 4) acc = (q0 + p[s * KP + (KP > 1 ? 1 : 0)]) + (p[s * KP + (KP > 2 ? 2 : 0)] + p[s * KP + (KP > 2 ? 3 : 0)]);
        else acc = ((q0 + p[(KP > 1 ? 1 : 0)]) + (p[(KP > 2 ? 2 : 0)] + p[(KP > 2 ? 3 : 0)])) +
                   ((p[(KP > 4 ? 4 : 0)] + p[(KP > 4 ? 5 : 0)]) + (p[(KP > 4 ? 6 : 0)] + p[(KP > 4 ? 7 : 0)]));
        if (gmax > 1) {
          const unsigned above = lane < 31 ? (H >> (lane + 1)) : 0u;
          const int dist = above ? __ffs((int)above) - 1 : 31 - lane;   // lanes of my row after me
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            if (o < gmax) {
              const double tt = __shfl_down_sync(0xffffffffu, acc, o);
              if (o <= dist) acc += tt;
            }
          }
        }
        if (row[s] >= 0) {
          if (!AHEAD) pc[s] = E::prefetch(op, row[s]);
          E::finish(op, row[s], acc, xw, pc[s]);
        }
      }
    }
  }
  if (!waited) {   // a warp without tiles still takes part in the push
    pdl_wait();
    if (GHOST) ghost_push(op.xpush, op.x, gw, nwarps, lane);
  }
}

// ------------------------------------------------------------------------------------------
// Round-1 TMA kernel (option kernel=1, A/B baseline): persistent CTAs walk CTA tiles (consecutive rows
// holding <= TILE nonzeros and <= NT rows); one elected thread issues three 1-D bulk copies per tile
// (values, column indices, row pointers) into a STAGES-deep ring; multiply into shared memory, CTA
// barrier, one thread per row sums in stored column order, CTA barrier.
template <int TILE, int MAXROWS>
struct TmaStage {
  double val[TILE + 8];
  int col[TILE + 8];
  int rp[MAXROWS + 8];
};

template <int NT, int TILE, int STAGES>
__global__ void __launch_bounds__(NT) spmv_tma_kernel(const SpmvOp op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef TmaStage<TILE, NT> Stage;
  Stage *stages = reinterpret_cast<Stage *>(smem_raw);
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ TileDesc sdesc[STAGES];
  __shared__ double red[NT / 32 + 1];
  const int tid = threadIdx.x;
  const TileDesc *__restrict__ tiles = op.tiles;
  const int ntiles = op.ntiles;
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = (first < ntiles) ? (ntiles - first + stride - 1) / stride : 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t pol = l2_policy_evict_first();
  TileDesc dnext = {0, 0, 0, 0};
  if (tid == 0 && my_tiles > 0) dnext = tiles[first];
  auto issue = [&](int j) {
    const TileDesc d = dnext;
    if (j + 1 < my_tiles) dnext = tiles[first + (j + 1) * stride];
    const int slot = j % STAGES;
    sdesc[slot] = d;
    if (d.n <= TILE) {
      Stage &S = stages[slot];
      const int s_al = d.s & ~3;
      const int cnt = (d.n + (d.s - s_al) + 3) & ~3;
      const int r_al = d.r0 & ~3;
      const int rcnt = (d.nrows + 1 + (d.r0 - r_al) + 3) & ~3;
      mbar_expect_tx(&full[slot], (uint32_t)(cnt * 12 + rcnt * 4));
      tma_load_1d(S.val, op.val + s_al, (uint32_t)(cnt * 8), &full[slot], pol);
      tma_load_1d(S.col, op.col + s_al, (uint32_t)(cnt * 4), &full[slot], pol);
      tma_load_1d(S.rp, op.rp + r_al, (uint32_t)(rcnt * 4), &full[slot], pol);
    } else {
      mbar_expect_tx(&full[slot], 0);  // long row: streamed straight from global by the whole CTA
    }
  };
  pdl_launch_dependents();
  if (tid == 0) {
    for (int j = 0; j < STAGES - 1 && j < my_tiles; ++j) issue(j);   // matrix data only: legal before pdl_wait
  }
  pdl_wait();
  ghost_push(op.xpush, op.x, (blockIdx.x * NT + tid) >> 5, (gridDim.x * NT) >> 5, tid & 31);
  if (tid == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
  __syncthreads();

  for (int it = 0; it < my_tiles; ++it) {
    const int slot = it % STAGES;
    const TileDesc d = sdesc[slot];
    Stage &S = stages[slot];
    if (d.n <= TILE) {
      const bool has_row = tid < d.nrows;
      EpiPre pre;
      if (has_row) pre = epi_prefetch(op, d.r0 + tid);
      mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
      const int o = d.s & 3;
      constexpr int kIter = TILE / NT;
      double xr[kIter];
#pragma unroll
      for (int k0 = 0; k0 < kIter; ++k0) {
        const int k = tid + k0 * NT;
        xr[k0] = 0.0;
        if (k < d.n) xr[k0] = gather_x(op, S.col[o + k]);
      }
      // the latency-critical gathers of THIS tile are queued ahead of the next tile's bulk copies
      if (tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);
#pragma unroll
      for (int k0 = 0; k0 < kIter; ++k0) {
        const int k = tid + k0 * NT;
        if (k < d.n) S.val[o + k] = S.val[o + k] * xr[k0];
      }
      __syncthreads();
      if (has_row) {
        const int ro = d.r0 & 3;
        int p = S.rp[ro + tid] - d.s + o;
        int q = S.rp[ro + tid + 1] - d.s + o;
        double xw = 0.0;
        if (op.wlast) { --q; xw = S.val[q]; }
        double sum = 0.0;
        for (; p < q; ++p) sum += S.val[p];
        epi_finish(op, d.r0 + tid, sum, xw, pre);
      }
    } else {
      if (tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);
      mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
      const int e = d.s + d.n;
      const int last = op.wlast ? e - 1 : e;
      double part = 0.0;
      for (int k = d.s + tid; k < last; k += NT) part += ld_stream(op.val + k) * gather_x(op, ld_stream(op.col + k));
#pragma unroll
      for (int w = 16; w > 0; w >>= 1) part += __shfl_xor_sync(0xffffffffu, part, w);
      if ((tid & 31) == 0) red[tid >> 5] = part;
      __syncthreads();
      if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < NT / 32; ++w) sum += red[w];
        const double xw = op.wlast ? op.val[last] * gather_x(op, op.col[last]) : 0.0;
        row_epilogue(op, d.r0, sum, xw);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
}

__device__ __forceinline__ void ew_apply(const EwOp &e, int i) {
  double v = e.alpha * e.a[e.gather ? e.gather[i] : i];
  if (e.b) v = v * e.b[i];
  if (e.dv) v = v / e.dv[i];
  const int o = e.scatter ? e.scatter[i] : i;
  if (e.mode == 1) e.out[o] = v;
  else e.out[o] += v;
}

__global__ void __launch_bounds__(kThreads) ew_kernel(const EwOp e) {
  pdl_launch_dependents();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += gridDim.x * blockDim.x) ew_apply(e, i);
}

// Dense collapsed tail: y = T x, T row-major n x n.  One CTA per row (grid-stride): every warp streams a
// contiguous eighth of the row with 8 independent loads per lane in flight, fixed summation order
// (8 interleaved partial sums per lane, a shuffle tree, then the warps in order) -> deterministic.
__global__ void __launch_bounds__(kThreads) dense_gemv_kernel(int n, const double *__restrict__ T, const double *__restrict__ x,
                                                              double *__restrict__ y) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double part[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NWARP = kThreads / 32;
  const int chunk = (n + NWARP - 1) / NWARP;
  const int j0 = w * chunk, j1 = min(n, j0 + chunk);
  for (int row = blockIdx.x; row < n; row += gridDim.x) {
    const double *t = T + (size_t)row * n;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int j = j0 + lane;
    for (; j + 7 * 32 < j1; j += 8 * 32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += __ldcs(t + j + u * 32) * x[j + u * 32];
    }
    for (int u = 0; j < j1; j += 32, ++u) acc[u] += __ldcs(t + j) * x[j];
    double s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) part[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0.0;
#pragma unroll
      for (int k = 0; k < NWARP; ++k) r += part[k];
      y[row] = r;
    }
    __syncthreads();
  }
}

// setup helpers of the dense tail: unit vector, strided column store
__global__ void unit_vector_kernel(int n, int j, double *v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = (i == j) ? 1.0 : 0.0;
}
__global__ void store_column_kernel(int n, int j, const double *v, double *T) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) T[(size_t)i * n + j] = v[i];
}

// ---- BLAS-1 of the outer Krylov method (KSPGMRES with classical Gram-Schmidt / KSPRICHARDSON around PCApply)
constexpr int kKspMaxVec = 32;     // restart + 2 <= 32 basis vectors per multi-dot / multi-axpy
struct KspCoef { double c[kKspMaxVec]; };
// partial[cta][i] = sum over the CTA's elements of w[j] * V[i][j], i < nv: ONE pass over w and the basis
// (fixed grid and fixed order inside a CTA -> deterministic)
__global__ void __launch_bounds__(kThreads) multi_dot_kernel(int n, const double *__restrict__ w, const double *__restrict__ V, long long ld,
                                                             int nv, double *__restrict__ partial) {
  __shared__ double red[kThreads / 32][kKspMaxVec];
  double acc[kKspMaxVec];
#pragma unroll
  for (int i = 0; i < kKspMaxVec; ++i) acc[i] = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double wj = w[j];
#pragma unroll
    for (int i = 0; i < kKspMaxVec; ++i)
      if (i < nv) acc[i] += wj * V[(size_t)i * ld + j];
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kKspMaxVec; ++i) {
    if (i < nv) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[wp][i] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) v += red[k][threadIdx.x];
    partial[(size_t)blockIdx.x * kKspMaxVec + threadIdx.x] = v;
  }
}
// out[i] = sum_cta partial[cta][i] in CTA order
__global__ void multi_dot_finish_kernel(int ncta, int nv, const double *__restrict__ partial, double *__restrict__ out) {
  const int i = threadIdx.x;
  if (i < nv) {
    double v = 0.0;
    for (int k = 0; k < ncta; ++k) v += partial[(size_t)k * kKspMaxVec + i];
    out[i] = v;
  }
}
// y[j] = beta * y[j] + sum_i co.c[i] * V[i][j]
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(int n, double *__restrict__ y, double beta, const double *__restrict__ V,
                                                              long long ld, int nv, const KspCoef co) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    double v = beta != 0.0 ? beta * y[j] : 0.0;
#pragma unroll
    for (int i = 0; i < kKspMaxVec; ++i)
      if (i < nv) v += co.c[i] * V[(size_t)i * ld + j];
    y[j] = v;
  }
}

// ---- peer-memory ghost exchange (one process per GPU with IPC-mapped arenas, or an in-process group)
struct PushOp {
  int n;                        // entries to push
  const int *idx;               // positions in x
  const double *x;
  int nranks, me;
  const int *send_off;          // [nranks + 1] prefix of send counts
  const unsigned long long *dst;        // [nranks] address (in MY address space) of my chunk inside peer p's ghost buffer
  const unsigned long long *peer_flags; // [nranks] address of peer p's flag block
  unsigned *my_flags;           // my own flag block (acks are written here by the consumers)
  const unsigned *epoch;
  unsigned *done;               // CTA counter of this instance
  int inst, ack_inst, ack_delta, max_inst;
  unsigned dstmask;
};
// flag block layout: [0..63] header (epoch at word 0), ready[max_inst][32], ack[max_inst][32]
__device__ __forceinline__ size_t flag_ready(int inst, int q) { return 64 + (size_t)inst * 32 + q; }
__device__ __forceinline__ size_t flag_ack(int max_inst, int inst, int q) { return 64 + (size_t)max_inst * 32 + (size_t)inst * 32 + q; }

__global__ void epoch_kernel(unsigned *epoch) { *epoch += 1; }

__global__ void __launch_bounds__(kThreads) push_kernel(const PushOp o) {
  const unsigned e = *o.epoch;
  if (threadIdx.x == 0 && o.ack_inst >= 0) {
    // the consumers must have finished reading the previous contents of their ghost buffer
    for (unsigned m = o.dstmask; m; m &= m - 1) {
      const int p = __ffs(m) - 1;
      const unsigned *ack = o.my_flags + flag_ack(o.max_inst, o.ack_inst, p);
      while ((int)(ld_acquire_sys(ack) - (e - (unsigned)o.ack_delta)) < 0) __nanosleep(20);
    }
  }
  __syncthreads();
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < o.n; j += gridDim.x * blockDim.x) {
    int p = 0;
    while (j >= o.send_off[p + 1]) ++p;
    double *dst = reinterpret_cast<double *>(o.dst[p]) + (j - o.send_off[p]);
    *dst = o.x[o.idx[j]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(o.done, 1u);
    if (prev == gridDim.x - 1) {   // last CTA: every chunk is written and fenced -> raise the flags
      *o.done = 0;
      __threadfence_system();
      for (unsigned m = o.dstmask; m; m &= m - 1) {
        const int p = __ffs(m) - 1;
        st_release_sys(reinterpret_cast<unsigned *>(o.peer_flags[p]) + flag_ready(o.inst, o.me), e);
      }
    }
  }
}

// p2p=2: the push on its own (the consuming operator has no rows on this rank, or an in-process rank group on ONE stream,
// where a consumer that waited inside its kernel for a later launch would never return)
__global__ void __launch_bounds__(kThreads) push2_kernel(const XPush *px, const double *x) {
  pdl_launch_dependents();
  pdl_wait();
  ghost_push(px, x, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, (gridDim.x * blockDim.x) >> 5, threadIdx.x & 31);
}
// p2p=2: enter cycle e = ++epoch, tell the ranks in `raise` (they push to me: everything this rank read in cycle e-1 is
// finished, stream order), and wait until the ranks in `wait` (I push to them) have entered cycle e as well
__global__ void epoch2_kernel(unsigned *flags, const unsigned long long *peer_flags, int me, unsigned raise, unsigned wait) {
  const unsigned e = *flags + 1;
  *flags = e;
  __threadfence_system();
  for (unsigned m = raise; m; m &= m - 1) {
    const int p = __ffs(m) - 1;
    st_release_sys(reinterpret_cast<unsigned *>(peer_flags[p]) + 32 + me, e);
  }
  for (unsigned m = wait; m; m &= m - 1) spin_until(flags + 32 + (__ffs(m) - 1), e, flags + 1);
}

// consumer -> producers: "I have finished reading the ghosts of instance inst" (runs after the SpMV)
__global__ void ack_kernel(const unsigned long long *peer_flags, const unsigned *epoch, int max_inst, int inst, int me, unsigned srcmask) {
  const unsigned e = *epoch;
  for (unsigned m = srcmask; m; m &= m - 1) {
    const int q = __ffs(m) - 1;
    st_release_sys(reinterpret_cast<unsigned *>(peer_flags[q]) + flag_ack(max_inst, inst, me), e);
  }
}

}  // namespace pfb
