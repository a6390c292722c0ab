// kernels.cuh -- sm_100a device code of the AIRG V-cycle apply.
//
// One SpMV "mega-op" covers every MatMult + Vec-AXPY chain of the reference's apply path
// (SURVEY.md section 2c): the CSR product of a row is reduced once and the row epilogue applies
// the fused vector updates.  The path is HBM-bound fp64/int32 work (no tensor cores by design).
//
//   spmv_tma_kernel     the kernel of the large levels: persistent CTAs, the matrix stream
//                       (values, column indices, row pointers of a tile) is brought into a
//                       shared-memory ring by 1-D TMA bulk copies signalled on mbarriers, so the
//                       HBM stream never drains while a tile is multiplied / reduced; the
//                       per-row epilogue operands are prefetched into registers before the
//                       tile's barrier is waited on.
//   spmv_stream_kernel  the first-generation smem-staged kernel (option kernel=0; A/B baseline).
//   tail_kernel         single CTA that runs the whole list of ops of the small coarse levels
//                       back to back with CTA barriers instead of kernel launches.
//   ew_kernel           diagonal inverses / scalings / permutations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfb {

constexpr int kThreads = 256;       // CTA size of the streaming kernels
constexpr int kTile = 2048;         // nnz per CTA tile of the stream / tail kernels (16 KB of fp64 products)
constexpr int kMaxRowsPerBlk = 1024;
constexpr int kTailThreads = 1024;  // CTA size of the single-CTA tail kernel

struct TileDesc { int r0, nrows, s, n; };  // first row, #rows, first nnz, #nnz (n > tile size: one long row)

struct SpmvOp {
  // CSR block + its row-block partition
  const int *rp, *col;
  const double *val;
  int m, nblk;
  const int *blk;
  const TileDesc *tiles;       // tile list of the TMA-pipelined kernel
  int ntiles;
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
  // peer-memory ghost exchange: before the first ghost read, wait until every source rank has pushed
  // its chunk of THIS exchange instance (ready[q] >= *epoch for the ranks q in srcmask)
  const unsigned *gw_ready; const unsigned *gw_epoch; unsigned gw_srcmask;
  // wide-tile kernel: up to 3 epilogue operand arrays travel with the tile as TMA bulk copies
  // (field ids: 1 aux, 2 D, 3 x_i (Neumann), 4 fd_a, 5 fd_m, 6 out (read-modify-write), 7 acc_src, 8 acc)
  unsigned char stg_field[3]; unsigned char n_stg; unsigned stg_mask;
  int wide;                    // this operator's tile list was built for spmv_tma_wide_kernel
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

struct DevOp {
  int kind;  // 0 spmv, 1 elementwise
  SpmvOp s;
  EwOp e;
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
// consumer side: called by ONE thread of a CTA before the CTA's first ghost read
__device__ __forceinline__ void ghost_wait(const unsigned *ready, const unsigned *epoch, unsigned srcmask) {
  if (!ready) return;
  const unsigned e = *epoch;
  for (unsigned m = srcmask; m; m &= m - 1) {
    const int q = __ffs(m) - 1;
    while ((int)(ld_acquire_sys(ready + q) - e) < 0) __nanosleep(20);
  }
}

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

__device__ __forceinline__ double gather_x(const SpmvOp &op, int c) {
  if (op.xg != nullptr && c >= op.nloc) return __ldcg(op.xg + (c - op.nloc));  // ghosts: written by peers, read through L2
  return op.x[c];
}

// Row epilogue, split in two so that the loads that depend only on the row index can be issued
// long before the row sum exists.
struct EpiPre { double aux, D, xi, fa, fm, out, accsrc, acc; };

__device__ __forceinline__ EpiPre epi_prefetch(const SpmvOp &op, int i) {
  EpiPre p;
  p.aux = op.aux ? op.aux[i] : 0.0;
  p.D = op.D ? op.D[i] : 1.0;
  p.xi = op.neumann ? op.x[i] : 0.0;
  p.fa = op.fd_its > 0 ? op.fd_a[i] : 0.0;
  p.fm = op.fd_its > 0 ? op.fd_m[i] : 0.0;
  p.out = op.out_mode == 2 ? op.out[i] : 0.0;
  p.accsrc = (op.acc_mode && op.acc_src) ? op.acc_src[i] : 0.0;
  p.acc = op.acc_mode == 2 ? op.acc[i] : 0.0;
  return p;
}

__device__ __forceinline__ const double *epi_field_ptr(const SpmvOp &op, int f) {
  switch (f) {
    case 1: return op.aux;
    case 2: return op.D;
    case 3: return op.x;
    case 4: return op.fd_a;
    case 5: return op.fd_m;
    case 6: return op.out;
    case 7: return op.acc_src;
    default: return op.acc;
  }
}
// epi_prefetch for the fields that were NOT staged (mask bit f-1 set = staged)
__device__ __forceinline__ EpiPre epi_prefetch_masked(const SpmvOp &op, int i, unsigned m) {
  EpiPre p;
  p.aux = (op.aux && !(m & 1u)) ? op.aux[i] : 0.0;
  p.D = (op.D && !(m & 2u)) ? op.D[i] : 1.0;
  p.xi = (op.neumann && !(m & 4u)) ? op.x[i] : 0.0;
  p.fa = (op.fd_its > 0 && !(m & 8u)) ? op.fd_a[i] : 0.0;
  p.fm = (op.fd_its > 0 && !(m & 16u)) ? op.fd_m[i] : 0.0;
  p.out = (op.out_mode == 2 && !(m & 32u)) ? op.out[i] : 0.0;
  p.accsrc = (op.acc_mode && op.acc_src && !(m & 64u)) ? op.acc_src[i] : 0.0;
  p.acc = (op.acc_mode == 2 && !(m & 128u)) ? op.acc[i] : 0.0;
  return p;
}
__device__ __forceinline__ void epi_set_field(EpiPre &p, int f, double v) {
  switch (f) {
    case 1: p.aux = v; break;
    case 2: p.D = v; break;
    case 3: p.xi = v; break;
    case 4: p.fa = v; break;
    case 5: p.fm = v; break;
    case 6: p.out = v; break;
    case 7: p.accsrc = v; break;
    default: p.acc = v; break;
  }
}

__device__ __forceinline__ void epi_finish(const SpmvOp &op, int i, double s, double xw, const EpiPre &p) {
  if (op.D) s = s / p.D;
  if (op.neumann) s = p.xi - s;
  double v = op.beta * s;
  if (op.aux) v = op.alpha * p.aux + v;
  if (op.wout) {
    for (int it = 0; it < op.fd_its; ++it) xw = xw + p.fm * (v - p.fa * xw);
    op.wout[i] = xw;
  }
  if (op.out_mode == 1) op.out[i] = v;
  else if (op.out_mode == 2) op.out[i] = p.out + v;
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
      row_epilogue(op, r, sum, xw);
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
      row_epilogue(op, r0, sum, xw);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) spmv_stream_kernel(const SpmvOp op) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double prod[kTile];
  __shared__ double red[kThreads / 32];
  if (threadIdx.x == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
  __syncthreads();
  for (int b = blockIdx.x; b < op.nblk; b += gridDim.x) process_block<kThreads>(op, b, prod, red);
}

// ------------------------------------------------------------------------------------------
// TMA-pipelined streaming SpMV (the kernel of the large levels).
//
// A persistent CTA walks its tiles (tile = consecutive rows holding <= TILE nonzeros and <= NT
// rows, fixed at upload).  For every tile ONE elected thread issues three 1-D bulk copies
// (cp.async.bulk, the TMA engine: SASS UBLKCP) that bring the tile's values, column indices and
// row pointers into a STAGES-deep shared-memory ring; completion is signalled on an mbarrier per
// stage.  While tile t is being multiplied/reduced, tiles t+1 .. t+STAGES-1 are in flight, so the
// HBM stream does not drain at the barriers of the multiply/reduce phases.  Every thread owns at
// most one row of the tile and issues the loads of that row's epilogue operands BEFORE it waits
// for the tile, so the reduce phase touches no global-memory latency.  x is gathered with
// ordinary loads (L1/L2; the nested CF ordering keeps the gathers near-sequential).
//
// Bulk copies need 16-byte aligned addresses and sizes: the copy starts at the tile's first
// nonzero rounded DOWN to a multiple of 4 entries and is rounded UP to a multiple of 4 (the
// arrays are over-allocated by a few entries at upload).
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

template <int TILE, int MAXROWS>
struct TmaStage {
  double val[TILE + 8];
  int col[TILE + 8];
  int rp[MAXROWS + 8];
};

template <int NT, int TILE, int STAGES, int MODE, int MINB>   // MODE 0: serial row sums, 1: row-mapped multiply, 2: g-lane row sums
__global__ void __launch_bounds__(NT, MINB) spmv_tma_kernel(const SpmvOp op) {
  constexpr bool ROWMAP = MODE == 1;
  constexpr bool GRED = MODE == 2;
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
  // producer (thread 0): issue the bulk copies of local tile j into ring slot j % STAGES
  // (the descriptor of the NEXT tile is fetched one issue ahead so the elected thread never
  // stalls on a global load between a tile's barrier and the next bulk copy)
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
  pdl_wait();   // from here on the vectors written by the previous kernels are read
  if (tid == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);   // peers' pushes (the local tiles are already in flight)
  __syncthreads();

  for (int it = 0; it < my_tiles; ++it) {
    const int slot = it % STAGES;
    const TileDesc d = sdesc[slot];
    // when to issue the bulk copies of tile it+STAGES-1 (its slot was freed by the barrier ending
    // iteration it-1): normally right after this tile's gathers have been queued
    const bool late_issue = !ROWMAP && d.n <= TILE;
    if (!late_issue && tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);
    Stage &S = stages[slot];
    if (d.n <= TILE) {
      // operands of my row's epilogue: in flight while the tile lands and is multiplied
      int g = 1;
      if (ROWMAP || GRED) {
        while (g < 32 && d.nrows * (g << 1) <= NT) g <<= 1;
      }
      const bool has_row = (ROWMAP || GRED) ? (tid < d.nrows * g && (tid & (g - 1)) == 0) : (tid < d.nrows);
      EpiPre pre;
      if (has_row) pre = epi_prefetch(op, d.r0 + ((ROWMAP || GRED) ? tid / g : tid));
      mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
      const int o = d.s & 3;
      if (ROWMAP) {
        // g lanes per row (g = largest power of two with nrows * g <= NT, at most 32): lanes of
        // neighbouring rows gather neighbouring x entries, so a warp's gather touches few lines
        const bool active = tid < d.nrows * g;
        const int row = tid / g, lg = tid & (g - 1);
        int p = 0, q = 0;
        if (active) {
          const int ro = d.r0 & 3;
          p = S.rp[ro + row] - d.s + o;
          q = S.rp[ro + row + 1] - d.s + o;
        }
        const int qs = op.wlast ? q - 1 : q;
        double sum = 0.0, xw = 0.0;
        for (int k = p + lg; k < qs; k += g) sum += S.val[k] * gather_x(op, S.col[k]);
        if (op.wlast && active && lg == ((qs - p) & (g - 1))) xw = S.val[qs] * gather_x(op, S.col[qs]);
        for (int w = g >> 1; w > 0; w >>= 1) {   // all lanes of the warp take part (inactive ones carry zeros)
          sum += __shfl_down_sync(0xffffffffu, sum, w, g);
          if (op.wlast) xw += __shfl_down_sync(0xffffffffu, xw, w, g);
        }
        if (active && lg == 0) epi_finish(op, d.r0 + row, sum, xw, pre);
      } else {
        constexpr int kIter = TILE / NT;
        double xr[kIter];
#pragma unroll
        for (int k0 = 0; k0 < kIter; ++k0) {
          const int k = tid + k0 * NT;
          xr[k0] = 0.0;
          if (k < d.n) {
            xr[k0] = gather_x(op, S.col[o + k]);
          }
        }
        // the latency-critical gathers of THIS tile are queued ahead of the next tile's bulk copies
        if (late_issue && tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);
#pragma unroll
        for (int k0 = 0; k0 < kIter; ++k0) {
          const int k = tid + k0 * NT;
          if (k < d.n) S.val[o + k] = S.val[o + k] * xr[k0];
        }
        __syncthreads();
        if (GRED) {
          // g lanes per row read the row's products at consecutive addresses (few bank conflicts) and
          // shuffle-reduce; the lane-0 thread of a group owns the row's epilogue
          const bool active = tid < d.nrows * g;
          const int row = tid / g, lg = tid & (g - 1);
          int p = 0, q = 0;
          if (active) {
            const int ro = d.r0 & 3;
            p = S.rp[ro + row] - d.s + o;
            q = S.rp[ro + row + 1] - d.s + o;
          }
          double xw = 0.0;
          if (op.wlast && active) { --q; xw = S.val[q]; }
          double sum = 0.0;
          for (int k = p + lg; k < q; k += g) sum += S.val[k];
          for (int w = g >> 1; w > 0; w >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, w, g);
          if (has_row) epi_finish(op, d.r0 + row, sum, xw, pre);
        } else if (has_row) {
          const int ro = d.r0 & 3;
          int p = S.rp[ro + tid] - d.s + o;
          int q = S.rp[ro + tid + 1] - d.s + o;
          double xw = 0.0;
          if (op.wlast) { --q; xw = S.val[q]; }
          double sum = 0.0;
          for (; p < q; ++p) sum += S.val[p];
          epi_finish(op, d.r0 + tid, sum, xw, pre);
        }
      }
    } else {
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
    // generic-proxy accesses to this slot are done; order them before the next bulk copy into it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Second TMA kernel: NOTHING in a tile's processing waits on global memory.
//   * matrix stream (values, columns, row pointers): TMA bulk copies, STAGES-deep ring (as above);
//   * x gathers of tile t+1: issued as 8-byte cp.async (LDGSTS) into a double-buffered shared array
//     while tile t is being reduced -- the dependent gather latency is off the critical path;
//   * epilogue operands of tile t+1: register prefetch one tile ahead;
//   * multiply + reduce fused: g lanes per row (g = largest power of two with rows*g <= NT, <= 32)
//     read values and gathered x from shared memory, shuffle-reduce, lane 0 runs the epilogue.
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, int TILE, int STAGES>
__global__ void __launch_bounds__(NT) spmv_tma2_kernel(const SpmvOp op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef TmaStage<TILE, NT> Stage;
  Stage *stages = reinterpret_cast<Stage *>(smem_raw);
  double *xring = reinterpret_cast<double *>(smem_raw + sizeof(Stage) * STAGES);   // [2][TILE]
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
      mbar_expect_tx(&full[slot], 0);
    }
  };
  // lanes per row of a tile
  auto lanes_per_row = [](int nrows) { int g = 1; while (g < 32 && nrows * (g << 1) <= NT) g <<= 1; return g; };
  // stage the gathers + epilogue operands of local tile j (its matrix data must have landed)
  EpiPre pre_next;
  auto stage_gathers = [&](int j) {
    const int slot = j % STAGES;
    const TileDesc d = sdesc[slot];
    mbar_wait(&full[slot], (uint32_t)((j / STAGES) & 1));
    if (d.n <= TILE) {
      const Stage &S = stages[slot];
      double *xs = xring + (j & 1) * TILE;
      const int o = d.s & 3;
      constexpr int kIter = TILE / NT;
#pragma unroll
      for (int k0 = 0; k0 < kIter; ++k0) {
        const int k = tid + k0 * NT;
        if (k < d.n) {
          const int c = S.col[o + k];
          const double *src = (op.xg != nullptr && c >= op.nloc) ? op.xg + (c - op.nloc) : op.x + c;
          cp_async8(xs + k, src);
        }
      }
      const int g = lanes_per_row(d.nrows);
      if (tid < d.nrows * g && (tid & (g - 1)) == 0) pre_next = epi_prefetch(op, d.r0 + tid / g);
    }
    cp_async_commit();
  };
  pdl_launch_dependents();
  if (tid == 0) {
    for (int j = 0; j < STAGES - 1 && j < my_tiles; ++j) issue(j);
  }
  pdl_wait();
  if (tid == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
  __syncthreads();
  if (my_tiles > 0) stage_gathers(0);

  for (int it = 0; it < my_tiles; ++it) {
    const int slot = it % STAGES;
    if (tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);  // slot freed by the barrier ending iteration it-1
    const TileDesc d = sdesc[slot];
    Stage &S = stages[slot];
    const EpiPre pre = pre_next;
    if (it + 1 < my_tiles) { stage_gathers(it + 1); cp_async_wait<1>(); }   // tile it's gathers are complete, tile it+1's in flight
    else cp_async_wait<0>();
    __syncthreads();
    if (d.n <= TILE) {
      const double *xs = xring + (it & 1) * TILE;
      const int o = d.s & 3;
      const int g = lanes_per_row(d.nrows);
      const bool active = tid < d.nrows * g;
      const int row = tid / g, lg = tid & (g - 1);
      int p = 0, q = 0;
      if (active) {
        const int ro = d.r0 & 3;
        p = S.rp[ro + row] - d.s;      // tile-relative
        q = S.rp[ro + row + 1] - d.s;
      }
      const int qs = op.wlast ? q - 1 : q;
      double sum = 0.0, xw = 0.0;
      for (int k = p + lg; k < qs; k += g) sum += S.val[o + k] * xs[k];
      if (op.wlast && active && lg == ((qs - p) & (g - 1))) xw = S.val[o + qs] * xs[qs];
      for (int w = g >> 1; w > 0; w >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, w, g);
        if (op.wlast) xw += __shfl_down_sync(0xffffffffu, xw, w, g);
      }
      if (active && lg == 0) epi_finish(op, d.r0 + row, sum, xw, pre);
    } else {
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

// ------------------------------------------------------------------------------------------
// Wide-tile TMA kernel for operators with very short rows (levels 1-4 of upwind problems: 1-3 nonzeros
// per row).  With <= NT rows per tile such a tile moves only ~5 KB while the per-tile latency chain
// (wait, gather, barrier, reduce, barrier) is ~1.5 us regardless of its size; here a tile holds up to
// MAXROWS = 4 NT rows, every thread reduces up to 4 rows, and the rows' epilogue operands arrive with the
// tile as extra bulk copies (contiguous row range) instead of per-thread register prefetches.
template <int TILE, int MAXROWS, int NOPD>
struct WideStage {
  double val[TILE + 8];
  double opd[NOPD][MAXROWS + 2];
  int col[TILE + 8];
  int rp[MAXROWS + 8];
};

template <int NT, int TILE, int MAXROWS, int STAGES>
__global__ void __launch_bounds__(NT) spmv_tma_wide_kernel(const SpmvOp op) {
  constexpr int NOPD = 3;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef WideStage<TILE, MAXROWS, NOPD> Stage;
  Stage *stages = reinterpret_cast<Stage *>(smem_raw);
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ TileDesc sdesc[STAGES];
  __shared__ int sshift[STAGES][NOPD];
  __shared__ double red[NT / 32 + 1];
  const int tid = threadIdx.x;
  const TileDesc *__restrict__ tiles = op.tiles;
  const int ntiles = op.ntiles;
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = (first < ntiles) ? (ntiles - first + stride - 1) / stride : 0;
  const int nstg = op.n_stg;
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
      uint32_t bytes = (uint32_t)(cnt * 12 + rcnt * 4);
      unsigned long long a_al[NOPD]; uint32_t ob[NOPD];
      for (int k = 0; k < nstg; ++k) {
        const unsigned long long a = (unsigned long long)(epi_field_ptr(op, op.stg_field[k]) + d.r0);
        a_al[k] = a & ~15ull;
        const int sh = (int)((a - a_al[k]) >> 3);
        sshift[slot][k] = sh;
        ob[k] = (uint32_t)(((d.nrows + sh + 1) & ~1) * 8);
        bytes += ob[k];
      }
      mbar_expect_tx(&full[slot], bytes);
      tma_load_1d(S.val, op.val + s_al, (uint32_t)(cnt * 8), &full[slot], pol);
      tma_load_1d(S.col, op.col + s_al, (uint32_t)(cnt * 4), &full[slot], pol);
      tma_load_1d(S.rp, op.rp + r_al, (uint32_t)(rcnt * 4), &full[slot], pol);
      for (int k = 0; k < nstg; ++k) tma_load_1d(S.opd[k], (const void *)a_al[k], ob[k], &full[slot], pol);
    } else {
      mbar_expect_tx(&full[slot], 0);
    }
  };
  pdl_launch_dependents();
  pdl_wait();   // the staged operands are vectors written by the previous kernels
  if (tid == 0) {
    ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask);
    for (int j = 0; j < STAGES - 1 && j < my_tiles; ++j) issue(j);
  }
  __syncthreads();

  for (int it = 0; it < my_tiles; ++it) {
    const int slot = it % STAGES;
    const TileDesc d = sdesc[slot];
    Stage &S = stages[slot];
    mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
    if (d.n <= TILE) {
      const int o = d.s & 3;
      constexpr int kIter = TILE / NT;
      double xr[kIter];
#pragma unroll
      for (int k0 = 0; k0 < kIter; ++k0) {
        const int k = tid + k0 * NT;
        xr[k0] = 0.0;
        if (k < d.n) xr[k0] = gather_x(op, S.col[o + k]);
      }
      if (tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);
#pragma unroll
      for (int k0 = 0; k0 < kIter; ++k0) {
        const int k = tid + k0 * NT;
        if (k < d.n) S.val[o + k] = S.val[o + k] * xr[k0];
      }
      __syncthreads();
      const int ro = d.r0 & 3;
      for (int r = tid; r < d.nrows; r += NT) {
        int p = S.rp[ro + r] - d.s + o;
        int q = S.rp[ro + r + 1] - d.s + o;
        double xw = 0.0;
        if (op.wlast) { --q; xw = S.val[q]; }
        double sum = 0.0;
        for (; p < q; ++p) sum += S.val[p];
        EpiPre pre = epi_prefetch_masked(op, d.r0 + r, op.stg_mask);
        for (int k = 0; k < nstg; ++k) epi_set_field(pre, op.stg_field[k], S.opd[k][sshift[slot][k] + r]);
        epi_finish(op, d.r0 + r, sum, xw, pre);
      }
    } else {
      if (tid == 0 && it + STAGES - 1 < my_tiles) issue(it + STAGES - 1);
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

// Dense collapsed tail: y = T x, T row-major n x n.  One warp per row, fixed summation order
// (8 interleaved partial sums per lane, then a shuffle tree) -> deterministic.
__global__ void __launch_bounds__(kThreads) dense_gemv_kernel(int n, const double *__restrict__ T, const double *__restrict__ x,
                                                              double *__restrict__ y) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < n; row += nwarps) {
    const double *t = T + (size_t)row * n;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int j = lane;
    for (; j + 7 * 32 < n; j += 8 * 32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += __ldcs(t + j + u * 32) * x[j + u * 32];
    }
    for (int u = 0; j < n; j += 32, ++u) acc[u] += __ldcs(t + j) * x[j];
    double s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
  }
}

// setup helpers of the dense tail: unit vector, strided column store
__global__ void unit_vector_kernel(int n, int j, double *v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = (i == j) ? 1.0 : 0.0;
}
__global__ void store_column_kernel(int n, int j, const double *v, double *T) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) T[(size_t)i * n + j] = v[i];
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

// consumer -> producers: "I have finished reading the ghosts of instance inst" (runs after the SpMV)
__global__ void ack_kernel(const unsigned long long *peer_flags, const unsigned *epoch, int max_inst, int inst, int me, unsigned srcmask) {
  const unsigned e = *epoch;
  for (unsigned m = srcmask; m; m &= m - 1) {
    const int q = __ffs(m) - 1;
    st_release_sys(reinterpret_cast<unsigned *>(peer_flags[q]) + flag_ack(max_inst, inst, me), e);
  }
}

// Single-CTA "tail": runs a whole list of ops (the small coarse levels: restrictions, coarse
// solve, prolongation + smoothing) back to back with CTA barriers instead of kernel launches.
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const DevOp *ops, int nops) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double prod[kTile];
  __shared__ double red[kTailThreads / 32];
  for (int o = 0; o < nops; ++o) {
    const DevOp &d = ops[o];
    if (d.kind == 0) {
      for (int b = 0; b < d.s.nblk; ++b) process_block<kTailThreads>(d.s, b, prod, red);
    } else {
      for (int i = threadIdx.x; i < d.e.n; i += kTailThreads) ew_apply(d.e, i);
    }
    __syncthreads();
  }
}

}  // namespace pfb
