// spmv_pipe.cu -- microbenchmark: which pipeline structure lets a CSR SpMV with an axpy epilogue
// (y = b - A x, the shape of the AIRG residual op) approach the HBM roofline on B200?
//
//   V0  the product kernel's structure (round 1): TMA ring for (val, col), x gathered with plain loads
//       inside the tile loop, products to smem, barrier, one thread per row, barrier.
//   V2  warp-specialised, everything asynchronous: a producer warp issues the TMA bulk copies (val, col,
//       b) several tiles ahead AND, two tiles behind that front, the x gathers as 8-byte cp.async into a
//       shared-memory ring (completion on an mbarrier via cp.async.mbarrier.arrive); consumer warps only
//       ever touch shared memory (fused multiply + g-lane row reduce) and store y; slots are recycled
//       through "empty" mbarriers -- no __syncthreads in the loop.
//
// Synthetic matrix: n rows, L nonzeros per row at columns row + d_k (banded, sorted), fp64 / int32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o spmv_pipe spmv_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma1d(void *dst, const void *src, uint32_t n, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory"); }
// the executing thread's prior cp.async operations arrive on the mbarrier when they complete (count pre-accounted)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }

// ------------------------------------------------------------------------------------------ V0
template <int NT, int TILE, int STAGES, int L>
__global__ void __launch_bounds__(NT) spmv_v0(const double *__restrict__ val, const int *__restrict__ col, const double *__restrict__ x,
                                             const double *__restrict__ b, double *__restrict__ y, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  struct Stage { double v[TILE]; int c[TILE]; };
  Stage *st = reinterpret_cast<Stage *>(smem);
  __shared__ __align__(8) uint64_t full[STAGES];
  const int tid = threadIdx.x;
  constexpr int ROWS = TILE / L;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int first = blockIdx.x, stride = gridDim.x;
  const int mine = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
  auto issue = [&](int j) {
    const int slot = j % STAGES;
    const size_t t = (size_t)(first + j * stride);
    mbar_expect(&full[slot], TILE * 12);
    tma1d(st[slot].v, val + t * TILE, TILE * 8, &full[slot]);
    tma1d(st[slot].c, col + t * TILE, TILE * 4, &full[slot]);
  };
  if (tid == 0) for (int j = 0; j < STAGES - 1 && j < mine; ++j) issue(j);
  __syncthreads();
  for (int it = 0; it < mine; ++it) {
    const int slot = it % STAGES;
    const size_t t = (size_t)(first + it * stride);
    Stage &S = st[slot];
    double bi[(ROWS + NT - 1) / NT];
#pragma unroll
    for (int r0 = 0; r0 < (ROWS + NT - 1) / NT; ++r0) bi[r0] = (tid + r0 * NT < ROWS) ? b[t * ROWS + tid + r0 * NT] : 0.0;   // epilogue operand prefetch
    mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
    double xr[TILE / NT];
#pragma unroll
    for (int k0 = 0; k0 < TILE / NT; ++k0) xr[k0] = x[S.c[tid + k0 * NT]];
    if (tid == 0 && it + STAGES - 1 < mine) issue(it + STAGES - 1);
#pragma unroll
    for (int k0 = 0; k0 < TILE / NT; ++k0) S.v[tid + k0 * NT] *= xr[k0];
    __syncthreads();
#pragma unroll
    for (int r0 = 0; r0 < (ROWS + NT - 1) / NT; ++r0) {
      const int r = tid + r0 * NT;
      if (r < ROWS) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < L; ++k) s += S.v[r * L + k];
        y[t * ROWS + r] = bi[r0] - s;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ V0F
// V0 plus, bit by bit, the generality of the product kernel (FEAT bitmask):
//   1 rows through a row-pointer tile (third bulk copy) and a run-time row loop
//   2 tile descriptors read from global memory (one-ahead prefetch), run-time tile extents and guards
//   4 L2 evict-first policy on the bulk copies
//   8 generic run-time-branched epilogue (the product's SpmvOp epilogue)
struct TileD { int r0, nrows, s, n; };
struct GenEpi {
  const double *D; int neumann; const double *aux; double alpha, beta; double *out; int out_mode; double *out2; double delta;
  double *acc; double gamma; const double *acc_src; int acc_mode; int wlast; double *wout; const double *fd_a, *fd_m; int fd_its;
  const double *x;
};
__device__ __forceinline__ void tma1d_pol(void *dst, const void *src, uint32_t n, uint64_t *b, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)), "l"(pol) : "memory");
}
template <int NT, int TILE, int STAGES, int L, int FEAT>
__global__ void __launch_bounds__(NT) spmv_v0f(const double *__restrict__ val, const int *__restrict__ col, const int *__restrict__ rp,
                                              const TileD *__restrict__ tiles, const double *__restrict__ x, const double *__restrict__ b,
                                              double *__restrict__ y, int ntiles, const GenEpi ge) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int ROWS = TILE / L;
  struct Stage { double v[TILE + 8]; int c[TILE + 8]; int r[ROWS + 8]; };
  Stage *st = reinterpret_cast<Stage *>(smem);
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ TileD sdesc[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint64_t pol = 0;
  if (FEAT & 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int first = blockIdx.x, stride = gridDim.x;
  const int mine = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
  TileD dnext = {0, 0, 0, 0};
  if ((FEAT & 2) && tid == 0 && mine > 0) dnext = tiles[first];
  auto issue = [&](int j) {
    const int slot = j % STAGES;
    TileD d;
    if (FEAT & 2) { d = dnext; if (j + 1 < mine) dnext = tiles[first + (j + 1) * stride]; }
    else { const int t = first + j * stride; d = TileD{t * ROWS, ROWS, t * TILE, TILE}; }
    sdesc[slot] = d;
    const int s_al = d.s & ~3, cnt = (d.n + (d.s - s_al) + 3) & ~3;
    const int r_al = d.r0 & ~3, rcnt = (d.nrows + 1 + (d.r0 - r_al) + 3) & ~3;
    mbar_expect(&full[slot], (uint32_t)(cnt * 12 + ((FEAT & 1) ? rcnt * 4 : 0)));
    if (FEAT & 4) {
      tma1d_pol(st[slot].v, val + s_al, cnt * 8, &full[slot], pol);
      tma1d_pol(st[slot].c, col + s_al, cnt * 4, &full[slot], pol);
      if (FEAT & 1) tma1d_pol(st[slot].r, rp + r_al, rcnt * 4, &full[slot], pol);
    } else {
      tma1d(st[slot].v, val + s_al, cnt * 8, &full[slot]);
      tma1d(st[slot].c, col + s_al, cnt * 4, &full[slot]);
      if (FEAT & 1) tma1d(st[slot].r, rp + r_al, rcnt * 4, &full[slot]);
    }
  };
  if (tid == 0) for (int j = 0; j < STAGES - 1 && j < mine; ++j) issue(j);
  __syncthreads();
  for (int it = 0; it < mine; ++it) {
    const int slot = it % STAGES;
    const TileD d = sdesc[slot];
    Stage &S = st[slot];
    constexpr int RPT = (ROWS + NT - 1) / NT;
    double bi[RPT], pD[RPT], pout[RPT];
#pragma unroll
    for (int r0 = 0; r0 < RPT; ++r0) {
      const int r = tid + r0 * NT;
      bi[r0] = 0.0; pD[r0] = 1.0; pout[r0] = 0.0;
      if (r < d.nrows) {
        if (FEAT & 8) {
          bi[r0] = ge.aux ? ge.aux[d.r0 + r] : 0.0;
          pD[r0] = ge.D ? ge.D[d.r0 + r] : 1.0;
          pout[r0] = ge.out_mode == 2 ? ge.out[d.r0 + r] : 0.0;
        } else {
          bi[r0] = b[d.r0 + r];
        }
      }
    }
    mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
    const int o = d.s & 3;
    double xr[TILE / NT];
#pragma unroll
    for (int k0 = 0; k0 < TILE / NT; ++k0) { const int k = tid + k0 * NT; xr[k0] = 0.0; if (k < d.n) xr[k0] = x[S.c[o + k]]; }
    if (tid == 0 && it + STAGES - 1 < mine) issue(it + STAGES - 1);
#pragma unroll
    for (int k0 = 0; k0 < TILE / NT; ++k0) { const int k = tid + k0 * NT; if (k < d.n) S.v[o + k] *= xr[k0]; }
    __syncthreads();
    const int ro = d.r0 & 3;
#pragma unroll
    for (int r0 = 0; r0 < RPT; ++r0) {
      const int r = tid + r0 * NT;
      if (r < d.nrows) {
        double s = 0.0;
        if (FEAT & 1) {
          int p = S.r[ro + r] - d.s + o;
          const int q = S.r[ro + r + 1] - d.s + o;
          for (; p < q; ++p) s += S.v[p];
        } else {
#pragma unroll
          for (int k = 0; k < L; ++k) s += S.v[o + r * L + k];
        }
        const int i = d.r0 + r;
        if (FEAT & 8) {
          if (ge.D) s = s / pD[r0];
          if (ge.neumann) s = ge.x[i] - s;
          double v = ge.beta * s;
          if (ge.aux) v = ge.alpha * bi[r0] + v;
          if (ge.wout) { double xw = 0.0; for (int q2 = 0; q2 < ge.fd_its; ++q2) xw = xw + ge.fd_m[i] * (v - ge.fd_a[i] * xw); ge.wout[i] = xw; }
          if (ge.out_mode == 1) ge.out[i] = v;
          else if (ge.out_mode == 2) ge.out[i] = pout[r0] + v;
          if (ge.out2) ge.out2[i] = ge.delta * v;
          if (ge.acc_mode) { const double t2 = ge.gamma * (ge.acc_src ? ge.acc_src[i] : v); if (ge.acc_mode == 1) ge.acc[i] = t2; else ge.acc[i] += t2; }
        } else {
          y[i] = bi[r0] - s;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ V2
// NT threads = 1 producer warp + (NT/32 - 1) consumer warps.  Ring of R slots.
template <int NT, int TILE, int R, int G, int L>
__global__ void __launch_bounds__(NT) spmv_v2(const double *__restrict__ val, const int *__restrict__ col, const double *__restrict__ x,
                                             const double *__restrict__ b, double *__restrict__ y, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int ROWS = TILE / L;
  struct Slot { double v[TILE]; double xg[TILE]; double bb[ROWS]; int c[TILE]; };
  Slot *sl = reinterpret_cast<Slot *>(smem);
  __shared__ __align__(8) uint64_t full_m[R], full_x[R], empty[R];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NCONS = NT - 32;
  if (tid == 0) {
    for (int s = 0; s < R; ++s) { mbar_init(&full_m[s], 1); mbar_init(&full_x[s], 32); mbar_init(&empty[s], NCONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int first = blockIdx.x, stride = gridDim.x;
  const int mine = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
  if (warp == 0) {
    // ---------------- producer warp: TMA front at tile j, gather front at tile j - G
    for (int j = 0; j < mine + G; ++j) {
      if (j < mine) {
        const int slot = j % R;
        if (j >= R) mbar_wait(&empty[slot], (uint32_t)(((j / R) - 1) & 1));     // consumers released the slot
        if (lane == 0) {
          const size_t t = (size_t)(first + j * stride);
          mbar_expect(&full_m[slot], TILE * 12 + ROWS * 8);
          tma1d(sl[slot].v, val + t * TILE, TILE * 8, &full_m[slot]);
          tma1d(sl[slot].c, col + t * TILE, TILE * 4, &full_m[slot]);
          tma1d(sl[slot].bb, b + t * ROWS, ROWS * 8, &full_m[slot]);
        }
      }
      const int jg = j - G;
      if (jg >= 0) {
        const int slot = jg % R;
        mbar_wait(&full_m[slot], (uint32_t)((jg / R) & 1));                       // its column indices have landed
        Slot &S = sl[slot];
#pragma unroll 8
        for (int k = lane; k < TILE; k += 32) cp_async8(&S.xg[k], x + S.c[k]);
        cp_async_arrive_noinc(&full_x[slot]);                                     // 32 arrivals, each after the lane's copies land
      }
    }
  } else {
    // ---------------- consumer warps: shared memory only
    const int ctid = tid - 32;
    constexpr int GL = (ROWS * 32 <= NCONS) ? 32 : (ROWS * 16 <= NCONS) ? 16 : (ROWS * 8 <= NCONS) ? 8 : (ROWS * 4 <= NCONS) ? 4 : (ROWS * 2 <= NCONS) ? 2 : 1;
    for (int it = 0; it < mine; ++it) {
      const int slot = it % R;
      const size_t t = (size_t)(first + it * stride);
      mbar_wait(&full_x[slot], (uint32_t)((it / R) & 1));
      Slot &S = sl[slot];
      for (int base = 0; base < ROWS * GL; base += NCONS) {
        const int u = base + ctid;
        const int row = u / GL, lg = u % GL;
        double s = 0.0;
        if (row < ROWS) {
#pragma unroll
          for (int k = lg; k < L; k += GL) s += S.v[row * L + k] * S.xg[row * L + k];
        }
#pragma unroll
        for (int w = GL >> 1; w > 0; w >>= 1) s += __shfl_down_sync(0xffffffffu, s, w, GL);
        if (row < ROWS && lg == 0) y[t * ROWS + row] = S.bb[row] - s;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
  }
}

// ------------------------------------------------------------------------------------------ reference
__global__ void spmv_ref(const double *val, const int *col, const double *x, const double *b, double *y, int n, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < L; ++k) s += val[(size_t)i * L + k] * x[col[(size_t)i * L + k]];
  y[i] = b[i] - s;
}

template <class K>
float time_kernel(K launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  CK(cudaGetLastError());
  return ms / 5;
}

static bool bisect = false;

template <int L>
void run(int n, int spread) {
  const size_t nnz = (size_t)n * L;
  std::vector<double> hv(nnz), hx(n), hb(n);
  std::vector<int> hc(nnz);
  srand(1);
  for (int i = 0; i < n; ++i) { hx[i] = (rand() % 1000) / 1000.0; hb[i] = (rand() % 1000) / 500.0; }
  for (int i = 0; i < n; ++i) {
    // sorted banded columns: i - spread*k/L (clamped), like an upwind cone
    for (int k = 0; k < L; ++k) {
      long c = (long)i - (long)spread * (L - 1 - k) / (L > 1 ? L - 1 : 1) - (rand() % 3);
      if (c < 0) c = 0;
      hc[(size_t)i * L + k] = (int)c;
      hv[(size_t)i * L + k] = 1.0 / (1 + (rand() % 7));
    }
  }
  double *val, *x, *b, *y, *yref; int *col;
  CK(cudaMalloc(&val, nnz * 8)); CK(cudaMalloc(&col, nnz * 4)); CK(cudaMalloc(&x, (size_t)n * 8)); CK(cudaMalloc(&b, (size_t)n * 8));
  CK(cudaMalloc(&y, (size_t)n * 8)); CK(cudaMalloc(&yref, (size_t)n * 8));
  CK(cudaMemcpy(val, hv.data(), nnz * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(col, hc.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x, hx.data(), (size_t)n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(b, hb.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  spmv_ref<<<(n + 255) / 256, 256>>>(val, col, x, b, yref, n, L);
  CK(cudaDeviceSynchronize());
  std::vector<double> href(n), hy(n);
  CK(cudaMemcpy(href.data(), yref, (size_t)n * 8, cudaMemcpyDeviceToHost));
  const double bytes = 12.0 * nnz + 8.0 * n /*x*/ + 8.0 * n /*b*/ + 8.0 * n /*y*/;
  auto check = [&](const char *name, float ms) {
    CK(cudaMemcpy(hy.data(), y, (size_t)n * 8, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < n; ++i) err = fmax(err, fabs(hy[i] - href[i]));
    printf("L=%-3d %-34s %8.3f ms  %7.0f GB/s   maxerr %.1e\n", L, name, ms, bytes / (ms * 1e-3) / 1e9, err);
    CK(cudaMemset(y, 0, (size_t)n * 8));
  };
  {
    float ms = time_kernel([&] { spmv_ref<<<(n + 255) / 256, 256>>>(val, col, x, b, y, n, L); });
    check("thread-per-row (scalar CSR)", ms);
  }
#define RUN_V0(NT, TILE, S, CPS) { const int ntiles = (int)(nnz / TILE); const size_t sm = (size_t)S * TILE * 12; \
    CK(cudaFuncSetAttribute(spmv_v0<NT, TILE, S, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_v0<NT, TILE, S, L>, NT, sm); \
    if (nb >= 1) { const int cps = nb < CPS ? nb : CPS; float ms = time_kernel([&] { spmv_v0<NT, TILE, S, L><<<148 * cps, NT, sm>>>(val, col, x, b, y, ntiles); }); \
    char nm[96]; snprintf(nm, 96, "V0 NT%d T%d S%d CTAs/SM %d", NT, TILE, S, cps); check(nm, ms); } }
#define RUN_V2(NT, TILE, R, G, CPS) { const int ntiles = (int)(nnz / TILE); const size_t sm = (size_t)R * (TILE * 20 + (TILE / L) * 8); \
    if (sm <= 220 * 1024) { CK(cudaFuncSetAttribute(spmv_v2<NT, TILE, R, G, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_v2<NT, TILE, R, G, L>, NT, sm); \
    if (nb >= 1) { const int cps = nb < CPS ? nb : CPS; float ms = time_kernel([&] { spmv_v2<NT, TILE, R, G, L><<<148 * cps, NT, sm>>>(val, col, x, b, y, ntiles); }); \
    char nm[96]; snprintf(nm, 96, "V2 NT%d T%d R%d G%d CTAs/SM %d", NT, TILE, R, G, cps); check(nm, ms); } } }
  RUN_V0(256, 1024, 2, 3) RUN_V0(256, 1024, 3, 4) RUN_V0(256, 2048, 3, 3)
  if (bisect) {
    constexpr int TILE = 1024, NT = 256, S = 2;
    const int ntiles = (int)(nnz / TILE);
    std::vector<int> hrp((size_t)n + 9);
    for (int i = 0; i <= n; ++i) hrp[i] = i * L;
    std::vector<TileD> ht((size_t)ntiles);
    for (int t = 0; t < ntiles; ++t) ht[t] = TileD{t * (TILE / L), TILE / L, t * TILE, TILE};
    int *rp; TileD *tiles;
    CK(cudaMalloc(&rp, hrp.size() * 4)); CK(cudaMalloc(&tiles, ht.size() * sizeof(TileD)));
    CK(cudaMemcpy(rp, hrp.data(), hrp.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(tiles, ht.data(), ht.size() * sizeof(TileD), cudaMemcpyHostToDevice));
    GenEpi ge{}; ge.aux = b; ge.alpha = 1.0; ge.beta = -1.0; ge.out = y; ge.out_mode = 1; ge.x = x;
    const size_t sm = (size_t)S * ((TILE + 8) * 12 + (TILE / L + 8) * 4 + 64);
#define RUN_F(FEAT) { CK(cudaFuncSetAttribute(spmv_v0f<NT, TILE, S, L, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_v0f<NT, TILE, S, L, FEAT>, NT, sm); const int cps = nb < 3 ? nb : 3; \
      float ms = time_kernel([&] { spmv_v0f<NT, TILE, S, L, FEAT><<<148 * cps, NT, sm>>>(val, col, rp, tiles, x, b, y, ntiles, ge); }); \
      char nm[96]; snprintf(nm, 96, "V0F feat %2d CTAs/SM %d (occ %d)", FEAT, cps, nb); check(nm, ms); }
    RUN_F(0) RUN_F(1) RUN_F(2) RUN_F(4) RUN_F(8) RUN_F(3) RUN_F(7) RUN_F(11) RUN_F(15)
    cudaFree(rp); cudaFree(tiles);
  }
  RUN_V2(256, 1024, 4, 2, 8) RUN_V2(256, 1024, 6, 3, 8) RUN_V2(256, 1024, 8, 4, 8)
  RUN_V2(256, 2048, 4, 2, 8) RUN_V2(256, 2048, 5, 2, 8)
  RUN_V2(128, 1024, 4, 2, 8) RUN_V2(128, 1024, 6, 3, 8) RUN_V2(128, 512, 6, 3, 8) RUN_V2(128, 512, 8, 4, 8)
  RUN_V2(512, 2048, 4, 2, 8) RUN_V2(512, 4096, 3, 1, 8)
  cudaFree(val); cudaFree(col); cudaFree(x); cudaFree(b); cudaFree(y); cudaFree(yref);
}

// per-launch overhead: the same V0 kernel on small problems, 200 dependent launches (stream order and a CUDA graph)
void launch_overhead() {
  constexpr int L = 8, NT = 256, TILE = 1024, S = 2;
  const size_t sm = (size_t)S * TILE * 12;
  CK(cudaFuncSetAttribute(spmv_v0<NT, TILE, S, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  printf("\nper-launch cost of V0 NT256 T1024 S2 (L=8), 200 back-to-back launches\n%-12s %-10s %-22s %-22s\n", "rows", "MB/launch", "stream: us/launch GB/s", "graph: us/launch GB/s");
  for (int lg = 14; lg <= 23; ++lg) {
    const int n = 1 << lg;
    const size_t nnz = (size_t)n * L;
    double *val, *x, *b, *y; int *col;
    CK(cudaMalloc(&val, nnz * 8)); CK(cudaMalloc(&col, nnz * 4)); CK(cudaMalloc(&x, (size_t)n * 8)); CK(cudaMalloc(&b, (size_t)n * 8)); CK(cudaMalloc(&y, (size_t)n * 8));
    CK(cudaMemset(val, 0, nnz * 8)); CK(cudaMemset(col, 0, nnz * 4)); CK(cudaMemset(x, 0, (size_t)n * 8)); CK(cudaMemset(b, 0, (size_t)n * 8));
    const int ntiles = (int)(nnz / TILE);
    const int grid = ntiles < 148 * 3 ? ntiles : 148 * 3;
    const double bytes = 12.0 * nnz + 24.0 * n;
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto body = [&] { for (int r = 0; r < 200; ++r) spmv_v0<NT, TILE, S, L><<<grid, NT, sm, st>>>(val, col, x, b, y, ntiles); };
    body(); CK(cudaStreamSynchronize(st));
    cudaEventRecord(e0, st); body(); cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1));
    float ms1; cudaEventElapsedTime(&ms1, e0, e1);
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal)); body(); CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    cudaEventRecord(e0, st); CK(cudaGraphLaunch(ge, st)); cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1));
    float ms2; cudaEventElapsedTime(&ms2, e0, e1);
    printf("%-12d %-10.1f %8.2f %10.0f   %8.2f %10.0f\n", n, bytes / 1e6, ms1 * 1e3 / 200, bytes / (ms1 * 1e-3 / 200) / 1e9, ms2 * 1e3 / 200, bytes / (ms2 * 1e-3 / 200) / 1e9);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    cudaFree(val); cudaFree(col); cudaFree(x); cudaFree(b); cudaFree(y);
  }
}

// real operator from a file written by dump_operator.py: int64 n, int64 nnz, int32 rp[n+1], int32 col[nnz], double val[nnz]
// (square, rows <= 1024 nonzeros).  Runs the scalar kernel and the V0F kernels with run-time tiles (<= 1024 nnz, <= 256 rows).
void run_real(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) { printf("cannot open %s\n", path); return; }
  long long n64, nnz64;
  if (fread(&n64, 8, 1, f) != 1 || fread(&nnz64, 8, 1, f) != 1) { printf("bad header\n"); return; }
  const int n = (int)n64; const size_t nnz = (size_t)nnz64;
  std::vector<int> hrp((size_t)n + 9, 0), hc(nnz + 8, 0); std::vector<double> hv(nnz + 8, 0.0);
  if (fread(hrp.data(), 4, (size_t)n + 1, f) != (size_t)n + 1 || fread(hc.data(), 4, nnz, f) != nnz || fread(hv.data(), 8, nnz, f) != nnz) { printf("short file\n"); return; }
  fclose(f);
  for (int i = n + 1; i < n + 9; ++i) hrp[i] = hrp[n];
  constexpr int TILE = 1024, NT = 256, S = 2, L = 4;     // L = 4 only sizes the stage for 256 rows
  std::vector<TileD> ht;
  for (int r = 0; r < n;) {
    int r0 = r; const int base = hrp[r0];
    if (hrp[r0 + 1] - base > TILE) { printf("row %d longer than a tile: not supported here\n", r0); return; }
    while (r < n && r - r0 < TILE / L && hrp[r + 1] - base <= TILE) ++r;
    ht.push_back(TileD{r0, r - r0, base, hrp[r] - base});
  }
  const int ntiles = (int)ht.size();
  std::vector<double> hx(n), hb(n);
  srand(2);
  for (int i = 0; i < n; ++i) { hx[i] = (rand() % 1000) / 1000.0; hb[i] = (rand() % 1000) / 500.0; }
  double *val, *x, *b, *y, *yref; int *col, *rp; TileD *tiles;
  CK(cudaMalloc(&val, (nnz + 8) * 8)); CK(cudaMalloc(&col, (nnz + 8) * 4)); CK(cudaMalloc(&rp, hrp.size() * 4)); CK(cudaMalloc(&tiles, ht.size() * sizeof(TileD)));
  CK(cudaMalloc(&x, (size_t)n * 8 + 64)); CK(cudaMalloc(&b, (size_t)n * 8 + 64)); CK(cudaMalloc(&y, (size_t)n * 8 + 64)); CK(cudaMalloc(&yref, (size_t)n * 8 + 64));
  CK(cudaMemcpy(val, hv.data(), (nnz + 8) * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(col, hc.data(), (nnz + 8) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(rp, hrp.data(), hrp.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(tiles, ht.data(), ht.size() * sizeof(TileD), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x, hx.data(), (size_t)n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(b, hb.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  // reference result on the host
  std::vector<double> href(n), hy(n);
  for (int i = 0; i < n; ++i) { double sacc = 0; for (int p = hrp[i]; p < hrp[i + 1]; ++p) sacc += hv[p] * hx[hc[p]]; href[i] = hb[i] - sacc; }
  const double bytes = 12.0 * nnz + 4.0 * n + 24.0 * n;
  printf("%s: n %d nnz %zu (%.2f per row), %d tiles\n", path, n, nnz, (double)nnz / n, ntiles);
  auto check = [&](const char *name, float ms) {
    CK(cudaMemcpy(hy.data(), y, (size_t)n * 8, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < n; ++i) err = fmax(err, fabs(hy[i] - href[i]));
    printf("  %-34s %8.3f ms  %7.0f GB/s   maxerr %.1e\n", name, ms, bytes / (ms * 1e-3) / 1e9, err);
    CK(cudaMemset(y, 0, (size_t)n * 8));
  };
  GenEpi ge{}; ge.aux = b; ge.alpha = 1.0; ge.beta = -1.0; ge.out = y; ge.out_mode = 1; ge.x = x;
  const size_t sm = (size_t)S * ((TILE + 8) * 12 + (TILE / L + 8) * 4 + 64);
#define RUN_R(FEAT) { CK(cudaFuncSetAttribute(spmv_v0f<NT, TILE, S, L, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_v0f<NT, TILE, S, L, FEAT>, NT, sm); const int cps = nb < 3 ? nb : 3; \
    const int grid = ntiles < 148 * cps ? ntiles : 148 * cps; \
    float ms = time_kernel([&] { spmv_v0f<NT, TILE, S, L, FEAT><<<grid, NT, sm>>>(val, col, rp, tiles, x, b, y, ntiles, ge); }); \
    char nm[96]; snprintf(nm, 96, "V0F feat %2d CTAs/SM %d", FEAT, cps); check(nm, ms); }
  RUN_R(3) RUN_R(7) RUN_R(11) RUN_R(15)
  cudaFree(val); cudaFree(col); cudaFree(rp); cudaFree(tiles); cudaFree(x); cudaFree(b); cudaFree(y); cudaFree(yref);
}

int main(int argc, char **argv) {
  if (argc > 1 && argv[1][0] == 'o') { launch_overhead(); return 0; }
  if (argc > 2 && argv[1][0] == 'r') { for (int a = 2; a < argc; ++a) run_real(argv[a]); return 0; }
  if (argc > 1 && argv[1][0] == 'b') bisect = true;
  const int n = 1 << 24;
  run<2>(n, 3000);
  run<8>(n / 2, 3000);
  run<32>(n / 8, 3000);
  return 0;
}
