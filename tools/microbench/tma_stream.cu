// tma_stream.cu -- microbenchmark: how fast can a persistent CTA ring of 1-D TMA bulk copies stream HBM on B200?
// (tile bytes x stages x CTAs/SM), next to a plain 128-bit-load streaming kernel.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma1d(void *dst, const void *src, uint32_t n, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}

template <int STAGES>
__global__ void tma_ring(const char *src, size_t ntiles, int tile_bytes, double *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const size_t first = blockIdx.x, stride = gridDim.x;
  const size_t mine = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
  auto issue = [&](size_t j) {
    const int slot = (int)(j % STAGES);
    mbar_expect(&full[slot], (uint32_t)tile_bytes);
    tma1d(smem + (size_t)slot * tile_bytes, src + (first + j * stride) * (size_t)tile_bytes, (uint32_t)tile_bytes, &full[slot]);
  };
  if (tid == 0) for (size_t j = 0; j < (size_t)(STAGES - 1) && j < mine; ++j) issue(j);
  __syncthreads();
  double acc = 0.0;
  for (size_t it = 0; it < mine; ++it) {
    const int slot = (int)(it % STAGES);
    if (tid == 0 && it + STAGES - 1 < mine) issue(it + STAGES - 1);
    mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
    const double *v = reinterpret_cast<const double *>(smem + (size_t)slot * tile_bytes);
    for (int k = tid; k < tile_bytes / 8; k += blockDim.x) acc += v[k];     // consume the whole tile from smem
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (acc == 1.234e300) sink[0] = acc;
}

__global__ void plain_stream(const double2 *src, size_t n2, double *sink) {
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = __ldcs(src + i);
    acc += v.x + v.y;
  }
  if (acc == 1.234e300) sink[0] = acc;
}

template <int STAGES>
float run_ring(const char *src, size_t bytes, int tile_bytes, int ctas_per_sm, int threads, double *sink) {
  const size_t ntiles = bytes / tile_bytes;
  const size_t smem = (size_t)STAGES * tile_bytes;
  if (smem > 220 * 1024) return -1.f;
  cudaFuncSetAttribute(tma_ring<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tma_ring<STAGES>, threads, smem);
  if (nb < ctas_per_sm) return -1.f;
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) tma_ring<STAGES><<<grid, threads, smem>>>(src, ntiles, tile_bytes, sink);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) tma_ring<STAGES><<<grid, threads, smem>>>(src, ntiles, tile_bytes, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  if (cudaGetLastError() != cudaSuccess) return -2.f;
  return (float)(5.0 * ntiles * tile_bytes / (ms * 1e-3) / 1e9);
}

int main() {
  const size_t bytes = (size_t)4 << 30;
  char *src; double *sink;
  cudaMalloc(&src, bytes); cudaMalloc(&sink, 8); cudaMemset(src, 1, bytes);
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int g : {148 * 8, 148 * 16, 148 * 32}) {
      plain_stream<<<g, 256>>>((const double2 *)src, bytes / 16, sink);
      cudaEventRecord(e0);
      for (int r = 0; r < 5; ++r) plain_stream<<<g, 256>>>((const double2 *)src, bytes / 16, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("plain ld.128 read-only  grid %5d x256 : %7.0f GB/s\n", g, 5.0 * bytes / (ms * 1e-3) / 1e9);
    }
  }
  printf("%-10s %-7s %-9s %-8s %s\n", "tile", "stages", "CTAs/SM", "threads", "GB/s");
  for (int tile : {4096, 8192, 16384, 32768, 65536})
    for (int cps : {1, 2, 3, 4, 6, 8})
      for (int threads : {128, 256}) {
        float a = run_ring<2>(src, bytes, tile, cps, threads, sink);
        float b = run_ring<3>(src, bytes, tile, cps, threads, sink);
        float c = run_ring<4>(src, bytes, tile, cps, threads, sink);
        float d = run_ring<6>(src, bytes, tile, cps, threads, sink);
        printf("%-10d S=2,3,4,6 CTAs/SM %d thr %3d : %7.0f %7.0f %7.0f %7.0f\n", tile, cps, threads, a, b, c, d);
      }
  return 0;
}
