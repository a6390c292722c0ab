// kernels.cuh -- sm_100a device code of the AIRG V-cycle apply.
//
// One SpMV "mega-op" covers every MatMult + Vec-AXPY chain of the reference's apply path
// (SURVEY.md section 2c): the CSR product of a row is reduced once and the row epilogue applies
// the fused vector updates.  The path is HBM-bound fp64/int32 work (no tensor cores by design):
//   * matrix values / column indices are streamed with coalesced, cache-streaming loads,
//     a fixed-size nnz tile per CTA ("row blocks" computed once at upload) so every CTA moves
//     the same number of bytes regardless of row lengths (segment-balanced CSR);
//   * products are staged in shared memory and each row is reduced by one thread in stored
//     column order (the reference's MatMult_SeqAIJ summation order);
//   * x gathers go through L1/L2 (the nested CF layout keeps them near-sequential);
//   * rows longer than a tile fall back to a whole-CTA reduction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfb {

constexpr int kThreads = 256;       // CTA size of the streaming kernels
constexpr int kTile = 2048;         // nnz per CTA tile (16 KB of fp64 products in smem)
constexpr int kMaxRowsPerBlk = 1024;
constexpr int kTailThreads = 1024;  // CTA size of the single-CTA tail kernel

struct SpmvOp {
  // CSR block + its row-block partition
  const int *rp, *col;
  const double *val;
  int m, nblk;
  const int *blk;
  // gather sources: column c < nloc reads x[c], otherwise xg[c - nloc] (ghost buffer)
  const double *x, *xg;
  int nloc;
  const int *rowmap;           // optional: CSR row r updates vector entry rowmap[r]
  const unsigned char *skip;   // optional: rows with skip[i] != 0 are left to the boundary pass
  // s = sum_j a_ij x_j ; optional s /= D[i] ; optional s = x[i] - s (Neumann I - D^-1 A)
  const double *D;
  int neumann;
  // v = alpha * aux[i] + beta * s
  const double *aux;
  double alpha, beta;
  double *out; int out_mode;   // 0 none, 1 out[i] = v, 2 out[i] += v
  double *out2; double delta;  // out2[i] = delta * v
  double *acc; double gamma; const double *acc_src; int acc_mode;  // acc[i] (=|+=) gamma * (acc_src ? acc_src[i] : v)
  // one-point prolongation companion: wout[i] = wcol[i] >= 0 ? wval[i] * x[wcol[i]] : 0
  const int *wcol; const double *wval; double *wout;
  // fully local F smooth (diagonal A_ff and diagonal inverse): x = wout value; repeat fd_its:
  // x += fd_m[i] * (v - fd_a[i] * x); wout[i] = x
  const double *fd_a, *fd_m; int fd_its;
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

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

__device__ __forceinline__ double gather_x(const SpmvOp &op, int c) {
  if (op.xg != nullptr && c >= op.nloc) return op.xg[c - op.nloc];
  return op.x[c];
}

__device__ __forceinline__ void row_epilogue(const SpmvOp &op, int r, double s) {
  const int i = op.rowmap ? op.rowmap[r] : r;
  if (op.skip && op.skip[i]) return;
  if (op.D) s = s / op.D[i];
  if (op.neumann) s = op.x[i] - s;
  double v = op.beta * s;
  if (op.aux) v = op.alpha * op.aux[i] + v;
  if (op.wout) {
    double xw = 0.0;
    const int wc = op.wcol[i];
    if (wc >= 0) xw = op.wval[i] * gather_x(op, wc);
    if (op.fd_its > 0) {
      const double a = op.fd_a[i], mm = op.fd_m[i];
      for (int it = 0; it < op.fd_its; ++it) xw = xw + mm * (v - a * xw);
    }
    op.wout[i] = xw;
  }
  if (op.out_mode == 1) op.out[i] = v;
  else if (op.out_mode == 2) op.out[i] += v;
  if (op.out2) op.out2[i] = op.delta * v;
  if (op.acc_mode) {
    const double t = op.gamma * (op.acc_src ? op.acc_src[i] : v);
    if (op.acc_mode == 1) op.acc[i] = t;
    else op.acc[i] += t;
  }
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
      const int q = __ldg(op.rp + r + 1) - s;
      double sum = 0.0;
      for (; p < q; ++p) sum += prod[p];
      row_epilogue(op, r, sum);
    }
    __syncthreads();
  } else {
    // a single long row: whole-CTA reduction
    double part = 0.0;
    for (int k = s + tid; k < e; k += NT) part += ld_stream(op.val + k) * gather_x(op, ld_stream(op.col + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int w = 0; w < NT / 32; ++w) sum += red[w];
      row_epilogue(op, r0, sum);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) spmv_stream_kernel(const SpmvOp op) {
  __shared__ double prod[kTile];
  __shared__ double red[kThreads / 32];
  for (int b = blockIdx.x; b < op.nblk; b += gridDim.x) process_block<kThreads>(op, b, prod, red);
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
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += gridDim.x * blockDim.x) ew_apply(e, i);
}

// Single-CTA "tail": runs a whole list of ops (the small coarse levels: restrictions, coarse
// solve, prolongation + smoothing) back to back with CTA barriers instead of kernel launches.
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const DevOp *ops, int nops) {
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
