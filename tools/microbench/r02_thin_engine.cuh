// Round-2 experiment kept for the record (not compiled into the product): the thin-warp engine, 2.76 ms per 4096^2 V-cycle
// against 2.56 ms for spmv_sv_kernel at 4 CTAs per SM.  Needs kernels.cuh's SpmvOp / EpiT.
// ------------------------------------------------------------------------------------------
// Thin-warp engine on the same storage: the direct engine cut down to ~40 registers so that 48 warps per
// SM are resident.  A warp walks its tile sub-tile by sub-tile, at most 4 slots per lane in flight
// (coalesced streaming loads of columns and values, gathers, fused multiply-adds), a segmented shuffle
// reduction per sub-tile, and the head lane of every row loads its epilogue operands and finishes the row.
// Nothing is software-pipelined: every latency (matrix stream, gathers, epilogue operands) is hidden by
// the other 47 warps of the SM -- the measured behaviour of all three engines is "throughput
// proportional to resident warps", so this one maximises them.
template <int EPI, int KP, bool GHOST>
__global__ void __launch_bounds__(256, 6) spmv_thin_kernel(const SpmvOp op) {
  typedef EpiT<EPI> E;
  constexpr bool XW = E::kXw;
  constexpr int CH = KP < 4 ? KP : 4;   // slots per lane in flight
  constexpr int NCH = KP / CH;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const WtDesc *__restrict__ wdesc = op.wdesc;
  const double *__restrict__ xv = op.x;
  const double *__restrict__ xg = op.xg;
  const int nloc = op.nloc;
  const bool wf = (EPI == EPI_AFCW || EPI == EPI_AFCW_LOCAL) || (EPI == EPI_GENERIC && op.wlast);
  pdl_launch_dependents();
  bool waited = false;
  for (int t = gw; t < op.nwt; t += nwarps) {
    const WtDesc d = wdesc[t];
    const int ns = d.geom & 0xff, gmax = d.geom >> 8;
    const int nslots = ns * KP;
    const unsigned char *b = op.blob + (size_t)d.off16 * 16;
    const double *val_g = reinterpret_cast<const double *>(b) + lane;
    const int *col_g = reinterpret_cast<const int *>(b + nslots * 256) + lane;
    const unsigned *heads = reinterpret_cast<const unsigned *>(b + nslots * 384);
    int rowbase = d.r0;
    for (int s = 0; s < ns; ++s) {
      const unsigned H = __ldg(heads + s);
      const bool head = (H >> lane) & 1u;
      double acc = 0.0, xw = 0.0;
#pragma unroll
      for (int h = 0; h < NCH; ++h) {
        int c[CH];
        double v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) c[j] = __ldcs(col_g + (s * KP + h * CH + j) * 32);
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __ldcs(val_g + (s * KP + h * CH + j) * 32);
        if (!waited) {   // from here on the vectors written by the previous kernels are read
          pdl_wait();
          waited = true;
          if (GHOST) { if (lane == 0) ghost_wait(op.gw_ready, op.gw_epoch, op.gw_srcmask); __syncwarp(); }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          double x;
          if (GHOST) x = (c[j] >= nloc) ? __ldcg(xg + (c[j] - nloc)) : xv[c[j]];
          else x = xv[c[j]];
          if (XW && h == 0 && j == 0 && wf && head) xw = v[j] * x;   // merged A_fc|W: the row's first entry is the W entry
          else acc += v[j] * x;
        }
      }
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
      if (head) {
        const int r = rowbase + __popc(H & ((1u << lane) - 1u));
        E::finish(op, r, acc, xw, E::prefetch(op, r));
      }
      rowbase += __popc(H);
    }
  }
  if (!waited) pdl_wait();
}

