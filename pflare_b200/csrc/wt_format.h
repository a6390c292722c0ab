// wt_format.h -- the warp-tile operator storage of spmv_wt_kernel (plain C++: shared by the CUDA library,
// which builds it on the host at upload, and by the CPU-side format test tests/c/wt_format_check.cpp).
//
// A tile holds consecutive rows with at most 32 * kWtKpl nonzeros and at most 32 * rq rows (rows are never
// split; an empty row is stored as one explicit 0.0 entry).  With kpl = ceil(nnz / 32), lane l owns the
// tile's nonzeros [kpl * l, kpl * (l + 1)); the blob stores them lane-interleaved so that every
// shared-memory access of the kernel is conflict-free:
//     double   val[kpl][32]     val[j][l]  = j-th nonzero of lane l      (padding: 0.0)
//     int      col[kpl][32]     same order                               (padding: column 0)
//     uint16_t ends[32]         bit j of ends[l] set  <=>  that nonzero is the last one of its row
// = kpl * 384 + 64 bytes, a multiple of 16: ONE cp.async.bulk per tile.  For the merged A_fc|W operator
// the row-ending entry is the W entry (its product is wout, not part of the row sum).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace pfb {

constexpr int kWtKpl = 8;                          // nonzeros per lane of a full tile
constexpr int kWtTileNnz = 32 * kWtKpl;            // 256
constexpr int kWtStageBytes = kWtKpl * 384 + 64;   // shared-memory bytes of one ring slot
struct alignas(16) WtDesc { unsigned off16; int r0; int nrows; int kpl; };   // blob offset in 16-byte units

struct WtHost {
  bool ok = false;          // false: some row is longer than a tile -> the operator stays on the CSR stream kernel
  int rq = 1;               // rows per lane of a tile
  std::vector<unsigned char> blob;
  std::vector<WtDesc> desc;
  int n_int = 0;            // interior tiles (no ghost column, i.e. no column >= n_local) come first
};

// rows per lane of a tile, from the operator's mean row length
inline int wt_rows_per_lane(double mean_len) {
  if (mean_len >= 7.0) return 1;
  if (mean_len >= 3.5) return 2;
  if (mean_len >= 1.75) return 4;
  return 8;
}

inline void build_wt(int m, int n_local, const int *ia, const int *ja, const double *a, WtHost *W) {
  W->ok = true;
  for (int i = 0; i < m; ++i)
    if (ia[i + 1] - ia[i] > kWtTileNnz) { W->ok = false; return; }
  W->rq = wt_rows_per_lane(m > 0 ? (double)ia[m] / m : 1.0);
  const int maxrows = 32 * W->rq;
  // pass 1: greedy tile boundaries (an empty row counts as one explicit zero entry)
  std::vector<int> tb;
  tb.push_back(0);
  for (int r = 0; r < m;) {
    int n = 0, r0 = r;
    while (r < m && r - r0 < maxrows) {
      const int len = std::max(ia[r + 1] - ia[r], 1);
      if (n + len > kWtTileNnz) break;
      n += len; ++r;
    }
    tb.push_back(r);
  }
  const size_t nt = tb.size() - 1;
  std::vector<int> tn(nt);                 // entries per tile
  std::vector<size_t> off(nt + 1, 0);      // blob offsets (bytes)
  std::vector<unsigned char> ghost(nt, 0);
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < nt; ++t) {
    int n = 0;
    bool g = false;
    for (int r = tb[t]; r < tb[t + 1]; ++r) {
      n += std::max(ia[r + 1] - ia[r], 1);
      for (int p = ia[r]; p < ia[r + 1] && !g; ++p) g = ja[p] >= n_local;
    }
    tn[t] = n; ghost[t] = g ? 1 : 0;
  }
  for (size_t t = 0; t < nt; ++t) off[t + 1] = off[t] + (size_t)((tn[t] + 31) / 32) * 384 + 64;
  W->blob.assign(off[nt] + 16, 0);
  std::vector<WtDesc> desc(nt);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t t = 0; t < nt; ++t) {
    const int kpl = (tn[t] + 31) / 32;
    unsigned char *b = W->blob.data() + off[t];
    double *val = reinterpret_cast<double *>(b);
    int *col = reinterpret_cast<int *>(b + (size_t)kpl * 256);
    unsigned short *ends = reinterpret_cast<unsigned short *>(b + (size_t)kpl * 384);
    int p = 0;   // position inside the tile: lane = p / kpl, slot = p % kpl
    for (int r = tb[t]; r < tb[t + 1]; ++r) {
      const int s0 = ia[r], s1 = ia[r + 1];
      if (s1 == s0) {          // empty row: explicit 0.0 * x[0]
        ends[p / kpl] |= (unsigned short)(1u << (p % kpl));
        ++p;
        continue;
      }
      for (int q = s0; q < s1; ++q, ++p) {
        const int lane = p / kpl, j = p % kpl;
        val[j * 32 + lane] = a[q];
        col[j * 32 + lane] = ja[q];
        if (q + 1 == s1) ends[lane] |= (unsigned short)(1u << j);
      }
    }
    desc[t] = WtDesc{(unsigned)(off[t] / 16), tb[t], tb[t + 1] - tb[t], kpl};
  }
  // interior tiles first: they can be multiplied while the ghost exchange is still in flight
  W->desc.clear();
  W->desc.reserve(nt);
  for (size_t t = 0; t < nt; ++t) if (!ghost[t]) W->desc.push_back(desc[t]);
  W->n_int = (int)W->desc.size();
  for (size_t t = 0; t < nt; ++t) if (ghost[t]) W->desc.push_back(desc[t]);
}

}  // namespace pfb
