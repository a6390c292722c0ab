// wt_format.h -- the warp-tile operator storage of spmv_wt_kernel (plain C++: shared by the CUDA library,
// which builds it on the host at upload, and by the CPU-side format test tests/c/wt_format_check.cpp).
//
// Row-aligned lanes.  Every operator picks kp in {1, 2, 4, 8} nonzero slots per lane.  A SUB-TILE is
// 32 lanes x kp slots: a row of `len` nonzeros (an empty row is stored as one explicit 0.0 entry) takes
// g = ceil(len / kp) consecutive lanes, its i-th nonzero sits in lane (first + i % g), slot i / g --
// consecutive nonzeros of a row are read by adjacent lanes and neighbouring rows by neighbouring lane
// groups, so one gather instruction touches few cache lines (round-2 ncu finding: the L1 data pipe, not
// DRAM, bounds the kernel when lanes own arbitrary runs of the CSR stream).  Rows are packed into a
// sub-tile until its 32 lanes are used up; unused slots hold 0.0 * x[0].  A TILE is up to 8 / kp
// consecutive sub-tiles (<= 8 slots per lane, <= 32 * 8 / kp rows) stored as ONE blob:
//     double   val[nslots][32]     val[s * kp + j][l] = slot j of lane l in sub-tile s
//     int      col[nslots][32]     same order
//     uint32_t heads[8]            bit l of heads[s] set  <=>  lane l is the first lane of a row
// = nslots * 384 + 32 bytes, a multiple of 16: one cp.async.bulk per tile, every shared-memory access of
// the kernel conflict-free.  For the merged A_fc|W operator the W entry of a row is stored FIRST (slot 0
// of the row's head lane): its product is wout, not part of the row sum.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace pfb {

constexpr int kWtSlots = 8;                           // slots per lane of a full tile
constexpr int kWtMaxRow = 32 * kWtSlots;              // longest row a tile can hold
constexpr int kWtStageBytes = kWtSlots * 384 + 32;    // shared-memory bytes of one ring slot
// blob offset in 16-byte units; first row; rows; geom = #sub-tiles | (max lanes per row << 8)
struct alignas(16) WtDesc { unsigned off16; int r0; int nrows; int geom; };

struct WtHost {
  bool ok = false;
  std::vector<int> long_rows;   // rows longer than a tile (kWtMaxRow): left out of the tiles, handled by the CSR stream kernel
  int kp = 8;               // slots per lane of a sub-tile (1, 2, 4 or 8); a tile holds <= 8 / kp sub-tiles
  std::vector<unsigned char> blob;
  std::vector<WtDesc> desc;
  int n_int = 0;            // interior tiles (no ghost column, i.e. no column >= n_local) come first
  int64_t slots = 0;        // stored slots (nonzeros + padding)
};

// sub-tiles for a given kp: rows [sb[2t], sb[2t+1]) share one 32-lane sub-tile; rows longer than kWtMaxRow are skipped
// (they end the running sub-tile; brk[t] = 1 marks a sub-tile that follows such a gap)
// (false: some row that fits a tile needs more than 32 lanes with this kp)
inline bool wt_pack(int m, const int *ia, int kp, std::vector<int> *sb, std::vector<unsigned char> *brk = nullptr) {
  sb->clear();
  if (brk) brk->clear();
  bool gap = false;
  for (int r = 0; r < m;) {
    if (ia[r + 1] - ia[r] > kWtMaxRow) { ++r; gap = true; continue; }
    const int r0 = r;
    int lanes = 0;
    while (r < m) {
      const int len = ia[r + 1] - ia[r];
      if (len > kWtMaxRow) break;
      const int g = (std::max(len, 1) + kp - 1) / kp;
      if (g > 32) return false;
      if (lanes + g > 32) break;
      lanes += g; ++r;
    }
    sb->push_back(r0); sb->push_back(r);
    if (brk) brk->push_back(gap ? 1 : 0);
    gap = false;
  }
  return true;
}

// wfirst: every row's LAST stored entry (the W entry of the merged A_fc|W operator) is moved to the front
inline void build_wt(int m, int n_local, const int *ia, const int *ja, const double *a, bool wfirst, WtHost *W) {
  W->ok = true;
  W->long_rows.clear();
  for (int i = 0; i < m; ++i)
    if (ia[i + 1] - ia[i] > kWtMaxRow) W->long_rows.push_back(i);
  // slots per lane: the largest kp whose padded size is within 6 % of the smallest
  std::vector<int> sb;
  std::vector<unsigned char> brk;
  int64_t best = -1, cost[4];
  const int kps[4] = {8, 4, 2, 1};
  for (int k = 0; k < 4; ++k) {
    cost[k] = wt_pack(m, ia, kps[k], &sb) ? (int64_t)(sb.size() / 2) * 32 * kps[k] : -1;
    if (cost[k] >= 0 && (best < 0 || cost[k] < best)) best = cost[k];
  }
  int kp = 8;
  for (int k = 0; k < 4; ++k)
    if (cost[k] >= 0 && (double)cost[k] <= 1.06 * (double)best) { kp = kps[k]; break; }
  W->kp = kp;
  wt_pack(m, ia, kp, &sb, &brk);
  const size_t nsub = sb.size() / 2;
  const int per = kWtSlots / kp;                         // sub-tiles per tile
  W->slots = (int64_t)nsub * 32 * kp;
  // tiles = runs of <= per consecutive sub-tiles with contiguous rows (a skipped long row ends the tile)
  std::vector<size_t> tfirst;                            // first sub-tile of every tile (+ sentinel)
  for (size_t q = 0; q < nsub; ++q)
    if (tfirst.empty() || brk[q] || q - tfirst.back() == (size_t)per) tfirst.push_back(q);
  const size_t nt = tfirst.size();
  tfirst.push_back(nsub);
  std::vector<size_t> off(nt + 1, 0);
  for (size_t t = 0; t < nt; ++t) off[t + 1] = off[t] + (tfirst[t + 1] - tfirst[t]) * kp * 384 + 32;
  W->blob.assign(off[nt] + 16, 0);
  std::vector<WtDesc> desc(nt);
  std::vector<unsigned char> ghost(nt, 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t t = 0; t < nt; ++t) {
    const int ns = (int)(tfirst[t + 1] - tfirst[t]);
    const int nslots = ns * kp;
    unsigned char *b = W->blob.data() + off[t];
    double *val = reinterpret_cast<double *>(b);
    int *col = reinterpret_cast<int *>(b + (size_t)nslots * 256);
    unsigned *heads = reinterpret_cast<unsigned *>(b + (size_t)nslots * 384);
    int gmax = 1;
    bool gh = false;
    for (int s = 0; s < ns; ++s) {
      const size_t st = tfirst[t] + s;
      int lane = 0;
      for (int r = sb[2 * st]; r < sb[2 * st + 1]; ++r) {
        const int s0 = ia[r], len = ia[r + 1] - s0;
        const int g = (std::max(len, 1) + kp - 1) / kp;
        gmax = std::max(gmax, g);
        heads[s] |= 1u << lane;
        for (int i = 0; i < len; ++i) {
          // i-th entry in storage order; with wfirst the row's last CSR entry comes first
          const int q = wfirst ? (i == 0 ? s0 + len - 1 : s0 + i - 1) : s0 + i;
          const int l = lane + i % g, j = s * kp + i / g;
          val[j * 32 + l] = a[q];
          col[j * 32 + l] = ja[q];
          gh = gh || ja[q] >= n_local;
        }
        lane += g;
      }
    }
    ghost[t] = gh ? 1 : 0;
    desc[t] = WtDesc{(unsigned)(off[t] / 16), sb[2 * tfirst[t]], sb[2 * (tfirst[t] + ns) - 1] - sb[2 * tfirst[t]], ns | (gmax << 8)};
  }
  // interior tiles first: they can be multiplied while the ghost exchange is still in flight
  W->desc.clear();
  W->desc.reserve(nt);
  for (size_t t = 0; t < nt; ++t) if (!ghost[t]) W->desc.push_back(desc[t]);
  W->n_int = (int)W->desc.size();
  for (size_t t = 0; t < nt; ++t) if (ghost[t]) W->desc.push_back(desc[t]);
}


// ------------------------------------------------------------------------------------------------------------------
// CHUNK format (operators with short rows, mean length < ~4.5): lanes own consecutive runs of the CSR stream.
// A tile holds consecutive rows with at most 32 * kWcKpl nonzeros and at most 32 * rq rows (rows are never split; an
// empty row is stored as one explicit 0.0 entry).  With kpl = ceil(nnz / 32), lane l owns the tile's nonzeros
// [kpl * l, kpl * (l + 1)), stored lane-interleaved:
//     double   val[kpl][32]     val[j][l]  = j-th nonzero of lane l      (padding: 0.0)
//     int      col[kpl][32]     same order                               (padding: column 0)
//     uint16_t ends[32]         bit j of ends[l] set  <=>  that nonzero is the last one of its row
// = kpl * 384 + 64 bytes.  No padding inside rows (row-aligned lanes would pad 1-3 nonzero rows by 30 % and more) and,
// with many short equal rows, neighbouring lanes still read neighbouring columns.  For the merged A_fc|W operator the
// row-ending entry is the W entry.  The tile descriptor is a WtDesc whose `geom` field holds kpl.
constexpr int kWcKpl = 8;                          // nonzeros per lane of a full tile
constexpr int kWcTileNnz = 32 * kWcKpl;            // 256
constexpr int kWcStageBytes = kWcKpl * 384 + 64;   // shared-memory bytes of one ring slot

struct WcHost {
  bool ok = false;
  std::vector<int> long_rows;   // rows longer than a tile: left out of the tiles, handled by the CSR stream kernel
  int rq = 1;               // rows per lane of a tile
  std::vector<unsigned char> blob;
  std::vector<WtDesc> desc;
  int n_int = 0;            // interior tiles (no ghost column, i.e. no column >= n_local) come first
};

// rows per lane of a tile, from the operator's mean row length
inline int wc_rows_per_lane(double mean_len) {
  if (mean_len >= 7.0) return 1;
  if (mean_len >= 3.5) return 2;
  if (mean_len >= 1.75) return 4;
  return 8;
}

inline void build_wc(int m, int n_local, const int *ia, const int *ja, const double *a, WcHost *W) {
  W->ok = true;
  W->long_rows.clear();
  W->rq = wc_rows_per_lane(m > 0 ? (double)ia[m] / m : 1.0);
  const int maxrows = 32 * W->rq;
  // pass 1: greedy tiles [tb[2t], tb[2t+1]) (an empty row counts as one explicit zero entry; a row longer than a tile is skipped)
  std::vector<int> tb;
  for (int r = 0; r < m;) {
    if (ia[r + 1] - ia[r] > kWcTileNnz) { W->long_rows.push_back(r); ++r; continue; }
    int n = 0, r0 = r;
    while (r < m && r - r0 < maxrows) {
      if (ia[r + 1] - ia[r] > kWcTileNnz) break;
      const int len = std::max(ia[r + 1] - ia[r], 1);
      if (n + len > kWcTileNnz) break;
      n += len; ++r;
    }
    tb.push_back(r0); tb.push_back(r);
  }
  const size_t nt = tb.size() / 2;
  std::vector<int> tn(nt);                 // entries per tile
  std::vector<size_t> off(nt + 1, 0);      // blob offsets (bytes)
  std::vector<unsigned char> ghost(nt, 0);
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < nt; ++t) {
    int n = 0;
    bool g = false;
    for (int r = tb[2 * t]; r < tb[2 * t + 1]; ++r) {
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
    for (int r = tb[2 * t]; r < tb[2 * t + 1]; ++r) {
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
    desc[t] = WtDesc{(unsigned)(off[t] / 16), tb[2 * t], tb[2 * t + 1] - tb[2 * t], kpl};
  }
  // interior tiles first: they can be multiplied while the ghost exchange is still in flight
  W->desc.clear();
  W->desc.reserve(nt);
  for (size_t t = 0; t < nt; ++t) if (!ghost[t]) W->desc.push_back(desc[t]);
  W->n_int = (int)W->desc.size();
  for (size_t t = 0; t < nt; ++t) if (ghost[t]) W->desc.push_back(desc[t]);
}


}  // namespace pfb
