// wt_format_check.cpp -- CPU check of the warp-tile operator storage (pflare_b200/csrc/wt_format.h) and of
// the row-sum algorithm spmv_wt_kernel runs on it (pflare_b200/csrc/kernels.cuh, stage B): the 32 lanes of a
// warp are emulated with arrays, shuffles with indexed reads.  Test infrastructure only (no GPU needed):
// it pins the layout (row-aligned lanes, sub-tiles, head masks, W entry first, explicit zero for empty
// rows, interior tiles first) and the segmented-reduction logic against a plain CSR product on random matrices.
//
//   g++ -O2 -std=c++17 -fopenmp -o wt_format_check wt_format_check.cpp && ./wt_format_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../pflare_b200/csrc/wt_format.h"

using namespace pfb;

static int ffs32(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }

// one tile, exactly the steps of the kernel's stage B (products -> lane sums -> segmented shuffle reduction -> head lanes)
static void tile_rows(const WtHost &W, const WtDesc &d, const std::vector<double> &x, bool wfirst, std::vector<double> &rowsum,
                      std::vector<double> &xw) {
  const int KP = W.kp, ns = d.geom & 0xff, gmax = d.geom >> 8, nslots = ns * KP;
  const unsigned char *b = W.blob.data() + (size_t)d.off16 * 16;
  const double *val = reinterpret_cast<const double *>(b);
  const int *col = reinterpret_cast<const int *>(b + (size_t)nslots * 256);
  const unsigned *heads = reinterpret_cast<const unsigned *>(b + (size_t)nslots * 384);
  double p[32][kWtSlots];
  for (int l = 0; l < 32; ++l)
    for (int k = 0; k < kWtSlots; ++k) p[l][k] = k < nslots ? val[k * 32 + l] * x[col[k * 32 + l]] : 0.0;
  int rowbase = 0;
  for (int s = 0; s < ns; ++s) {
    const unsigned H = heads[s];
    double acc[32], xwv[32];
    int dist[32];
    for (int l = 0; l < 32; ++l) {
      const bool head = (H >> l) & 1u;
      acc[l] = 0.0; xwv[l] = 0.0;
      for (int j = 0; j < KP; ++j) {
        if (j == 0 && wfirst && head) xwv[l] = p[l][s * KP];
        else acc[l] += p[l][s * KP + j];
      }
      const unsigned above = l < 31 ? (H >> (l + 1)) : 0u;
      dist[l] = above ? ffs32(above) - 1 : 31 - l;
    }
    for (int o = 1; o < 32; o <<= 1) {
      if (o < gmax) {
        double t[32];
        for (int l = 0; l < 32; ++l) t[l] = acc[l + o < 32 ? l + o : l];   // shfl_down: own value when out of range
        for (int l = 0; l < 32; ++l)
          if (o <= dist[l]) acc[l] += t[l];
      }
    }
    for (int l = 0; l < 32; ++l)
      if ((H >> l) & 1u) {
        const int r = rowbase + __builtin_popcount(H & ((1u << l) - 1u));
        rowsum[d.r0 + r] = acc[l];
        xw[d.r0 + r] = xwv[l];
      }
    rowbase += __builtin_popcount(H);
  }
}

static int run_case(int m, int n, double mean_len, double p_empty, bool wlast, int n_ghost, unsigned seed, int len_cap = kWtMaxRow) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<int> ia(1, 0), ja;
  std::vector<double> a;
  for (int i = 0; i < m; ++i) {
    int len = 0;
    if (U(rng) >= p_empty) {
      len = 1 + (int)(-std::log(1.0 - U(rng) * 0.999) * (mean_len - 1.0));
      if (len > len_cap) len = len_cap;
    }
    if (wlast && len == 0) len = 1;   // the merged A_fc|W operator always has the W entry
    for (int k = 0; k < len; ++k) { ja.push_back((int)(U(rng) * (n + n_ghost)) % (n + n_ghost)); a.push_back(U(rng) - 0.5); }
    ia.push_back((int)ja.size());
  }
  std::vector<double> x((size_t)n + n_ghost);
  for (double &v : x) v = U(rng);
  WtHost W;
  build_wt(m, n, ia.data(), ja.data(), a.data(), wlast, &W);
  if (!W.ok) { printf("FAIL: builder refused a matrix without long rows\n"); return 1; }
  std::vector<double> rs((size_t)m, 1e300), xw((size_t)m, 1e300);
  std::vector<int> covered((size_t)m, 0);
  size_t bytes = 0;
  for (size_t t = 0; t < W.desc.size(); ++t) {
    const WtDesc &d = W.desc[t];
    const int ns = d.geom & 0xff;
    if (ns < 1 || ns * W.kp > kWtSlots || d.nrows < 1 || d.nrows > 32 * ns) { printf("FAIL: tile %zu geometry ns %d kp %d rows %d\n", t, ns, W.kp, d.nrows); return 1; }
    for (int r = 0; r < d.nrows; ++r) covered[d.r0 + r]++;
    // interior tiles first
    bool ghost = false;
    for (int r = d.r0; r < d.r0 + d.nrows; ++r)
      for (int q = ia[r]; q < ia[r + 1]; ++q) ghost |= ja[q] >= n;
    if (ghost != ((int)t >= W.n_int)) { printf("FAIL: tile %zu interior/boundary order\n", t); return 1; }
    bytes += (size_t)ns * W.kp * 384 + 32;
    tile_rows(W, d, x, wlast, rs, xw);
  }
  if (bytes + 16 != W.blob.size()) { printf("FAIL: blob size %zu vs tiles %zu\n", W.blob.size(), bytes); return 1; }
  for (int i : W.long_rows) {   // rows left to the CSR stream kernel
    if (ia[i + 1] - ia[i] <= kWtMaxRow) { printf("FAIL: row %d listed as long\n", i); return 1; }
    covered[i]++;
    double sl = 0.0;
    const int q1 = wlast ? ia[i + 1] - 1 : ia[i + 1];
    for (int q = ia[i]; q < q1; ++q) sl += a[q] * x[ja[q]];
    rs[i] = sl;
    if (wlast) xw[i] = a[q1] * x[ja[q1]];
  }
  double maxerr = 0.0;
  for (int i = 0; i < m; ++i) {
    if (covered[i] != 1) { printf("FAIL: row %d covered %d times\n", i, covered[i]); return 1; }
    double s = 0.0, w = 0.0;
    const int q1 = wlast ? ia[i + 1] - 1 : ia[i + 1];
    for (int q = ia[i]; q < q1; ++q) s += a[q] * x[ja[q]];
    if (wlast) w = a[q1] * x[ja[q1]];
    maxerr = std::fmax(maxerr, std::fabs(s - rs[i]));
    if (wlast && w != xw[i]) { printf("FAIL: row %d W product %g vs %g\n", i, xw[i], w); return 1; }
  }
  if (maxerr > 1e-13) { printf("FAIL: m %d mean %.1f wlast %d: max row-sum error %.3e\n", m, mean_len, (int)wlast, maxerr); return 1; }
  return 0;
}

int main() {
  int bad = 0, n = 0;
  const double means[] = {1.0, 1.4, 2.0, 3.0, 5.0, 8.0, 8.5, 13.0, 33.0, 70.0, 200.0};
  for (double ml : means)
    for (int wl = 0; wl < 2; ++wl)
      for (int gh = 0; gh < 2; ++gh) {
        bad += run_case(3001, 2500, ml, wl ? 0.0 : 0.07, wl != 0, gh ? 300 : 0, (unsigned)(1000 * ml) + wl * 7 + gh);
        ++n;
      }
  bad += run_case(1, 1, 1.0, 0.0, false, 0, 5); ++n;          // one row
  bad += run_case(40, 10, 1.0, 1.0, false, 0, 6); ++n;        // every row empty
  bad += run_case(257, 64, 256.0, 0.0, false, 0, 7); ++n;     // rows of (up to) a whole tile
  bad += run_case(900, 700, 150.0, 0.02, false, 40, 8, 700); ++n;   // some rows longer than a tile: skipped by the tiles, listed
  bad += run_case(900, 700, 120.0, 0.0, true, 0, 9, 700); ++n;
  // a row longer than a tile is left out of the tiles and listed (it then runs on the CSR stream kernel)
  {
    const int L = kWtMaxRow + 5;
    std::vector<int> ia = {0, 2, 2 + L, 2 + L + 3}, ja((size_t)L + 5, 0);
    std::vector<double> a((size_t)L + 5, 1.0);
    WtHost W;
    build_wt(3, 1, ia.data(), ja.data(), a.data(), false, &W);
    int rows = 0;
    for (const WtDesc &d : W.desc) rows += d.nrows;
    if (!W.ok || W.long_rows.size() != 1 || W.long_rows[0] != 1 || rows != 2 || W.desc.size() != 2) { printf("FAIL: long row handling\n"); ++bad; }
    ++n;
  }
  printf("%s: %d cases, %d failed\n", bad ? "WT_FORMAT_FAIL" : "WT_FORMAT_OK", n, bad);
  return bad ? 1 : 0;
}
