// wc_format_check.cpp -- CPU check of the CHUNK operator storage (pflare_b200/csrc/wt_format.h) and of
// the row-sum algorithm spmv_wc_kernel runs on it (pflare_b200/csrc/kernels.cuh, stage B): the 32 lanes of a
// warp are emulated with arrays, shuffles with indexed reads.  Test infrastructure only (no GPU needed):
// it pins the layout (lane-interleaved values / columns, per-lane row-end masks, explicit zero for empty
// rows, interior tiles first) and the segmented-scan logic against a plain CSR product on random matrices.
//
//   g++ -O2 -std=c++17 -fopenmp -o wc_format_check wc_format_check.cpp && ./wc_format_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../pflare_b200/csrc/wt_format.h"

using namespace pfb;

static int clz32(unsigned v) { return v ? __builtin_clz(v) : 32; }

// one tile, exactly the steps of the kernel (products -> tails -> segmented scan -> row walk)
static void tile_rows(const WcHost &W, const WtDesc &d, const std::vector<double> &x, bool wlast, std::vector<double> &rowsum,
                      std::vector<double> &xw) {
  const unsigned char *b = W.blob.data() + (size_t)d.off16 * 16;
  const double *val = reinterpret_cast<const double *>(b);
  const int *col = reinterpret_cast<const int *>(b + (size_t)d.geom * 256);
  const unsigned short *ends = reinterpret_cast<const unsigned short *>(b + (size_t)d.geom * 384);
  double p[32][kWcKpl];
  unsigned e[32];
  double tail[32], sc[32], acc[32];
  int rb[32], cnt[32];
  for (int l = 0; l < 32; ++l) {
    for (int k = 0; k < kWcKpl; ++k) p[l][k] = k < d.geom ? val[k * 32 + l] * x[col[k * 32 + l]] : 0.0;
    e[l] = ends[l];
    const int last = 31 - clz32(e[l]);
    tail[l] = 0.0;
    for (int k = 0; k < kWcKpl; ++k)
      if (k > last) tail[l] += p[l][k];
  }
  unsigned has = 0;
  for (int l = 0; l < 32; ++l)
    if (e[l]) has |= 1u << l;
  int dist[32];
  for (int l = 0; l < 32; ++l) {
    const unsigned upto = has & (0xffffffffu >> (31 - l));
    dist[l] = l - (upto ? 31 - clz32(upto) : 0);
    sc[l] = tail[l];
    cnt[l] = __builtin_popcount(e[l]);
    rb[l] = cnt[l];
  }
  for (int o = 1; o < 32; o <<= 1) {
    double t[32];
    int ti[32];
    for (int l = 0; l < 32; ++l) { t[l] = sc[l >= o ? l - o : l]; ti[l] = rb[l >= o ? l - o : l]; }   // shfl_up: own value when out of range
    for (int l = 0; l < 32; ++l) {
      if (o <= dist[l]) sc[l] += t[l];
      if (l >= o) rb[l] += ti[l];
    }
  }
  for (int l = 0; l < 32; ++l) { acc[l] = l == 0 ? 0.0 : sc[l - 1]; rb[l] -= cnt[l]; }
  for (int l = 0; l < 32; ++l)
    for (int k = 0; k < kWcKpl; ++k) {
      const bool end = (e[l] >> k) & 1u;
      if (wlast) {
        if (end) { rowsum[d.r0 + rb[l]] = acc[l]; xw[d.r0 + rb[l]] = p[l][k]; ++rb[l]; acc[l] = 0.0; }
        else acc[l] += p[l][k];
      } else {
        acc[l] += p[l][k];
        if (end) { rowsum[d.r0 + rb[l]] = acc[l]; ++rb[l]; acc[l] = 0.0; }
      }
    }
}

static int run_case(int m, int n, double mean_len, double p_empty, bool wlast, int n_ghost, unsigned seed, int len_cap = kWcTileNnz) {
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
  WcHost W;
  build_wc(m, n, ia.data(), ja.data(), a.data(), &W);
  if (!W.ok) { printf("FAIL: builder refused a matrix without long rows\n"); return 1; }
  std::vector<double> rs((size_t)m, 1e300), xw((size_t)m, 1e300);
  std::vector<int> covered((size_t)m, 0);
  size_t bytes = 0;
  for (size_t t = 0; t < W.desc.size(); ++t) {
    const WtDesc &d = W.desc[t];
    if (d.geom < 1 || d.geom > kWcKpl || d.nrows < 1 || d.nrows > 32 * W.rq) { printf("FAIL: tile %zu geometry kpl %d rows %d\n", t, d.geom, d.nrows); return 1; }
    for (int r = 0; r < d.nrows; ++r) covered[d.r0 + r]++;
    // interior tiles first
    bool ghost = false;
    for (int r = d.r0; r < d.r0 + d.nrows; ++r)
      for (int q = ia[r]; q < ia[r + 1]; ++q) ghost |= ja[q] >= n;
    if (ghost != ((int)t >= W.n_int)) { printf("FAIL: tile %zu interior/boundary order\n", t); return 1; }
    bytes += (size_t)d.geom * 384 + 64;
    tile_rows(W, d, x, wlast, rs, xw);
  }
  if (bytes + 16 != W.blob.size()) { printf("FAIL: blob size %zu vs tiles %zu\n", W.blob.size(), bytes); return 1; }
  for (int i : W.long_rows) {   // rows left to the CSR stream kernel
    if (ia[i + 1] - ia[i] <= kWcTileNnz) { printf("FAIL: row %d listed as long\n", i); return 1; }
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
    const int L = kWcTileNnz + 5;
    std::vector<int> ia = {0, 2, 2 + L, 2 + L + 3}, ja((size_t)L + 5, 0);
    std::vector<double> a((size_t)L + 5, 1.0);
    WcHost W;
    build_wc(3, 1, ia.data(), ja.data(), a.data(), &W);
    int rows = 0;
    for (const WtDesc &d : W.desc) rows += d.nrows;
    if (!W.ok || W.long_rows.size() != 1 || W.long_rows[0] != 1 || rows != 2 || W.desc.size() != 2) { printf("FAIL: long row handling\n"); ++bad; }
    ++n;
  }
  printf("%s: %d cases, %d failed\n", bad ? "WC_FORMAT_FAIL" : "WC_FORMAT_OK", n, bad);
  return bad ? 1 : 0;
}
