// hiergen native helpers (OpenMP): sparse matrix products used by the synthetic
// hierarchy generator.  Input generation only -- never on the measured path.
//
//  hg_spgemm_*      : C = A*B, CSR, sorted columns (stands in for PETSc MatMatMult,
//                     e.g. /root/reference/src/AIR_Operators_Setup.F90:788-793,1021-1025)
//  hg_masked_powers : sum_{t>s+1} coeff[t] * T_{t-1}, T_k = mask_S(T_{k-1} * A), the
//                     fixed-sparsity matrix powers of
//                     /root/reference/src/Gmres_Poly.F90:1177-1310 (row-wise, same order)
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

typedef int64_t i64;
typedef int32_t i32;

extern "C" {

int hg_num_threads() { return omp_get_max_threads(); }

// Phase 1: exact row counts of C = A*B.
void hg_spgemm_count(i64 m, const i64* ai, const i32* aj, const i64* bi, const i32* bj, i64* crow_nnz) {
#pragma omp parallel
  {
    std::vector<i32> buf;
#pragma omp for schedule(dynamic, 512)
    for (i64 r = 0; r < m; ++r) {
      buf.clear();
      for (i64 p = ai[r]; p < ai[r + 1]; ++p) {
        i32 k = aj[p];
        buf.insert(buf.end(), bj + bi[k], bj + bi[k + 1]);
      }
      std::sort(buf.begin(), buf.end());
      crow_nnz[r] = std::unique(buf.begin(), buf.end()) - buf.begin();
    }
  }
}

// Phase 2: numeric fill; ci is the exclusive prefix sum of the counts.
void hg_spgemm_fill(i64 m, const i64* ai, const i32* aj, const double* av, const i64* bi, const i32* bj,
                    const double* bv, const i64* ci, i32* cj, double* cv) {
#pragma omp parallel
  {
    std::vector<std::pair<i32, double>> buf;
#pragma omp for schedule(dynamic, 512)
    for (i64 r = 0; r < m; ++r) {
      buf.clear();
      for (i64 p = ai[r]; p < ai[r + 1]; ++p) {
        i32 k = aj[p];
        double a = av[p];
        for (i64 q = bi[k]; q < bi[k + 1]; ++q) buf.emplace_back(bj[q], a * bv[q]);
      }
      std::stable_sort(buf.begin(), buf.end(),
                       [](const std::pair<i32, double>& x, const std::pair<i32, double>& y) { return x.first < y.first; });
      i64 o = ci[r];
      size_t t = 0;
      while (t < buf.size()) {
        i32 c = buf[t].first;
        double s = 0.0;
        while (t < buf.size() && buf[t].first == c) s += buf[t++].second;
        cj[o] = c;
        cv[o] = s;
        ++o;
      }
    }
  }
}

// Fixed-sparsity matrix powers.  S (si,sj,sv) = A^s with its own sparsity; A (ai,aj,av).
// acc (size nnz(S)) receives sum_{term=s+2..ncoef} coeff[term-1] * T, T the masked power.
void hg_masked_powers(i64 n, const i64* si, const i32* sj, const double* sv, const i64* ai, const i32* aj,
                      const double* av, int ncoef, int sparsity_order, const double* coeff, double* acc) {
#pragma omp parallel
  {
    std::vector<i32> m1;      // index into row pattern
    std::vector<double> mv;   // matching A value
    std::vector<i64> mstart;  // per j_loc start into m1/mv
    std::vector<double> prev, pw;
#pragma omp for schedule(dynamic, 256)
    for (i64 r = 0; r < n; ++r) {
      const i64 s0 = si[r], nc = si[r + 1] - s0;
      const i32* cols = sj + s0;
      m1.clear(); mv.clear(); mstart.assign(nc + 1, 0);
      for (i64 jl = 0; jl < nc; ++jl) {
        mstart[jl] = (i64)m1.size();
        const i32 k = cols[jl];
        i64 p = 0, q = ai[k];
        const i64 qe = ai[k + 1];
        while (p < nc && q < qe) {  // intersect two sorted lists
          if (cols[p] < aj[q]) ++p;
          else if (cols[p] > aj[q]) ++q;
          else { m1.push_back((i32)p); mv.push_back(av[q]); ++p; ++q; }
        }
      }
      mstart[nc] = (i64)m1.size();
      prev.assign(sv + s0, sv + s0 + nc);
      pw.assign(nc, 0.0);
      for (i64 t = 0; t < nc; ++t) acc[s0 + t] = 0.0;
      for (int term = sparsity_order + 2; term <= ncoef; ++term) {
        std::fill(pw.begin(), pw.end(), 0.0);
        for (i64 jl = 0; jl < nc; ++jl) {
          const double pj = prev[jl];
          for (i64 t = mstart[jl]; t < mstart[jl + 1]; ++t) pw[m1[t]] += pj * mv[t];
        }
        const double c = coeff[term - 1];
        if (c != 0.0)
          for (i64 t = 0; t < nc; ++t) acc[s0 + t] += c * pw[t];
        prev.swap(pw);
      }
    }
  }
}

}  // extern "C"
