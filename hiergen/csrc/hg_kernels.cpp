// hiergen native helpers (OpenMP): sparse matrix products used by the synthetic
// hierarchy generator.  Input generation only -- never on the measured path.
//
//  hg_spgemm_*      : C = A*B, CSR, sorted columns (stands in for PETSc MatMatMult,
//                     e.g. /root/reference/src/AIR_Operators_Setup.F90:788-793,1021-1025)
//  hg_masked_powers : sum_{t>s+1} coeff[t] * T_{t-1}, T_k = mask_S(T_{k-1} * A), the
//                     fixed-sparsity matrix powers of
//                     /root/reference/src/Gmres_Poly.F90:1177-1310 (row-wise, same order)
#include <algorithm>
#include <cmath>
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

// One-pass C = A*B: rows are processed in chunks, every chunk keeps its result until the caller
// has allocated the output (hg_spgemm_fetch).  Same per-row arithmetic as count + fill.
struct SpgemmResult {
  i64 m = 0;
  int chunk = 4096;
  std::vector<std::vector<i32>> cols;
  std::vector<std::vector<double>> vals;
  std::vector<i64> rownnz;
};

void* hg_spgemm_run(i64 m, i64 ncols_b, const i64* ai, const i32* aj, const double* av, const i64* bi, const i32* bj,
                    const double* bv, i64* total_nnz) {
  SpgemmResult* R = new SpgemmResult();
  R->m = m;
  const i64 nchunk = (m + R->chunk - 1) / R->chunk;
  R->cols.resize((size_t)nchunk); R->vals.resize((size_t)nchunk); R->rownnz.assign((size_t)m, 0);
#pragma omp parallel
  {
    // sparse accumulator: dense value array + marker + list of touched columns; contributions to a
    // column are added in encounter order (the order a stable sort by column would group them in)
    std::vector<double> acc((size_t)std::max<i64>(ncols_b, 1), 0.0);
    std::vector<i32> mark((size_t)std::max<i64>(ncols_b, 1), -1);
    std::vector<i32> touched;
#pragma omp for schedule(dynamic, 1)
    for (i64 ch = 0; ch < nchunk; ++ch) {
      std::vector<i32>& oc = R->cols[(size_t)ch];
      std::vector<double>& ov = R->vals[(size_t)ch];
      const i64 r0 = ch * R->chunk, r1 = std::min<i64>(m, r0 + R->chunk);
      for (i64 r = r0; r < r1; ++r) {
        touched.clear();
        const i32 tag = (i32)r;   // unique per row: the marker never needs a reset
        for (i64 p = ai[r]; p < ai[r + 1]; ++p) {
          const i32 k = aj[p];
          const double a = av[p];
          for (i64 q = bi[k]; q < bi[k + 1]; ++q) {
            const i32 c = bj[q];
            if (mark[c] != tag) { mark[c] = tag; acc[c] = a * bv[q]; touched.push_back(c); }
            else acc[c] += a * bv[q];
          }
        }
        std::sort(touched.begin(), touched.end());
        for (i32 c : touched) { oc.push_back(c); ov.push_back(acc[c]); }
        R->rownnz[(size_t)r] = (i64)touched.size();
      }
    }
  }
  i64 tot = 0;
  for (i64 r = 0; r < m; ++r) tot += R->rownnz[(size_t)r];
  *total_nnz = tot;
  return R;
}

void hg_spgemm_fetch(void* h, i64* ci, i32* cj, double* cv) {
  SpgemmResult* R = (SpgemmResult*)h;
  ci[0] = 0;
  for (i64 r = 0; r < R->m; ++r) ci[r + 1] = ci[r] + R->rownnz[(size_t)r];
  const i64 nchunk = (i64)R->cols.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (i64 ch = 0; ch < nchunk; ++ch) {
    const i64 o = ci[ch * R->chunk];
    const std::vector<i32>& oc = R->cols[(size_t)ch];
    if (!oc.empty()) {
      memcpy(cj + o, oc.data(), oc.size() * sizeof(i32));
      memcpy(cv + o, R->vals[(size_t)ch].data(), oc.size() * sizeof(double));
    }
  }
  delete R;
}

// remove_small_from_sparse (/root/reference/src/PETSc_Helper.F90:207-412): keep |v| >= rowtol,
// rowtol = tol * max|row| (relative 1: incl. diagonal, -1: excl. diagonal) or tol (relative 0);
// drop_diagonal 0 never / -1 always / 1 allowed; optional lumping of the dropped entries onto the diagonal.
// Pass 1 (out arrays NULL) returns the kept count per row in rowcnt; pass 2 fills.
int hg_drop_small(i64 m, const i32* ai, const i32* aj, const double* av, double tol, int relative, int lump, int drop_diagonal,
                  i64* rowcnt, const i64* oi, i32* oj, double* ov) {
  int bad = 0;
#pragma omp parallel for schedule(static, 2048) reduction(| : bad)
  for (i64 r = 0; r < m; ++r) {
    const i64 p0 = ai[r], p1 = ai[r + 1];
    double rowtol = tol;
    if (relative == 1) {
      double mx = 0.0;
      for (i64 p = p0; p < p1; ++p) mx = std::max(mx, std::fabs(av[p]));
      rowtol = tol * mx;
    } else if (relative == -1) {
      double mx = -1.0; bool any = false;
      for (i64 p = p0; p < p1; ++p) if (aj[p] != r) { mx = std::max(mx, std::fabs(av[p])); any = true; }
      rowtol = any ? tol * mx : -tol * 1.7976931348623157e308;
    }
    double lumpsum = 0.0;
    i64 cnt = 0, dpos = -1;
    i64 o = oi ? oi[r] : 0;
    for (i64 p = p0; p < p1; ++p) {
      const bool isd = aj[p] == r;
      bool keep = std::fabs(av[p]) >= rowtol;
      if (drop_diagonal == -1) keep = keep && !isd;
      else if (drop_diagonal == 0) keep = keep || isd;
      if (keep) {
        if (oj) { oj[o] = aj[p]; ov[o] = av[p]; if (isd) dpos = o; ++o; }
        ++cnt;
      } else if (lump) {
        lumpsum += av[p];
      }
    }
    if (lump && oj && lumpsum != 0.0) {
      if (dpos >= 0) ov[dpos] += lumpsum; else bad = 1;
    }
    if (rowcnt) rowcnt[r] = cnt;
  }
  return bad;
}

// out = a[rows, :][:, col_map >= 0] with columns relabelled by col_map (MatCreateSubMatrix with sorted index sets).
// rows == NULL means all rows.  Pass 1 (oj == NULL): per-row counts; pass 2: fill.
void hg_extract(i64 nrows, const i32* rows, const i32* ai, const i32* aj, const double* av, const i64* col_map, i64* rowcnt,
                const i64* oi, i32* oj, double* ov) {
#pragma omp parallel for schedule(static, 4096)
  for (i64 r = 0; r < nrows; ++r) {
    const i64 src = rows ? rows[r] : r;
    i64 cnt = 0, o = oi ? oi[r] : 0;
    for (i64 p = ai[src]; p < ai[src + 1]; ++p) {
      const i64 c = col_map[aj[p]];
      if (c < 0) continue;
      if (oj) { oj[o] = (i32)c; ov[o] = av[p]; ++o; }
      ++cnt;
    }
    if (rowcnt) rowcnt[r] = cnt;
  }
}

// PMISR Luby loop (/root/reference/src/PMISR_Module.F90:271-671), synchronous rounds: an unassigned node joins the set
// (becomes F, cf = -1) when its measure is below every unassigned neighbour's; its neighbours become assigned (later C).
// Exact ties: the larger index loses (:519-521).  cf: 0 unassigned on entry (non-zero entries are kept).
void hg_pmisr(i64 n, const i32* si, const i32* sj, const double* measure, signed char* cf) {
  std::vector<unsigned char> assigned((size_t)n), inset((size_t)n);
  i64 remaining = 0;
  for (i64 i = 0; i < n; ++i) {
    assigned[(size_t)i] = cf[i] != 0;
    if (!assigned[(size_t)i] && std::fabs(measure[i]) < 1.0) { cf[i] = -1; assigned[(size_t)i] = 1; }
    if (!assigned[(size_t)i]) ++remaining;
  }
  while (remaining > 0) {
    i64 nsel = 0;
#pragma omp parallel for schedule(static, 4096) reduction(+ : nsel)
    for (i64 i = 0; i < n; ++i) {
      inset[(size_t)i] = 0;
      if (assigned[(size_t)i]) continue;
      bool ok = true;
      for (i64 p = si[i]; p < si[i + 1] && ok; ++p) {
        const i32 j = sj[p];
        if (!assigned[(size_t)j] && !(measure[i] < measure[j])) ok = false;
      }
      if (ok) { inset[(size_t)i] = 1; ++nsel; }
    }
    if (nsel == 0) {  // ties
#pragma omp parallel for schedule(static, 4096) reduction(+ : nsel)
      for (i64 i = 0; i < n; ++i) {
        inset[(size_t)i] = 0;
        if (assigned[(size_t)i]) continue;
        bool ok = true;
        for (i64 p = si[i]; p < si[i + 1] && ok; ++p) {
          const i32 j = sj[p];
          if (assigned[(size_t)j]) continue;
          if (measure[j] < measure[i]) ok = false;
          else if (measure[j] == measure[i] && j < i) ok = false;
        }
        if (ok) { inset[(size_t)i] = 1; ++nsel; }
      }
      if (nsel == 0) break;  // cannot happen on a symmetric strength matrix
    }
#pragma omp parallel for schedule(static, 4096)
    for (i64 i = 0; i < n; ++i) {
      if (!inset[(size_t)i]) continue;
      cf[i] = -1;
      assigned[(size_t)i] = 1;
      for (i64 p = si[i]; p < si[i + 1]; ++p) assigned[(size_t)sj[p]] = 1;   // benign race: every writer stores 1
    }
    remaining = 0;
#pragma omp parallel for schedule(static, 4096) reduction(+ : remaining)
    for (i64 i = 0; i < n; ++i) remaining += assigned[(size_t)i] ? 0 : 1;
  }
  for (i64 i = 0; i < n; ++i)
    if (cf[i] == 0) cf[i] = 1;
}

// per row: sum |a_ij| over F columns j != i, divided by |a_ii| (0 when the diagonal is missing or the row is not F)
void hg_diag_dom_ratio(i64 n, const i32* ai, const i32* aj, const double* av, const signed char* cf, double* ratio) {
#pragma omp parallel for schedule(static, 4096)
  for (i64 i = 0; i < n; ++i) {
    ratio[i] = 0.0;
    if (cf[i] != -1) continue;
    double diag = 0.0, offs = 0.0;
    for (i64 p = ai[i]; p < ai[i + 1]; ++p) {
      const i32 j = aj[p];
      if (cf[j] != -1) continue;
      if (j == i) diag += std::fabs(av[p]); else offs += std::fabs(av[p]);
    }
    if (diag != 0.0) ratio[i] = offs / diag;
  }
}


// lAIR restrictor (incomplete SAI, /root/reference/src/SAI_Z.F90:24-640 with incomplete = .TRUE.): for every C row i
// with F neighbourhood J = sparsity row i (sorted F-local indices), solve the square system
//     A_ff(J, J)^T z = -A_cf(i, J)^T          (dense LU with partial pivoting, as the reference's gesv for |J| <= 40;
//                                              larger neighbourhoods use the same direct solve here)
// and store Z(i, J) = z.  zv has the sparsity pattern (si, sj).
void hg_lair_z(i64 nc, const i32* si, const i32* sj, const i32* ffi, const i32* ffj, const double* ffv, const i32* cfi,
               const i32* cfj, const double* cfv, double* zv) {
#pragma omp parallel
  {
    std::vector<double> M, rhs;
    std::vector<int> piv;
#pragma omp for schedule(dynamic, 256)
    for (i64 i = 0; i < nc; ++i) {
      const int p0 = si[i], n = si[i + 1] - si[i];
      if (n == 0) continue;
      const i32* J = sj + p0;
      M.assign((size_t)n * n, 0.0);
      rhs.assign((size_t)n, 0.0);
      // M = A_ff(J, J)^T  (row-major: M[r * n + c] = A_ff(J[c], J[r]))
      for (int c = 0; c < n; ++c) {
        const int row = J[c];
        int q = 0;
        for (int p = ffi[row]; p < ffi[row + 1]; ++p) {
          const int col = ffj[p];
          while (q < n && J[q] < col) ++q;
          if (q < n && J[q] == col) M[(size_t)q * n + c] = ffv[p];
        }
      }
      {  // rhs = -A_cf(i, J)
        int q = 0;
        for (int p = cfi[i]; p < cfi[i + 1]; ++p) {
          const int col = cfj[p];
          while (q < n && J[q] < col) ++q;
          if (q < n && J[q] == col) rhs[q] = -cfv[p];
        }
      }
      // Gaussian elimination with partial pivoting
      for (int k = 0; k < n; ++k) {
        int pv = k;
        double best = std::fabs(M[(size_t)k * n + k]);
        for (int r = k + 1; r < n; ++r)
          if (std::fabs(M[(size_t)r * n + k]) > best) { best = std::fabs(M[(size_t)r * n + k]); pv = r; }
        if (pv != k) {
          for (int c = 0; c < n; ++c) std::swap(M[(size_t)k * n + c], M[(size_t)pv * n + c]);
          std::swap(rhs[k], rhs[pv]);
        }
        const double d = M[(size_t)k * n + k];
        if (d == 0.0) continue;
        for (int r = k + 1; r < n; ++r) {
          const double f = M[(size_t)r * n + k] / d;
          if (f == 0.0) continue;
          for (int c = k + 1; c < n; ++c) M[(size_t)r * n + c] -= f * M[(size_t)k * n + c];
          rhs[r] -= f * rhs[k];
        }
      }
      for (int k = n - 1; k >= 0; --k) {
        double v = rhs[k];
        for (int c = k + 1; c < n; ++c) v -= M[(size_t)k * n + c] * rhs[c];
        const double d = M[(size_t)k * n + k];
        rhs[k] = d != 0.0 ? v / d : 0.0;
      }
      for (int q = 0; q < n; ++q) zv[p0 + q] = rhs[q];
    }
  }
}
}  // extern "C"
