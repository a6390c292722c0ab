// pflare_b200.cu -- host side of the B200-native AIRG V-cycle apply behind the C-ABI of
// include/pflare_b200.h.  See DESIGN.md for the layout; reference citations are relative to
// the PFLARE source tree.
//
// What this file does, in order:
//   set_*            : copy the operators the reference's setup built (host CSR, natural
//                      numbering, exactly the objects listed in SURVEY.md section 8b "Upload hook").
//   finalize_setup   : (1) nested CF ordering -- level l vector = [F_l | level l+1 vector], so
//                      the identity blocks of R=[Z I] and P=[W;I] (src/Grid_Transfer.F90:329-461,
//                      588-815) become no-ops and every VecISCopy gather/scatter
//                      (src/FC_Smooth.F90:161-417) disappears; (2) R -> Z, P -> W, all operators
//                      relabelled into that ordering and uploaded once; (3) the V-cycle is
//                      compiled into a fixed program of fused SpMV ops
//                      (PCMG Kaskade wiring: src/AIR_MG_Setup.F90:967-1156; F/C smoothing:
//                      src/FC_Smooth.F90:421-640; Horner: src/Gmres_Poly.F90:1418-1484; Newton:
//                      src/Gmres_Poly_Newton.F90:763-875; Neumann: src/Neumann_Poly.F90:19-55);
//                      (4) the program is captured in a CUDA graph; small coarse levels run in
//                      one single-CTA kernel.
//   apply            : permute in, launch the graph, permute out.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/pflare_b200.h"
#include "kernels.cuh"
#include "comm.h"

using namespace pfb;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) return fail(100 + (int)e_, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// PFLARE_TOL_ZERO: single-precision literal 1e-12 widened to double (src/Pflare_Parameters.F90:206)
const double kTolZero = (double)1e-12f;

// TMA-pipelined kernel variants: {threads (= max rows per tile), nnz per tile, pipeline stages}
struct Variant { int nt, tile, stages; };
const Variant kVariants[] = {{0, 0, 0}, {256, 1024, 2}, {256, 1024, 3}, {256, 2048, 2}, {256, 2048, 3}, {512, 2048, 2}, {512, 4096, 2}, {128, 512, 3}, {128, 1024, 4},
                             {256, 1024, 2}, {256, 1024, 3}, {256, 2048, 2}, {512, 2048, 2}, {128, 512, 3},   // 9..13: row-mapped multiply
                             {256, 1024, 2}, {256, 1024, 2}, {256, 1024, 2}, {256, 1024, 3}, {128, 512, 2}, {128, 512, 3}};  // 14..19: register-capped for more resident CTAs
const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));

struct HostCSR {
  bool set = false;
  int m = 0, n = 0;
  std::vector<int> ia, ja;
  std::vector<double> a;
  // off-diagonal block (MPIAIJ), compressed ghost columns
  int n_ghost = 0;
  int64_t cstart = 0;
  std::vector<int> oia, oja;
  std::vector<double> oa;
  std::vector<int64_t> garray;
  int64_t nnz() const { return (int64_t)ja.size(); }
};

struct DevCSR {
  int m = 0, n = 0;
  int64_t nnz = 0;
  int64_t nx = 0;  // distinct columns referenced (byte model)
  int64_t nnz_model = -1;  // nnz counted by the reference's work model when it differs from nnz
  int *rp = nullptr, *col = nullptr;
  double *val = nullptr;
  int nblk = 0;
  int *blk = nullptr;
  int ntiles = 0;
  TileDesc *tiles = nullptr;
  bool valid() const { return rp != nullptr; }
};

struct Inv {
  int kind = 0;  // 0 none, 1 assembled, 2 diagonal, 3 polynomial
  HostCSR h;
  DevCSR d;
  std::vector<double> hdiag;
  double *ddiag = nullptr;
  int type = 0, diag_scale = 0;
  std::vector<double> re, im;
};

struct Level {
  bool set = false;
  int64_t rstart = 0;
  int n = 0, nf = 0, nc = 0;
  std::vector<int> is_f, is_c, smooth;
  HostCSR H[9];
  Inv inv_ff, inv_cc;
  // device
  DevCSR Z, W, Afc, Afcw, Aff, Acf, Acc, Coarse;   // Afcw = A_fc with the one-point W entry appended to every row
  bool w_onepoint = false;
  bool aff_diag_only = false;
  double *aff_diag = nullptr;  // diagonal of A_ff (F-local order): MF_VEC_DIAG and the fused local smooth
  double *acc_diag = nullptr;  // diagonal of A_cc (nested order)
  double *coarse_diag = nullptr;
  double *bc_save = nullptr;   // copy of b_c when the level has C smooths
  int64_t off = 0;             // offset of this level's vector in the nested arrays
  std::vector<int> pos;        // natural index -> nested position (relative to off)
  int *d_pos = nullptr, *d_inv = nullptr;
  bool any_c = false;
};

enum { OPK_SPMV = 0, OPK_EW = 1 };

struct Op {
  int kind = OPK_SPMV;
  SpmvOp s{};
  EwOp e{};
  int level = 0;
  int tag = 0;  // 1 restrict, 2 coarse, 3 A_fc(+W), 4 A_ff residual, 5 inverse, 6 elementwise, 7 fused local smooth, 8 A_cf, 9 A_cc
  double bytes = 0, nnz = 0;
};

struct Ctx {
  int rank = 0, nranks = 1, device = 0, no_levels = 0;
  std::vector<Level> L;  // 1-based
  bool finalized = false;
  cudaStream_t stream = nullptr;
  std::vector<void *> allocs;
  double dev_bytes = 0;
  // nested vectors + scratch
  double *xb = nullptr, *bb = nullptr;
  double *scr[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *io_b = nullptr, *io_x = nullptr;  // staging for host-pointer calls
  int maxn = 0;
  // program
  std::vector<Op> prog;
  int tail_begin = -1, tail_end = -1;  // [begin,end) range of ops executed by the tail kernel
  DevOp *d_tail = nullptr;
  int tail_levels = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int graph_kernels = 0;
  // options
  int use_graph = 1, fuse = 1;
  int tail_rows = 2048;
  int64_t tail_nnz = 40000;
  int kernel = 1;        // 0: smem-staged stream kernel, 1..: TMA-pipelined variants (kVariants)
  int tile_kernel = 1;   // the variant the uploaded tile lists were built for
  int ctas_per_sm = 0;   // 0 = from the occupancy calculator
  int dbg_seq = 0;       // measurement only (wrong results): sequential instead of indexed x gathers
  int num_sms = 148;
  std::unique_ptr<Comm> comm;
};

template <class T>
int dev_alloc(Ctx *c, T **p, size_t n) {
  void *q = nullptr;
  size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&q, bytes));
  c->allocs.push_back(q);
  c->dev_bytes += (double)bytes;
  *p = (T *)q;
  return 0;
}

template <class T>
int dev_upload(Ctx *c, T **p, const std::vector<T> &v) {
  int rc = dev_alloc(c, p, v.size());
  if (rc) return rc;
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// Row-block partition: consecutive rows with <= kTile nnz and <= kMaxRowsPerBlk rows per block;
// a row longer than kTile forms its own block.
std::vector<int> make_blocks(const std::vector<int> &ia, int m, int tile = kTile, int maxrows = kMaxRowsPerBlk) {
  std::vector<int> blk;
  blk.push_back(0);
  int r = 0;
  while (r < m) {
    int r0 = r;
    int64_t base = ia[r0];
    if (ia[r0 + 1] - base > tile) {
      r = r0 + 1;
    } else {
      while (r < m && (r - r0) < maxrows && ia[r + 1] - base <= tile) ++r;
    }
    blk.push_back(r);
  }
  return blk;
}

// over-allocating upload: the TMA kernel's bulk copies round their extent up to 16 bytes
template <class T>
int dev_upload_padded(Ctx *c, T **p, const std::vector<T> &v, size_t pad) {
  int rc = dev_alloc(c, p, v.size() + pad);
  if (rc) return rc;
  CUDA_TRY(cudaMemset(*p, 0, (v.size() + pad) * sizeof(T)));
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int upload_csr(Ctx *c, const HostCSR &h, DevCSR *d) {
  if (h.nnz() >= (int64_t)2147483647) return fail(3, "operator has >= 2^31 nonzeros (32-bit PetscInt only)");
  d->m = h.m;
  d->n = h.n;
  d->nnz = h.nnz();
  std::vector<unsigned char> seen((size_t)std::max(h.n, 1), 0);
  int64_t nx = 0;
  for (int cidx : h.ja)
    if (!seen[cidx]) { seen[cidx] = 1; ++nx; }
  d->nx = nx;
  int rc;
  if ((rc = dev_upload_padded(c, &d->rp, h.ia, 8))) return rc;
  if ((rc = dev_upload_padded(c, &d->col, h.ja, 8))) return rc;
  if ((rc = dev_upload_padded(c, &d->val, h.a, 8))) return rc;
  std::vector<int> blk = make_blocks(h.ia, h.m);
  d->nblk = (int)blk.size() - 1;
  if ((rc = dev_upload(c, &d->blk, blk))) return rc;
  const Variant &V = kVariants[c->tile_kernel];
  std::vector<int> tb = make_blocks(h.ia, h.m, V.tile, V.nt);
  std::vector<TileDesc> tiles(tb.size() - 1);
  for (size_t t = 0; t + 1 < tb.size(); ++t) tiles[t] = TileDesc{tb[t], tb[t + 1] - tb[t], h.ia[tb[t]], h.ia[tb[t + 1]] - h.ia[tb[t]]};
  d->ntiles = (int)tiles.size();
  if ((rc = dev_upload(c, &d->tiles, tiles))) return rc;
  return 0;
}

// rows permuted by rowpos (new row index), columns relabelled by colpos; columns sorted per row.
HostCSR remap(const HostCSR &A, const int *rowpos, const int *colpos, int new_n) {
  HostCSR B;
  B.set = true;
  B.m = A.m;
  B.n = new_n;
  B.ia.assign((size_t)A.m + 1, 0);
  for (int i = 0; i < A.m; ++i) {
    int r = rowpos ? rowpos[i] : i;
    B.ia[(size_t)r + 1] = A.ia[i + 1] - A.ia[i];
  }
  for (int i = 0; i < A.m; ++i) B.ia[i + 1] += B.ia[i];
  B.ja.resize(A.ja.size());
  B.a.resize(A.a.size());
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < A.m; ++i) {
    int r = rowpos ? rowpos[i] : i;
    int o = B.ia[r];
    const int p0 = A.ia[i], p1 = A.ia[i + 1];
    bool sorted = true;
    int prev = -1;
    for (int p = p0; p < p1; ++p) {
      int cc = colpos ? colpos[A.ja[p]] : A.ja[p];
      B.ja[o + (p - p0)] = cc;
      B.a[o + (p - p0)] = A.a[p];
      if (cc < prev) sorted = false;
      prev = cc;
    }
    if (!sorted) {
      const int len = p1 - p0;
      std::vector<std::pair<int, double>> tmp((size_t)len);
      for (int k = 0; k < len; ++k) tmp[k] = {B.ja[o + k], B.a[o + k]};
      std::sort(tmp.begin(), tmp.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) { return x.first < y.first; });
      for (int k = 0; k < len; ++k) { B.ja[o + k] = tmp[k].first; B.a[o + k] = tmp[k].second; }
    }
  }
  return B;
}

std::vector<double> extract_diag(const HostCSR &A) {
  std::vector<double> d((size_t)A.m, 0.0);
  for (int i = 0; i < A.m; ++i)
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (A.ja[p] == i) d[i] = A.a[p];
  return d;
}

bool is_diag_only(const HostCSR &A) {
  for (int i = 0; i < A.m; ++i)
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (A.ja[p] != i) return false;
  return true;
}

// ------------------------------------------------------------------ byte model
double spmv_bytes(const DevCSR &A, int n_aux_reads, int w) {
  return 12.0 * (double)A.nnz + 4.0 * ((double)A.m + 1) + 8.0 * (double)A.nx + 8.0 * (double)A.m * w + 8.0 * (double)A.m * n_aux_reads;
}

// ------------------------------------------------------------------ program construction
struct Builder {
  Ctx *c;
  std::vector<Op> *out;
  int level = 0;

  SpmvOp base(const DevCSR &A, const double *x) {
    SpmvOp s{};
    s.rp = A.rp; s.col = A.col; s.val = A.val; s.m = A.m; s.nblk = A.nblk; s.blk = A.blk;
    s.tiles = A.tiles; s.ntiles = A.ntiles; s.dbg_seq = c->dbg_seq;
    s.x = x; s.nloc = A.n; s.beta = 1.0;
    return s;
  }
  void push_spmv(const SpmvOp &s, const DevCSR &A, int tag, int aux_reads, int w, double extra_bytes = 0) {
    Op o;
    o.kind = OPK_SPMV; o.s = s; o.level = level; o.tag = tag;
    o.bytes = spmv_bytes(A, aux_reads, w) + extra_bytes;
    o.nnz = (double)(A.nnz_model >= 0 ? A.nnz_model : A.nnz);
    out->push_back(o);
  }
  // out (=|+=) alpha * a .* b ./ dv
  void push_ew(int n, const double *a, const double *b, const double *dv, double alpha, double *dst, int mode,
               const int *gather = nullptr, const int *scatter = nullptr) {
    Op o;
    o.kind = OPK_EW; o.level = level; o.tag = 6;
    o.e.n = n; o.e.a = a; o.e.b = b; o.e.dv = dv; o.e.alpha = alpha; o.e.out = dst; o.e.mode = mode;
    o.e.gather = gather; o.e.scatter = scatter;
    o.bytes = 8.0 * n * (1 + (b ? 1 : 0) + (dv ? 1 : 0) + (mode == 2 ? 2 : 1)) + 4.0 * n * ((gather ? 1 : 0) + (scatter ? 1 : 0));
    out->push_back(o);
  }

  // dst (=|+=) inverse * src.   mode 1 = set, 2 = add.  A/Adiag = the matrix a polynomial applies.
  int emit_inv(const Inv &I, const DevCSR &A, const double *Adiag, int n, const double *src, double *dst, int mode) {
    double **S = c->scr;
    if (I.kind == 1) {
      SpmvOp s = base(I.d, src);
      s.out = dst; s.out_mode = mode;
      push_spmv(s, I.d, 5, 0, mode == 2 ? 2 : 1);
      return 0;
    }
    if (I.kind == 2) {
      push_ew(n, src, I.ddiag, nullptr, 1.0, dst, mode);
      return 0;
    }
    if (I.kind != 3) return fail(4, "level %d: approximate inverse not set", level);
    if (!A.valid()) return fail(4, "level %d: matrix-free polynomial needs its matrix (set_csr)", level);
    const bool neumann = I.type == PFLARE_B200_INV_NEUMANN;
    const bool scaled = neumann || I.diag_scale;
    const double *rhs = src;
    if (scaled) {  // MF_VEC_RHS = x ./ D  (src/Gmres_Poly.F90:1406-1407, Gmres_Poly_Newton.F90:747-748)
      push_ew(n, src, nullptr, Adiag, 1.0, S[2], 1);
      rhs = S[2];
    }
    const int nc = (int)I.re.size();
    if (I.type == PFLARE_B200_INV_NEWTON || I.type == PFLARE_B200_INV_NEWTON_NO_EXTRA) {
      // petsc_newton, src/Gmres_Poly_Newton.F90:763-875.  t ping-pongs S[3]/S[4]; u = S[5].
      const double *re = I.re.data(), *im = I.im.data();
      double *y = (mode == 1) ? dst : S[6];
      bool y_init = false;
      const double *t = rhs;
      int flip = 0;
      auto next_t = [&]() { double *p = S[3 + flip]; flip ^= 1; return p; };
      auto acc_mode = [&]() { int mth = y_init ? 2 : 1; y_init = true; return mth; };
      int i = 1;
      while (i <= nc - 1) {
        if (im[i - 1] == 0.0) {
          if (std::fabs(re[i - 1]) < kTolZero) { i += 1; continue; }
          const double th = re[i - 1];
          double *tn = next_t();
          SpmvOp s = base(A, t);
          if (scaled) s.D = Adiag;
          s.neumann = neumann;
          s.aux = t; s.alpha = 1.0; s.beta = -1.0 / th;
          s.out = tn; s.out_mode = 1;
          s.acc = y; s.gamma = 1.0 / th; s.acc_src = t; s.acc_mode = acc_mode();
          push_spmv(s, A, 5, 1 + (scaled ? 1 : 0), 1, 16.0 * n);
          t = tn;
          i += 1;
        } else {
          const double sq = re[i - 1] * re[i - 1] + im[i - 1] * im[i - 1];
          if (sq < kTolZero) { i += 2; continue; }
          SpmvOp s = base(A, t);
          if (scaled) s.D = Adiag;
          s.neumann = neumann;
          s.aux = t; s.alpha = 2.0 * re[i - 1]; s.beta = -1.0;
          s.out = S[5]; s.out_mode = 1;
          s.acc = y; s.gamma = 1.0 / sq; s.acc_src = nullptr; s.acc_mode = acc_mode();
          push_spmv(s, A, 5, 1 + (scaled ? 1 : 0), 1, 16.0 * n);
          if (i <= nc - 2) {
            double *tn = next_t();
            SpmvOp s2 = base(A, S[5]);
            if (scaled) s2.D = Adiag;
            s2.neumann = neumann;
            s2.aux = t; s2.alpha = 1.0; s2.beta = -1.0 / sq;
            s2.out = tn; s2.out_mode = 1;
            push_spmv(s2, A, 5, 1 + (scaled ? 1 : 0), 1);
            t = tn;
          }
          i += 2;
        }
      }
      if (im[nc - 1] == 0.0 && std::fabs(re[nc - 1]) > kTolZero) {
        push_ew(n, t, nullptr, nullptr, 1.0 / re[nc - 1], y, acc_mode());
      }
      if (!y_init) push_ew(n, rhs, nullptr, nullptr, 0.0, y, 1);  // all roots skipped: y = 0
      if (mode == 2) push_ew(n, y, nullptr, nullptr, 1.0, dst, 2);
      return 0;
    }
    // Horner, src/Gmres_Poly.F90:1418-1484 (Neumann: all coefficients 1, A' = I - D^-1 A)
    const double *co = I.re.data();
    std::vector<int> steps;
    for (int order = nc - 2; order >= 0; --order)
      if (co[order] != 0.0) steps.push_back(order);
    if (steps.empty()) {
      push_ew(n, rhs, nullptr, nullptr, co[nc - 1], dst, mode);
      return 0;
    }
    push_ew(n, rhs, nullptr, nullptr, co[nc - 1], S[3], 1);  // y = c_n x
    const double *ycur = S[3];
    int flip = 1;
    for (size_t k = 0; k < steps.size(); ++k) {
      const bool last = (k + 1 == steps.size());
      double *ynext = last ? dst : S[3 + flip];
      flip ^= 1;
      SpmvOp s = base(A, ycur);
      if (scaled) s.D = Adiag;
      s.neumann = neumann;
      s.aux = rhs; s.alpha = co[steps[k]]; s.beta = 1.0;
      s.out = ynext; s.out_mode = last ? mode : 1;
      push_spmv(s, A, 5, 1 + (scaled ? 1 : 0), (last && mode == 2) ? 2 : 1);
      ycur = ynext;
    }
    return 0;
  }

  // x_f = W x_c as its own op (used when the first smoothing run is not an F smooth)
  void emit_prolong(Level &Lv, double *xf, const double *xc) {
    SpmvOp s = base(Lv.W, xc);
    s.out = xf; s.out_mode = 1;
    push_spmv(s, Lv.W, 3, 0, 1);
  }

  int emit_f_smooths(Level &Lv, bool first_smooth, bool prolong_pending, int its) {
    double *xb = c->xb + Lv.off, *bb = c->bb + Lv.off;
    double *xf = xb, *xc = xb + Lv.nf;
    const double *bf = bb;
    double **S = c->scr;
    (void)first_smooth;
    const bool fuse_w = prolong_pending && Lv.w_onepoint && c->fuse;
    if (prolong_pending && !fuse_w) emit_prolong(Lv, xf, xc);
    // fully local variant: A_ff diagonal and diagonal inverse -> the whole F smooth is row-local
    const bool local = c->fuse && fuse_w && Lv.aff_diag_only && Lv.inv_ff.kind == 2;
    {
      // rhs = b_f - A_fc x_c (src/FC_Smooth.F90:533-538); with the fused one-point prolongation the
      // merged A_fc|W operator also produces x_f = W x_c (MatInterpolate) from the same gathers
      const DevCSR &A = fuse_w ? Lv.Afcw : Lv.Afc;
      SpmvOp s = base(A, xc);
      s.aux = bf; s.alpha = 1.0; s.beta = -1.0;
      double extra = 0;
      if (fuse_w) { s.wlast = 1; s.wout = xf; extra += 8.0 * Lv.nf; }
      if (local) {
        s.fd_a = Lv.aff_diag; s.fd_m = Lv.inv_ff.ddiag; s.fd_its = its;
        extra += 16.0 * Lv.nf;
        push_spmv(s, A, 7, 1, 0, extra);
        return 0;
      }
      s.out = S[0]; s.out_mode = 1;
      push_spmv(s, A, 3, 1, 1, extra);
    }
    for (int f = 0; f < its; ++f) {
      SpmvOp s = base(Lv.Aff, xf);  // r = rhs - A_ff x_f     (src/FC_Smooth.F90:544-549)
      s.aux = S[0]; s.alpha = 1.0; s.beta = -1.0;
      s.out = S[1]; s.out_mode = 1;
      push_spmv(s, Lv.Aff, 4, 1, 1);
      int rc = emit_inv(Lv.inv_ff, Lv.Aff, Lv.aff_diag, Lv.nf, S[1], xf, 2);  // x_f += M_ff r  (:552-557)
      if (rc) return rc;
    }
    return 0;
  }

  int emit_c_smooths(Level &Lv, bool prolong_pending, int its) {
    double *xb = c->xb + Lv.off;
    double *xf = xb, *xc = xb + Lv.nf;
    double **S = c->scr;
    if (prolong_pending) emit_prolong(Lv, xf, xc);
    if (!Lv.Acf.valid() || !Lv.Acc.valid() || Lv.inv_cc.kind == 0)
      return fail(4, "level %d: C-point smoothing requested but A_cf/A_cc/inv_A_cc not set", level);
    {
      SpmvOp s = base(Lv.Acf, xf);  // rhs_c = b_c - A_cf x_f  (src/FC_Smooth.F90:606-610)
      s.aux = Lv.bc_save; s.alpha = 1.0; s.beta = -1.0;
      s.out = S[0]; s.out_mode = 1;
      push_spmv(s, Lv.Acf, 8, 1, 1);
    }
    for (int k = 0; k < its; ++k) {
      SpmvOp s = base(Lv.Acc, xc);  // r_c = rhs_c - A_cc x_c  (:616-621)
      s.aux = S[0]; s.alpha = 1.0; s.beta = -1.0;
      s.out = S[1]; s.out_mode = 1;
      push_spmv(s, Lv.Acc, 9, 1, 1);
      int rc = emit_inv(Lv.inv_cc, Lv.Acc, Lv.acc_diag, Lv.nc, S[1], xc, 2);  // x_c += M_cc r_c (:624-629)
      if (rc) return rc;
    }
    return 0;
  }

  // mg_FC_point_richardson (src/FC_Smooth.F90:421-495); prolong = x_f must first be produced by W x_c
  int emit_fc_richardson(Level &Lv, bool prolong) {
    bool first = true, pending = prolong;
    for (int sm : Lv.smooth) {
      if (sm == 0) break;
      int rc = sm > 0 ? emit_f_smooths(Lv, first, pending, sm) : emit_c_smooths(Lv, pending, -sm);
      if (rc) return rc;
      first = false;
      pending = false;
    }
    if (pending) emit_prolong(Lv, c->xb + Lv.off, c->xb + Lv.off + Lv.nf);
    return 0;
  }
};

template <int NT, int TILE, int STAGES, bool ROWMAP = false, int MINB = 1>
int launch_tma(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  auto kern = spmv_tma_kernel<NT, TILE, STAGES, ROWMAP, MINB>;
  const size_t smem = sizeof(TmaStage<TILE, NT>) * STAGES;
  static int per_sm = 0;   // resident CTAs per SM of this instantiation (occupancy calculator, once)
  if (per_sm == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NT, smem));
    per_sm = std::max(nb, 1);
  }
  if (dry) return 0;
  const int want = c->ctas_per_sm > 0 ? std::min(c->ctas_per_sm, per_sm) : per_sm;
  const int grid = std::min(s.ntiles, c->num_sms * want);   // persistent: a multiple of the SM count
  kern<<<grid, NT, smem, st>>>(s);
  return 0;
}

int launch_op(Ctx *c, const Op &o, cudaStream_t st, bool dry = false) {
  if (o.kind == OPK_SPMV) {
    if (!dry && (o.s.m == 0 || o.s.nblk == 0)) return 0;
    int rc = 0;
    const int k = c->kernel;
    switch (k) {
      case 0: {
        if (dry) return 0;
        int grid = std::min(o.s.nblk, c->num_sms * 8 * 4);
        spmv_stream_kernel<<<grid, kThreads, 0, st>>>(o.s);
        break;
      }
      case 1: rc = launch_tma<256, 1024, 2>(c, o.s, st, dry); break;
      case 2: rc = launch_tma<256, 1024, 3>(c, o.s, st, dry); break;
      case 3: rc = launch_tma<256, 2048, 2>(c, o.s, st, dry); break;
      case 4: rc = launch_tma<256, 2048, 3>(c, o.s, st, dry); break;
      case 5: rc = launch_tma<512, 2048, 2>(c, o.s, st, dry); break;
      case 6: rc = launch_tma<512, 4096, 2>(c, o.s, st, dry); break;
      case 7: rc = launch_tma<128, 512, 3>(c, o.s, st, dry); break;
      case 8: rc = launch_tma<128, 1024, 4>(c, o.s, st, dry); break;
      case 9: rc = launch_tma<256, 1024, 2, true>(c, o.s, st, dry); break;
      case 10: rc = launch_tma<256, 1024, 3, true>(c, o.s, st, dry); break;
      case 11: rc = launch_tma<256, 2048, 2, true>(c, o.s, st, dry); break;
      case 12: rc = launch_tma<512, 2048, 2, true>(c, o.s, st, dry); break;
      case 13: rc = launch_tma<128, 512, 3, true>(c, o.s, st, dry); break;
      case 14: rc = launch_tma<256, 1024, 2, false, 5>(c, o.s, st, dry); break;
      case 15: rc = launch_tma<256, 1024, 2, false, 6>(c, o.s, st, dry); break;
      case 16: rc = launch_tma<256, 1024, 2, false, 8>(c, o.s, st, dry); break;
      case 17: rc = launch_tma<256, 1024, 3, false, 5>(c, o.s, st, dry); break;
      case 18: rc = launch_tma<128, 512, 2, false, 16>(c, o.s, st, dry); break;
      case 19: rc = launch_tma<128, 512, 3, false, 12>(c, o.s, st, dry); break;
      default: return fail(2, "unknown kernel variant %d", k);
    }
    if (rc) return rc;
    if (dry) return 0;
  } else {
    if (dry) return 0;
    if (o.e.n == 0) return 0;
    int grid = std::min((o.e.n + kThreads - 1) / kThreads, c->num_sms * 8);
    ew_kernel<<<grid, kThreads, 0, st>>>(o.e);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int run_program(Ctx *c, cudaStream_t st, int *nkernels) {
  int nk = 0;
  const int n = (int)c->prog.size();
  for (int i = 0; i < n; ++i) {
    if (i == c->tail_begin && c->tail_end > c->tail_begin) {
      tail_kernel<<<1, kTailThreads, 0, st>>>(c->d_tail, c->tail_end - c->tail_begin);
      CUDA_TRY(cudaGetLastError());
      ++nk;
      i = c->tail_end - 1;
      continue;
    }
    const Op &o = c->prog[i];
    if ((o.kind == OPK_SPMV && (o.s.m == 0 || o.s.nblk == 0)) || (o.kind == OPK_EW && o.e.n == 0)) continue;
    int rc = launch_op(c, o, st);
    if (rc) return rc;
    ++nk;
  }
  if (nkernels) *nkernels = nk;
  return 0;
}

int build_graph(Ctx *c) {
  {  // kernel attributes / occupancy of the selected SpMV variant, outside the capture
    Op dummy;
    int rc0 = launch_op(c, dummy, c->stream, true);
    if (rc0) return rc0;
  }
  if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
  if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
  CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  int nk = 0;
  int rc = run_program(c, c->stream, &nk);
  cudaError_t e = cudaStreamEndCapture(c->stream, &c->graph);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(100 + (int)e, "graph capture failed: %s", cudaGetErrorString(e));
  CUDA_TRY(cudaGraphInstantiate(&c->gexec, c->graph, 0));
  c->graph_kernels = nk;
  return 0;
}

int build_program(Ctx *c) {
  c->prog.clear();
  c->tail_begin = c->tail_end = -1;
  const int NL = c->no_levels;
  if (NL < 2) return 0;
  Builder B{c, &c->prog};
  // which levels go to the single-CTA tail: the longest suffix of small levels
  int ltail = NL + 1;
  for (int l = NL; l >= 1; --l) {
    Level &Lv = c->L[l];
    int64_t mx = 0;
    for (const DevCSR *A : {&Lv.Z, &Lv.W, &Lv.Afc, &Lv.Afcw, &Lv.Aff, &Lv.Acf, &Lv.Acc, &Lv.Coarse, &Lv.inv_ff.d, &Lv.inv_cc.d}) mx = std::max(mx, A->nnz);
    if (Lv.n <= c->tail_rows && mx <= c->tail_nnz) ltail = l; else break;
  }
  c->tail_levels = (ltail <= NL) ? NL - ltail + 1 : 0;
  // down: b_{l+1} = b_c + Z b_f  (MatRestrict with R = [Z I])
  for (int l = 1; l <= NL - 1; ++l) {
    Level &Lv = c->L[l];
    B.level = l;
    if (l == ltail) c->tail_begin = (int)c->prog.size();
    if (Lv.any_c) B.push_ew(Lv.nc, c->bb + Lv.off + Lv.nf, nullptr, nullptr, 1.0, Lv.bc_save, 1);
    SpmvOp s = B.base(Lv.Z, c->bb + Lv.off);
    s.out = c->bb + Lv.off + Lv.nf; s.out_mode = 2;
    B.push_spmv(s, Lv.Z, 1, 0, 2);
  }
  // coarse solve: x_L = inv_A_ff(L) b_L  (mg_coarse_shell_apply, src/FC_Smooth.F90:29-49)
  {
    Level &Lv = c->L[NL];
    B.level = NL;
    if (NL == ltail) c->tail_begin = (int)c->prog.size();
    int rc = B.emit_inv(Lv.inv_ff, Lv.Coarse, Lv.coarse_diag, Lv.n, c->bb + Lv.off, c->xb + Lv.off, 1);
    if (rc) return rc;
  }
  // up: x_l = P x_{l+1}; one mg_FC_point_richardson
  for (int l = NL - 1; l >= 1; --l) {
    Level &Lv = c->L[l];
    B.level = l;
    int rc = B.emit_fc_richardson(Lv, true);
    if (rc) return rc;
    if (l == ltail) c->tail_end = (int)c->prog.size();
  }
  if (ltail == NL && c->tail_begin >= 0) c->tail_end = c->tail_begin;  // coarse level alone: not worth a tail
  if (c->tail_begin >= 0 && c->tail_end > c->tail_begin) {
    std::vector<DevOp> ops;
    for (int i = c->tail_begin; i < c->tail_end; ++i) {
      DevOp d{};
      d.kind = c->prog[i].kind; d.s = c->prog[i].s; d.e = c->prog[i].e;
      ops.push_back(d);
    }
    int rc = dev_upload(c, &c->d_tail, ops);
    if (rc) return rc;
  } else {
    c->tail_begin = c->tail_end = -1;
    c->tail_levels = 0;
  }
  return 0;
}

int check_handle(void *h, Ctx **c) {
  if (!h) return fail(1, "null handle");
  *c = (Ctx *)h;
  CUDA_TRY(cudaSetDevice((*c)->device));
  return 0;
}

}  // namespace

// ====================================================================== C-ABI
extern "C" {

const char *pflare_b200_last_error(void) { return g_err.c_str(); }

int pflare_b200_get_unique_id(void *id) {
  std::string err;
  if (!Comm::unique_id(id, &err)) return fail(20, "%s", err.c_str());
  return 0;
}

int pflare_b200_create(void **handle, int rank, int nranks, const void *unique_id, int device, int no_levels) {
  if (!handle) return fail(1, "null handle pointer");
  *handle = nullptr;
  if (no_levels < 1) return fail(2, "no_levels must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(10, "no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(10, "device %d out of range (%d devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  Ctx *c = new Ctx();
  c->rank = rank; c->nranks = nranks; c->device = device; c->no_levels = no_levels;
  c->L.resize((size_t)no_levels + 1);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  if (nranks > 1) {
    if (!unique_id) { delete c; return fail(20, "nranks > 1 needs a unique id"); }
    std::string err;
    c->comm.reset(Comm::create(rank, nranks, unique_id, &err));
    if (!c->comm) { delete c; return fail(20, "communicator: %s", err.c_str()); }
  }
  *handle = c;
  return 0;
}

int pflare_b200_set_level(void *handle, int our_level, int64_t rstart, int n_local, int n_fine, const int *is_fine,
                          int n_coarse, const int *is_coarse, const int *smooth_order, int n_smooth) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  if (our_level < c->no_levels && n_fine + n_coarse != n_local) return fail(2, "level %d: n_fine + n_coarse != n_local", our_level);
  Level &Lv = c->L[our_level];
  Lv.set = true; Lv.rstart = rstart; Lv.n = n_local; Lv.nf = n_fine; Lv.nc = n_coarse;
  Lv.is_f.assign(is_fine, is_fine + n_fine);
  Lv.is_c.assign(is_coarse, is_coarse + n_coarse);
  Lv.smooth.assign(smooth_order, smooth_order + n_smooth);
  Lv.any_c = false;
  for (int s : Lv.smooth) { if (s == 0) break; if (s < 0) Lv.any_c = true; }
  c->finalized = false;
  return 0;
}

int pflare_b200_set_csr(void *handle, int our_level, int which, int m, int n_local_cols, int64_t cstart, const int *di,
                        const int *dj, const double *da, int n_ghost, const int *oi, const int *oj, const double *oa,
                        const int64_t *garray) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  if (which < 0 || which > 8) return fail(2, "bad operator selector %d", which);
  Level &Lv = c->L[our_level];
  HostCSR *H = &Lv.H[which];
  if (which == PFLARE_B200_INV_AFF) { Lv.inv_ff.kind = 1; H = &Lv.inv_ff.h; }
  if (which == PFLARE_B200_INV_ACC) { Lv.inv_cc.kind = 1; H = &Lv.inv_cc.h; }
  H->set = true; H->m = m; H->n = n_local_cols; H->cstart = cstart;
  H->ia.assign(di, di + m + 1);
  H->ja.assign(dj, dj + di[m]);
  H->a.assign(da, da + di[m]);
  H->n_ghost = n_ghost;
  H->oia.clear(); H->oja.clear(); H->oa.clear(); H->garray.clear();
  if (n_ghost > 0) {
    H->oia.assign(oi, oi + m + 1);
    H->oja.assign(oj, oj + oi[m]);
    H->oa.assign(oa, oa + oi[m]);
    H->garray.assign(garray, garray + n_ghost);
  }
  c->finalized = false;
  return 0;
}

int pflare_b200_set_diag(void *handle, int our_level, int which, int n, const double *d) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  if (which != PFLARE_B200_INV_AFF && which != PFLARE_B200_INV_ACC) return fail(2, "set_diag: which must be INV_AFF or INV_ACC");
  Inv &I = which == PFLARE_B200_INV_AFF ? c->L[our_level].inv_ff : c->L[our_level].inv_cc;
  I.kind = 2;
  I.hdiag.assign(d, d + n);
  c->finalized = false;
  return 0;
}

int pflare_b200_set_poly(void *handle, int our_level, int which, int inverse_type, int ncoef, const double *coeffs_re,
                         const double *coeffs_im, int diag_scale) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  if (which != PFLARE_B200_INV_AFF && which != PFLARE_B200_INV_ACC) return fail(2, "set_poly: which must be INV_AFF or INV_ACC");
  if (ncoef < 1) return fail(2, "set_poly: ncoef must be >= 1");
  switch (inverse_type) {
    case PFLARE_B200_INV_POWER: case PFLARE_B200_INV_ARNOLDI: case PFLARE_B200_INV_NEWTON:
    case PFLARE_B200_INV_NEWTON_NO_EXTRA: case PFLARE_B200_INV_NEUMANN: break;
    default: return fail(2, "inverse type %d cannot be applied matrix-free (src/PCPFLAREINV.c:708-711)", inverse_type);
  }
  Inv &I = which == PFLARE_B200_INV_AFF ? c->L[our_level].inv_ff : c->L[our_level].inv_cc;
  I.kind = 3; I.type = inverse_type; I.diag_scale = diag_scale;
  I.re.assign(coeffs_re, coeffs_re + ncoef);
  if (coeffs_im) I.im.assign(coeffs_im, coeffs_im + ncoef); else I.im.assign((size_t)ncoef, 0.0);
  c->finalized = false;
  return 0;
}

int pflare_b200_finalize_setup(void *handle) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  const int NL = c->no_levels;
  if (c->nranks > 1) return fail(30, "multi-rank finalize not available in this build");
  for (int l = 1; l <= NL; ++l)
    if (!c->L[l].set) return fail(2, "level %d was never set", l);
  // sizes must chain: n_{l+1} == n_coarse(l)
  for (int l = 1; l < NL; ++l)
    if (c->L[l + 1].n != c->L[l].nc) return fail(2, "level %d has %d rows but level %d has %d C points", l + 1, c->L[l + 1].n, l, c->L[l].nc);
  // (1) nested positions, coarsest first
  c->maxn = 0;
  for (int l = NL; l >= 1; --l) {
    Level &Lv = c->L[l];
    Lv.pos.resize((size_t)Lv.n);
    if (l == NL) {
      std::iota(Lv.pos.begin(), Lv.pos.end(), 0);
    } else {
      std::vector<unsigned char> mark((size_t)Lv.n, 0);
      for (int j = 0; j < Lv.nf; ++j) {
        int i = Lv.is_f[j];
        if (i < 0 || i >= Lv.n || mark[i]) return fail(2, "level %d: IS_fine is not a valid index set", l);
        mark[i] = 1; Lv.pos[i] = j;
      }
      const std::vector<int> &pc = c->L[l + 1].pos;
      for (int k = 0; k < Lv.nc; ++k) {
        int i = Lv.is_c[k];
        if (i < 0 || i >= Lv.n || mark[i]) return fail(2, "level %d: IS_coarse overlaps IS_fine or is out of range", l);
        mark[i] = 1; Lv.pos[i] = Lv.nf + pc[k];
      }
    }
    c->maxn = std::max(c->maxn, Lv.n);
  }
  c->L[1].off = 0;
  for (int l = 1; l < NL; ++l) c->L[l + 1].off = c->L[l].off + c->L[l].nf;
  // (2) operators
  for (int l = 1; l <= NL; ++l) {
    Level &Lv = c->L[l];
    if (l < NL) {
      const std::vector<int> &pc = c->L[l + 1].pos;  // coarse index -> nested position on level l+1
      std::vector<int> fpos((size_t)Lv.n, -1);
      for (int j = 0; j < Lv.nf; ++j) fpos[Lv.is_f[j]] = j;
      // R = [Z I] -> Z with rows in level l+1 nested order, columns F-local
      const HostCSR &R = Lv.H[PFLARE_B200_R];
      if (!R.set || R.m != Lv.nc || R.n != Lv.n) return fail(2, "level %d: restrictor missing or wrong shape", l);
      HostCSR Z; Z.set = true; Z.m = Lv.nc; Z.n = Lv.nf; Z.ia.assign((size_t)Lv.nc + 1, 0);
      for (int i = 0; i < Lv.nc; ++i) {
        bool ident = false; int cnt = 0;
        for (int p = R.ia[i]; p < R.ia[i + 1]; ++p) {
          int col = R.ja[p];
          if (fpos[col] >= 0) ++cnt;
          else if (col == Lv.is_c[i] && R.a[p] == 1.0 && !ident) ident = true;
          else return fail(5, "level %d: restrictor row %d is not of the form [Z I]", l, i);
        }
        if (!ident) return fail(5, "level %d: restrictor row %d has no identity entry", l, i);
        Z.ia[(size_t)pc[i] + 1] = cnt;
      }
      for (int i = 0; i < Lv.nc; ++i) Z.ia[i + 1] += Z.ia[i];
      Z.ja.resize((size_t)Z.ia[Lv.nc]); Z.a.resize((size_t)Z.ia[Lv.nc]);
      for (int i = 0; i < Lv.nc; ++i) {
        int o = Z.ia[pc[i]];
        for (int p = R.ia[i]; p < R.ia[i + 1]; ++p)
          if (fpos[R.ja[p]] >= 0) { Z.ja[o] = fpos[R.ja[p]]; Z.a[o] = R.a[p]; ++o; }
      }
      if ((rc = upload_csr(c, Z, &Lv.Z))) return rc;
      // P = [W; I] -> W with F-local rows, columns in level l+1 nested order
      const HostCSR &P = Lv.H[PFLARE_B200_P];
      if (!P.set || P.m != Lv.n || P.n != Lv.nc) return fail(2, "level %d: prolongator missing or wrong shape", l);
      for (int k = 0; k < Lv.nc; ++k) {
        int i = Lv.is_c[k];
        if (P.ia[i + 1] - P.ia[i] != 1 || P.ja[P.ia[i]] != k || P.a[P.ia[i]] != 1.0)
          return fail(5, "level %d: prolongator C row %d is not an identity row", l, k);
      }
      HostCSR Wn; Wn.set = true; Wn.m = Lv.nf; Wn.n = Lv.nc; Wn.ia.assign((size_t)Lv.nf + 1, 0);
      for (int j = 0; j < Lv.nf; ++j) Wn.ia[j + 1] = Wn.ia[j] + (P.ia[Lv.is_f[j] + 1] - P.ia[Lv.is_f[j]]);
      Wn.ja.resize((size_t)Wn.ia[Lv.nf]); Wn.a.resize((size_t)Wn.ia[Lv.nf]);
      bool onept = true;
      for (int j = 0; j < Lv.nf; ++j) {
        int i = Lv.is_f[j], o = Wn.ia[j];
        if (P.ia[i + 1] - P.ia[i] > 1) onept = false;
        for (int p = P.ia[i]; p < P.ia[i + 1]; ++p) { Wn.ja[o] = P.ja[p]; Wn.a[o] = P.a[p]; ++o; }
      }
      HostCSR W = remap(Wn, nullptr, pc.data(), Lv.nc);
      if ((rc = upload_csr(c, W, &Lv.W))) return rc;
      Lv.w_onepoint = onept;
      // A_fc, A_ff
      const HostCSR &Afc = Lv.H[PFLARE_B200_AFC], &Aff = Lv.H[PFLARE_B200_AFF];
      if (!Afc.set || Afc.m != Lv.nf || Afc.n != Lv.nc) return fail(2, "level %d: A_fc missing or wrong shape", l);
      if (!Aff.set || Aff.m != Lv.nf || Aff.n != Lv.nf) return fail(2, "level %d: A_ff missing or wrong shape", l);
      HostCSR Afc2 = remap(Afc, nullptr, pc.data(), Lv.nc);
      if ((rc = upload_csr(c, Afc2, &Lv.Afc))) return rc;
      if (onept) {
        // merged A_fc|W: every row gets its W entry (or an explicit 0.0 * x_c[0]) appended as LAST entry
        HostCSR M; M.set = true; M.m = Lv.nf; M.n = Lv.nc; M.ia.assign((size_t)Lv.nf + 1, 0);
        for (int j = 0; j < Lv.nf; ++j) M.ia[j + 1] = M.ia[j] + (Afc2.ia[j + 1] - Afc2.ia[j]) + 1;
        M.ja.resize((size_t)M.ia[Lv.nf]); M.a.resize((size_t)M.ia[Lv.nf]);
        for (int j = 0; j < Lv.nf; ++j) {
          int o = M.ia[j];
          for (int p = Afc2.ia[j]; p < Afc2.ia[j + 1]; ++p) { M.ja[o] = Afc2.ja[p]; M.a[o] = Afc2.a[p]; ++o; }
          if (W.ia[j + 1] > W.ia[j]) { M.ja[o] = W.ja[W.ia[j]]; M.a[o] = W.a[W.ia[j]]; }
          else { M.ja[o] = 0; M.a[o] = 0.0; }
        }
        if ((rc = upload_csr(c, M, &Lv.Afcw))) return rc;
        Lv.Afcw.nnz_model = Afc2.nnz() + W.nnz();
      }
      if ((rc = upload_csr(c, Aff, &Lv.Aff))) return rc;
      Lv.aff_diag_only = is_diag_only(Aff);
      if ((rc = dev_upload(c, &Lv.aff_diag, extract_diag(Aff)))) return rc;
      // inverse of A_ff
      Inv &I = Lv.inv_ff;
      if (I.kind == 1) {
        if (I.h.m != Lv.nf || I.h.n != Lv.nf) return fail(2, "level %d: inv_A_ff has the wrong shape", l);
        if ((rc = upload_csr(c, I.h, &I.d))) return rc;
      } else if (I.kind == 2) {
        if ((int)I.hdiag.size() != Lv.nf) return fail(2, "level %d: diagonal inv_A_ff has the wrong size", l);
        if ((rc = dev_upload(c, &I.ddiag, I.hdiag))) return rc;
      } else if (I.kind == 0) {
        return fail(2, "level %d: inv_A_ff not set", l);
      }
      // C-point smoothing operators
      if (Lv.any_c) {
        const HostCSR &Acf = Lv.H[PFLARE_B200_ACF], &Acc = Lv.H[PFLARE_B200_ACC];
        if (!Acf.set || !Acc.set) return fail(2, "level %d: C smoothing requested but A_cf / A_cc not set", l);
        HostCSR Acf2 = remap(Acf, pc.data(), nullptr, Lv.nf);
        HostCSR Acc2 = remap(Acc, pc.data(), pc.data(), Lv.nc);
        if ((rc = upload_csr(c, Acf2, &Lv.Acf))) return rc;
        if ((rc = upload_csr(c, Acc2, &Lv.Acc))) return rc;
        if ((rc = dev_upload(c, &Lv.acc_diag, extract_diag(Acc2)))) return rc;
        Inv &J = Lv.inv_cc;
        if (J.kind == 1) {
          HostCSR M2 = remap(J.h, pc.data(), pc.data(), Lv.nc);
          if ((rc = upload_csr(c, M2, &J.d))) return rc;
        } else if (J.kind == 2) {
          std::vector<double> d2((size_t)Lv.nc);
          for (int k = 0; k < Lv.nc; ++k) d2[pc[k]] = J.hdiag[k];
          if ((rc = dev_upload(c, &J.ddiag, d2))) return rc;
        } else if (J.kind == 0) {
          return fail(2, "level %d: inv_A_cc not set", l);
        }
        if ((rc = dev_alloc(c, &Lv.bc_save, (size_t)Lv.nc))) return rc;
      }
    } else {
      // coarsest level: inv_A_ff(no_levels) (+ coarse_matrix for a matrix-free polynomial)
      Inv &I = Lv.inv_ff;
      const HostCSR &Cm = Lv.H[PFLARE_B200_COARSE];
      if (Cm.set) {
        if ((rc = upload_csr(c, Cm, &Lv.Coarse))) return rc;
        if ((rc = dev_upload(c, &Lv.coarse_diag, extract_diag(Cm)))) return rc;
      }
      if (I.kind == 1) {
        if (I.h.m != Lv.n) return fail(2, "coarse inverse has the wrong shape");
        if ((rc = upload_csr(c, I.h, &I.d))) return rc;
      } else if (I.kind == 2) {
        if ((int)I.hdiag.size() != Lv.n) return fail(2, "diagonal coarse inverse has the wrong size");
        if ((rc = dev_upload(c, &I.ddiag, I.hdiag))) return rc;
      } else if (I.kind == 3) {
        if (!Cm.set) return fail(2, "matrix-free coarse inverse needs coarse_matrix (PFLARE_B200_COARSE)");
      } else {
        return fail(2, "coarse inverse (inv_A_ff on the coarsest level) not set");
      }
    }
  }
  // (3) vectors
  const size_t n1 = (size_t)c->L[1].n;
  if ((rc = dev_alloc(c, &c->xb, n1))) return rc;
  if ((rc = dev_alloc(c, &c->bb, n1))) return rc;
  for (int k = 0; k < 7; ++k)
    if ((rc = dev_alloc(c, &c->scr[k], (size_t)c->maxn))) return rc;
  if ((rc = dev_alloc(c, &c->io_b, (size_t)c->maxn))) return rc;
  if ((rc = dev_alloc(c, &c->io_x, (size_t)c->maxn))) return rc;
  CUDA_TRY(cudaMemset(c->xb, 0, std::max<size_t>(n1, 1) * 8));
  CUDA_TRY(cudaMemset(c->bb, 0, std::max<size_t>(n1, 1) * 8));
  {
    Level &L1 = c->L[1];
    std::vector<int> inv((size_t)L1.n);
    for (int i = 0; i < L1.n; ++i) inv[L1.pos[i]] = i;
    if ((rc = dev_upload(c, &L1.d_pos, L1.pos))) return rc;
    if ((rc = dev_upload(c, &L1.d_inv, inv))) return rc;
  }
  // (4) program + graph
  if ((rc = build_program(c))) return rc;
  if (c->use_graph && NL >= 2) {
    if ((rc = build_graph(c))) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->finalized = true;
  return 0;
}

static int ensure_level_perm(Ctx *c, Level &Lv) {
  if (Lv.d_pos) return 0;
  std::vector<int> inv((size_t)Lv.n);
  for (int i = 0; i < Lv.n; ++i) inv[Lv.pos[i]] = i;
  int rc;
  if ((rc = dev_upload(c, &Lv.d_pos, Lv.pos))) return rc;
  if ((rc = dev_upload(c, &Lv.d_inv, inv))) return rc;
  return 0;
}

static int launch_ew_now(Ctx *c, int n, const double *a, double *out, const int *gather, const int *scatter) {
  Op o; o.kind = OPK_EW;
  o.e.n = n; o.e.a = a; o.e.alpha = 1.0; o.e.out = out; o.e.mode = 1; o.e.gather = gather; o.e.scatter = scatter;
  return launch_op(c, o, c->stream);
}

int pflare_b200_apply(void *handle, const double *b, double *x, int on_device) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->finalized) return fail(6, "apply called before finalize_setup");
  if (c->no_levels < 2) return fail(6, "apply needs >= 2 levels (the reference falls back to PCJACOBI, src/AIR_MG_Setup.F90:1167-1174)");
  Level &L1 = c->L[1];
  const double *bd = b;
  double *xd = x;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(c->io_b, b, (size_t)L1.n * 8, cudaMemcpyHostToDevice, c->stream));
    bd = c->io_b; xd = c->io_x;
  }
  if ((rc = launch_ew_now(c, L1.n, bd, c->bb, L1.d_inv, nullptr))) return rc;  // bb[p] = b[inv[p]]
  if (c->use_graph && c->gexec) {
    CUDA_TRY(cudaGraphLaunch(c->gexec, c->stream));
  } else {
    if ((rc = run_program(c, c->stream, nullptr))) return rc;
  }
  if ((rc = launch_ew_now(c, L1.n, c->xb, xd, L1.d_pos, nullptr))) return rc;  // x[i] = xb[pos[i]]
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(x, c->io_x, (size_t)L1.n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pflare_b200_inv_apply(void *handle, int our_level, int which, const double *x, double *y, int on_device) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->finalized) return fail(6, "inv_apply called before finalize_setup");
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  Level &Lv = c->L[our_level];
  const bool coarse = our_level == c->no_levels;
  const Inv *I; const DevCSR *A; const double *Ad; int n; const int *perm = nullptr, *iperm = nullptr;
  if (which == PFLARE_B200_INV_AFF) {
    I = &Lv.inv_ff; A = coarse ? &Lv.Coarse : &Lv.Aff; Ad = coarse ? Lv.coarse_diag : Lv.aff_diag; n = coarse ? Lv.n : Lv.nf;
  } else if (which == PFLARE_B200_INV_ACC) {
    if (coarse) return fail(2, "no inv_A_cc on the coarsest level");
    I = &Lv.inv_cc; A = &Lv.Acc; Ad = Lv.acc_diag; n = Lv.nc;
    Level &Ln = c->L[our_level + 1];
    if ((rc = ensure_level_perm(c, Ln))) return rc;
    perm = Ln.d_pos; iperm = Ln.d_inv;
  } else {
    return fail(2, "inv_apply: which must be INV_AFF or INV_ACC");
  }
  if (I->kind == 0) return fail(4, "that inverse was not set");
  const double *xd = x; double *yd = y;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(c->io_b, x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    xd = c->io_b; yd = c->io_x;
  }
  // work in bb (input) / xb (output) as scratch: both hold >= n entries
  double *in = c->bb, *out = c->xb;
  if ((rc = launch_ew_now(c, n, xd, in, iperm, nullptr))) return rc;
  std::vector<Op> ops;
  Builder B{c, &ops};
  B.level = our_level;
  if ((rc = B.emit_inv(*I, *A, Ad, n, in, out, 1))) return rc;
  for (const Op &o : ops)
    if ((rc = launch_op(c, o, c->stream))) return rc;
  if ((rc = launch_ew_now(c, n, out, yd, perm, nullptr))) return rc;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(y, c->io_x, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pflare_b200_fc_smooth(void *handle, int our_level, const double *b, double *x, int on_device) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->finalized) return fail(6, "fc_smooth called before finalize_setup");
  if (our_level < 1 || our_level >= c->no_levels) return fail(2, "fc_smooth: our_level %d has no smoother", our_level);
  Level &Lv = c->L[our_level];
  if ((rc = ensure_level_perm(c, Lv))) return rc;
  const double *bd = b; double *xd = x;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(c->io_b, b, (size_t)Lv.n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->io_x, x, (size_t)Lv.n * 8, cudaMemcpyHostToDevice, c->stream));
    bd = c->io_b; xd = c->io_x;
  }
  if ((rc = launch_ew_now(c, Lv.n, bd, c->bb + Lv.off, Lv.d_inv, nullptr))) return rc;
  if ((rc = launch_ew_now(c, Lv.n, xd, c->xb + Lv.off, Lv.d_inv, nullptr))) return rc;
  std::vector<Op> ops;
  Builder B{c, &ops};
  B.level = our_level;
  if (Lv.any_c) B.push_ew(Lv.nc, c->bb + Lv.off + Lv.nf, nullptr, nullptr, 1.0, Lv.bc_save, 1);
  if ((rc = B.emit_fc_richardson(Lv, false))) return rc;
  for (const Op &o : ops)
    if ((rc = launch_op(c, o, c->stream))) return rc;
  if ((rc = launch_ew_now(c, Lv.n, c->xb + Lv.off, xd, Lv.d_pos, nullptr))) return rc;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(x, c->io_x, (size_t)Lv.n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pflare_b200_get_stream(void *handle, void **stream) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  *stream = (void *)c->stream;
  return 0;
}

int pflare_b200_synchronize(void *handle) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

int pflare_b200_get_is(void *handle, int our_level, int which_is, int *out) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  const std::vector<int> &v = which_is == 0 ? c->L[our_level].is_f : c->L[our_level].is_c;
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(int));
  return 0;
}

int pflare_b200_get_garray(void *handle, int our_level, int which, int64_t *out, int *n_ghost) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (our_level < 1 || our_level > c->no_levels || which < 0 || which > 8) return fail(2, "bad selector");
  Level &Lv = c->L[our_level];
  const HostCSR *H = &Lv.H[which];
  if (which == PFLARE_B200_INV_AFF) H = &Lv.inv_ff.h;
  if (which == PFLARE_B200_INV_ACC) H = &Lv.inv_cc.h;
  *n_ghost = H->n_ghost;
  if (out && H->n_ghost) memcpy(out, H->garray.data(), sizeof(int64_t) * (size_t)H->n_ghost);
  return 0;
}

int pflare_b200_get_stats(void *handle, double *stats, int nstats) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int nk = 0;
  for (size_t i = 0; i < c->prog.size(); ++i) {
    const Op &o = c->prog[i];
    v[1] += o.bytes; v[2] += o.nnz;
    v[5] = std::max(v[5], o.bytes);
    const bool in_tail = (int)i >= c->tail_begin && (int)i < c->tail_end;
    const bool empty = (o.kind == OPK_SPMV && (o.s.m == 0 || o.s.nblk == 0)) || (o.kind == OPK_EW && o.e.n == 0);
    if (!in_tail && !empty) ++nk;
  }
  if (c->tail_end > c->tail_begin) ++nk;
  v[0] = nk + 2;  // + permute in / out
  v[1] += 2.0 * 20.0 * (c->no_levels >= 1 ? c->L[1].n : 0);
  v[3] = c->dev_bytes;
  v[6] = c->tail_levels;
  for (int i = 0; i < nstats && i < 8; ++i) stats[i] = v[i];
  return 0;
}

int pflare_b200_profile_apply(void *handle, const double *b_dev, double *x_dev, int max_ops, float *ms, double *bytes,
                              int *level, int *kind, int *n_ops) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->finalized || c->no_levels < 2) return fail(6, "profile_apply needs a finalized hierarchy with >= 2 levels");
  Level &L1 = c->L[1];
  const int n = (int)c->prog.size();
  std::vector<cudaEvent_t> ev((size_t)n + 1);
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  if ((rc = launch_ew_now(c, L1.n, b_dev, c->bb, L1.d_inv, nullptr))) return rc;
  CUDA_TRY(cudaEventRecord(ev[0], c->stream));
  for (int i = 0; i < n; ++i) {
    if ((rc = launch_op(c, c->prog[i], c->stream))) return rc;
    CUDA_TRY(cudaEventRecord(ev[(size_t)i + 1], c->stream));
  }
  if ((rc = launch_ew_now(c, L1.n, c->xb, x_dev, L1.d_pos, nullptr))) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  int cnt = 0;
  for (int i = 0; i < n && cnt < max_ops; ++i) {
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, ev[i], ev[(size_t)i + 1]));
    ms[cnt] = t; bytes[cnt] = c->prog[i].bytes; level[cnt] = c->prog[i].level; kind[cnt] = c->prog[i].tag;
    ++cnt;
  }
  *n_ops = cnt;
  for (auto &e : ev) cudaEventDestroy(e);
  return 0;
}

int pflare_b200_set_option(void *handle, const char *key, double value) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  std::string k(key ? key : "");
  if (k == "graph") c->use_graph = value != 0;
  else if (k == "fuse") c->fuse = value != 0;
  else if (k == "tail_rows") c->tail_rows = (int)value;
  else if (k == "tail_nnz") c->tail_nnz = (int64_t)value;
  else if (k == "kernel") {
    const int v = (int)value;
    if (v < 0 || v >= kNumVariants) return fail(2, "kernel variant must be 0..%d", kNumVariants - 1);
    if (c->finalized && v != 0 && v != c->tile_kernel) return fail(2, "kernel variant %d needs other tile lists: set it before finalize_setup", v);
    c->kernel = v;
    if (!c->finalized && v != 0) c->tile_kernel = v;
  }
  else if (k == "ctas_per_sm") c->ctas_per_sm = (int)value;
  else if (k == "dbg_seq_gather") c->dbg_seq = (int)value;
  else return fail(2, "unknown option '%s'", k.c_str());
  if (c->finalized) {
    if ((rc = build_program(c))) return rc;
    if (c->use_graph && c->no_levels >= 2) { if ((rc = build_graph(c))) return rc; }
  }
  return 0;
}

int pflare_b200_destroy(void **handle) {
  if (!handle || !*handle) return 0;
  Ctx *c = (Ctx *)*handle;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->gexec) cudaGraphExecDestroy(c->gexec);
  if (c->graph) cudaGraphDestroy(c->graph);
  for (void *p : c->allocs) cudaFree(p);
  if (c->stream) cudaStreamDestroy(c->stream);
  c->comm.reset();
  delete c;
  *handle = nullptr;
  return 0;
}

}  // extern "C"
