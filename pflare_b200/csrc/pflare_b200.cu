// pflare_b200.cu -- host side of the B200-native AIRG V-cycle apply behind the C-ABI of
// include/pflare_b200.h.  See DESIGN.md for the layout; reference citations are relative to
// the PFLARE source tree.
//
// What this file does, in order:
//   set_*            : copy the operators the reference's setup built (host CSR in PETSc MPIAIJ
//                      layout, natural numbering: the objects listed in SURVEY.md section 8b).
//   finalize_setup   : (1) nested CF ordering -- level l vector = [F_l | level l+1 vector], so
//                      the identity blocks of R=[Z I] and P=[W;I] (src/Grid_Transfer.F90:329-461,
//                      588-815) become no-ops and every VecISCopy gather/scatter
//                      (src/FC_Smooth.F90:161-417) disappears; (2) R -> Z, P -> W, all operators
//                      relabelled into that ordering with their ghost columns appended, uploaded
//                      once; (3) multi-rank: ghost-exchange plans (dist.h) and agglomeration of
//                      the coarse levels onto rank 0 (a child context running the serial path);
//                      (4) the V-cycle is compiled into a fixed program of fused SpMV ops
//                      (PCMG Kaskade wiring: src/AIR_MG_Setup.F90:967-1156; F/C smoothing:
//                      src/FC_Smooth.F90:421-640; Horner: src/Gmres_Poly.F90:1418-1484; Newton:
//                      src/Gmres_Poly_Newton.F90:763-875; Neumann: src/Neumann_Poly.F90:19-55);
//                      (5) the program is captured in a CUDA graph; small coarse levels run in
//                      one single-CTA kernel.
//   apply            : permute in, launch the graph, permute out.
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pflare_b200.h"
#include "kernels.cuh"
#include "comm.h"
#include "dist.h"

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

// SpMV kernels (option "kernel"): 0 = first-generation smem-staged stream kernel (also the fallback for
// operators with rows longer than a warp tile), 1 = round-1 TMA kernel with CTA tiles (A/B baseline),
// 2 = warp-tile kernel (default).  Geometry of the round-1 kernel's CTA tiles:
constexpr int kR1Threads = 256, kR1Tile = 1024, kR1Stages = 2;
constexpr int kWtWarps = 8;   // warps per CTA of the warp-tile kernel

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

// which index space an operator's ghost columns live in (decides how the OWNER translates a
// requested index into a position of the vector segment the operator reads)
enum { SP_F = 0,      // global F numbering of a level: position = F-local index
       SP_VNEST = 1,  // global natural numbering of a level whose local vector is read in nested order
       SP_VF = 2,     // global natural numbering of a level, only F points allowed, position = F-local index
       SP_NAT = 3 };  // global natural numbering of a level whose local vector is read in NATURAL order (outer Krylov operator)

struct DevPlan {
  GhostPlan plan;
  int space_kind = SP_F, space_level = 0;
  std::vector<int64_t> garray;
  bool global_any = false;  // some rank has ghosts for this operator -> every rank runs its exchange
  int *d_send_idx = nullptr;
  double *d_sendbuf = nullptr;
  double *d_xg = nullptr;
  // peer-memory exchange
  int64_t xg_off = 0;                        // byte offset of d_xg inside this rank's arena
  int *d_send_off = nullptr;                 // [P + 1]
  unsigned long long *d_dst = nullptr;       // [P] where my chunk goes inside peer p's ghost buffer (my address space)
  unsigned dstmask = 0, srcmask = 0;
  int first_inst = -1, last_inst = -1;       // exchange instances of this plan inside the cycle program
};

struct DevCSR {
  int m = 0, n = 0;        // n = local columns (ghost columns are numbered n, n+1, ...)
  int64_t nnz = 0;
  int64_t nx = 0;  // distinct columns referenced (byte model)
  int64_t nnz_model = -1;  // nnz counted by the reference's work model when it differs from nnz
  int *rp = nullptr, *col = nullptr;
  double *val = nullptr;
  int nblk = 0;
  int *blk = nullptr;
  int ntiles = 0;
  int ntiles_int = 0;      // the first ntiles_int tiles reference no ghost column (interior), the rest do (boundary)
  TileDesc *tiles = nullptr;
  // warp-tile storage (kernel 2): one blob per tile + descriptor list, interior tiles first
  unsigned char *blob = nullptr;
  WtDesc *wdesc = nullptr;
  int nwt = 0, nwt_int = 0, kp = 8, fmt = 2;
  bool wt = false;         // this operator runs on the warp-tile kernels
  // rows longer than a warp tile (> 256 nonzeros): compact CSR of just those rows for the stream kernel
  int nlong = 0;
  int *lrp = nullptr, *lcol = nullptr, *lrow = nullptr, *lblk = nullptr;
  double *lval = nullptr;
  DevPlan *xp = nullptr;   // ghost exchange (multi-rank)
  bool is_set = false;
  bool valid() const { return is_set; }
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
  DevCSR Znat;                 // level 1 with the fused entry permutation: Z with NATURAL column indices (gathers from the caller's b)
  DevCSR Pn;                   // full smoothing: P = [W; I] in nested ordering (x_l += P x_{l+1}); Coarse = A_l on every level
  bool w_onepoint = false;
  bool aff_diag_only = false;
  double *aff_diag = nullptr;  // diagonal of A_ff (F-local order): MF_VEC_DIAG and the fused local smooth
  double *acc_diag = nullptr;  // diagonal of A_cc (nested order)
  double *coarse_diag = nullptr;
  double *bc_save = nullptr;   // copy of b_c when the level has C smooths
  int64_t xoff = 0, boff = 0;  // offsets of this level's x / b vector in the nested arrays (equal in the Kaskade cycle)
  std::vector<int> pos;        // natural index -> nested position (relative to off)
  std::vector<int> fpos;       // natural index -> F-local index or -1
  int *d_pos = nullptr, *d_inv = nullptr;
  bool any_c = false;
};

enum { OPK_SPMV = 0, OPK_EW = 1, OPK_XCHG = 2, OPK_GATHER0 = 3, OPK_SCATTER0 = 4, OPK_CHILD = 5, OPK_DENSE = 6, OPK_EPOCH = 7, OPK_ACK = 8, OPK_XWAIT = 9 };
const int kMaxInst = 4096;   // exchange instances per cycle the flag block has room for

// Placeholders for the caller's vectors inside the cycle program (entry / exit permutation fused into the level-1
// ops): patched with the real pointers of each apply; the ops that carry them run outside the CUDA graph.
double *const kUserB = reinterpret_cast<double *>((uintptr_t)8);
double *const kUserX = reinterpret_cast<double *>((uintptr_t)16);

struct Op {
  bool user = false;            // references kUserB / kUserX
  int kind = OPK_SPMV;
  SpmvOp s{};
  EwOp e{};
  DevPlan *xp = nullptr;        // OPK_XCHG / OPK_ACK: which plan; xsrc = the vector segment being exchanged
  const double *xsrc = nullptr;
  int inst = -1, ack_inst = -1, ack_delta = 0;   // peer-memory exchange instance (and the one whose acks it waits for)
  bool async = false;           // OPK_XCHG: run on the side stream, joined by the following OPK_XWAIT
  bool push_here = false;       // p2p=2: the consuming SpMV kernel of this rank pushes the entries itself (no launch for the exchange)
  int level = 0;
  int tag = 0;  // 1 restrict, 2 coarse, 3 A_fc(+W), 4 A_ff residual, 5 inverse, 6 elementwise, 7 fused local smooth, 8 A_cf, 9 A_cc, 10 exchange, 11 dense tail
  double bytes = 0, nnz = 0;
};

struct Cluster;

struct Ctx {
  int rank = 0, nranks = 1, device = 0, no_levels = 0;
  std::vector<Level> L;  // 1-based
  bool finalized = false, planned = false;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::vector<void *> allocs;
  double dev_bytes = 0;
  // nested vectors + scratch
  double *xb = nullptr, *bb = nullptr;
  double *scr[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *io_b = nullptr, *io_x = nullptr;  // staging for host-pointer calls
  int maxn = 0;
  // program
  std::vector<Op> prog;
  int tail_levels = 0;                  // levels collapsed into the dense tail
  // dense collapsed tail: the sub-cycle of levels >= dense_level as ONE n x n matrix (x_l = T b_l)
  std::vector<Op> dense_prog;           // the ops it replaces (run once per unit vector at setup)
  int dense_level = 0, dense_n = 0;
  double *dense_T = nullptr;
  size_t dense_cap = 0;                 // entries allocated for dense_T
  bool dense_built = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int graph_kernels = 0;
  // options
  int use_graph = 1, fuse = 1;
  int coarse_its = 1;   // -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N: Richardson iterations of the coarse solve (KSP_NORM_NONE: exactly N)
  int fuse_epi = 1;      // compile-time specialised epilogue classes (0: every op runs the generic epilogue)
  int fuse_perm = 0;     // (measured slower: natural-order accesses interleave F and C points -> half-used sectors) natural <-> nested permutation of b / x fused into the level-1 ops (serial Kaskade contexts)
  bool io_fused = false; // decided at finalize_setup
  int n_head = 0, tail_begin = -1;   // program ops [0, n_head) and [tail_begin, end) reference the caller's vectors: launched per apply, outside the graph
  int full_smooth = 0;   // -pc_air_full_smoothing_up_and_down: PCMG multiplicative V(1,1), inv_A_ff(l) ~ A_l^-1 on all unknowns
  int dense_rows = 4096; // levels with <= this many rows are collapsed into one dense matrix (0 = off)
  int kernel = 2;        // 0: smem-staged stream kernel, 1: round-1 TMA kernel (CTA tiles), 2: warp-tile kernel
  int wt_format = 0;     // kernel 2 operator storage: 0 = by mean row length (fmt_split), 1 = chunk format, 2 = row-aligned lanes
  double fmt_split = 8.0;   // measured on the 4096^2 cycle: 3 -> 2.82 ms, 4.5 -> 2.65 ms, 8 -> 2.62 ms
  int engine = 1;        // kernel 2, row-aligned storage: 1 = direct engine (spmv_sv_kernel), 0 = TMA-ring engine (spmv_wt_kernel, A/B) // 0 = TMA-ring engine (spmv_wt_kernel), 1 = direct engine (spmv_sv_kernel), 2 = thin-warp engine (spmv_thin_kernel)
  int sv_pf = 0;         // (retired A/B switch, always 0: see set_option)
  int wt_stages = 2;     // ring depth of the warp-tile kernel (2 or 3 tiles per warp; 2 leaves more of the SM's L1 to the gathers)
  int ctas_per_sm = 0;   // 0 = from the occupancy calculator
  int max_ctas = 0;      // > 0: cap on the persistent grid (tests: forces many tiles per CTA / warp)
  int pdl = 1;           // programmatic dependent launch between the kernels of the cycle
  int64_t agg_rows = 262144;  // levels with <= this many GLOBAL rows are agglomerated onto rank 0 (multi-rank)
  int num_sms = 148;
  // multi-rank
  std::unique_ptr<Comm> comm;            // NCCL (one process per GPU)
  std::unique_ptr<HostComm> hostcomm;    // setup-time host communicator (callbacks / in-process / NCCL-staged)
  Cluster *cluster = nullptr;            // in-process group of ranks (lockstep execution on one stream)
  std::deque<DevPlan> plans;
  std::vector<Ranges> rangeV, rangeF;    // per level (1-based): ownership of natural rows / of F points
  // outer Krylov method (pflare_b200_ksp_*): the system matrix in natural ordering + work vectors
  HostCSR ksp_A;
  DevCSR ksp_Ad;
  double *ksp_buf = nullptr; size_t ksp_cap = 0;      // basis V (restart + 1 vectors) + w, z, u
  double *ksp_partial = nullptr, *ksp_dots = nullptr; // multi-dot partials / results
  double *ksp_hdots = nullptr;                        // pinned host copy of the dots
  int l_agg = 0;                         // first agglomerated level (no_levels + 1: none)
  int n_dist = 0;                        // number of distributed levels with an F/C structure (l < min(l_agg, NL))
  std::unique_ptr<Ctx> child;            // rank 0: serial hierarchy of the agglomerated levels
  double *child_b = nullptr, *child_x = nullptr;  // rank 0: global natural vectors of level l_agg
  double ghost_bytes = 0; int xchg_groups = 0;
  // peer-memory ghost exchange (option p2p): one arena per rank = [flag block | ghost buffers], mapped by every peer
  int overlap = 1;       // NCCL exchange on a side stream, overlapped with the interior tiles of the SpMV
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int p2p = 2;   // 0: NCCL send/recv; 1: peer-memory push kernel + flags + acks (CUDA IPC); 2: push fused into the consuming SpMV kernel,
                 //    one ghost buffer per exchange instance, one started[] flag per cycle instead of acks
  bool p2p_ready = false;
  bool fused_push() const { return p2p == 2 && p2p_ready; }
  std::vector<int> inst_plan;            // p2p=2: plan of every exchange instance of the cycle program
  std::vector<int> inst_plan_built;      //        ... the layout arena2 was built for
  char *arena2 = nullptr;                //        ghost buffers of the instances (mapped by every peer)
  std::vector<int64_t> inst_off;         //        byte offset of instance i inside arena2
  std::vector<void *> peer_arena2;
  std::vector<bool> peer_ipc2;
  XPush *d_xpush = nullptr;              //        [n_inst] descriptors
  unsigned all_dst = 0, all_src = 0;     //        union of the plans' masks
  char *arena = nullptr; size_t arena_bytes = 0;
  std::vector<void *> peer_arena;        // [P] peers' arenas in my address space (IPC-mapped or same process)
  std::vector<bool> peer_ipc;
  unsigned long long *d_peer_flags = nullptr;  // [P]
  unsigned *d_done = nullptr;            // [kMaxInst] CTA counters of the push kernels
  int n_inst = 0;
  unsigned *flags() const { return reinterpret_cast<unsigned *>(arena); }
};

template <class T>
int dev_alloc(Ctx *c, T **p, size_t n) {
  void *q = nullptr;
  size_t bytes = std::max<size_t>(n, 1) * sizeof(T) + 64;   // slack: bulk copies round their extent up to 16 bytes
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

// Row-block partition: consecutive rows with <= tile nnz and <= maxrows rows per block;
// a row longer than the tile forms its own block.
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

// Upload one device-ordered operator.  h.n = local columns; ghost columns (if any) are numbered
// h.n .. h.n + h.n_ghost - 1 inside h.ja and described by h.garray (global ids in `space`).
int upload_csr(Ctx *c, const HostCSR &h, DevCSR *d, int space_kind = SP_F, int space_level = 0, bool wfirst = false) {
  if (h.nnz() >= (int64_t)2147483647) return fail(3, "operator has >= 2^31 nonzeros (32-bit PetscInt only)");
  d->m = h.m;
  d->n = h.n;
  d->nnz = h.nnz();
  d->is_set = true;
  std::vector<unsigned char> seen((size_t)std::max(h.n + h.n_ghost, 1), 0);
  int64_t nx = 0;
  for (int cidx : h.ja)
    if (!seen[cidx]) { seen[cidx] = 1; ++nx; }
  d->nx = nx;
  if (c->nranks > 1) {
    c->plans.emplace_back();
    DevPlan &P = c->plans.back();
    P.plan.init(c->nranks);
    P.space_kind = space_kind; P.space_level = space_level;
    P.garray = h.garray;
    P.plan.n_ghost = h.n_ghost;
    d->xp = &P;
  } else if (h.n_ghost > 0) {
    return fail(2, "operator has ghost columns but the context has a single rank");
  }
  if (c->device < 0) {  // host-only planning context (the warp-tile layout is still built when asked for: CPU-side checks)
    if (getenv("PFLARE_B200_PLAN_BUILDS_TILES") && c->kernel == 2) {
      WtHost W; build_wt(h.m, h.n, h.ia.data(), h.ja.data(), h.a.data(), wfirst, &W);
      WcHost V; build_wc(h.m, h.n, h.ia.data(), h.ja.data(), h.a.data(), &V);
    }
    return 0;
  }
  int rc;
  d->wt = false;
  if (c->kernel == 2) {
    // short rows: chunk format (no padding inside rows); longer rows: row-aligned lanes (coalesced gathers)
    const double mean_len = h.m > 0 ? (double)h.nnz() / h.m : 0.0;
    const bool chunk = c->wt_format == 1 || (c->wt_format == 0 && mean_len < c->fmt_split);
    std::vector<int> long_rows;
    if (chunk) {
      WcHost W;
      build_wc(h.m, h.n, h.ia.data(), h.ja.data(), h.a.data(), &W);
      d->wt = true; d->kp = W.rq; d->fmt = 1;
      d->nwt = (int)W.desc.size(); d->nwt_int = W.n_int;
      if ((rc = dev_upload(c, &d->blob, W.blob))) return rc;
      if ((rc = dev_upload(c, &d->wdesc, W.desc))) return rc;
      long_rows.swap(W.long_rows);
    } else {
      WtHost W;
      build_wt(h.m, h.n, h.ia.data(), h.ja.data(), h.a.data(), wfirst, &W);
      d->wt = true; d->kp = W.kp; d->fmt = 2;
      d->nwt = (int)W.desc.size(); d->nwt_int = W.n_int;
      if ((rc = dev_upload(c, &d->blob, W.blob))) return rc;
      if ((rc = dev_upload(c, &d->wdesc, W.desc))) return rc;
      long_rows.swap(W.long_rows);
    }
    d->ntiles = d->nwt; d->ntiles_int = d->nwt_int;
    d->nlong = (int)long_rows.size();
    if (d->nlong > 0) {
      // the few rows longer than a tile: compact CSR (entries in the operator's order, W entry last) + row map + one block per row
      std::vector<int> lrp(1, 0), lcol, lblk((size_t)d->nlong + 1);
      std::vector<double> lval;
      for (int k = 0; k < d->nlong; ++k) {
        const int r = long_rows[(size_t)k];
        lcol.insert(lcol.end(), h.ja.begin() + h.ia[r], h.ja.begin() + h.ia[r + 1]);
        lval.insert(lval.end(), h.a.begin() + h.ia[r], h.a.begin() + h.ia[r + 1]);
        lrp.push_back((int)lcol.size());
        lblk[(size_t)k] = k;
      }
      lblk[(size_t)d->nlong] = d->nlong;
      if ((rc = dev_upload_padded(c, &d->lrp, lrp, 8))) return rc;
      if ((rc = dev_upload_padded(c, &d->lcol, lcol, 8))) return rc;
      if ((rc = dev_upload_padded(c, &d->lval, lval, 8))) return rc;
      if ((rc = dev_upload(c, &d->lrow, long_rows))) return rc;
      if ((rc = dev_upload(c, &d->lblk, lblk))) return rc;
    }
    return 0;
  }
  // CSR stream (kernel 0 / 1, or a row longer than a warp tile)
  if ((rc = dev_upload_padded(c, &d->rp, h.ia, 8))) return rc;
  if ((rc = dev_upload_padded(c, &d->col, h.ja, 8))) return rc;
  if ((rc = dev_upload_padded(c, &d->val, h.a, 8))) return rc;
  std::vector<int> blk = make_blocks(h.ia, h.m);
  d->nblk = (int)blk.size() - 1;
  if ((rc = dev_upload(c, &d->blk, blk))) return rc;
  std::vector<int> tb = make_blocks(h.ia, h.m, kR1Tile, kR1Threads);
  std::vector<TileDesc> tiles(tb.size() - 1);
  for (size_t t = 0; t + 1 < tb.size(); ++t) tiles[t] = TileDesc{tb[t], tb[t + 1] - tb[t], h.ia[tb[t]], h.ia[tb[t + 1]] - h.ia[tb[t]]};
  d->ntiles = (int)tiles.size();
  d->ntiles_int = d->ntiles;
  if (h.n_ghost > 0) {
    // interior tiles first: they can be multiplied while the ghost exchange is still in flight
    std::vector<TileDesc> inner, bnd;
    for (const TileDesc &t : tiles) {
      bool ghost = false;
      for (int k = t.s; k < t.s + t.n && !ghost; ++k) ghost = h.ja[k] >= h.n;
      (ghost ? bnd : inner).push_back(t);
    }
    d->ntiles_int = (int)inner.size();
    tiles = inner;
    tiles.insert(tiles.end(), bnd.begin(), bnd.end());
  }
  if ((rc = dev_upload(c, &d->tiles, tiles))) return rc;
  return 0;
}

// Device-ordered copy of an MPIAIJ operator: rows permuted by rowpos (new row index), diag-block
// columns relabelled by colpos into [0, nloc), off-diag (ghost) columns appended as nloc + g.
// Per row: diag entries first, then ghost entries (the order MatMult_MPIAIJ sums them), each sorted.
HostCSR remap(const HostCSR &A, const int *rowpos, const int *colpos, int nloc) {
  HostCSR B;
  B.set = true;
  B.m = A.m;
  B.n = nloc;
  B.n_ghost = A.n_ghost;
  B.garray = A.garray;
  const bool og = A.n_ghost > 0;
  B.ia.assign((size_t)A.m + 1, 0);
  for (int i = 0; i < A.m; ++i) {
    int r = rowpos ? rowpos[i] : i;
    B.ia[(size_t)r + 1] = A.ia[i + 1] - A.ia[i] + (og ? A.oia[i + 1] - A.oia[i] : 0);
  }
  for (int i = 0; i < A.m; ++i) B.ia[i + 1] += B.ia[i];
  B.ja.resize((size_t)B.ia[A.m]);
  B.a.resize((size_t)B.ia[A.m]);
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
    const int len = p1 - p0;
    if (!sorted) {
      std::vector<std::pair<int, double>> tmp((size_t)len);
      for (int k = 0; k < len; ++k) tmp[k] = {B.ja[o + k], B.a[o + k]};
      std::sort(tmp.begin(), tmp.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) { return x.first < y.first; });
      for (int k = 0; k < len; ++k) { B.ja[o + k] = tmp[k].first; B.a[o + k] = tmp[k].second; }
    }
    if (og) {
      int q = o + len;
      for (int p = A.oia[i]; p < A.oia[i + 1]; ++p, ++q) { B.ja[q] = nloc + A.oja[p]; B.a[q] = A.oa[p]; }
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
  if (A.n_ghost > 0 && !A.oja.empty()) return false;
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
bool op_is_empty(const Op &o);
struct Builder {
  Ctx *c;
  std::vector<Op> *out;
  int level = 0;
  bool use_p2p = false;   // only the cycle program owns exchange instances; ad-hoc op lists go through NCCL / copies
  bool io = false;        // cycle program of a context with fused entry / exit permutation: level-1 ops read b / write x in natural order

  SpmvOp base(const DevCSR &A, const double *x) {
    SpmvOp s{};
    s.rp = A.rp; s.col = A.col; s.val = A.val; s.m = A.m; s.nblk = A.nblk; s.blk = A.blk;
    s.tiles = A.tiles; s.ntiles = A.ntiles;
    s.blob = A.blob; s.wdesc = A.wdesc; s.nwt = A.wt ? A.nwt : 0; s.kp = A.kp; s.fmt = A.fmt;
    s.x = x; s.nloc = A.n; s.beta = 1.0;
    s.xg = A.xp ? A.xp->d_xg : nullptr;
    return s;
  }
  // epilogue class of an op (kernels.cuh: the compile-time specialisations of the warp-tile kernel)
  static int classify(const SpmvOp &s) {
    if (s.D || s.neumann || s.out2) return EPI_GENERIC;
    if (s.wlast) {
      if (!s.aux || !s.wout || s.acc_mode) return EPI_GENERIC;
      if (s.fd_its > 0) return s.out_mode == 0 ? EPI_AFCW_LOCAL : EPI_GENERIC;
      return (s.out_mode == 1 && !s.wout_idx) ? EPI_AFCW : EPI_GENERIC;
    }
    if (s.wout) return EPI_GENERIC;
    if (s.acc_mode) return (s.aux && !s.aux_idx && s.out_mode == 1) ? EPI_AXPBY_ACC : EPI_GENERIC;
    if (!s.aux) {
      if (s.beta != 1.0) return EPI_GENERIC;
      return s.out_mode == 2 ? EPI_ADD : (s.out_mode == 1 ? EPI_SET : EPI_GENERIC);
    }
    return s.out_mode == 1 ? EPI_AXPBY : EPI_GENERIC;
  }
  void push_spmv(const SpmvOp &s, const DevCSR &A, int tag, int aux_reads, int w, double extra_bytes = 0) {
    const bool xchg = A.xp && A.xp->global_any;
    const bool p2p = xchg && c->p2p_ready && use_p2p;
    int inst = -1;
    if (xchg) {  // ghost scatter of MatMult_MPIAIJ: pack + point-to-point exchange of s.x
      Op x;
      x.kind = OPK_XCHG; x.xp = A.xp; x.xsrc = s.x; x.level = level; x.tag = 10;
      x.bytes = 8.0 * (A.xp->plan.n_ghost + A.xp->plan.n_send());
      if (p2p) {
        inst = c->n_inst++;
        x.inst = inst;
        x.ack_inst = A.xp->last_inst; x.ack_delta = 0;        // same cycle; the first instance is fixed up at the end
        if (A.xp->first_inst < 0) A.xp->first_inst = inst;
        A.xp->last_inst = inst;
      }
      out->push_back(x);
    }
    Op o;
    o.kind = OPK_SPMV; o.s = s; o.level = level; o.tag = tag;
    o.s.epi = c->fuse_epi ? classify(s) : EPI_GENERIC;
    o.bytes = spmv_bytes(A, aux_reads, w) + extra_bytes;
    o.nnz = (double)(A.nnz_model >= 0 ? A.nnz_model : A.nnz);
    if (xchg && !p2p && use_p2p && c->overlap && (A.wt || c->kernel == 1)) {
      // split-phase MatMult_MPIAIJ: exchange on the side stream || interior tiles; then the boundary tiles
      out->back().async = true;
      const double fi = A.ntiles > 0 ? (double)A.ntiles_int / A.ntiles : 1.0;
      Op oi = o, ob = o;
      oi.s.ntiles = A.ntiles_int; oi.bytes = o.bytes * fi; oi.nnz = o.nnz * fi;
      ob.s.ntiles = A.ntiles - A.ntiles_int; ob.bytes = o.bytes - oi.bytes; ob.nnz = o.nnz - oi.nnz;
      if (A.wt) { oi.s.nwt = A.nwt_int; ob.s.wdesc = A.wdesc + A.nwt_int; ob.s.nwt = A.nwt - A.nwt_int; }
      else ob.s.tiles = A.tiles + A.ntiles_int;
      Op wt; wt.kind = OPK_XWAIT; wt.level = level; wt.tag = 10;
      out->push_back(oi); out->push_back(wt); out->push_back(ob);
      push_long_rows(o, A);
      return;
    }
    if (p2p) {
      o.s.gw_ready = c->flags() + 64 + (size_t)inst * 32;
      o.s.gw_epoch = c->flags();
      o.s.gw_srcmask = A.xp->srcmask;
    }
    if (p2p && c->p2p == 2) {
      // fused exchange: xg / xpush are patched by setup_fused_push() once the instance buffers exist
      o.inst = inst;
      o.s.gw_first = A.wt ? A.nwt_int : 0;
      o.push_here = !c->cluster && !op_is_empty(o);
      (*out)[out->size() - 1].push_here = o.push_here;   // the OPK_XCHG pushed above
      int pid = 0;
      for (const DevPlan &D : c->plans) { if (&D == A.xp) break; ++pid; }
      c->inst_plan.push_back(pid);
      out->push_back(o);
      push_long_rows(o, A);
      return;
    }
    out->push_back(o);
    push_long_rows(o, A);
    if (p2p) {   // tell the producers that this rank is done with the ghosts of this instance
      Op a;
      a.kind = OPK_ACK; a.xp = A.xp; a.inst = inst; a.level = level; a.tag = 10;
      out->push_back(a);
    }
  }
  // The rows longer than a warp tile are not in the operator's tiles: the same op (same epilogue, disjoint rows) runs once
  // more over a compact CSR of just those rows on the stream kernel.  Multi-rank programs carry the (possibly empty) op
  // on every rank so that the ranks' op lists keep the same shape.
  void push_long_rows(const Op &main, const DevCSR &A) {
    if (!A.wt || (A.nlong == 0 && c->nranks == 1)) return;
    Op l = main;
    l.s.blob = nullptr; l.s.wdesc = nullptr; l.s.nwt = 0; l.s.tiles = nullptr;
    l.s.rp = A.lrp; l.s.col = A.lcol; l.s.val = A.lval; l.s.rowmap = A.lrow;
    l.s.m = A.nlong; l.s.nblk = A.nlong; l.s.blk = A.lblk; l.s.ntiles = A.nlong;
    l.s.epi = EPI_GENERIC;
    if (c->p2p != 2) l.s.gw_ready = nullptr;   // p2p=1: the main op already waited for (and acknowledges) the ghosts
    l.s.xpush = nullptr; l.s.gw_first = 0; l.push_here = false;   // p2p=2: waits itself (own buffer per instance, no ack), never pushes
    l.bytes = 0; l.nnz = 0;
    out->push_back(l);
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
      out->back().nnz = (double)n;   // the reference's work model counts a MATDIAGONAL MatMult as n nonzeros
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
    double *xb = c->xb + Lv.xoff, *bb = c->bb + Lv.boff;
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
      const bool io1 = io && level == 1;
      if (io1) { s.aux = kUserB; s.aux_idx = Lv.d_inv; }   // b_f read straight from the caller's b (natural ordering)
      double extra = 0;
      if (fuse_w) { s.wlast = 1; s.wout = xf; extra += 8.0 * Lv.nf; }
      if (local) {
        s.fd_a = Lv.aff_diag; s.fd_m = Lv.inv_ff.ddiag; s.fd_its = its;
        extra += 16.0 * Lv.nf;
        if (io1) { s.wout_idx = Lv.d_inv; s.xnat = kUserX; }   // x_f written straight into the caller's x
        push_spmv(s, A, 7, 1, 0, extra);
        out->back().nnz += 2.0 * its * Lv.nf;   // the its x (diagonal A_ff + diagonal M_ff) products folded into this op
        out->back().user = io1;
        return 0;
      }
      s.out = S[0]; s.out_mode = 1;
      push_spmv(s, A, 3, 1, 1, extra);
      out->back().user = io1;
    }
    for (int f = 0; f < its; ++f) {
      SpmvOp s = base(Lv.Aff, xf);  // r = rhs - A_ff x_f     (src/FC_Smooth.F90:544-549)
      s.aux = S[0]; s.alpha = 1.0; s.beta = -1.0;
      s.out = S[1]; s.out_mode = 1;
      push_spmv(s, Lv.Aff, 4, 1, 1);
      int rc = emit_inv(Lv.inv_ff, Lv.Aff, Lv.aff_diag, Lv.nf, S[1], xf, 2);  // x_f += M_ff r  (:552-557)
      if (rc) return rc;
    }
    if (io && level == 1) {
      // exit permutation of the F points: folded into the last update of x_f when that is `x_f += M r` with an
      // assembled M, otherwise one scatter over the F points
      Op &last = out->back();
      if (last.kind == OPK_SPMV && last.s.epi == EPI_ADD && last.s.out == xf) {
        last.s.wout_idx = Lv.d_inv; last.s.xnat = kUserX; last.user = true;
      } else {
        push_ew(Lv.nf, xf, nullptr, nullptr, 1.0, kUserX, 1, nullptr, Lv.d_inv);
        out->back().user = true;
      }
    }
    return 0;
  }

  int emit_c_smooths(Level &Lv, bool prolong_pending, int its) {
    double *xb = c->xb + Lv.xoff;
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
    if (pending) emit_prolong(Lv, c->xb + Lv.xoff, c->xb + Lv.xoff + Lv.nf);
    return 0;
  }
};

// ------------------------------------------------------------------ in-process rank group
// Several logical ranks driven by ONE process on ONE stream, executed in lockstep (op i of every
// rank, then op i+1 ...).  Ghost exchanges become device-to-device copies between the ranks'
// buffers.  It is the single-process way to run a partitioned hierarchy (and the way the
// multi-rank path is exercised on a one-GPU box); one process per GPU uses NCCL instead.
struct SharedComm;
struct Cluster {
  std::vector<std::unique_ptr<Ctx>> ranks;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  bool finalized = false;
  std::shared_ptr<SharedComm> shared;
};

// setup-time host communicator between the threads of a Cluster (one thread per rank)
struct SharedComm {
  int P;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, generation = 0;
  std::vector<const int64_t *> a_send;
  struct V { const char *sbuf; const int64_t *scnt, *sdsp; };
  std::vector<V> v_send;
  bool aborted = false;   // a rank failed: every barrier returns at once (no rank is left waiting for it)
  explicit SharedComm(int p) : P(p), a_send(p), v_send(p) {}
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return false;
    const int gen = generation;
    if (++arrived == P) { arrived = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != generation || aborted; });
    return !aborted;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};
struct SharedRankComm : HostComm {
  std::shared_ptr<SharedComm> sc;
  int r;
  SharedRankComm(std::shared_ptr<SharedComm> s, int rank) : sc(s), r(rank) {}
  int rank() const override { return r; }
  int size() const override { return sc->P; }
  int alltoall(const int64_t *send, int64_t *recv, std::string *err) override {
    sc->a_send[r] = send;
    if (!sc->barrier()) { if (err) *err = "another rank of the group failed"; return 1; }
    for (int p = 0; p < sc->P; ++p) recv[p] = sc->a_send[p][r];
    if (!sc->barrier()) { if (err) *err = "another rank of the group failed"; return 1; }
    return 0;
  }
  int alltoallv(const char *sbuf, const int64_t *scnt, const int64_t *sdsp, char *rbuf, const int64_t *rcnt, const int64_t *rdsp,
                std::string *err) override {
    sc->v_send[r] = {sbuf, scnt, sdsp};
    if (!sc->barrier()) { if (err) *err = "another rank of the group failed"; return 1; }
    int bad = 0;
    for (int p = 0; p < sc->P; ++p) {
      const SharedComm::V &v = sc->v_send[p];
      if (rcnt[p] != v.scnt[r]) { bad = 1; continue; }
      if (rcnt[p]) memcpy(rbuf + rdsp[p], v.sbuf + v.sdsp[r], (size_t)rcnt[p]);
    }
    if (bad) { sc->abort(); if (err) *err = "receive counts disagree with the senders'"; return 1; }
    if (!sc->barrier()) { if (err) *err = "another rank of the group failed"; return 1; }
    return 0;
  }
};

// setup-time host communicator over NCCL (bytes staged through device buffers); used when one
// process per GPU runs without user-supplied MPI-style callbacks
struct NcclHostComm : HostComm {
  Comm *comm;
  cudaStream_t st;
  NcclHostComm(Comm *c, cudaStream_t s) : comm(c), st(s) {}
  int rank() const override { return comm->rank(); }
  int size() const override { return comm->size(); }
  int alltoall(const int64_t *send, int64_t *recv, std::string *err) override {
    const int P = size();
    std::vector<int64_t> cnt(P, 8), dsp(P);
    for (int p = 0; p < P; ++p) dsp[p] = 8 * p;
    return alltoallv((const char *)send, cnt.data(), dsp.data(), (char *)recv, cnt.data(), dsp.data(), err);
  }
  int alltoallv(const char *sbuf, const int64_t *scnt, const int64_t *sdsp, char *rbuf, const int64_t *rcnt, const int64_t *rdsp,
                std::string *err) override {
    const int P = size();
    int64_t st_ = 0, rt_ = 0;
    for (int p = 0; p < P; ++p) { st_ = std::max(st_, sdsp[p] + scnt[p]); rt_ = std::max(rt_, rdsp[p] + rcnt[p]); }
    char *ds = nullptr, *dr = nullptr;
    auto cu = [&](cudaError_t e, const char *what) {
      if (e != cudaSuccess) { if (err) *err = std::string(what) + ": " + cudaGetErrorString(e); return 1; }
      return 0;
    };
    if (cu(cudaMalloc(&ds, (size_t)std::max<int64_t>(st_, 1)), "cudaMalloc")) return 1;
    if (cu(cudaMalloc(&dr, (size_t)std::max<int64_t>(rt_, 1)), "cudaMalloc")) return 1;
    if (st_ && cu(cudaMemcpyAsync(ds, sbuf, (size_t)st_, cudaMemcpyHostToDevice, st), "H2D")) return 1;
    bool ok = comm->group_start(err);
    for (int p = 0; p < P && ok; ++p) {
      if (p == rank()) continue;
      if (scnt[p]) ok = comm->send(ds + sdsp[p], (size_t)scnt[p], 1, p, st, err);
      if (ok && rcnt[p]) ok = comm->recv(dr + rdsp[p], (size_t)rcnt[p], 1, p, st, err);
    }
    ok = ok && comm->group_end(err);
    if (!ok) return 1;
    const int me = rank();
    if (scnt[me] && cu(cudaMemcpyAsync(dr + rdsp[me], ds + sdsp[me], (size_t)scnt[me], cudaMemcpyDeviceToDevice, st), "D2D")) return 1;
    if (rt_ && cu(cudaMemcpyAsync(rbuf, dr, (size_t)rt_, cudaMemcpyDeviceToHost, st), "D2H")) return 1;
    if (cu(cudaStreamSynchronize(st), "sync")) return 1;
    cudaFree(ds); cudaFree(dr);
    return 0;
  }
};

// ------------------------------------------------------------------ launching
// Every kernel of the cycle goes through here: with option pdl=1 the launch carries the
// programmatic-stream-serialization attribute (PDL), so a kernel's prologue (mbarrier init, tile
// descriptors, the first TMA bulk copies of constant matrix data) overlaps the predecessor's drain.
template <class K, class... Args>
cudaError_t launch_k(bool pdl, K kern, int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
// resident CTAs per SM of one kernel instantiation (attribute + occupancy calculator, once per process)
template <class K>
int kernel_per_sm(K kern, int threads, size_t smem, int *per_sm) {
  if (*per_sm == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
    *per_sm = std::max(nb, 1);
  }
  return 0;
}

int launch_r1(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  auto kern = spmv_tma_kernel<kR1Threads, kR1Tile, kR1Stages>;
  const size_t smem = sizeof(TmaStage<kR1Tile, kR1Threads>) * kR1Stages;
  static int per_sm = 0;
  int rc = kernel_per_sm(kern, kR1Threads, smem, &per_sm);
  if (rc || dry) return rc;
  const int want = c->ctas_per_sm > 0 ? std::min(c->ctas_per_sm, per_sm) : per_sm;
  int grid = std::min(s.ntiles, c->num_sms * want);   // persistent: a multiple of the SM count
  if (c->max_ctas > 0) grid = std::min(grid, c->max_ctas);
  CUDA_TRY(launch_k(c->pdl != 0, kern, grid, kR1Threads, smem, st, s));
  return 0;
}

// warp-tile kernel: one instantiation per (epilogue class, rows per lane, ghost columns, ring depth)
template <int EPI, int KP, bool GH, int ST>
int launch_wt_inst(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  auto kern = spmv_wt_kernel<EPI, KP, GH, kWtWarps, ST>;
  const size_t smem = (size_t)kWtWarps * (ST * kWtStageBytes + (kWtSlots / KP) * 32 * 8 * (EpiT<EPI>::kXw ? 2 : 1));
  static int per_sm = 0;
  int rc = kernel_per_sm(kern, kWtWarps * 32, smem, &per_sm);
  if (rc || dry) return rc;
  const int want = c->ctas_per_sm > 0 ? std::min(c->ctas_per_sm, per_sm) : per_sm;
  int grid = std::min((s.nwt + kWtWarps - 1) / kWtWarps, c->num_sms * want);   // persistent
  if (c->max_ctas > 0) grid = std::min(grid, c->max_ctas);
  CUDA_TRY(launch_k(c->pdl != 0, kern, grid, kWtWarps * 32, smem, st, s));
  return 0;
}
// direct engine (no shared memory): one instantiation per (epilogue class, slots per lane, ghost columns)
template <int EPI, int KP, bool GH, int PF>
int launch_sv_inst(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  auto kern = spmv_sv_kernel<EPI, KP, GH, PF>;
  static int per_sm = 0;
  int rc = kernel_per_sm(kern, 256, 0, &per_sm);
  if (rc || dry) return rc;
  const int want = c->ctas_per_sm > 0 ? std::min(c->ctas_per_sm, per_sm) : per_sm;
  int grid = std::min((s.nwt + 7) / 8, c->num_sms * want);   // persistent, 8 warps per CTA
  if (c->max_ctas > 0) grid = std::min(grid, c->max_ctas);
  CUDA_TRY(launch_k(c->pdl != 0, kern, grid, 256, 0, st, s));
  return 0;
}
template <int EPI, int KP>
int launch_wt_kp(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  const bool gh = s.xg != nullptr;
  if (c->engine == 1) return gh ? launch_sv_inst<EPI, KP, true, 0>(c, s, st, dry) : launch_sv_inst<EPI, KP, false, 0>(c, s, st, dry);
  return gh ? launch_wt_inst<EPI, KP, true, 2>(c, s, st, dry) : launch_wt_inst<EPI, KP, false, 2>(c, s, st, dry);
}
// chunk-format engine: one instantiation per (epilogue class, rows per lane, ghost columns); TMA ring of 2 tiles
template <int EPI, int RQ, bool GH>
int launch_wc_inst(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  constexpr int ST = 2;
  auto kern = spmv_wc_kernel<EPI, RQ, GH, kWtWarps, ST>;
  const size_t smem = (size_t)kWtWarps * (ST * kWcStageBytes + (EpiT<EPI>::kXw ? RQ * 32 * 8 : 0));
  static int per_sm = 0;
  int rc = kernel_per_sm(kern, kWtWarps * 32, smem, &per_sm);
  if (rc || dry) return rc;
  const int want = c->ctas_per_sm > 0 ? std::min(c->ctas_per_sm, per_sm) : per_sm;
  int grid = std::min((s.nwt + kWtWarps - 1) / kWtWarps, c->num_sms * want);   // persistent
  if (c->max_ctas > 0) grid = std::min(grid, c->max_ctas);
  CUDA_TRY(launch_k(c->pdl != 0, kern, grid, kWtWarps * 32, smem, st, s));
  return 0;
}
template <int EPI>
int launch_wc_epi(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  const bool gh = s.xg != nullptr;
  switch (s.kp) {
    case 1: return gh ? launch_wc_inst<EPI, 1, true>(c, s, st, dry) : launch_wc_inst<EPI, 1, false>(c, s, st, dry);
    case 2: return gh ? launch_wc_inst<EPI, 2, true>(c, s, st, dry) : launch_wc_inst<EPI, 2, false>(c, s, st, dry);
    case 4: return gh ? launch_wc_inst<EPI, 4, true>(c, s, st, dry) : launch_wc_inst<EPI, 4, false>(c, s, st, dry);
    case 8: return gh ? launch_wc_inst<EPI, 8, true>(c, s, st, dry) : launch_wc_inst<EPI, 8, false>(c, s, st, dry);
  }
  return fail(7, "internal: rows per lane %d", s.kp);
}
template <int EPI>
int launch_wt_epi(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  if (s.fmt == 1) return launch_wc_epi<EPI>(c, s, st, dry);
  switch (s.kp) {
    case 1: return launch_wt_kp<EPI, 1>(c, s, st, dry);
    case 2: return launch_wt_kp<EPI, 2>(c, s, st, dry);
    case 4: return launch_wt_kp<EPI, 4>(c, s, st, dry);
    case 8: return launch_wt_kp<EPI, 8>(c, s, st, dry);
  }
  return fail(7, "internal: slots per lane %d", s.kp);
}
int launch_wt(Ctx *c, const SpmvOp &s, cudaStream_t st, bool dry) {
  switch (s.epi) {
    case EPI_GENERIC: return launch_wt_epi<EPI_GENERIC>(c, s, st, dry);
    case EPI_ADD: return launch_wt_epi<EPI_ADD>(c, s, st, dry);
    case EPI_SET: return launch_wt_epi<EPI_SET>(c, s, st, dry);
    case EPI_AXPBY: return launch_wt_epi<EPI_AXPBY>(c, s, st, dry);
    case EPI_AFCW: return launch_wt_epi<EPI_AFCW>(c, s, st, dry);
    case EPI_AFCW_LOCAL: return launch_wt_epi<EPI_AFCW_LOCAL>(c, s, st, dry);
    case EPI_AXPBY_ACC: return launch_wt_epi<EPI_AXPBY_ACC>(c, s, st, dry);
  }
  return fail(7, "internal: epilogue class %d", s.epi);
}

__global__ void __launch_bounds__(kThreads) pack_kernel(int n, const int *__restrict__ idx, const double *__restrict__ x, double *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = x[idx[i]];
}

int launch_pack(Ctx *c, const Op &o, cudaStream_t st) {
  const int n = o.xp->plan.n_send();
  if (n == 0) return 0;
  int grid = std::min((n + kThreads - 1) / kThreads, c->num_sms * 8);
  pack_kernel<<<grid, kThreads, 0, st>>>(n, o.xp->d_send_idx, o.xsrc, o.xp->d_sendbuf);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// dry = true: only set the kernel attributes / occupancy of the instantiation this op needs (outside any capture)
int launch_op(Ctx *c, const Op &o, cudaStream_t st, bool dry = false) {
  if (o.kind == OPK_SPMV) {
    if (o.s.m == 0 || o.s.ntiles == 0) return 0;
    int rc = 0;
    if (o.s.blob) {
      rc = launch_wt(c, o.s, st, dry);
    } else if (c->kernel == 1 && o.s.tiles) {
      rc = launch_r1(c, o.s, st, dry);
    } else if (o.s.rowmap) {   // the rows longer than a warp tile: one warp per row
      if (dry) return 0;
      int grid = std::min((o.s.m + 7) / 8, c->num_sms * 8);
      if (c->max_ctas > 0) grid = std::min(grid, c->max_ctas);
      CUDA_TRY(launch_k(c->pdl != 0, spmv_longrow_kernel, grid, 256, 0, st, o.s));
    } else {
      if (dry) return 0;
      int grid = std::min(o.s.nblk, c->num_sms * 8 * 4);
      CUDA_TRY(launch_k(c->pdl != 0, spmv_stream_kernel, grid, kThreads, 0, st, o.s));
    }
    if (rc || dry) return rc;
  } else if (o.kind == OPK_EW) {
    if (dry) return 0;
    if (o.e.n == 0) return 0;
    int grid = std::min((o.e.n + kThreads - 1) / kThreads, c->num_sms * 8);
    CUDA_TRY(launch_k(c->pdl != 0, ew_kernel, grid, kThreads, 0, st, o.e));
  } else if (o.kind == OPK_DENSE) {
    if (dry) return 0;
    const int n = c->dense_n;
    int grid = std::min(n, c->num_sms * 8);   // one CTA per row
    CUDA_TRY(launch_k(c->pdl != 0, dense_gemv_kernel, grid, kThreads, 0, st, n, (const double *)c->dense_T, (const double *)o.e.a, o.e.out));
  } else {
    if (dry) return 0;
    return fail(7, "internal: op kind %d cannot be launched directly", o.kind);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// kernel attributes of every instantiation an op list needs (must not run inside a stream capture)
int prepare_ops(Ctx *c, const std::vector<Op> &ops) {
  for (const Op &o : ops)
    if (o.kind == OPK_SPMV) { int rc = launch_op(c, o, c->stream, true); if (rc) return rc; }
  return 0;
}

bool op_is_empty(const Op &o) {
  return (o.kind == OPK_SPMV && (o.s.m == 0 || o.s.ntiles == 0)) || (o.kind == OPK_EW && o.e.n == 0);
}

int launch_ew_now(Ctx *c, int n, const double *a, double *out, const int *gather, const int *scatter, cudaStream_t st) {
  Op o; o.kind = OPK_EW;
  o.e.n = n; o.e.a = a; o.e.alpha = 1.0; o.e.out = out; o.e.mode = 1; o.e.gather = gather; o.e.scatter = scatter;
  return launch_op(c, o, st);
}

int run_program(Ctx *c, cudaStream_t st, int *nkernels, int begin = 0, int end = -1, const double *ub = nullptr, double *ux = nullptr);

// the agglomerated coarse levels on rank 0: natural-order vectors in, natural-order vectors out
int run_child(Ctx *c, cudaStream_t st) {
  Ctx *ch = c->child.get();
  if (!ch) return 0;
  Level &L1 = ch->L[1];
  int rc;
  if ((rc = launch_ew_now(ch, L1.n, c->child_b, ch->bb, nullptr, L1.d_pos, st))) return rc;
  if ((rc = run_program(ch, st, nullptr))) return rc;
  if ((rc = launch_ew_now(ch, L1.n, ch->xb, c->child_x, L1.d_pos, nullptr, st))) return rc;
  return 0;
}

// Execute op lists of a group of ranks in lockstep (R.size() == 1: a serial context, or one NCCL
// rank of a multi-process run).  All lists have the same length and op kinds by construction.
int exec_ops(const std::vector<Ctx *> &R, const std::vector<const std::vector<Op> *> &progs, int begin, int end, cudaStream_t st) {
  const int nr = (int)R.size();
  for (int i = begin; i < end; ++i) {
    const int kind = (*progs[0])[i].kind;
    if (kind == OPK_SPMV || kind == OPK_EW || kind == OPK_DENSE) {
      for (int r = 0; r < nr; ++r) {
        const Op &o = (*progs[r])[i];
        if (op_is_empty(o)) continue;
        int rc = launch_op(R[r], o, st);
        if (rc) return rc;
      }
    } else if (kind == OPK_EPOCH) {
      for (int r = 0; r < nr; ++r) {
        Ctx *c = R[r];
        if (c->p2p == 2) {
          // enter the cycle, tell the ranks that push to me, and (one process per GPU) wait until the ranks I push to have
          // entered it too: their ghost buffers of the previous cycle are then free.  An in-process group runs on one
          // stream: all epoch kernels are queued before the first push, nothing to wait for.
          epoch2_kernel<<<1, 1, 0, st>>>(c->flags(), c->d_peer_flags, c->rank, c->all_src, nr > 1 ? 0u : c->all_dst);
        } else {
          epoch_kernel<<<1, 1, 0, st>>>(c->flags());
        }
      }
      CUDA_TRY(cudaGetLastError());
    } else if (kind == OPK_ACK) {
      for (int r = 0; r < nr; ++r) {
        const Op &o = (*progs[r])[i];
        if (!o.xp->srcmask) continue;
        ack_kernel<<<1, 1, 0, st>>>(R[r]->d_peer_flags, R[r]->flags(), kMaxInst, o.inst, R[r]->rank, o.xp->srcmask);
      }
      CUDA_TRY(cudaGetLastError());
    } else if (kind == OPK_XCHG && (*progs[0])[i].inst >= 0 && R[0]->p2p == 2) {
      // fused exchange: normally no launch at all (the consuming kernel pushes); a stand-alone push only where this
      // rank's consumer has no rows, and for in-process groups (one stream: a consumer must never wait for a later launch)
      for (int r = 0; r < nr; ++r) {
        Ctx *c = R[r];
        const Op &o = (*progs[r])[i];
        if (o.push_here || !o.xp->dstmask) continue;
        const int units = (o.xp->plan.n_send() + 31) / 32;
        const int grid = std::max(1, std::min((units + 7) / 8, 64));
        CUDA_TRY(launch_k(c->pdl != 0, push2_kernel, grid, kThreads, 0, st, (const XPush *)(c->d_xpush + o.inst), o.xsrc));
      }
    } else if (kind == OPK_XCHG && (*progs[0])[i].inst >= 0) {
      // peer-memory exchange: every rank pushes its chunks straight into the consumers' ghost buffers and
      // raises their flags; the consumers' SpMV kernels wait on the flags (kernels.cuh: push_kernel / ghost_wait)
      for (int r = 0; r < nr; ++r) {
        Ctx *c = R[r];
        const Op &o = (*progs[r])[i];
        const DevPlan *D = o.xp;
        if (!D->dstmask) continue;
        PushOp po;
        po.n = D->plan.n_send(); po.idx = D->d_send_idx; po.x = o.xsrc; po.nranks = c->nranks; po.me = c->rank;
        po.send_off = D->d_send_off; po.dst = D->d_dst; po.peer_flags = c->d_peer_flags; po.my_flags = c->flags();
        po.epoch = c->flags(); po.done = c->d_done + o.inst; po.inst = o.inst; po.ack_inst = o.ack_inst; po.ack_delta = o.ack_delta;
        po.max_inst = kMaxInst; po.dstmask = D->dstmask;
        const int grid = std::max(1, std::min((po.n + kThreads - 1) / kThreads, 64));
        push_kernel<<<grid, kThreads, 0, st>>>(po);
      }
      CUDA_TRY(cudaGetLastError());
    } else if (kind == OPK_XCHG) {
      for (int r = 0; r < nr; ++r) {
        int rc = launch_pack(R[r], (*progs[r])[i], st);
        if (rc) return rc;
      }
      if (nr > 1) {  // in-process group: copy every peer's packed chunk into my ghost buffer
        for (int r = 0; r < nr; ++r) {
          const DevPlan *mine = (*progs[r])[i].xp;
          for (int p = 0; p < nr; ++p) {
            const int cnt = mine->plan.recv_count[p];
            if (!cnt) continue;
            const DevPlan *theirs = (*progs[p])[i].xp;
            CUDA_TRY(cudaMemcpyAsync(mine->d_xg + mine->plan.recv_off[p], theirs->d_sendbuf + theirs->plan.send_off[r], (size_t)cnt * 8,
                                     cudaMemcpyDeviceToDevice, st));
          }
        }
      } else if (R[0]->comm) {
        Ctx *c = R[0];
        const DevPlan *P = (*progs[0])[i].xp;
        std::string err;
        if ((*progs[0])[i].async) {   // fork: the pack was launched on the main stream; the NCCL group runs on the side stream
          CUDA_TRY(cudaEventRecord(c->ev_fork, st));
          CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
          bool ok2 = c->comm->group_start(&err);
          for (int p = 0; p < c->nranks && ok2; ++p) {
            if (P->plan.send_count[p]) ok2 = c->comm->send(P->d_sendbuf + P->plan.send_off[p], (size_t)P->plan.send_count[p], 8, p, c->side, &err);
            if (ok2 && P->plan.recv_count[p]) ok2 = c->comm->recv(P->d_xg + P->plan.recv_off[p], (size_t)P->plan.recv_count[p], 8, p, c->side, &err);
          }
          ok2 = ok2 && c->comm->group_end(&err);
          if (!ok2) return fail(21, "ghost exchange: %s", err.c_str());
          CUDA_TRY(cudaEventRecord(c->ev_join, c->side));
          continue;
        }
        bool ok = c->comm->group_start(&err);
        for (int p = 0; p < c->nranks && ok; ++p) {
          if (P->plan.send_count[p]) ok = c->comm->send(P->d_sendbuf + P->plan.send_off[p], (size_t)P->plan.send_count[p], 8, p, st, &err);
          if (ok && P->plan.recv_count[p]) ok = c->comm->recv(P->d_xg + P->plan.recv_off[p], (size_t)P->plan.recv_count[p], 8, p, st, &err);
        }
        ok = ok && c->comm->group_end(&err);
        if (!ok) return fail(21, "ghost exchange: %s", err.c_str());
      } else {
        return fail(21, "ghost exchange requested without a communicator");
      }
    } else if (kind == OPK_XWAIT) {
      if (nr == 1 && R[0]->comm) CUDA_TRY(cudaStreamWaitEvent(st, R[0]->ev_join, 0));   // join (in-process groups exchanged synchronously)
    } else if (kind == OPK_GATHER0 || kind == OPK_SCATTER0) {
      const bool gather = kind == OPK_GATHER0;
      if (nr > 1) {
        Ctx *c0 = R[0];
        for (int r = 0; r < nr; ++r) {
          Ctx *c = R[r];
          const Level &LB = c->L[c->l_agg];
          if (!LB.n) continue;
          double *piece = gather ? c->bb + LB.boff : c->xb + LB.xoff;
          double *glob = (gather ? c0->child_b : c0->child_x) + c0->rangeV[c0->l_agg].start[r];
          CUDA_TRY(cudaMemcpyAsync(gather ? glob : piece, gather ? piece : glob, (size_t)LB.n * 8, cudaMemcpyDeviceToDevice, st));
        }
      } else {
        Ctx *c = R[0];
        const Level &LB = c->L[c->l_agg];
        double *piece = gather ? c->bb + LB.boff : c->xb + LB.xoff;
        std::string err;
        bool ok = true;
        if (c->rank == 0) {
          double *glob = gather ? c->child_b : c->child_x;
          const Ranges &rg = c->rangeV[c->l_agg];
          ok = c->comm->group_start(&err);
          for (int p = 1; p < c->nranks && ok; ++p) {
            const size_t cnt = (size_t)(rg.start[p + 1] - rg.start[p]);
            if (!cnt) continue;
            ok = gather ? c->comm->recv(glob + rg.start[p], cnt, 8, p, st, &err) : c->comm->send(glob + rg.start[p], cnt, 8, p, st, &err);
          }
          ok = ok && c->comm->group_end(&err);
          if (ok && LB.n) CUDA_TRY(cudaMemcpyAsync(gather ? glob : piece, gather ? piece : glob, (size_t)LB.n * 8, cudaMemcpyDeviceToDevice, st));
        } else if (LB.n) {
          ok = c->comm->group_start(&err);
          ok = ok && (gather ? c->comm->send(piece, (size_t)LB.n, 8, 0, st, &err) : c->comm->recv(piece, (size_t)LB.n, 8, 0, st, &err));
          ok = ok && c->comm->group_end(&err);
        }
        if (!ok) return fail(21, "coarse-level agglomeration exchange: %s", err.c_str());
      }
    } else if (kind == OPK_CHILD) {
      for (int r = 0; r < nr; ++r) {
        int rc = run_child(R[r], st);
        if (rc) return rc;
      }
    }
  }
  return 0;
}

// the caller's vectors substituted for the placeholders of a `user` op
Op patch_user(const Op &o, const double *ub, double *ux) {
  Op q = o;
  if (q.s.x == kUserB) q.s.x = ub;
  if (q.s.aux == kUserB) q.s.aux = ub;
  if (q.s.xnat == kUserX) q.s.xnat = ux;
  if (q.e.a == kUserB) q.e.a = ub;
  if (q.e.out == kUserX) q.e.out = ux;
  return q;
}

// ops [begin, end) of the cycle program; `user` ops need the caller's vectors (ub, ux)
int run_program(Ctx *c, cudaStream_t st, int *nkernels, int begin, int end, const double *ub, double *ux) {
  int nk = 0;
  const int n = end < 0 ? (int)c->prog.size() : end;
  std::vector<Ctx *> R{c};
  std::vector<const std::vector<Op> *> P{&c->prog};
  for (int i = begin; i < n; ++i) {
    const Op &o = c->prog[i];
    if (op_is_empty(o)) continue;
    int rc;
    if (o.user) {
      if (!ub || !ux) return fail(7, "internal: op %d needs the caller's vectors", i);
      std::vector<Op> one{patch_user(o, ub, ux)};
      std::vector<const std::vector<Op> *> P1{&one};
      rc = exec_ops(R, P1, 0, 1, st);
    } else {
      rc = exec_ops(R, P, i, i + 1, st);
    }
    if (rc) return rc;
    ++nk;
  }
  if (nkernels) *nkernels = nk;
  return 0;
}

int build_graph(Ctx *c) {
  if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
  if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
  CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  int nk = 0;
  int rc = run_program(c, c->stream, &nk, c->n_head, c->tail_begin);
  cudaError_t e = cudaStreamEndCapture(c->stream, &c->graph);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(100 + (int)e, "graph capture failed: %s", cudaGetErrorString(e));
  CUDA_TRY(cudaGraphInstantiate(&c->gexec, c->graph, 0));
  c->graph_kernels = nk;
  return 0;
}

int setup_fused_push(Ctx *c);

int build_program(Ctx *c) {
  c->prog.clear();
  c->tail_levels = 0;
  const int NL = c->no_levels;
  if (NL < 2) return 0;
  const bool agg = c->l_agg <= NL;
  const int LB = agg ? c->l_agg : NL;  // bottom level of this context's nested vectors
  Builder B{c, &c->prog};
  B.use_p2p = true;
  B.io = c->io_fused;
  c->n_head = 0; c->tail_begin = -1;
  c->n_inst = 0;
  c->inst_plan.clear();
  for (DevPlan &D : c->plans) D.first_inst = D.last_inst = -1;
  if (c->p2p_ready) { Op e; e.kind = OPK_EPOCH; e.tag = 10; c->prog.push_back(e); }
  // dense collapsed tail (serial contexts): the longest suffix of levels with <= dense_rows rows
  int ldense = NL + 1, dense_begin = -1, dense_end = -1;
  if (c->nranks == 1 && c->dense_rows > 0 && c->device >= 0) {
    for (int l = NL; l >= 1; --l) { if (c->L[l].n <= c->dense_rows) ldense = l; else break; }
    if (c->io_fused && ldense < 2) ldense = 2;   // level 1 carries the fused entry / exit permutation
    if (ldense > NL - 1) ldense = NL + 1;   // >= 2 levels
  }
  // -pc_air_full_smoothing_up_and_down: PCMG multiplicative V(1,1) (src/AIR_MG_Setup.F90:978-1074), going down
  //   x_l = M_l b_l (Richardson from a zero guess) ; r_l = b_l - A_l x_l ; b_{l+1} = R r_l = r_c + Z r_f
  // A_l is streamed once for the residual (its C rows are written straight into b_{l+1}), Z once.
  for (int l = 1; c->full_smooth && l <= LB - 1; ++l) {
    Level &Lv = c->L[l];
    B.level = l;
    if (l == ldense) dense_begin = (int)c->prog.size();
    double *xl = c->xb + Lv.xoff, *bl = c->bb + Lv.boff, *rl = bl + Lv.n;
    int rc = B.emit_inv(Lv.inv_ff, Lv.Coarse, Lv.coarse_diag, Lv.n, bl, xl, 1);
    if (rc) return rc;
    {
      SpmvOp s = B.base(Lv.Coarse, xl);
      s.aux = bl; s.alpha = 1.0; s.beta = -1.0;
      s.out = rl; s.out_mode = 1;
      B.push_spmv(s, Lv.Coarse, 4, 1, 1);
    }
    {
      SpmvOp s = B.base(Lv.Z, rl);
      s.out = rl + Lv.nf; s.out_mode = 2;
      B.push_spmv(s, Lv.Z, 1, 0, 2);
    }
  }
  // down: b_{l+1} = b_c + Z b_f  (MatRestrict with R = [Z I])
  for (int l = 1; !c->full_smooth && l <= LB - 1; ++l) {
    Level &Lv = c->L[l];
    B.level = l;
    if (l == ldense) dense_begin = (int)c->prog.size();
    if (l == 1 && c->io_fused) {
      // entry permutation fused into the level-1 restriction: b_2 = b_c + Z b_f with b read in natural ordering
      // (Z stored with natural column indices, b_c picked through the nested -> natural index list)
      SpmvOp s = B.base(Lv.Znat, kUserB);
      s.aux = kUserB; s.aux_idx = Lv.d_inv + Lv.nf; s.alpha = 1.0; s.beta = 1.0;
      s.out = c->bb + Lv.boff + Lv.nf; s.out_mode = 1;
      B.push_spmv(s, Lv.Znat, 1, 1, 1);
      c->prog.back().user = true;
      c->n_head = (int)c->prog.size();
      continue;
    }
    if (Lv.any_c) B.push_ew(Lv.nc, c->bb + Lv.boff + Lv.nf, nullptr, nullptr, 1.0, Lv.bc_save, 1);
    SpmvOp s = B.base(Lv.Z, c->bb + Lv.boff);
    s.out = c->bb + Lv.boff + Lv.nf; s.out_mode = 2;
    B.push_spmv(s, Lv.Z, 1, 0, 2);
  }
  if (agg) {
    // agglomerated coarse levels: gather b of level l_agg on rank 0, serial sub-cycle there, scatter x back
    Op g; g.kind = OPK_GATHER0; g.level = LB; g.tag = 10; g.bytes = 8.0 * c->L[LB].n;
    Op ch; ch.kind = OPK_CHILD; ch.level = LB; ch.tag = 2;
    Op sc; sc.kind = OPK_SCATTER0; sc.level = LB; sc.tag = 10; sc.bytes = 8.0 * c->L[LB].n;
    c->prog.push_back(g); c->prog.push_back(ch); c->prog.push_back(sc);
  } else {
    // coarse solve: x_L = inv_A_ff(L) b_L  (mg_coarse_shell_apply, src/FC_Smooth.F90:29-49)
    Level &Lv = c->L[NL];
    B.level = NL;
    int rc = B.emit_inv(Lv.inv_ff, Lv.Coarse, Lv.coarse_diag, Lv.n, c->bb + Lv.boff, c->xb + Lv.xoff, 1);
    if (rc) return rc;
    // -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N (src/AIR_MG_Setup.F90:1094-1102: KSP_NORM_NONE, zero initial guess):
    // N - 1 more sweeps x_L += inv_A_ff(L) (b_L - A_L x_L)
    for (int it = 1; it < c->coarse_its; ++it) {
      if (!Lv.Coarse.valid()) return fail(4, "mg_coarse_ksp_max_it > 1 needs coarse_matrix(no_levels) (set_csr(..., PFLARE_B200_COARSE))");
      SpmvOp s = B.base(Lv.Coarse, c->xb + Lv.xoff);
      s.aux = c->bb + Lv.boff; s.alpha = 1.0; s.beta = -1.0;
      s.out = c->scr[1]; s.out_mode = 1;
      B.push_spmv(s, Lv.Coarse, 4, 1, 1);
      if ((rc = B.emit_inv(Lv.inv_ff, Lv.Coarse, Lv.coarse_diag, Lv.n, c->scr[1], c->xb + Lv.xoff, 2))) return rc;
    }
  }
  // up: x_l = P x_{l+1}; one mg_FC_point_richardson
  for (int l = LB - 1; !c->full_smooth && l >= 1; --l) {
    Level &Lv = c->L[l];
    B.level = l;
    if (l == 1 && c->io_fused) {
      // exit permutation of the C points of level 1 (= every deeper level's result): one scatter into the caller's x
      c->tail_begin = (int)c->prog.size();
      B.push_ew(Lv.nc, c->xb + Lv.xoff + Lv.nf, nullptr, nullptr, 1.0, kUserX, 1, nullptr, Lv.d_inv + Lv.nf);
      c->prog.back().user = true;
    }
    int rc = B.emit_fc_richardson(Lv, true);
    if (rc) return rc;
    if (l == ldense) dense_end = (int)c->prog.size();
  }
  // full smoothing, going up: x_l += P x_{l+1} (MatInterpolateAdd) ; x_l += M_l (b_l - A_l x_l)
  for (int l = LB - 1; c->full_smooth && l >= 1; --l) {
    Level &Lv = c->L[l];
    B.level = l;
    double *xl = c->xb + Lv.xoff, *bl = c->bb + Lv.boff;
    {
      SpmvOp s = B.base(Lv.Pn, c->xb + c->L[l + 1].xoff);
      s.out = xl; s.out_mode = 2;
      B.push_spmv(s, Lv.Pn, 3, 0, 2);
    }
    {
      SpmvOp s = B.base(Lv.Coarse, xl);
      s.aux = bl; s.alpha = 1.0; s.beta = -1.0;
      s.out = c->scr[1]; s.out_mode = 1;
      B.push_spmv(s, Lv.Coarse, 4, 1, 1);
    }
    int rc = B.emit_inv(Lv.inv_ff, Lv.Coarse, Lv.coarse_diag, Lv.n, c->scr[1], xl, 2);
    if (rc) return rc;
    if (l == ldense) dense_end = (int)c->prog.size();
  }
  if (dense_begin >= 0 && dense_end > dense_begin + 1) {
    // replace the ops of levels >= ldense by ONE dense product x_l = T b_l; T is built at setup by
    // running exactly these ops on the unit vectors (build_dense_tail)
    Level &Ld = c->L[ldense];
    std::vector<Op> sub(c->prog.begin() + dense_begin, c->prog.begin() + dense_end);
    Op d; d.kind = OPK_DENSE; d.level = ldense; d.tag = 11;
    d.e.n = Ld.n; d.e.a = c->bb + Ld.boff; d.e.out = c->xb + Ld.xoff;
    for (const Op &o : sub) { d.bytes += o.bytes; d.nnz += o.nnz; }
    c->prog.erase(c->prog.begin() + dense_begin, c->prog.begin() + dense_end);
    c->prog.insert(c->prog.begin() + dense_begin, d);
    if (c->tail_begin > dense_begin) c->tail_begin -= dense_end - dense_begin - 1;
    if (c->dense_level != ldense || c->dense_n != Ld.n) c->dense_built = false;
    c->dense_prog.swap(sub);
    c->dense_level = ldense; c->dense_n = Ld.n;
    c->tail_levels = NL - ldense + 1;
  } else {
    c->dense_prog.clear(); c->dense_level = 0; c->dense_n = 0;
  }
  if (c->fused_push()) {
    int rc = setup_fused_push(c);
    if (rc) return rc;
  }
  if (c->p2p_ready) {
    if (c->n_inst > kMaxInst) return fail(25, "the cycle needs %d ghost exchanges, the flag block holds %d: set option p2p=0", c->n_inst, kMaxInst);
    // the first exchange of a plan in a cycle reuses the ghost buffer of its LAST exchange of the previous cycle
    for (Op &o : c->prog)
      if (o.kind == OPK_XCHG && o.inst >= 0 && o.ack_inst < 0) { o.ack_inst = o.xp->last_inst; o.ack_delta = 1; }
  }
  c->ghost_bytes = 0; c->xchg_groups = 0;
  for (const Op &o : c->prog)
    if (o.kind == OPK_XCHG) { c->ghost_bytes += 8.0 * o.xp->plan.n_send(); ++c->xchg_groups; }
    else if (o.kind == OPK_GATHER0 || o.kind == OPK_SCATTER0) { c->ghost_bytes += o.bytes; ++c->xchg_groups; }
  if (c->device >= 0) {
    int rc;
    if ((rc = prepare_ops(c, c->prog))) return rc;
    if ((rc = prepare_ops(c, c->dense_prog))) return rc;
  }
  return 0;
}

// T = the linear map b_l -> x_l of the sub-cycle over levels >= dense_level, column by column
int build_dense_tail(Ctx *c) {
  if (c->dense_prog.empty() || c->dense_built) return 0;
  const int n = c->dense_n;
  Level &Ld = c->L[c->dense_level];
  int rc;
  if (c->dense_cap < (size_t)n * n) {   // (a superseded smaller matrix stays in the context's allocation list until destroy)
    if ((rc = dev_alloc(c, &c->dense_T, (size_t)n * n))) return rc;
    c->dense_cap = (size_t)n * n;
  }
  std::vector<Ctx *> R{c};
  std::vector<const std::vector<Op> *> P{&c->dense_prog};
  const int grid = std::min((n + kThreads - 1) / kThreads, c->num_sms * 8);
  for (int j = 0; j < n; ++j) {
    unit_vector_kernel<<<grid, kThreads, 0, c->stream>>>(n, j, c->bb + Ld.boff);
    if ((rc = exec_ops(R, P, 0, (int)c->dense_prog.size(), c->stream))) return rc;
    store_column_kernel<<<grid, kThreads, 0, c->stream>>>(n, j, c->xb + Ld.xoff, c->dense_T);
    if ((j & 255) == 255) CUDA_TRY(cudaStreamSynchronize(c->stream));   // bound the launch queue
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->dense_built = true;
  return 0;
}

int check_handle(void *h, Ctx **c) {
  if (!h) return fail(1, "null handle");
  *c = (Ctx *)h;
  if ((*c)->device >= 0) CUDA_TRY(cudaSetDevice((*c)->device));
  return 0;
}

// ------------------------------------------------------------------ setup: collectives
int allgather_blob(Ctx *c, const std::vector<char> &mine, std::vector<std::vector<char>> *all) {
  const int P = c->nranks;
  if (P == 1) { all->assign(1, mine); return 0; }
  std::vector<std::vector<char>> out((size_t)P, mine);
  std::string err;
  if (exchange_blobs(c->hostcomm.get(), out, all, &err)) return fail(22, "setup exchange failed: %s", err.c_str());
  return 0;
}

// merged global-column CSR of the local rows of an uploaded MPIAIJ operator (wire format of the
// agglomeration gather)
void serialize_global_csr(const HostCSR &A, Writer *w) {
  w->put<int32_t>(A.set ? 1 : 0);
  if (!A.set) return;
  std::vector<int64_t> ia((size_t)A.m + 1, 0), ja;
  std::vector<double> a;
  const bool og = A.n_ghost > 0;
  for (int i = 0; i < A.m; ++i) {
    std::vector<std::pair<int64_t, double>> row;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) row.push_back({A.cstart + A.ja[p], A.a[p]});
    if (og)
      for (int p = A.oia[i]; p < A.oia[i + 1]; ++p) row.push_back({A.garray[A.oja[p]], A.oa[p]});
    std::sort(row.begin(), row.end(), [](const std::pair<int64_t, double> &x, const std::pair<int64_t, double> &y) { return x.first < y.first; });
    for (auto &e : row) { ja.push_back(e.first); a.push_back(e.second); }
    ia[(size_t)i + 1] = (int64_t)ja.size();
  }
  w->put<int64_t>(A.m);
  w->put_vec(ia); w->put_vec(ja); w->put_vec(a);
}

struct GlobCSR { bool set = false; int64_t m = 0; std::vector<int64_t> ia, ja; std::vector<double> a; };
bool read_global_csr(Reader *r, GlobCSR *g) {
  g->set = r->get<int32_t>() != 0;
  if (!g->set) return r->ok;
  g->m = r->get<int64_t>();
  r->get_vec(&g->ia); r->get_vec(&g->ja); r->get_vec(&g->a);
  return r->ok;
}

void serialize_inv(const Inv &I, Writer *w) {
  w->put<int32_t>(I.kind);
  if (I.kind == 1) serialize_global_csr(I.h, w);
  else if (I.kind == 2) w->put_vec(I.hdiag);
  else if (I.kind == 3) { w->put<int32_t>(I.type); w->put<int32_t>(I.diag_scale); w->put_vec(I.re); w->put_vec(I.im); }
}

int finalize_ctx(Ctx *c);
int set_csr_impl(Ctx *c, int our_level, int which, int m, int n_local_cols, int64_t cstart, const int *di, const int *dj,
                 const double *da, int n_ghost, const int *oi, const int *oj, const double *oa, const int64_t *garray);

// Rank 0: assemble the serial hierarchy of the agglomerated levels from every rank's rows.
int build_child(Ctx *c, const std::vector<std::vector<char>> &blobs) {
  const int P = c->nranks, NL = c->no_levels, LA = c->l_agg;
  c->child.reset(new Ctx());
  Ctx *ch = c->child.get();
  ch->rank = 0; ch->nranks = 1; ch->device = c->device; ch->no_levels = NL - LA + 1;
  ch->L.resize((size_t)ch->no_levels + 1);
  ch->num_sms = c->num_sms; ch->stream = c->stream; ch->own_stream = false;
  ch->coarse_its = c->coarse_its;
  ch->use_graph = 0; ch->fuse = c->fuse; ch->fuse_epi = c->fuse_epi; ch->full_smooth = c->full_smooth; ch->dense_rows = c->dense_rows; ch->pdl = c->pdl;
  ch->kernel = c->kernel; ch->wt_format = c->wt_format; ch->fmt_split = c->fmt_split; ch->engine = c->engine; ch->sv_pf = c->sv_pf; ch->wt_stages = c->wt_stages; ch->ctas_per_sm = c->ctas_per_sm; ch->max_ctas = c->max_ctas;
  std::vector<Reader> rd;
  for (int p = 0; p < P; ++p) rd.emplace_back(blobs[p]);
  for (int l = LA; l <= NL; ++l) {
    const int cl = l - LA + 1;
    Level &Lv = ch->L[cl];
    Lv.set = true; Lv.rstart = 0;
    // concatenated pieces, rank order == global order
    GlobCSR parts[9];
    struct InvAcc { int kind = 0; GlobCSR g; std::vector<double> diag; int type = 0, ds = 0; std::vector<double> re, im; } iv[2];
    std::vector<int> is_f, is_c, smooth;
    int64_t n = 0;
    for (int p = 0; p < P; ++p) {
      Reader &r = rd[p];
      const int64_t pn = r.get<int64_t>();
      std::vector<int> f, cc, sm;
      r.get_vec(&f); r.get_vec(&cc); r.get_vec(&sm);
      const int64_t base = c->rangeV[l].start[p];
      for (int v : f) is_f.push_back((int)(base + v));
      for (int v : cc) is_c.push_back((int)(base + v));
      if (p == 0) smooth = sm;
      n += pn;
      for (int w = 0; w < 9; ++w) {
        GlobCSR g;
        if (!read_global_csr(&r, &g)) return fail(23, "agglomeration: malformed operator blob from rank %d", p);
        if (!g.set) continue;
        GlobCSR &acc = parts[w];
        if (!acc.set) { acc.set = true; acc.ia.assign(1, 0); }
        const int64_t shift = acc.ia.back();
        for (int64_t i = 1; i <= g.m; ++i) acc.ia.push_back(g.ia[(size_t)i] + shift);
        acc.ja.insert(acc.ja.end(), g.ja.begin(), g.ja.end());
        acc.a.insert(acc.a.end(), g.a.begin(), g.a.end());
        acc.m += g.m;
      }
      for (int k = 0; k < 2; ++k) {
        const int kind = r.get<int32_t>();
        InvAcc &A = iv[k];
        if (kind) A.kind = kind;
        if (kind == 1) {
          GlobCSR g;
          if (!read_global_csr(&r, &g)) return fail(23, "agglomeration: malformed inverse blob from rank %d", p);
          if (!A.g.set) { A.g.set = true; A.g.ia.assign(1, 0); }
          const int64_t shift = A.g.ia.back();
          for (int64_t i = 1; i <= g.m; ++i) A.g.ia.push_back(g.ia[(size_t)i] + shift);
          A.g.ja.insert(A.g.ja.end(), g.ja.begin(), g.ja.end());
          A.g.a.insert(A.g.a.end(), g.a.begin(), g.a.end());
          A.g.m += g.m;
        } else if (kind == 2) {
          std::vector<double> d; r.get_vec(&d);
          A.diag.insert(A.diag.end(), d.begin(), d.end());
        } else if (kind == 3) {
          A.type = r.get<int32_t>(); A.ds = r.get<int32_t>();
          std::vector<double> re, im; r.get_vec(&re); r.get_vec(&im);
          if (A.re.empty()) { A.re = re; A.im = im; }
        }
      }
      if (!r.ok) return fail(23, "agglomeration: truncated blob from rank %d", p);
    }
    Lv.n = (int)n; Lv.nf = (int)is_f.size(); Lv.nc = (int)is_c.size();
    Lv.is_f = is_f; Lv.is_c = is_c; Lv.smooth = smooth;
    Lv.any_c = false;
    for (int s : Lv.smooth) { if (s == 0) break; if (s < 0) Lv.any_c = true; }
    auto column_count = [&](int w) -> int {
      switch (w) {
        case PFLARE_B200_INV_AFF: if (c->full_smooth) return Lv.n;   // fall through
        case PFLARE_B200_AFF: case PFLARE_B200_ACF: return cl == ch->no_levels ? Lv.n : Lv.nf;
        case PFLARE_B200_AFC: case PFLARE_B200_ACC: case PFLARE_B200_P: case PFLARE_B200_INV_ACC: return Lv.nc;
        default: return Lv.n;  // R, COARSE
      }
    };
    auto give = [&](int w, const GlobCSR &g) -> int {
      std::vector<int> ia(g.ia.begin(), g.ia.end()), ja(g.ja.begin(), g.ja.end());
      if (ia.empty()) ia.assign(1, 0);
      return set_csr_impl(ch, cl, w, (int)g.m, column_count(w), 0, ia.data(), ja.data(), g.a.data(), 0, nullptr, nullptr, nullptr, nullptr);
    };
    int rc;
    for (int w = 0; w < 9; ++w)
      if (parts[w].set && (rc = give(w, parts[w]))) return rc;
    for (int k = 0; k < 2; ++k) {
      Inv &I = k == 0 ? Lv.inv_ff : Lv.inv_cc;
      InvAcc &A = iv[k];
      if (A.kind == 1) { if ((rc = give(k == 0 ? PFLARE_B200_INV_AFF : PFLARE_B200_INV_ACC, A.g))) return rc; }
      else if (A.kind == 2) { I.kind = 2; I.hdiag = A.diag; }
      else if (A.kind == 3) { I.kind = 3; I.type = A.type; I.diag_scale = A.ds; I.re = A.re; I.im = A.im; }
    }
  }
  return finalize_ctx(ch);
}

// Map every peer's arena and tell every producer where its chunks go (X4).  One process per GPU: CUDA IPC
// handles; in-process rank group: the raw pointers.
int setup_p2p(Ctx *c) {
  const int P = c->nranks;
  int rc;
  struct Hello { int32_t same_process_tag; int32_t pad; uint64_t raw; cudaIpcMemHandle_t ipc; };
  Hello me{};
  me.same_process_tag = c->cluster ? 1 : 0;
  me.raw = (uint64_t)(uintptr_t)c->arena;
  if (!c->cluster) CUDA_TRY(cudaIpcGetMemHandle(&me.ipc, c->arena));
  std::vector<std::vector<char>> out((size_t)P), in;
  for (int p = 0; p < P; ++p) {
    Writer w;
    w.put(me);
    for (DevPlan &D : c->plans) w.put<int64_t>(D.xg_off + 8 * (int64_t)D.plan.recv_off[p]);   // where p's chunk lands in MY arena
    out[p] = std::move(w.buf);
  }
  std::string err;
  if (exchange_blobs(c->hostcomm.get(), out, &in, &err)) return fail(22, "peer-memory setup exchange failed: %s", err.c_str());
  c->peer_arena.assign((size_t)P, nullptr);
  c->peer_ipc.assign((size_t)P, false);
  std::vector<unsigned long long> pf((size_t)P, 0);
  std::vector<std::vector<int64_t>> dst_off((size_t)P);
  for (int p = 0; p < P; ++p) {
    Reader r(in[p]);
    const Hello h = r.get<Hello>();
    for (size_t k = 0; k < c->plans.size(); ++k) dst_off[p].push_back(r.get<int64_t>());
    if (!r.ok) return fail(24, "malformed peer-memory hello from rank %d", p);
    if (p == c->rank) c->peer_arena[p] = c->arena;
    else if (h.same_process_tag) c->peer_arena[p] = (void *)(uintptr_t)h.raw;
    else {
      void *ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail(26, "cudaIpcOpenMemHandle(rank %d): %s -- set option p2p=0 to use NCCL", p, cudaGetErrorString(e));
      c->peer_arena[p] = ptr; c->peer_ipc[p] = true;
    }
    pf[p] = (unsigned long long)(uintptr_t)c->peer_arena[p];
  }
  if ((rc = dev_upload(c, &c->d_peer_flags, pf))) return rc;
  if ((rc = dev_alloc(c, &c->d_done, (size_t)kMaxInst))) return rc;
  CUDA_TRY(cudaMemset(c->d_done, 0, kMaxInst * sizeof(unsigned)));
  size_t k = 0;
  for (DevPlan &D : c->plans) {
    std::vector<int> so((size_t)P + 1, 0);
    std::vector<unsigned long long> dst((size_t)P, 0);
    for (int p = 0; p < P; ++p) {
      so[p] = D.plan.send_off[p];
      dst[p] = pf[p] + (unsigned long long)dst_off[p][k];
    }
    // send_off[] is only meaningful for peers with a non-zero count: rebuild it as a proper prefix
    int acc = 0;
    for (int p = 0; p < P; ++p) { so[p] = acc; acc += D.plan.send_count[p]; }
    so[P] = acc;
    if ((rc = dev_upload(c, &D.d_send_off, so))) return rc;
    if ((rc = dev_upload(c, &D.d_dst, dst))) return rc;
    ++k;
  }
  c->p2p_ready = true;
  return 0;
}

// p2p=2, after the cycle program assigned the exchange instances: one ghost buffer per instance (arena2, mapped by the
// peers like the flag arena), the per-instance destination table and the XPush descriptors; then the program's ops get
// their buffer / descriptor pointers.  Collective (every rank builds the same instance list); skipped when the instance
// layout is the one arena2 was built for (an option change that keeps the op list's exchanges).
int setup_fused_push(Ctx *c) {
  const int P = c->nranks, NI = c->n_inst;
  int rc;
  if (NI > kMaxInst) return fail(25, "the cycle needs %d ghost exchanges, the flag block holds %d: set option p2p=0", NI, kMaxInst);
  if (!c->arena2 || c->inst_plan != c->inst_plan_built) {
    for (size_t p = 0; p < c->peer_arena2.size(); ++p)
      if (c->peer_ipc2[p] && c->peer_arena2[p]) cudaIpcCloseMemHandle(c->peer_arena2[p]);
    c->peer_arena2.clear(); c->peer_ipc2.clear();
    c->inst_off.assign((size_t)NI + 1, 0);
    for (int i = 0; i < NI; ++i)
      c->inst_off[(size_t)i + 1] = c->inst_off[(size_t)i] + (int64_t)((((size_t)c->plans[(size_t)c->inst_plan[(size_t)i]].plan.n_ghost * 8) + 255) & ~(size_t)255);
    if ((rc = dev_alloc(c, &c->arena2, (size_t)c->inst_off[(size_t)NI] + 256))) return rc;
    CUDA_TRY(cudaMemset(c->arena2, 0, (size_t)c->inst_off[(size_t)NI] + 256));
    struct Hello { int32_t same_process_tag; int32_t n_inst; uint64_t raw; cudaIpcMemHandle_t ipc; };
    Hello me{};
    me.same_process_tag = c->cluster ? 1 : 0;
    me.n_inst = NI;
    me.raw = (uint64_t)(uintptr_t)c->arena2;
    if (!c->cluster) CUDA_TRY(cudaIpcGetMemHandle(&me.ipc, c->arena2));
    std::vector<std::vector<char>> out((size_t)P), in;
    for (int p = 0; p < P; ++p) {
      Writer w;
      w.put(me);
      for (int i = 0; i < NI; ++i)   // where p's chunk of instance i lands in MY arena2
        w.put<int64_t>(c->inst_off[(size_t)i] + 8 * (int64_t)c->plans[(size_t)c->inst_plan[(size_t)i]].plan.recv_off[p]);
      out[p] = std::move(w.buf);
    }
    std::string err;
    if (exchange_blobs(c->hostcomm.get(), out, &in, &err)) return fail(22, "peer-memory setup exchange failed: %s", err.c_str());
    c->peer_arena2.assign((size_t)P, nullptr);
    c->peer_ipc2.assign((size_t)P, false);
    std::vector<unsigned long long> dst((size_t)std::max(NI, 1) * P, 0);
    for (int p = 0; p < P; ++p) {
      Reader r(in[p]);
      const Hello h = r.get<Hello>();
      if (!r.ok || h.n_inst != NI) return fail(24, "rank %d built a cycle with %d ghost exchanges, this rank %d", p, r.ok ? h.n_inst : -1, NI);
      if (p == c->rank) c->peer_arena2[p] = c->arena2;
      else if (h.same_process_tag) c->peer_arena2[p] = (void *)(uintptr_t)h.raw;
      else {
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h.ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(26, "cudaIpcOpenMemHandle(rank %d): %s -- set option p2p=0 to use NCCL", p, cudaGetErrorString(e));
        c->peer_arena2[p] = ptr; c->peer_ipc2[p] = true;
      }
      for (int i = 0; i < NI; ++i) {
        const int64_t off = r.get<int64_t>();
        dst[(size_t)i * P + p] = (unsigned long long)(uintptr_t)c->peer_arena2[p] + (unsigned long long)off;
      }
      if (!r.ok) return fail(24, "malformed peer-memory hello from rank %d", p);
    }
    unsigned long long *d_dst = nullptr;
    if ((rc = dev_upload(c, &d_dst, dst))) return rc;
    std::vector<XPush> xp((size_t)std::max(NI, 1));
    for (int i = 0; i < NI; ++i) {
      const DevPlan &D = c->plans[(size_t)c->inst_plan[(size_t)i]];
      XPush &x = xp[(size_t)i];
      x.n = D.plan.n_send(); x.nranks = P; x.me = c->rank; x.inst = i;
      x.idx = D.d_send_idx; x.send_off = D.d_send_off; x.dst = d_dst + (size_t)i * P;
      x.peer_flags = c->d_peer_flags; x.epoch = c->flags(); x.done = c->d_done + i; x.dstmask = D.dstmask;
    }
    if ((rc = dev_upload(c, &c->d_xpush, xp))) return rc;
    c->inst_plan_built = c->inst_plan;
    c->all_dst = c->all_src = 0;
    for (const DevPlan &D : c->plans) { c->all_dst |= D.dstmask; c->all_src |= D.srcmask; }
  }
  auto patch = [&](std::vector<Op> &ops) {
    for (Op &o : ops)
      if (o.kind == OPK_SPMV && o.inst >= 0) {
        o.s.xg = reinterpret_cast<const double *>(c->arena2 + c->inst_off[(size_t)o.inst]);
        o.s.xpush = o.push_here ? c->d_xpush + o.inst : nullptr;
      }
  };
  patch(c->prog);
  patch(c->dense_prog);
  return 0;
}

void destroy_ctx(Ctx *c);

// Everything a previous finalize_setup built on the device (operators, vectors, plans, dense tail, graph, the
// child hierarchy of the agglomerated levels) is released, so that set_* + finalize_setup on a live handle
// (the reference's re-setup, src/PCAIR_Shell.F90:148-162) rebuilds from the new host operators and never
// mixes old and new state.
void release_device_state(Ctx *c) {
  if (c->device >= 0) {
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
    if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
    for (size_t p = 0; p < c->peer_arena.size(); ++p)
      if (c->peer_ipc[p] && c->peer_arena[p]) cudaIpcCloseMemHandle(c->peer_arena[p]);
    for (size_t p = 0; p < c->peer_arena2.size(); ++p)
      if (c->peer_ipc2[p] && c->peer_arena2[p]) cudaIpcCloseMemHandle(c->peer_arena2[p]);
    for (void *p : c->allocs) cudaFree(p);
  }
  if (c->child) destroy_ctx(c->child.release());
  c->peer_arena.clear(); c->peer_ipc.clear();
  c->peer_arena2.clear(); c->peer_ipc2.clear(); c->arena2 = nullptr; c->d_xpush = nullptr;
  c->inst_plan.clear(); c->inst_plan_built.clear(); c->inst_off.clear();
  c->allocs.clear(); c->dev_bytes = 0;
  c->plans.clear(); c->prog.clear(); c->dense_prog.clear();
  c->dense_T = nullptr; c->dense_cap = 0; c->dense_built = false; c->dense_level = 0; c->dense_n = 0;
  c->xb = c->bb = c->io_b = c->io_x = nullptr;
  for (double *&p : c->scr) p = nullptr;
  c->child_b = c->child_x = nullptr;
  c->arena = nullptr; c->arena_bytes = 0; c->p2p_ready = false; c->d_peer_flags = nullptr; c->d_done = nullptr;
  c->ksp_Ad = DevCSR(); c->ksp_buf = nullptr; c->ksp_cap = 0; c->ksp_partial = c->ksp_dots = nullptr;
  for (Level &Lv : c->L) {
    for (DevCSR *A : {&Lv.Z, &Lv.W, &Lv.Afc, &Lv.Afcw, &Lv.Aff, &Lv.Acf, &Lv.Acc, &Lv.Coarse, &Lv.Pn, &Lv.Znat, &Lv.inv_ff.d, &Lv.inv_cc.d}) *A = DevCSR();
    Lv.inv_ff.ddiag = Lv.inv_cc.ddiag = nullptr;
    Lv.aff_diag = Lv.acc_diag = Lv.coarse_diag = Lv.bc_save = nullptr;
    Lv.d_pos = Lv.d_inv = nullptr;
  }
  c->finalized = c->planned = false;
}

// ------------------------------------------------------------------ setup: the layout
int finalize_ctx(Ctx *c) {
  const int NL = c->no_levels, P = c->nranks;
  int rc;
  for (int l = 1; l <= NL; ++l)
    if (!c->L[l].set) return fail(2, "level %d was never set", l);
  if (c->planned || c->finalized || !c->allocs.empty()) release_device_state(c);
  c->plans.clear();
  // sizes must chain locally: n_{l+1} == n_coarse(l)
  for (int l = 1; l < NL; ++l)
    if (c->L[l + 1].n != c->L[l].nc)
      return fail(2, "level %d has %d rows but level %d has %d C points on this rank: the library keeps PETSc's ownership on every level "
                     "(x_c of level l is x of level l+1); hierarchies repartitioned by -pc_air_processor_agglom are not accepted -- run the "
                     "reference's setup with -pc_air_processor_agglom 0 (the library agglomerates its coarse levels itself, option agg_rows)",
                  l + 1, c->L[l + 1].n, l, c->L[l].nc);

  // ---- X1: ownership ranges of every level + the structural flags that shape the program
  std::vector<int64_t> onept((size_t)NL + 1, 1), affdiag((size_t)NL + 1, 1);
  for (int l = 1; l < NL; ++l) {
    Level &Lv = c->L[l];
    const HostCSR &Pm = Lv.H[PFLARE_B200_P], &Aff = Lv.H[PFLARE_B200_AFF];
    if (!Pm.set || Pm.m != Lv.n) return fail(2, "level %d: prolongator missing or wrong shape", l);
    if (c->full_smooth) {
      const HostCSR &Al = Lv.H[PFLARE_B200_COARSE];
      if (!Al.set || Al.m != Lv.n || Al.n != Lv.n) return fail(2, "level %d: full smoothing needs coarse_matrix(level) (PFLARE_B200_COARSE)", l);
      affdiag[l] = 0;
      continue;
    }
    if (!Aff.set || Aff.m != Lv.nf) return fail(2, "level %d: A_ff missing or wrong shape", l);
    for (int j = 0; j < Lv.nf && onept[l]; ++j) {
      const int i = Lv.is_f[j];
      if (i < 0 || i >= Lv.n) return fail(2, "level %d: IS_fine out of range", l);
      int cnt = Pm.ia[i + 1] - Pm.ia[i];
      if (Pm.n_ghost > 0) cnt += Pm.oia[i + 1] - Pm.oia[i];
      if (cnt > 1) onept[l] = 0;
    }
    affdiag[l] = is_diag_only(Aff) ? 1 : 0;
  }
  c->rangeV.assign((size_t)NL + 2, Ranges());
  c->rangeF.assign((size_t)NL + 2, Ranges());
  {
    Writer w;
    for (int l = 1; l <= NL; ++l) { w.put<int64_t>(c->L[l].n); w.put<int64_t>(c->L[l].nf); w.put<int64_t>(onept[l]); w.put<int64_t>(affdiag[l]); }
    std::vector<std::vector<char>> all;
    if ((rc = allgather_blob(c, w.buf, &all))) return rc;
    std::vector<std::vector<int64_t>> cn((size_t)NL + 1, std::vector<int64_t>(P)), cf((size_t)NL + 1, std::vector<int64_t>(P));
    for (int p = 0; p < P; ++p) {
      Reader r(all[p]);
      for (int l = 1; l <= NL; ++l) {
        cn[l][p] = r.get<int64_t>(); cf[l][p] = r.get<int64_t>();
        if (!r.get<int64_t>()) onept[l] = 0;
        if (!r.get<int64_t>()) affdiag[l] = 0;
      }
      if (!r.ok) return fail(22, "setup exchange: rank %d uploaded a different number of levels", p);
    }
    for (int l = 1; l <= NL; ++l) { c->rangeV[l].from_counts(cn[l]); c->rangeF[l].from_counts(cf[l]); }
    for (int l = 1; l <= NL; ++l)
      if (c->rangeV[l].start[c->rank] != c->L[l].rstart && P > 1)
        return fail(2, "level %d: rstart %lld does not match the ranks' row counts (%lld)", l, (long long)c->L[l].rstart, (long long)c->rangeV[l].start[c->rank]);
  }

  // ---- entry / exit permutation fused into the level-1 ops: stand-alone serial Kaskade contexts whose level 1 is a
  // pure F smooth (fuse_perm = 1: only where the permutation kernels cost something, n_1 > 16384; 2: always)
  {
    bool pure_f = NL >= 2 && !c->L[1].any_c && !c->L[1].smooth.empty() && c->L[1].smooth[0] > 0;
    c->io_fused = pure_f && P == 1 && !c->cluster && c->own_stream && !c->full_smooth && c->device >= 0 &&
                  (c->fuse_perm == 2 || (c->fuse_perm == 1 && c->L[1].n > 16384));
  }

  // ---- coarse-level agglomeration (multi-rank): levels with few global rows move to rank 0
  c->l_agg = NL + 1;
  if (P > 1 && NL >= 2 && c->agg_rows > 0)
    for (int l = 2; l <= NL; ++l)
      if (c->rangeV[l].total() <= c->agg_rows) { c->l_agg = l; break; }
  const bool agg = c->l_agg <= NL;
  const int LB = agg ? c->l_agg : NL;
  if (agg) {
    // X2: every rank ships its rows of the agglomerated levels to rank 0
    Writer w;
    for (int l = c->l_agg; l <= NL; ++l) {
      Level &Lv = c->L[l];
      w.put<int64_t>(Lv.n);
      w.put_vec(Lv.is_f); w.put_vec(Lv.is_c); w.put_vec(Lv.smooth);
      for (int k = 0; k < 9; ++k) serialize_global_csr(Lv.H[k], &w);
      serialize_inv(Lv.inv_ff, &w);
      serialize_inv(Lv.inv_cc, &w);
    }
    std::vector<std::vector<char>> out((size_t)P), in;
    out[0] = std::move(w.buf);
    std::string err;
    if (exchange_blobs(c->hostcomm.get(), out, &in, &err)) return fail(22, "agglomeration exchange failed: %s", err.c_str());
    if (c->rank == 0) {
      if ((rc = build_child(c, in))) return rc;
      const size_t N = (size_t)c->rangeV[c->l_agg].total();
      if (c->device >= 0) {
        if ((rc = dev_alloc(c, &c->child_b, N))) return rc;
        if ((rc = dev_alloc(c, &c->child_x, N))) return rc;
      }
    }
  }

  // ---- (1) nested positions of the distributed levels, bottom level (LB) stays in natural order
  c->maxn = 0;
  for (int l = LB; l >= 1; --l) {
    Level &Lv = c->L[l];
    Lv.pos.resize((size_t)Lv.n);
    Lv.fpos.assign((size_t)Lv.n, -1);
    if (l == LB) {
      std::iota(Lv.pos.begin(), Lv.pos.end(), 0);
    } else {
      std::vector<unsigned char> mark((size_t)Lv.n, 0);
      for (int j = 0; j < Lv.nf; ++j) {
        int i = Lv.is_f[j];
        if (i < 0 || i >= Lv.n || mark[i]) return fail(2, "level %d: IS_fine is not a valid index set", l);
        mark[i] = 1; Lv.pos[i] = j; Lv.fpos[i] = j;
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
  // Kaskade cycle: x_c(l) and x_{l+1} (b_c(l), b_{l+1}) are the same memory.  Full smoothing keeps every level's
  // x_l and b_l: X = [x_1 | x_2 | ...], B = [b_1 | r_f(1) | b_2 | r_f(2) | b_3 ...] so that the residual
  // r_l = b_l - A_l x_l, written at b_l + n_l, drops its C part straight into b_{l+1}.
  c->L[1].xoff = c->L[1].boff = 0;
  for (int l = 1; l < LB; ++l) {
    c->L[l + 1].xoff = c->L[l].xoff + (c->full_smooth ? c->L[l].n : c->L[l].nf);
    c->L[l + 1].boff = c->L[l].boff + (c->full_smooth ? c->L[l].n + c->L[l].nf : c->L[l].nf);
  }

  // ---- (2) operators of the distributed levels
  for (int l = 1; l <= LB; ++l) {
    Level &Lv = c->L[l];
    if (l < LB) {
      const std::vector<int> &pc = c->L[l + 1].pos;  // local coarse index -> nested position on level l+1
      const std::vector<int> &fpos = Lv.fpos;
      // R = [Z I] -> Z with rows in level l+1 nested order, columns F-local (+ ghost F points of other ranks)
      const HostCSR &R = Lv.H[PFLARE_B200_R];
      if (!R.set || R.m != Lv.nc || R.n != Lv.n) return fail(2, "level %d: restrictor missing or wrong shape", l);
      {
        HostCSR Zn; Zn.set = true; Zn.m = Lv.nc; Zn.n = Lv.nf; Zn.ia.assign((size_t)Lv.nc + 1, 0);
        for (int i = 0; i < Lv.nc; ++i) {
          bool ident = false;
          for (int p = R.ia[i]; p < R.ia[i + 1]; ++p) {
            int col = R.ja[p];
            if (fpos[col] >= 0) { Zn.ja.push_back(fpos[col]); Zn.a.push_back(R.a[p]); }
            else if (col == Lv.is_c[i] && R.a[p] == 1.0 && !ident) ident = true;
            else return fail(5, "level %d: restrictor row %d is not of the form [Z I]", l, i);
          }
          if (!ident) return fail(5, "level %d: restrictor row %d has no identity entry", l, i);
          Zn.ia[(size_t)i + 1] = (int)Zn.ja.size();
        }
        Zn.n_ghost = R.n_ghost; Zn.oia = R.oia; Zn.oja = R.oja; Zn.oa = R.oa; Zn.garray = R.garray;
        HostCSR Z = remap(Zn, pc.data(), nullptr, Lv.nf);
        if (l == 1 && c->io_fused) {
          // natural column indices: the level-1 restriction gathers from the caller's b (is_f is increasing, so rows stay sorted)
          for (int &cj : Z.ja) cj = Lv.is_f[cj];
          Z.n = Lv.n;
          if ((rc = upload_csr(c, Z, &Lv.Znat, SP_VF, l))) return rc;
        } else if ((rc = upload_csr(c, Z, &Lv.Z, SP_VF, l))) return rc;
      }
      // P = [W; I] -> W with F-local rows, columns in level l+1 nested order
      const HostCSR &Pm = Lv.H[PFLARE_B200_P];
      if (Pm.n != Lv.nc) return fail(2, "level %d: prolongator has the wrong number of local columns", l);
      for (int k = 0; k < Lv.nc; ++k) {
        int i = Lv.is_c[k];
        const bool extra = Pm.n_ghost > 0 && Pm.oia[i + 1] > Pm.oia[i];
        if (extra || Pm.ia[i + 1] - Pm.ia[i] != 1 || Pm.ja[Pm.ia[i]] != k || Pm.a[Pm.ia[i]] != 1.0)
          return fail(5, "level %d: prolongator C row %d is not an identity row", l, k);
      }
      if (c->full_smooth) {
        // x_l += P x_{l+1} (MatInterpolateAdd): P kept whole, rows and columns in nested order
        {
          HostCSR Pn = remap(Pm, Lv.pos.data(), pc.data(), Lv.nc);
          if ((rc = upload_csr(c, Pn, &Lv.Pn, SP_VNEST, l + 1))) return rc;
        }
        // the level matrix and its approximate inverse act on ALL unknowns of the level
        {
          HostCSR A2 = remap(Lv.H[PFLARE_B200_COARSE], Lv.pos.data(), Lv.pos.data(), Lv.n);
          if (c->device >= 0 && (rc = dev_upload(c, &Lv.coarse_diag, extract_diag(A2)))) return rc;
          if ((rc = upload_csr(c, A2, &Lv.Coarse, SP_VNEST, l))) return rc;
        }
        Inv &I = Lv.inv_ff;
        if (I.kind == 1) {
          if (I.h.m != Lv.n || I.h.n != Lv.n) return fail(2, "level %d: full smoothing: inv_A_ff must have the shape of coarse_matrix(level)", l);
          HostCSR M2 = remap(I.h, Lv.pos.data(), Lv.pos.data(), Lv.n);
          if ((rc = upload_csr(c, M2, &I.d, SP_VNEST, l))) return rc;
        } else if (I.kind == 2) {
          if ((int)I.hdiag.size() != Lv.n) return fail(2, "level %d: diagonal inv_A_ff has the wrong size", l);
          std::vector<double> d2((size_t)Lv.n);
          for (int k = 0; k < Lv.n; ++k) d2[Lv.pos[k]] = I.hdiag[k];
          if (c->device >= 0 && (rc = dev_upload(c, &I.ddiag, d2))) return rc;
        } else if (I.kind == 0) {
          return fail(2, "level %d: inv_A_ff not set", l);
        }
        continue;
      }
      HostCSR Wn; Wn.set = true; Wn.m = Lv.nf; Wn.n = Lv.nc; Wn.ia.assign((size_t)Lv.nf + 1, 0);
      const bool pg = Pm.n_ghost > 0;
      if (pg) { Wn.n_ghost = Pm.n_ghost; Wn.garray = Pm.garray; Wn.oia.assign((size_t)Lv.nf + 1, 0); }
      for (int j = 0; j < Lv.nf; ++j) {
        const int i = Lv.is_f[j];
        for (int p = Pm.ia[i]; p < Pm.ia[i + 1]; ++p) { Wn.ja.push_back(Pm.ja[p]); Wn.a.push_back(Pm.a[p]); }
        Wn.ia[(size_t)j + 1] = (int)Wn.ja.size();
        if (pg) {
          for (int p = Pm.oia[i]; p < Pm.oia[i + 1]; ++p) { Wn.oja.push_back(Pm.oja[p]); Wn.oa.push_back(Pm.oa[p]); }
          Wn.oia[(size_t)j + 1] = (int)Wn.oja.size();
        }
      }
      HostCSR W = remap(Wn, nullptr, pc.data(), Lv.nc);
      if ((rc = upload_csr(c, W, &Lv.W, SP_VNEST, l + 1))) return rc;
      Lv.w_onepoint = onept[l] != 0;
      // A_fc, A_ff
      const HostCSR &Afc = Lv.H[PFLARE_B200_AFC], &Aff = Lv.H[PFLARE_B200_AFF];
      if (!Afc.set || Afc.m != Lv.nf || Afc.n != Lv.nc) return fail(2, "level %d: A_fc missing or wrong shape", l);
      if (Aff.n != Lv.nf) return fail(2, "level %d: A_ff has the wrong number of local columns", l);
      HostCSR Afc2 = remap(Afc, nullptr, pc.data(), Lv.nc);
      if ((rc = upload_csr(c, Afc2, &Lv.Afc, SP_VNEST, l + 1))) return rc;
      if (Lv.w_onepoint) {
        // merged A_fc|W: every row gets its W entry (or an explicit 0.0 * x_c[0]) appended as LAST entry;
        // the ghost columns of the two operators are merged into one sorted list
        HostCSR M; M.set = true; M.m = Lv.nf; M.n = Lv.nc; M.ia.assign((size_t)Lv.nf + 1, 0);
        std::vector<int64_t> &ug = M.garray;
        ug.resize(Afc2.garray.size() + W.garray.size());
        ug.resize((size_t)(std::set_union(Afc2.garray.begin(), Afc2.garray.end(), W.garray.begin(), W.garray.end(), ug.begin()) - ug.begin()));
        M.n_ghost = (int)ug.size();
        auto ghost_map = [&](const std::vector<int64_t> &g) {
          std::vector<int> mp(g.size());
          for (size_t k = 0; k < g.size(); ++k) mp[k] = (int)(std::lower_bound(ug.begin(), ug.end(), g[k]) - ug.begin());
          return mp;
        };
        const std::vector<int> ma = ghost_map(Afc2.garray), mw = ghost_map(W.garray);
        for (int j = 0; j < Lv.nf; ++j) M.ia[j + 1] = M.ia[j] + (Afc2.ia[j + 1] - Afc2.ia[j]) + 1;
        M.ja.resize((size_t)M.ia[Lv.nf]); M.a.resize((size_t)M.ia[Lv.nf]);
        for (int j = 0; j < Lv.nf; ++j) {
          int o = M.ia[j];
          for (int p = Afc2.ia[j]; p < Afc2.ia[j + 1]; ++p) {
            const int cc = Afc2.ja[p];
            M.ja[o] = cc < Lv.nc ? cc : Lv.nc + ma[cc - Lv.nc]; M.a[o] = Afc2.a[p]; ++o;
          }
          if (W.ia[j + 1] > W.ia[j]) {
            const int cc = W.ja[W.ia[j]];
            M.ja[o] = cc < Lv.nc ? cc : Lv.nc + mw[cc - Lv.nc]; M.a[o] = W.a[W.ia[j]];
          } else { M.ja[o] = 0; M.a[o] = 0.0; }
        }
        if (Lv.nc == 0 && M.n_ghost == 0)   // no column to point the padding entry at
          for (size_t k = 0; k < M.ja.size(); ++k) M.ja[k] = 0;
        if ((rc = upload_csr(c, M, &Lv.Afcw, SP_VNEST, l + 1, true))) return rc;
        Lv.Afcw.nnz_model = Afc2.nnz() + W.nnz();
      }
      {
        HostCSR Aff2 = remap(Aff, nullptr, nullptr, Lv.nf);
        if ((rc = upload_csr(c, Aff2, &Lv.Aff, SP_F, l))) return rc;
      }
      Lv.aff_diag_only = affdiag[l] != 0;
      if (c->device >= 0 && (rc = dev_upload(c, &Lv.aff_diag, extract_diag(Aff)))) return rc;
      // inverse of A_ff
      Inv &I = Lv.inv_ff;
      if (I.kind == 1) {
        if (I.h.m != Lv.nf || I.h.n != Lv.nf) return fail(2, "level %d: inv_A_ff has the wrong shape", l);
        HostCSR M2 = remap(I.h, nullptr, nullptr, Lv.nf);
        if ((rc = upload_csr(c, M2, &I.d, SP_F, l))) return rc;
      } else if (I.kind == 2) {
        if ((int)I.hdiag.size() != Lv.nf) return fail(2, "level %d: diagonal inv_A_ff has the wrong size", l);
        if (c->device >= 0 && (rc = dev_upload(c, &I.ddiag, I.hdiag))) return rc;
      } else if (I.kind == 0) {
        return fail(2, "level %d: inv_A_ff not set", l);
      }
      // C-point smoothing operators
      if (Lv.any_c) {
        const HostCSR &Acf = Lv.H[PFLARE_B200_ACF], &Acc = Lv.H[PFLARE_B200_ACC];
        if (!Acf.set || !Acc.set) return fail(2, "level %d: C smoothing requested but A_cf / A_cc not set", l);
        HostCSR Acf2 = remap(Acf, pc.data(), nullptr, Lv.nf);
        HostCSR Acc2 = remap(Acc, pc.data(), pc.data(), Lv.nc);
        if ((rc = upload_csr(c, Acf2, &Lv.Acf, SP_F, l))) return rc;
        if ((rc = upload_csr(c, Acc2, &Lv.Acc, SP_VNEST, l + 1))) return rc;
        if (c->device >= 0 && (rc = dev_upload(c, &Lv.acc_diag, extract_diag(Acc2)))) return rc;
        Inv &J = Lv.inv_cc;
        if (J.kind == 1) {
          HostCSR M2 = remap(J.h, pc.data(), pc.data(), Lv.nc);
          if ((rc = upload_csr(c, M2, &J.d, SP_VNEST, l + 1))) return rc;
        } else if (J.kind == 2) {
          std::vector<double> d2((size_t)Lv.nc);
          for (int k = 0; k < Lv.nc; ++k) d2[pc[k]] = J.hdiag[k];
          if (c->device >= 0 && (rc = dev_upload(c, &J.ddiag, d2))) return rc;
        } else if (J.kind == 0) {
          return fail(2, "level %d: inv_A_cc not set", l);
        }
        if (c->device >= 0 && (rc = dev_alloc(c, &Lv.bc_save, (size_t)Lv.nc))) return rc;
      }
    } else if (!agg) {
      // coarsest level: inv_A_ff(no_levels) (+ coarse_matrix for a matrix-free polynomial)
      Inv &I = Lv.inv_ff;
      const HostCSR &Cm = Lv.H[PFLARE_B200_COARSE];
      if (Cm.set) {
        HostCSR C2 = remap(Cm, nullptr, nullptr, Lv.n);
        if ((rc = upload_csr(c, C2, &Lv.Coarse, SP_VNEST, l))) return rc;
        if (c->device >= 0 && (rc = dev_upload(c, &Lv.coarse_diag, extract_diag(Cm)))) return rc;
      }
      if (I.kind == 1) {
        if (I.h.m != Lv.n) return fail(2, "coarse inverse has the wrong shape");
        HostCSR M2 = remap(I.h, nullptr, nullptr, Lv.n);
        if ((rc = upload_csr(c, M2, &I.d, SP_VNEST, l))) return rc;
      } else if (I.kind == 2) {
        if ((int)I.hdiag.size() != Lv.n) return fail(2, "diagonal coarse inverse has the wrong size");
        if (c->device >= 0 && (rc = dev_upload(c, &I.ddiag, I.hdiag))) return rc;
      } else if (I.kind == 3) {
        if (!Cm.set) return fail(2, "matrix-free coarse inverse needs coarse_matrix (PFLARE_B200_COARSE)");
      } else {
        return fail(2, "coarse inverse (inv_A_ff on the coarsest level) not set");
      }
    }
  }

  // ---- the outer Krylov method's system matrix (natural ordering: KSP vectors are the caller's vectors)
  if (c->ksp_A.set) {
    if (c->ksp_A.m != c->L[1].n) return fail(2, "ksp operator has %d rows, level 1 has %d", c->ksp_A.m, c->L[1].n);
    HostCSR A2 = remap(c->ksp_A, nullptr, nullptr, c->L[1].n);
    if ((rc = upload_csr(c, A2, &c->ksp_Ad, SP_NAT, 1))) return rc;
  }

  // ---- X3: ghost exchange plans (multi-rank): who needs which entries of whose vector segments
  if (P > 1) {
    std::vector<std::vector<char>> out((size_t)P), in;
    const int nplans = (int)c->plans.size();
    for (int p = 0; p < P; ++p) {
      Writer w;
      w.put<int32_t>(nplans);
      for (DevPlan &D : c->plans) w.put<int8_t>(D.plan.n_ghost > 0 ? 1 : 0);
      out[p] = std::move(w.buf);
    }
    std::string err;
    int id = 0;
    for (DevPlan &D : c->plans) {
      const Ranges &sp = D.space_kind == SP_F ? c->rangeF[D.space_level] : c->rangeV[D.space_level];
      if (plan_requests(id, D.garray, sp, c->rank, &D.plan, &out, &err)) return fail(24, "ghost plan of operator %d: %s", id, err.c_str());
      ++id;
    }
    if (exchange_blobs(c->hostcomm.get(), out, &in, &err)) return fail(22, "ghost plan exchange failed: %s", err.c_str());
    for (int q = 0; q < P; ++q) {
      Reader r(in[q]);
      if (r.get<int32_t>() != nplans) return fail(24, "rank %d uploaded a different set of operators", q);
      for (DevPlan &D : c->plans)
        if (r.get<int8_t>()) D.global_any = true;
      while (r.ok && !r.done()) {
        const int op = r.get<int32_t>(), cnt = r.get<int32_t>();
        if (!r.ok || op < 0 || op >= nplans || cnt < 0) return fail(24, "malformed ghost request from rank %d", q);
        DevPlan &D = c->plans[(size_t)op];
        D.plan.send_off[q] = D.plan.n_send();
        D.plan.send_count[q] = cnt;
        const Level &Ls = c->L[D.space_level];
        for (int k = 0; k < cnt; ++k) {
          const int idx = r.get<int32_t>();
          int posn = -1;
          if (D.space_kind == SP_F) { if (idx >= 0 && idx < Ls.nf) posn = idx; }
          else if (idx >= 0 && idx < Ls.n) posn = D.space_kind == SP_VNEST ? Ls.pos[idx] : (D.space_kind == SP_NAT ? idx : Ls.fpos[idx]);
          if (posn < 0) return fail(24, "rank %d requested an entry this rank cannot serve (operator %d)", q, op);
          D.plan.send_idx.push_back(posn);
        }
      }
      if (!r.ok) return fail(24, "truncated ghost request from rank %d", q);
    }
    for (DevPlan &D : c->plans)
      for (int q = 0; q < P; ++q) {
        if (D.plan.send_count[q]) D.dstmask |= 1u << q;
        if (D.plan.recv_count[q]) D.srcmask |= 1u << q;
      }
    if (c->device >= 0) {
      // one arena per rank: [flag block | ghost buffers of every operator]; peers map it (peer-memory exchange)
      const size_t flag_bytes = (64 + 2 * (size_t)kMaxInst * 32) * sizeof(unsigned);
      size_t off = (flag_bytes + 255) & ~(size_t)255;
      for (DevPlan &D : c->plans) { D.xg_off = (int64_t)off; off += (((size_t)D.plan.n_ghost * 8) + 255) & ~(size_t)255; }
      c->arena_bytes = off;
      if ((rc = dev_alloc(c, &c->arena, c->arena_bytes))) return rc;
      CUDA_TRY(cudaMemset(c->arena, 0, c->arena_bytes));
      for (DevPlan &D : c->plans) {
        if ((rc = dev_upload(c, &D.d_send_idx, D.plan.send_idx))) return rc;
        if ((rc = dev_alloc(c, &D.d_sendbuf, (size_t)D.plan.n_send()))) return rc;
        D.d_xg = reinterpret_cast<double *>(c->arena + D.xg_off);
      }
      if (c->p2p && P <= 32) {   // p2p = 1 and 2 share the flag arena; 2 adds the per-instance buffers after the program is built
        if ((rc = setup_p2p(c))) return rc;
      }
    }
  }
  c->planned = true;
  if (c->device < 0) {  // host-only planning context: the program is still built (op list, counters), nothing runs
    if ((rc = build_program(c))) return rc;
    return 0;
  }

  // ---- (3) vectors
  const size_t n1 = (size_t)(c->L[LB].xoff + c->L[LB].n), nb1 = (size_t)(c->L[LB].boff + c->L[LB].n);
  if ((rc = dev_alloc(c, &c->xb, n1 + 8))) return rc;   // + padding: the A_fc|W padding entry may point one past an empty C block
  if ((rc = dev_alloc(c, &c->bb, nb1 + 8))) return rc;
  for (int k = 0; k < 7; ++k)
    if ((rc = dev_alloc(c, &c->scr[k], (size_t)c->maxn))) return rc;
  if ((rc = dev_alloc(c, &c->io_b, (size_t)c->maxn))) return rc;
  if ((rc = dev_alloc(c, &c->io_x, (size_t)c->maxn))) return rc;
  CUDA_TRY(cudaMemset(c->xb, 0, (n1 + 8) * 8));
  CUDA_TRY(cudaMemset(c->bb, 0, (nb1 + 8) * 8));
  {
    Level &L1 = c->L[1];
    std::vector<int> inv((size_t)L1.n);
    for (int i = 0; i < L1.n; ++i) inv[L1.pos[i]] = i;
    if ((rc = dev_upload(c, &L1.d_pos, L1.pos))) return rc;
    if ((rc = dev_upload(c, &L1.d_inv, inv))) return rc;
  }
  // ---- (4) program + graph
  if ((rc = build_program(c))) return rc;
  if ((rc = build_dense_tail(c))) return rc;
  if (c->use_graph && NL >= 2 && !c->cluster) {
    if ((rc = build_graph(c))) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->finalized = true;
  return 0;
}

int ensure_level_perm(Ctx *c, Level &Lv) {
  if (Lv.d_pos) return 0;
  std::vector<int> inv((size_t)Lv.n);
  for (int i = 0; i < Lv.n; ++i) inv[Lv.pos[i]] = i;
  int rc;
  if ((rc = dev_upload(c, &Lv.d_pos, Lv.pos))) return rc;
  if ((rc = dev_upload(c, &Lv.d_inv, inv))) return rc;
  return 0;
}

int set_csr_impl(Ctx *c, int our_level, int which, int m, int n_local_cols, int64_t cstart, const int *di, const int *dj,
                 const double *da, int n_ghost, const int *oi, const int *oj, const double *oa, const int64_t *garray) {
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

// one context's apply on the handle's stream (serial context or one NCCL rank)
int apply_ctx(Ctx *c, const double *b, double *x, int on_device) {
  Level &L1 = c->L[1];
  const double *bd = b;
  double *xd = x;
  int rc;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(c->io_b, b, (size_t)L1.n * 8, cudaMemcpyHostToDevice, c->stream));
    bd = c->io_b; xd = c->io_x;
  }
  if (c->io_fused) {
    // the level-1 ops read b / write x in natural ordering themselves: [head ops] [graph body] [tail ops]
    if ((rc = run_program(c, c->stream, nullptr, 0, c->n_head, bd, xd))) return rc;
    if (c->use_graph && c->gexec) {
      CUDA_TRY(cudaGraphLaunch(c->gexec, c->stream));
    } else {
      if ((rc = run_program(c, c->stream, nullptr, c->n_head, c->tail_begin))) return rc;
    }
    if ((rc = run_program(c, c->stream, nullptr, c->tail_begin, -1, bd, xd))) return rc;
  } else {
    if ((rc = launch_ew_now(c, L1.n, bd, c->bb, nullptr, L1.d_pos, c->stream))) return rc;  // bb[pos[i]] = b[i] (contiguous reads, a few sequential write streams)
    if (c->use_graph && c->gexec) {
      CUDA_TRY(cudaGraphLaunch(c->gexec, c->stream));
    } else {
      if ((rc = run_program(c, c->stream, nullptr))) return rc;
    }
    if ((rc = launch_ew_now(c, L1.n, c->xb, xd, L1.d_pos, nullptr, c->stream))) return rc;  // x[i] = xb[pos[i]]
  }
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(x, c->io_x, (size_t)L1.n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->p2p_ready) {   // word 1 of the flag block: a bounded wait for a peer's ghosts gave up (kernels.cuh: spin_until)
      unsigned tmo = 0;
      CUDA_TRY(cudaMemcpy(&tmo, c->flags() + 1, sizeof(unsigned), cudaMemcpyDeviceToHost));
      if (tmo) return fail(27, "peer-memory ghost exchange timed out waiting for another rank: the result of this apply is not valid");
    }
  }
  return 0;
}

// ops of one inverse apply (PCPFLAREINV / seam 3) on a context; in/out are c->bb / c->xb
int build_inv_ops(Ctx *c, int our_level, int which, std::vector<Op> *ops, int *n_out, const int **perm, const int **iperm,
                  const double *src = nullptr, double *dst = nullptr) {
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  if (c->l_agg <= c->no_levels && our_level >= c->l_agg) return fail(2, "level %d is agglomerated on rank 0; its inverse cannot be applied on its own", our_level);
  Level &Lv = c->L[our_level];
  const bool coarse = our_level == c->no_levels;
  const Inv *I; const DevCSR *A; const double *Ad; int n;
  *perm = *iperm = nullptr;
  int rc;
  if (which == PFLARE_B200_INV_AFF && c->full_smooth && !coarse) {
    // full smoothing: inv_A_ff(level) acts on all unknowns of the level (natural <-> nested permutation needed)
    I = &Lv.inv_ff; A = &Lv.Coarse; Ad = Lv.coarse_diag; n = Lv.n;
    if ((rc = ensure_level_perm(c, Lv))) return rc;
    *perm = Lv.d_pos; *iperm = Lv.d_inv;
  } else if (which == PFLARE_B200_INV_AFF) {
    I = &Lv.inv_ff; A = coarse ? &Lv.Coarse : &Lv.Aff; Ad = coarse ? Lv.coarse_diag : Lv.aff_diag; n = coarse ? Lv.n : Lv.nf;
  } else if (which == PFLARE_B200_INV_ACC) {
    if (coarse) return fail(2, "no inv_A_cc on the coarsest level");
    I = &Lv.inv_cc; A = &Lv.Acc; Ad = Lv.acc_diag; n = Lv.nc;
    Level &Ln = c->L[our_level + 1];
    if ((rc = ensure_level_perm(c, Ln))) return rc;
    *perm = Ln.d_pos; *iperm = Ln.d_inv;
  } else {
    return fail(2, "inv_apply: which must be INV_AFF or INV_ACC");
  }
  if (I->kind == 0) return fail(4, "that inverse was not set");
  Builder B{c, ops};
  B.level = our_level;
  // INV_AFF needs no permutation: the caller's vectors can be used in place (src / dst given)
  const bool direct = src && dst && *perm == nullptr;
  if ((rc = B.emit_inv(*I, *A, Ad, n, direct ? src : c->bb, direct ? dst : c->xb, 1))) return rc;
  *n_out = n;
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

static int create_ctx(Ctx **out, int rank, int nranks, int device, int no_levels, cudaStream_t shared_stream) {
  if (no_levels < 1) return fail(2, "no_levels must be >= 1");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(2, "bad rank %d of %d", rank, nranks);
  std::unique_ptr<Ctx> c(new Ctx());
  c->rank = rank; c->nranks = nranks; c->device = device; c->no_levels = no_levels;
  c->L.resize((size_t)no_levels + 1);
  if (device >= 0) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      return fail(10, "no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device >= ndev) return fail(10, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (shared_stream) { c->stream = shared_stream; c->own_stream = false; }
    else CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (nranks > 1 && !shared_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
  }
  *out = c.release();
  return 0;
}

int pflare_b200_create(void **handle, int rank, int nranks, const void *unique_id, int device, int no_levels) {
  if (!handle) return fail(1, "null handle pointer");
  *handle = nullptr;
  Ctx *c = nullptr;
  int rc = create_ctx(&c, rank, nranks, device, no_levels, nullptr);
  if (rc) return rc;
  if (nranks > 1 && device >= 0) {
    if (!unique_id) { delete c; return fail(20, "nranks > 1 needs a unique id"); }
    std::string err;
    c->comm.reset(Comm::create(rank, nranks, unique_id, &err));
    if (!c->comm) { delete c; return fail(20, "communicator: %s", err.c_str()); }
  }
  *handle = c;
  return 0;
}

int pflare_b200_set_host_exchange(void *handle, void *alltoall_fn, void *alltoallv_fn, void *ctx) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!alltoall_fn || !alltoallv_fn) return fail(2, "set_host_exchange: null callback");
  c->hostcomm.reset(new CallbackComm(c->rank, c->nranks, (pfb_alltoall_fn)alltoall_fn, (pfb_alltoallv_fn)alltoallv_fn, ctx));
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
  return set_csr_impl(c, our_level, which, m, n_local_cols, cstart, di, dj, da, n_ghost, oi, oj, oa, garray);
}

// ---- PetscInt = 64-bit builds (--with-64-bit-indices, Makefile:75 of the reference): the same two upload calls with
// 64-bit index arrays.  A rank's LOCAL block is narrowed to 32 bits on entry (rows, local columns and nonzeros of one
// rank must stay below 2^31; global sizes are unrestricted -- rstart / cstart / garray are 64-bit in both variants).
static bool narrow(const int64_t *src, int64_t n, std::vector<int> *dst) {
  dst->resize((size_t)std::max<int64_t>(n, 0));
  for (int64_t i = 0; i < n; ++i) {
    if (src[i] < 0 || src[i] > 2147483647LL) return false;
    (*dst)[(size_t)i] = (int)src[i];
  }
  return true;
}

int pflare_b200_set_level_i64(void *handle, int our_level, int64_t rstart, int64_t n_local, int64_t n_fine, const int64_t *is_fine,
                              int64_t n_coarse, const int64_t *is_coarse, const int64_t *smooth_order, int64_t n_smooth) {
  if (n_local > 2147483647LL) return fail(3, "level %d: a rank's local block must have < 2^31 rows", our_level);
  std::vector<int> f, cc, sm((size_t)std::max<int64_t>(n_smooth, 0));
  if (!narrow(is_fine, n_fine, &f) || !narrow(is_coarse, n_coarse, &cc)) return fail(3, "level %d: local index out of the 32-bit range", our_level);
  for (int64_t i = 0; i < n_smooth; ++i) sm[(size_t)i] = (int)smooth_order[i];
  return pflare_b200_set_level(handle, our_level, rstart, (int)n_local, (int)n_fine, f.data(), (int)n_coarse, cc.data(), sm.data(), (int)n_smooth);
}

int pflare_b200_set_csr_i64(void *handle, int our_level, int which, int64_t m, int64_t n_local_cols, int64_t cstart, const int64_t *di,
                            const int64_t *dj, const double *da, int64_t n_ghost, const int64_t *oi, const int64_t *oj, const double *oa,
                            const int64_t *garray) {
  if (m > 2147483647LL || n_local_cols > 2147483647LL || n_ghost > 2147483647LL) return fail(3, "a rank's local block must have < 2^31 rows / columns");
  std::vector<int> i1, j1, i2, j2;
  if (!narrow(di, m + 1, &i1) || !narrow(dj, di[m], &j1)) return fail(3, "operator %d of level %d: local block has >= 2^31 nonzeros or columns", which, our_level);
  if (n_ghost > 0 && (!narrow(oi, m + 1, &i2) || !narrow(oj, oi[m], &j2))) return fail(3, "operator %d of level %d: off-diagonal block out of the 32-bit range", which, our_level);
  return pflare_b200_set_csr(handle, our_level, which, (int)m, (int)n_local_cols, cstart, i1.data(), j1.data(), da, (int)n_ghost,
                             n_ghost > 0 ? i2.data() : nullptr, n_ghost > 0 ? j2.data() : nullptr, oa, garray);
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
  if (c->cluster) return fail(2, "this handle belongs to an in-process rank group: call pflare_b200_cluster_finalize");
  if (c->nranks > 1 && !c->hostcomm) {
    if (!c->comm) return fail(20, "multi-rank setup needs NCCL (device >= 0 + unique id) or pflare_b200_set_host_exchange callbacks");
    c->hostcomm.reset(new NcclHostComm(c->comm.get(), c->stream));
  }
  return finalize_ctx(c);
}

int pflare_b200_apply(void *handle, const double *b, double *x, int on_device) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (c->device < 0) return fail(10, "host-only planning context: no CUDA device bound, and this library has no CPU fallback");
  if (!c->finalized) return fail(6, "apply called before finalize_setup");
  if (c->cluster) return fail(2, "this handle belongs to an in-process rank group: call pflare_b200_cluster_apply");
  if (c->no_levels < 2) return fail(6, "apply needs >= 2 levels (the reference falls back to PCJACOBI, src/AIR_MG_Setup.F90:1167-1174)");
  return apply_ctx(c, b, x, on_device);
}

int pflare_b200_inv_apply(void *handle, int our_level, int which, const double *x, double *y, int on_device) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (c->device < 0) return fail(10, "host-only planning context: no CUDA device bound");
  if (!c->finalized) return fail(6, "inv_apply called before finalize_setup");
  if (c->cluster) return fail(2, "this handle belongs to an in-process rank group: call pflare_b200_cluster_inv_apply");
  const int n_guess = (our_level >= 1 && our_level <= c->no_levels)
                          ? (which == PFLARE_B200_INV_ACC ? c->L[our_level].nc : ((our_level == c->no_levels || c->full_smooth) ? c->L[our_level].n : c->L[our_level].nf)) : 0;
  const double *xd = x; double *yd = y;
  if (!on_device) {
    if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
    CUDA_TRY(cudaMemcpyAsync(c->io_b, x, (size_t)n_guess * 8, cudaMemcpyHostToDevice, c->stream));
    xd = c->io_b; yd = c->io_x;
  }
  std::vector<Op> ops;
  int n = 0; const int *perm, *iperm;
  const bool direct = which == PFLARE_B200_INV_AFF && !(c->full_smooth && our_level < c->no_levels);   // no permutation: run on the caller's (or the staging) vectors in place
  if ((rc = build_inv_ops(c, our_level, which, &ops, &n, &perm, &iperm, direct ? xd : nullptr, direct ? yd : nullptr))) return rc;
  if (!direct && (rc = launch_ew_now(c, n, xd, c->bb, iperm, nullptr, c->stream))) return rc;
  std::vector<Ctx *> R{c};
  std::vector<const std::vector<Op> *> Pp{&ops};
  if ((rc = exec_ops(R, Pp, 0, (int)ops.size(), c->stream))) return rc;
  if (!direct && (rc = launch_ew_now(c, n, c->xb, yd, perm, nullptr, c->stream))) return rc;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(y, c->io_x, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pflare_b200_fc_smooth(void *handle, int our_level, const double *b, double *x, int on_device) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (c->device < 0) return fail(10, "host-only planning context: no CUDA device bound");
  if (!c->finalized) return fail(6, "fc_smooth called before finalize_setup");
  if (c->cluster) return fail(2, "fc_smooth is not available on an in-process rank group");
  const int LB = c->l_agg <= c->no_levels ? c->l_agg : c->no_levels;
  if (our_level < 1 || our_level >= LB) return fail(2, "fc_smooth: our_level %d has no smoother on this context", our_level);
  if (c->full_smooth) return fail(2, "fc_smooth: with full_smoothing_up_and_down the level smoother is a plain Richardson sweep with inv_A_ff (use inv_apply)");
  Level &Lv = c->L[our_level];
  if ((rc = ensure_level_perm(c, Lv))) return rc;
  const double *bd = b; double *xd = x;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(c->io_b, b, (size_t)Lv.n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->io_x, x, (size_t)Lv.n * 8, cudaMemcpyHostToDevice, c->stream));
    bd = c->io_b; xd = c->io_x;
  }
  if ((rc = launch_ew_now(c, Lv.n, bd, c->bb + Lv.boff, Lv.d_inv, nullptr, c->stream))) return rc;
  if ((rc = launch_ew_now(c, Lv.n, xd, c->xb + Lv.xoff, Lv.d_inv, nullptr, c->stream))) return rc;
  std::vector<Op> ops;
  Builder B{c, &ops};
  B.level = our_level;
  if (Lv.any_c) B.push_ew(Lv.nc, c->bb + Lv.boff + Lv.nf, nullptr, nullptr, 1.0, Lv.bc_save, 1);
  if ((rc = B.emit_fc_richardson(Lv, false))) return rc;
  std::vector<Ctx *> R{c};
  std::vector<const std::vector<Op> *> Pp{&ops};
  if ((rc = exec_ops(R, Pp, 0, (int)ops.size(), c->stream))) return rc;
  if ((rc = launch_ew_now(c, Lv.n, c->xb + Lv.xoff, xd, Lv.d_pos, nullptr, c->stream))) return rc;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(x, c->io_x, (size_t)Lv.n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

// ---------------------------------------------------------------------- outer Krylov method on the device
// KSPSolve as the reference's drivers run it around PCApply (tests/Makefile:537-546, 1128-1134, 1322-1323):
// KSPGMRES (restart 30, classical Gram-Schmidt, left or right preconditioning) or preconditioned KSPRICHARDSON,
// KSPConvergedDefault (||r|| <= max(rtol ||b||, atol); zero rhs with a nonzero guess: rtol ||r0||).  Vectors,
// SpMV, the V-cycle and all BLAS-1 stay on the device; per iteration only the k + 2 Gram-Schmidt scalars cross
// to the host (as in PETSc's own GPU back ends).  Multi-rank: one ncclAllReduce per reduction.
int pflare_b200_ksp_set_operator(void *handle, int m, int n_local_cols, int64_t cstart, const int *di, const int *dj, const double *da,
                                 int n_ghost, const int *oi, const int *oj, const double *oa, const int64_t *garray) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  HostCSR *H = &c->ksp_A;
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

namespace {
struct Ksp {
  Ctx *c;
  int n;
  cudaStream_t st;
  double *V, *w, *z, *u;   // basis (ld = n), work vectors
  int ncta;
  // y = alpha * aux + beta * A x   (aux may be null)
  int matmult(const double *x, double *y, const double *aux, double alpha, double beta) {
    std::vector<Op> ops;
    Builder B{c, &ops};
    B.level = 1;
    SpmvOp s = B.base(c->ksp_Ad, x);
    s.aux = aux; s.alpha = alpha; s.beta = beta;
    s.out = y; s.out_mode = 1;
    B.push_spmv(s, c->ksp_Ad, 4, aux ? 1 : 0, 1);
    std::vector<Ctx *> R{c};
    std::vector<const std::vector<Op> *> P{&ops};
    return exec_ops(R, P, 0, (int)ops.size(), st);
  }
  // y = M x : the handle's preconditioner (V-cycle, or the single inverse of a PCPFLAREINV handle)
  int pcapply(const double *x, double *y) {
    if (c->no_levels >= 2) return apply_ctx(c, x, y, 1);
    std::vector<Op> ops;
    int nn = 0; const int *perm, *iperm;
    int rc = build_inv_ops(c, 1, PFLARE_B200_INV_AFF, &ops, &nn, &perm, &iperm, x, y);
    if (rc) return rc;
    std::vector<Ctx *> R{c};
    std::vector<const std::vector<Op> *> P{&ops};
    return exec_ops(R, P, 0, (int)ops.size(), st);
  }
  // out[i] = <a, Vb_i>, i < nv (global sums)
  int dots(const double *a, const double *Vb, int nv, double *out) {
    multi_dot_kernel<<<ncta, kThreads, 0, st>>>(n, a, Vb, (long long)n, nv, c->ksp_partial);
    multi_dot_finish_kernel<<<1, kKspMaxVec, 0, st>>>(ncta, nv, c->ksp_partial, c->ksp_dots);
    CUDA_TRY(cudaGetLastError());
    if (c->nranks > 1) {
      std::string err;
      if (!c->comm || !c->comm->allreduce_sum(c->ksp_dots, c->ksp_dots, (size_t)nv, st, &err)) return fail(21, "ksp reduction: %s", err.c_str());
    }
    CUDA_TRY(cudaMemcpyAsync(c->ksp_hdots, c->ksp_dots, sizeof(double) * (size_t)nv, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < nv; ++i) out[i] = c->ksp_hdots[i];
    return 0;
  }
  int norm(const double *a, double *out) {
    double v = 0.0;
    int rc = dots(a, a, 1, &v);
    *out = std::sqrt(v);
    return rc;
  }
  // y = beta * y + sum_i co[i] * Vb_i
  int axpys(double *y, double beta, const double *Vb, int nv, const double *co) {
    KspCoef k;
    for (int i = 0; i < kKspMaxVec; ++i) k.c[i] = i < nv ? co[i] : 0.0;
    multi_axpy_kernel<<<ncta, kThreads, 0, st>>>(n, y, beta, Vb, (long long)n, nv, k);
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
};
}  // namespace

int pflare_b200_ksp_solve(void *handle, int ksp_type, int pc_side, double rtol, double atol, int max_it, int restart, const double *b,
                          double *x, int on_device, int *its_out, int *reason, double *rnorm_out) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (c->device < 0) return fail(10, "host-only planning context: no CUDA device bound, and this library has no CPU fallback");
  if (!c->finalized) return fail(6, "ksp_solve called before finalize_setup");
  if (c->cluster) return fail(2, "ksp_solve is not available on an in-process rank group");
  if (!c->ksp_Ad.valid()) return fail(6, "ksp_solve needs the system matrix: call pflare_b200_ksp_set_operator before finalize_setup");
  if (restart < 1 || restart + 2 > kKspMaxVec) return fail(2, "restart must be in 1..%d", kKspMaxVec - 2);
  if (ksp_type != 0 && ksp_type != 1) return fail(2, "ksp_type must be 0 (gmres) or 1 (richardson)");
  const int n = c->L[1].n;
  const size_t need = (size_t)(restart + 1 + 3) * (size_t)std::max(n, 1) + (on_device ? 0 : 2 * (size_t)std::max(n, 1));
  if (c->ksp_cap < need) {
    if ((rc = dev_alloc(c, &c->ksp_buf, need))) return rc;
    c->ksp_cap = need;
  }
  Ksp K;
  K.c = c; K.n = n; K.st = c->stream;
  K.ncta = std::max(1, std::min((n + kThreads - 1) / kThreads, c->num_sms * 4));
  if (!c->ksp_partial) {
    if ((rc = dev_alloc(c, &c->ksp_partial, (size_t)c->num_sms * 4 * kKspMaxVec))) return rc;
    if ((rc = dev_alloc(c, &c->ksp_dots, (size_t)kKspMaxVec))) return rc;
  }
  if (!c->ksp_hdots) CUDA_TRY(cudaMallocHost((void **)&c->ksp_hdots, sizeof(double) * kKspMaxVec));
  K.V = c->ksp_buf;
  K.w = K.V + (size_t)(restart + 1) * n;
  K.z = K.w + n;
  K.u = K.z + n;
  const double *bd = b;
  double *xd = x;
  if (!on_device) {
    double *hb = K.u + n, *hx = hb + n;
    CUDA_TRY(cudaMemcpyAsync(hb, b, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(hx, x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    bd = hb; xd = hx;
  }
  const bool left = pc_side == 0;
  int its = 0, why = 0;
  double ref = -1.0, rn = 0.0;
  const double one = 1.0;
  if (ksp_type == 1) {
    // KSPRICHARDSON, unpreconditioned norm: x += M (b - A x)
    double bn;
    if ((rc = K.matmult(xd, K.w, bd, 1.0, -1.0))) return rc;            // r = b - A x
    if ((rc = K.norm(bd, &bn))) return rc;
    if ((rc = K.norm(K.w, &rn))) return rc;
    ref = bn > 0 ? bn : rn;
    for (int it = 0;; ++it) {
      if (rn <= std::max(rtol * ref, atol)) { why = rn <= atol && !(rn <= rtol * ref) ? 3 : 2; its = it; break; }
      if (it == max_it) { why = -3; its = max_it; break; }
      if ((rc = K.pcapply(K.w, K.z))) return rc;
      if ((rc = K.axpys(xd, 1.0, K.z, 1, &one))) return rc;             // x += z
      if ((rc = K.matmult(xd, K.w, bd, 1.0, -1.0))) return rc;
      if ((rc = K.norm(K.w, &rn))) return rc;
    }
  } else {
    std::vector<double> H((size_t)(restart + 1) * restart, 0.0), g((size_t)restart + 1), cs((size_t)restart), sn((size_t)restart), hcol((size_t)restart + 2);
    auto Hat = [&](int i, int k) -> double & { return H[(size_t)i * restart + k]; };
    while (true) {
      // r = b - A x (left: M r)
      if ((rc = K.matmult(xd, K.w, bd, 1.0, -1.0))) return rc;
      double *r = K.w;
      if (left) { if ((rc = K.pcapply(K.w, K.z))) return rc; r = K.z; }
      double beta;
      if ((rc = K.norm(r, &beta))) return rc;
      if (ref < 0) {
        double bn;
        if (left) { if ((rc = K.pcapply(bd, K.u))) return rc; if ((rc = K.norm(K.u, &bn))) return rc; }
        else if ((rc = K.norm(bd, &bn))) return rc;
        ref = bn > 0 ? bn : beta;
        rn = beta;
        if (beta <= std::max(rtol * ref, atol)) { why = 2; break; }
      }
      { const double inv = 1.0 / beta; if ((rc = K.axpys(K.V, 0.0, r, 1, &inv))) return rc; }   // V_0 = r / beta
      std::fill(H.begin(), H.end(), 0.0);
      std::fill(g.begin(), g.end(), 0.0);
      g[0] = beta;
      int kdone = 0;
      bool conv = false;
      for (int k = 0; k < restart; ++k) {
        double *Vk = K.V + (size_t)k * n, *Vk1 = K.V + (size_t)(k + 1) * n;
        if (left) { if ((rc = K.matmult(Vk, K.z, nullptr, 0.0, 1.0))) return rc; if ((rc = K.pcapply(K.z, K.w))) return rc; }
        else { if ((rc = K.pcapply(Vk, K.z))) return rc; if ((rc = K.matmult(K.z, K.w, nullptr, 0.0, 1.0))) return rc; }
        // classical Gram-Schmidt: all projections from the same w, then one update
        if ((rc = K.dots(K.w, K.V, k + 1, hcol.data()))) return rc;
        for (int i = 0; i <= k; ++i) { Hat(i, k) = hcol[i]; hcol[i] = -hcol[i]; }
        if ((rc = K.axpys(K.w, 1.0, K.V, k + 1, hcol.data()))) return rc;
        double hn;
        if ((rc = K.norm(K.w, &hn))) return rc;
        Hat(k + 1, k) = hn;
        if (hn != 0.0) { const double inv = 1.0 / hn; if ((rc = K.axpys(Vk1, 0.0, K.w, 1, &inv))) return rc; }
        for (int i = 0; i < k; ++i) {
          const double t = cs[i] * Hat(i, k) + sn[i] * Hat(i + 1, k);
          Hat(i + 1, k) = -sn[i] * Hat(i, k) + cs[i] * Hat(i + 1, k);
          Hat(i, k) = t;
        }
        const double d = std::hypot(Hat(k, k), Hat(k + 1, k));
        cs[k] = Hat(k, k) / d; sn[k] = Hat(k + 1, k) / d;
        Hat(k, k) = d; Hat(k + 1, k) = 0.0;
        g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
        ++its; kdone = k + 1;
        rn = std::fabs(g[k + 1]);
        if (rn <= std::max(rtol * ref, atol)) { conv = true; break; }
        if (its >= max_it) break;
      }
      // y = H^-1 g ; x += V y (right: x += M V y)
      std::vector<double> y((size_t)kdone, 0.0);
      for (int i = kdone - 1; i >= 0; --i) {
        double v = g[i];
        for (int j = i + 1; j < kdone; ++j) v -= Hat(i, j) * y[j];
        y[i] = v / Hat(i, i);
      }
      if (kdone > 0) {
        if (left) { if ((rc = K.axpys(xd, 1.0, K.V, kdone, y.data()))) return rc; }
        else {
          if ((rc = K.axpys(K.u, 0.0, K.V, kdone, y.data()))) return rc;
          if ((rc = K.pcapply(K.u, K.z))) return rc;
          if ((rc = K.axpys(xd, 1.0, K.z, 1, &one))) return rc;
        }
      }
      if (conv) { why = 2; break; }
      if (its >= max_it) { why = -3; break; }
    }
  }
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(x, xd, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (its_out) *its_out = its;
  if (reason) *reason = why;
  if (rnorm_out) *rnorm_out = rn;
  return 0;
}

int pflare_b200_get_stream(void *handle, void **stream) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  *stream = (void *)c->stream;
  return 0;
}

int pflare_b200_synchronize(void *handle) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (c->device >= 0) CUDA_TRY(cudaStreamSynchronize(c->stream));
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

static DevCSR *pick_dev(Ctx *c, int our_level, int which) {
  Level &Lv = c->L[our_level];
  switch (which) {
    case PFLARE_B200_AFF: return &Lv.Aff;
    case PFLARE_B200_AFC: return &Lv.Afc;
    case PFLARE_B200_ACF: return &Lv.Acf;
    case PFLARE_B200_ACC: return &Lv.Acc;
    case PFLARE_B200_INV_AFF: return &Lv.inv_ff.d;
    case PFLARE_B200_INV_ACC: return &Lv.inv_cc.d;
    case PFLARE_B200_R: return &Lv.Z;
    case PFLARE_B200_P: return &Lv.W;
    case PFLARE_B200_COARSE: return &Lv.Coarse;
  }
  return nullptr;
}

int pflare_b200_get_ghost_plan(void *handle, int our_level, int which, int *send_count, int *recv_count, int *recv_off,
                               int *send_idx, int *n_send_idx) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->planned) return fail(6, "get_ghost_plan called before finalize_setup");
  if (our_level < 1 || our_level > c->no_levels) return fail(2, "our_level %d out of range", our_level);
  DevCSR *A = pick_dev(c, our_level, which);
  if (!A || !A->valid() || !A->xp) return fail(2, "operator %d of level %d has no exchange plan on this context", which, our_level);
  const GhostPlan &P = A->xp->plan;
  for (int p = 0; p < c->nranks; ++p) {
    if (send_count) send_count[p] = P.send_count[p];
    if (recv_count) recv_count[p] = P.recv_count[p];
    if (recv_off) recv_off[p] = P.recv_off[p];
  }
  if (n_send_idx) *n_send_idx = P.n_send();
  if (send_idx && P.n_send()) memcpy(send_idx, P.send_idx.data(), sizeof(int) * (size_t)P.n_send());
  return 0;
}

int pflare_b200_get_layout(void *handle, int *l_agg, int64_t *global_rows, int n_levels) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->planned) return fail(6, "get_layout called before finalize_setup");
  if (l_agg) *l_agg = c->l_agg;
  for (int l = 1; l <= c->no_levels && l <= n_levels; ++l) global_rows[l - 1] = c->rangeV[l].total();
  return 0;
}

static void collect_stats(Ctx *c, double *v) {
  int nk = 0;
  for (size_t i = 0; i < c->prog.size(); ++i) {
    const Op &o = c->prog[i];
    v[1] += (o.kind == OPK_XCHG || o.kind == OPK_GATHER0 || o.kind == OPK_SCATTER0 || (o.kind == OPK_EW && o.user)) ? 0.0 : o.bytes;   // exchanges and the exit scatter are not in the SURVEY.md 8d model
    v[2] += o.nnz;
    if (o.kind == OPK_SPMV || o.kind == OPK_EW) v[5] = std::max(v[5], o.bytes);
    const bool silent = o.kind == OPK_CHILD || o.kind == OPK_XWAIT ||
                        (o.kind == OPK_XCHG && c->p2p == 2 && o.inst >= 0 && (o.push_here || !o.xp->dstmask));   // fused push: no launch of its own
    if (!op_is_empty(o) && !silent) ++nk;
  }
  v[0] += nk;
  v[3] += c->dev_bytes;
  v[4] += c->ghost_bytes;
  v[6] += c->tail_levels;
  v[7] += c->xchg_groups;
  if (c->child) { collect_stats(c->child.get(), v); v[0] += 2; }
}

int pflare_b200_get_stats(void *handle, double *stats, int nstats) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  collect_stats(c, v);
  if (!c->io_fused) v[0] += 2;  // + permute in / out (their bytes are not part of the SURVEY.md 8d model)
  for (int i = 0; i < nstats && i < 8; ++i) stats[i] = v[i];
  return 0;
}

int pflare_b200_profile_apply(void *handle, const double *b_dev, double *x_dev, int max_ops, float *ms, double *bytes,
                              int *level, int *kind, int *n_ops) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if (!c->finalized || c->no_levels < 2 || c->cluster) return fail(6, "profile_apply needs a finalized stand-alone hierarchy with >= 2 levels");
  Level &L1 = c->L[1];
  const int n = (int)c->prog.size();
  std::vector<cudaEvent_t> ev((size_t)n + 3);
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  CUDA_TRY(cudaEventRecord(ev[0], c->stream));
  if (!c->io_fused && (rc = launch_ew_now(c, L1.n, b_dev, c->bb, nullptr, L1.d_pos, c->stream))) return rc;   // entry permutation
  CUDA_TRY(cudaEventRecord(ev[1], c->stream));
  for (int i = 0; i < n; ++i) {
    if ((rc = run_program(c, c->stream, nullptr, i, i + 1, b_dev, x_dev))) return rc;
    CUDA_TRY(cudaEventRecord(ev[(size_t)i + 2], c->stream));
  }
  if (!c->io_fused && (rc = launch_ew_now(c, L1.n, c->xb, x_dev, L1.d_pos, nullptr, c->stream))) return rc;   // exit permutation
  CUDA_TRY(cudaEventRecord(ev[(size_t)n + 2], c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  int cnt = 0;
  for (int i = -1; i <= n && cnt < max_ops; ++i) {
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, ev[(size_t)(i + 1)], ev[(size_t)(i + 2)]));
    const bool perm = i < 0 || i == n;
    ms[cnt] = t;
    bytes[cnt] = perm ? (c->io_fused ? 0.0 : 20.0 * L1.n) : c->prog[i].bytes;
    level[cnt] = perm ? 1 : c->prog[i].level;
    kind[cnt] = perm ? 6 : c->prog[i].tag;
    ++cnt;
  }
  *n_ops = cnt;
  for (auto &e : ev) cudaEventDestroy(e);
  return 0;
}

static int set_option_ctx(Ctx *c, const std::string &k, double value) {
  if (k == "graph") c->use_graph = value != 0;
  else if (k == "fuse") c->fuse = value != 0;
  else if (k == "epi_classes") c->fuse_epi = value != 0;
  else if (k == "fuse_perm") {
    if (c->finalized || c->planned) return fail(2, "fuse_perm must be set before finalize_setup");
    c->fuse_perm = value != 0;
  }
  else if (k == "full_smoothing_up_and_down") {
    if (c->finalized || c->planned) return fail(2, "full_smoothing_up_and_down must be set before finalize_setup");
    c->full_smooth = value != 0;
  }
  else if (k == "wt_format") {
    if (c->finalized || c->planned) return fail(2, "wt_format must be set before finalize_setup");
    if (value != 0 && value != 1 && value != 2) return fail(2, "wt_format must be 0 (auto), 1 (chunk) or 2 (row-aligned lanes)");
    c->wt_format = (int)value;
  }
  else if (k == "fmt_split") {
    if (c->finalized || c->planned) return fail(2, "fmt_split must be set before finalize_setup");
    c->fmt_split = value;
  }
  else if (k == "engine") {
    if (value != 0 && value != 1) return fail(2, "engine must be 0 (TMA ring) or 1 (direct)");
    c->engine = (int)value;
  }
  else if (k == "sv_pf") {
    // the direct engine with the next tile's column indices loaded one tile ahead (template parameter PF of spmv_sv_kernel,
    // 80 registers, 3 CTAs per SM) was measured slower (8.84 vs 8.55 ms on 3D 256^3) and is no longer instantiated
    if (value != 0) return fail(2, "sv_pf=1 was measured slower than the default direct engine and is not compiled into the library");
  }
  else if (k == "mg_coarse_ksp_max_it") {
    if (value < 1 || value > 1000) return fail(2, "mg_coarse_ksp_max_it must be between 1 and 1000");
    c->coarse_its = (int)value;
    c->dense_built = false;                       // the collapsed tail contains the coarse solve
    if (c->child) c->child->dense_built = false;
  }
  else if (k == "wt_stages") {
    if (value != 2) return fail(2, "wt_stages: only the 2-deep ring is compiled in (3 and 4 deep measured slower: 2.79 / 2.91 / 3.96 ms)");
    c->wt_stages = (int)value;
  }
  else if (k == "dense_rows") {
    if (value > 16384) return fail(2, "dense_rows is limited to 16384 (the collapsed tail is a dense n x n fp64 matrix)");
    c->dense_rows = (int)value;
  }
  else if (k == "agg_rows") {
    if (c->finalized || c->planned) return fail(2, "agg_rows must be set before finalize_setup");
    c->agg_rows = (int64_t)value;
  }
  else if (k == "kernel") {
    const int v = (int)value;
    if (v < 0 || v > 2) return fail(2, "kernel must be 0 (stream), 1 (round-1 TMA, CTA tiles) or 2 (warp tiles)");
    if ((c->finalized || c->planned) && v != c->kernel) return fail(2, "the kernel decides the operator storage: set it before finalize_setup");
    c->kernel = v;
  }
  else if (k == "ctas_per_sm") c->ctas_per_sm = (int)value;
  else if (k == "max_ctas") c->max_ctas = (int)value;
  else if (k == "pdl") c->pdl = value != 0;
  else if (k == "overlap") c->overlap = value != 0;
  else if (k == "p2p") {
    if (c->finalized || c->planned) return fail(2, "p2p must be set before finalize_setup");
    if (value != 0 && value != 1 && value != 2) return fail(2, "p2p: 0 (NCCL send/recv), 1 (push kernel + acks) or 2 (push fused into the consuming kernel)");
    c->p2p = (int)value;
  }
  else return fail(2, "unknown option '%s'", k.c_str());
  if (c->child) {
    Ctx *ch = c->child.get();
    ch->coarse_its = c->coarse_its; ch->fuse = c->fuse; ch->fuse_epi = c->fuse_epi; ch->engine = c->engine; ch->sv_pf = c->sv_pf; ch->wt_stages = c->wt_stages; ch->ctas_per_sm = c->ctas_per_sm; ch->max_ctas = c->max_ctas;
    ch->dense_rows = c->dense_rows; ch->pdl = c->pdl;
  }
  return 0;
}

static int rebuild_after_option(Ctx *c) {
  int rc;
  if (c->child && c->child->finalized) {
    if ((rc = build_program(c->child.get()))) return rc;
    if ((rc = build_dense_tail(c->child.get()))) return rc;
  }
  if (c->finalized) {
    if ((rc = build_program(c))) return rc;
    if ((rc = build_dense_tail(c))) return rc;
    if (c->use_graph && c->no_levels >= 2 && !c->cluster) { if ((rc = build_graph(c))) return rc; }
  }
  return 0;
}

int pflare_b200_set_option(void *handle, const char *key, double value) {
  Ctx *c; int rc = check_handle(handle, &c); if (rc) return rc;
  if ((rc = set_option_ctx(c, std::string(key ? key : ""), value))) return rc;
  return rebuild_after_option(c);
}

}  // extern "C"
namespace {
void destroy_ctx(Ctx *c) {
  if (c->device >= 0) {
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->gexec) cudaGraphExecDestroy(c->gexec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->child) { destroy_ctx(c->child.release()); }
    for (size_t p = 0; p < c->peer_arena.size(); ++p)
      if (c->peer_ipc[p] && c->peer_arena[p]) cudaIpcCloseMemHandle(c->peer_arena[p]);
    for (size_t p = 0; p < c->peer_arena2.size(); ++p)
      if (c->peer_ipc2[p] && c->peer_arena2[p]) cudaIpcCloseMemHandle(c->peer_arena2[p]);
    for (void *p : c->allocs) cudaFree(p);
    if (c->ksp_hdots) cudaFreeHost(c->ksp_hdots);
    if (c->side) { cudaStreamDestroy(c->side); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); }
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
  }
  c->comm.reset();
  delete c;
}
}  // namespace
extern "C" {

int pflare_b200_destroy(void **handle) {
  if (!handle || !*handle) return 0;
  Ctx *c = (Ctx *)*handle;
  if (c->cluster) return fail(2, "this handle belongs to an in-process rank group: call pflare_b200_cluster_destroy");
  destroy_ctx(c);
  *handle = nullptr;
  return 0;
}

// ---------------------------------------------------------------------- in-process rank group
int pflare_b200_cluster_create(void **cluster, int nranks, int device, int no_levels) {
  if (!cluster) return fail(1, "null cluster pointer");
  *cluster = nullptr;
  if (nranks < 1) return fail(2, "nranks must be >= 1");
  std::unique_ptr<Cluster> cl(new Cluster());
  cl->device = device;
  if (device >= 0) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      return fail(10, "no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamCreateWithFlags(&cl->stream, cudaStreamNonBlocking));
  }
  cl->shared.reset(new SharedComm(nranks));
  for (int r = 0; r < nranks; ++r) {
    Ctx *c = nullptr;
    int rc = create_ctx(&c, r, nranks, device, no_levels, cl->stream);
    if (rc) return rc;
    c->cluster = cl.get();
    c->hostcomm.reset(new SharedRankComm(cl->shared, r));
    cl->ranks.emplace_back(c);
  }
  *cluster = cl.release();
  return 0;
}

int pflare_b200_cluster_rank(void *cluster, int rank, void **handle) {
  Cluster *cl = (Cluster *)cluster;
  if (!cl || rank < 0 || rank >= (int)cl->ranks.size()) return fail(2, "bad cluster / rank");
  *handle = cl->ranks[(size_t)rank].get();
  return 0;
}

static int cluster_exec(Cluster *cl, cudaStream_t st) {
  std::vector<Ctx *> R;
  std::vector<const std::vector<Op> *> Pp;
  for (auto &c : cl->ranks) { R.push_back(c.get()); Pp.push_back(&c->prog); }
  const size_t n = Pp[0]->size();
  for (auto *p : Pp)
    if (p->size() != n) return fail(7, "internal: rank programs differ in length");
  return exec_ops(R, Pp, 0, (int)n, st);
}

int pflare_b200_cluster_finalize(void *cluster) {
  Cluster *cl = (Cluster *)cluster;
  if (!cl) return fail(1, "null cluster");
  const int P = (int)cl->ranks.size();
  std::vector<int> rcs((size_t)P, 0);
  std::vector<std::string> errs((size_t)P);
  std::vector<std::thread> th;
  for (int r = 0; r < P; ++r)
    th.emplace_back([&, r] {
      Ctx *c = cl->ranks[(size_t)r].get();
      if (c->device >= 0) cudaSetDevice(c->device);
      rcs[(size_t)r] = finalize_ctx(c);
      if (rcs[(size_t)r]) { errs[(size_t)r] = g_err; cl->shared->abort(); }   // wake the ranks waiting in a setup collective
    });
  for (auto &t : th) t.join();
  cl->shared->aborted = false;   // a later, corrected finalize may run again
  cl->shared->arrived = 0;
  for (int r = 0; r < P; ++r)
    if (rcs[(size_t)r] && errs[(size_t)r].find("another rank") == std::string::npos) return fail(rcs[(size_t)r], "rank %d: %s", r, errs[(size_t)r].c_str());
  for (int r = 0; r < P; ++r)
    if (rcs[(size_t)r]) return fail(rcs[(size_t)r], "rank %d: %s", r, errs[(size_t)r].c_str());
  cl->finalized = true;
  if (cl->device >= 0 && cl->ranks[0]->use_graph && cl->ranks[0]->no_levels >= 2) {
    Ctx *c0 = cl->ranks[0].get();
    if (cl->gexec) { cudaGraphExecDestroy(cl->gexec); cl->gexec = nullptr; }
    if (cl->graph) { cudaGraphDestroy(cl->graph); cl->graph = nullptr; }
    CUDA_TRY(cudaStreamBeginCapture(cl->stream, cudaStreamCaptureModeThreadLocal));
    int rc = cluster_exec(cl, cl->stream);
    cudaError_t e = cudaStreamEndCapture(cl->stream, &cl->graph);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(100 + (int)e, "graph capture failed: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaGraphInstantiate(&cl->gexec, cl->graph, 0));
  }
  if (cl->device >= 0) CUDA_TRY(cudaStreamSynchronize(cl->stream));
  return 0;
}

// b[r], x[r]: rank r's local rows (natural local ordering), host or device pointers
int pflare_b200_cluster_apply(void *cluster, const double *const *b, double *const *x, int on_device) {
  Cluster *cl = (Cluster *)cluster;
  if (!cl || !cl->finalized) return fail(6, "cluster_apply called before cluster_finalize");
  if (cl->device < 0) return fail(10, "host-only planning group: no CUDA device bound, and this library has no CPU fallback");
  CUDA_TRY(cudaSetDevice(cl->device));
  if (cl->ranks[0]->no_levels < 2) return fail(6, "apply needs >= 2 levels");
  int rc;
  cudaStream_t st = cl->stream;
  for (size_t r = 0; r < cl->ranks.size(); ++r) {
    Ctx *c = cl->ranks[r].get();
    Level &L1 = c->L[1];
    const double *bd = b[r];
    if (!on_device) { CUDA_TRY(cudaMemcpyAsync(c->io_b, b[r], (size_t)L1.n * 8, cudaMemcpyHostToDevice, st)); bd = c->io_b; }
    if ((rc = launch_ew_now(c, L1.n, bd, c->bb, nullptr, L1.d_pos, st))) return rc;
  }
  if (cl->gexec && cl->ranks[0]->use_graph) CUDA_TRY(cudaGraphLaunch(cl->gexec, st));
  else if ((rc = cluster_exec(cl, st))) return rc;
  for (size_t r = 0; r < cl->ranks.size(); ++r) {
    Ctx *c = cl->ranks[r].get();
    Level &L1 = c->L[1];
    double *xd = on_device ? x[r] : c->io_x;
    if ((rc = launch_ew_now(c, L1.n, c->xb, xd, L1.d_pos, nullptr, st))) return rc;
    if (!on_device) CUDA_TRY(cudaMemcpyAsync(x[r], c->io_x, (size_t)L1.n * 8, cudaMemcpyDeviceToHost, st));
  }
  if (!on_device) CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int pflare_b200_cluster_inv_apply(void *cluster, int our_level, int which, const double *const *xin, double *const *y, int on_device) {
  Cluster *cl = (Cluster *)cluster;
  if (!cl || !cl->finalized) return fail(6, "cluster_inv_apply called before cluster_finalize");
  if (cl->device < 0) return fail(10, "host-only planning group: no CUDA device bound");
  CUDA_TRY(cudaSetDevice(cl->device));
  const size_t P = cl->ranks.size();
  std::vector<std::vector<Op>> ops(P);
  std::vector<int> n(P, 0);
  std::vector<const int *> perm(P), iperm(P);
  std::vector<Ctx *> R;
  std::vector<const std::vector<Op> *> Pp;
  int rc;
  cudaStream_t st = cl->stream;
  for (size_t r = 0; r < P; ++r) {
    Ctx *c = cl->ranks[r].get();
    if ((rc = build_inv_ops(c, our_level, which, &ops[r], &n[r], &perm[r], &iperm[r]))) return rc;
    R.push_back(c); Pp.push_back(&ops[r]);
    const double *xd = xin[r];
    if (!on_device) { CUDA_TRY(cudaMemcpyAsync(c->io_b, xin[r], (size_t)n[r] * 8, cudaMemcpyHostToDevice, st)); xd = c->io_b; }
    if ((rc = launch_ew_now(c, n[r], xd, c->bb, iperm[r], nullptr, st))) return rc;
  }
  for (size_t r = 1; r < P; ++r)
    if (ops[r].size() != ops[0].size()) return fail(7, "internal: rank op lists differ in length");
  if ((rc = exec_ops(R, Pp, 0, (int)ops[0].size(), st))) return rc;
  for (size_t r = 0; r < P; ++r) {
    Ctx *c = cl->ranks[r].get();
    double *yd = on_device ? y[r] : c->io_x;
    if ((rc = launch_ew_now(c, n[r], c->xb, yd, perm[r], nullptr, st))) return rc;
    if (!on_device) CUDA_TRY(cudaMemcpyAsync(y[r], c->io_x, (size_t)n[r] * 8, cudaMemcpyDeviceToHost, st));
  }
  if (!on_device) CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int pflare_b200_cluster_set_option(void *cluster, const char *key, double value) {
  Cluster *cl = (Cluster *)cluster;
  if (!cl) return fail(1, "null cluster");
  for (auto &c : cl->ranks) {
    int rc = set_option_ctx(c.get(), std::string(key ? key : ""), value);
    if (rc) return rc;
    if ((rc = rebuild_after_option(c.get()))) return rc;
  }
  if (cl->finalized && cl->gexec) {  // re-capture
    cudaGraphExecDestroy(cl->gexec); cl->gexec = nullptr;
    cudaGraphDestroy(cl->graph); cl->graph = nullptr;
    if (cl->ranks[0]->use_graph) {
      CUDA_TRY(cudaStreamBeginCapture(cl->stream, cudaStreamCaptureModeThreadLocal));
      int rc = cluster_exec(cl, cl->stream);
      cudaError_t e = cudaStreamEndCapture(cl->stream, &cl->graph);
      if (rc) return rc;
      if (e != cudaSuccess) return fail(100 + (int)e, "graph capture failed: %s", cudaGetErrorString(e));
      CUDA_TRY(cudaGraphInstantiate(&cl->gexec, cl->graph, 0));
    }
  }
  return 0;
}

int pflare_b200_cluster_get_stream(void *cluster, void **stream) {
  Cluster *cl = (Cluster *)cluster;
  if (!cl) return fail(1, "null cluster");
  *stream = (void *)cl->stream;
  return 0;
}

int pflare_b200_cluster_destroy(void **cluster) {
  if (!cluster || !*cluster) return 0;
  Cluster *cl = (Cluster *)*cluster;
  if (cl->device >= 0) {
    cudaSetDevice(cl->device);
    if (cl->stream) cudaStreamSynchronize(cl->stream);
    if (cl->gexec) cudaGraphExecDestroy(cl->gexec);
    if (cl->graph) cudaGraphDestroy(cl->graph);
  }
  for (auto &c : cl->ranks) { Ctx *p = c.release(); p->cluster = nullptr; destroy_ctx(p); }
  if (cl->device >= 0 && cl->stream) cudaStreamDestroy(cl->stream);
  delete cl;
  *cluster = nullptr;
  return 0;
}

}  // extern "C"
