// dist.h -- host-side (no CUDA) logic of the multi-rank layout: PETSc MPIAIJ ownership ranges,
// ghost-exchange plans and the wire format used to agglomerate coarse levels on one rank.
//
// Reference behaviour being reproduced (paths relative to the PFLARE tree):
//   * row ownership = contiguous blocks per rank on every level; every operator arrives as the
//     (diag block, off-diag block, garray) triple of MatMPIAIJGetSeqAIJ
//     (src/Grid_Transferk.kokkos.cxx:30-42, src/PMISR_Module.F90:174-180);
//   * every MatMult does a ghost scatter (VecScatterBegin/End inside MatMult_MPIAIJ) -- here a
//     pack kernel + grouped point-to-point exchange whose plan is built once, below;
//   * processor agglomeration (src/AIR_MG_Setup.F90:645-907, src/Repartition.F90): coarse levels
//     live on fewer ranks -- here levels below a row threshold are gathered onto rank 0.
//
// Everything in this file is plain C++ so that it can be exercised without a GPU (tests use a
// world_size-2 gloo group as the host communicator).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace pfb {

// Host communicator used at SETUP time only (MPI_Alltoall / MPI_Alltoallv semantics on bytes).
struct HostComm {
  virtual ~HostComm() {}
  virtual int rank() const = 0;
  virtual int size() const = 0;
  // one int64 to / from every rank
  virtual int alltoall(const int64_t *send, int64_t *recv, std::string *err) = 0;
  virtual int alltoallv(const char *sbuf, const int64_t *scnt, const int64_t *sdsp, char *rbuf, const int64_t *rcnt,
                        const int64_t *rdsp, std::string *err) = 0;
};

// user-supplied callbacks (C-ABI pflare_b200_set_host_exchange): return 0 on success
typedef int (*pfb_alltoall_fn)(void *ctx, const int64_t *send, int64_t *recv);
typedef int (*pfb_alltoallv_fn)(void *ctx, const char *sbuf, const int64_t *scnt, const int64_t *sdsp, char *rbuf,
                                const int64_t *rcnt, const int64_t *rdsp);

struct CallbackComm : HostComm {
  int rank_, size_;
  pfb_alltoall_fn a2a;
  pfb_alltoallv_fn a2av;
  void *ctx;
  CallbackComm(int r, int s, pfb_alltoall_fn f1, pfb_alltoallv_fn f2, void *c) : rank_(r), size_(s), a2a(f1), a2av(f2), ctx(c) {}
  int rank() const override { return rank_; }
  int size() const override { return size_; }
  int alltoall(const int64_t *send, int64_t *recv, std::string *err) override {
    int rc = a2a(ctx, send, recv);
    if (rc && err) *err = "host alltoall callback failed";
    return rc;
  }
  int alltoallv(const char *sbuf, const int64_t *scnt, const int64_t *sdsp, char *rbuf, const int64_t *rcnt, const int64_t *rdsp,
                std::string *err) override {
    int rc = a2av(ctx, sbuf, scnt, sdsp, rbuf, rcnt, rdsp);
    if (rc && err) *err = "host alltoallv callback failed";
    return rc;
  }
};

// Exchange variable-length byte strings with every rank (sizes first, then payloads).
inline int exchange_blobs(HostComm *hc, const std::vector<std::vector<char>> &out, std::vector<std::vector<char>> *in,
                          std::string *err) {
  const int P = hc->size();
  std::vector<int64_t> scnt(P), rcnt(P), sdsp(P), rdsp(P);
  for (int p = 0; p < P; ++p) scnt[p] = (int64_t)out[p].size();
  int rc = hc->alltoall(scnt.data(), rcnt.data(), err);
  if (rc) return rc;
  int64_t st = 0, rt = 0;
  for (int p = 0; p < P; ++p) { sdsp[p] = st; st += scnt[p]; rdsp[p] = rt; rt += rcnt[p]; }
  std::vector<char> sbuf((size_t)std::max<int64_t>(st, 1)), rbuf((size_t)std::max<int64_t>(rt, 1));
  for (int p = 0; p < P; ++p)
    if (scnt[p]) memcpy(sbuf.data() + sdsp[p], out[p].data(), (size_t)scnt[p]);
  rc = hc->alltoallv(sbuf.data(), scnt.data(), sdsp.data(), rbuf.data(), rcnt.data(), rdsp.data(), err);
  if (rc) return rc;
  in->assign(P, std::vector<char>());
  for (int p = 0; p < P; ++p) (*in)[p].assign(rbuf.begin() + rdsp[p], rbuf.begin() + rdsp[p] + rcnt[p]);
  return 0;
}

// Contiguous ownership ranges of one index space: rank p owns [start[p], start[p+1]).
struct Ranges {
  std::vector<int64_t> start;
  void from_counts(const std::vector<int64_t> &cnt) {
    start.assign(cnt.size() + 1, 0);
    for (size_t p = 0; p < cnt.size(); ++p) start[p + 1] = start[p] + cnt[p];
  }
  int64_t total() const { return start.empty() ? 0 : start.back(); }
  int owner(int64_t g) const {  // rank owning global index g
    return (int)(std::upper_bound(start.begin(), start.end(), g) - start.begin()) - 1;
  }
};

// Ghost exchange plan of ONE operator: which entries of my x segment every peer needs (send side,
// already translated to positions in my local segment) and how my ghost buffer is filled (recv
// side; ghost slot order == garray order, so every peer's chunk is contiguous).
struct GhostPlan {
  int n_ghost = 0;
  std::vector<int> recv_count, recv_off, send_count, send_off;  // per rank
  std::vector<int> send_idx;
  int n_send() const { return (int)send_idx.size(); }
  bool any() const { return n_ghost > 0 || !send_idx.empty(); }
  void init(int P) {
    recv_count.assign(P, 0); recv_off.assign(P, 0); send_count.assign(P, 0); send_off.assign(P, 0);
    send_idx.clear(); n_ghost = 0;
  }
};

// Recv side of a plan from a (sorted) garray; appends the request [op_id, count, owner-local ids]
// for every owner to that owner's outgoing blob.
inline int plan_requests(int op_id, const std::vector<int64_t> &garray, const Ranges &space, int my_rank, GhostPlan *plan,
                         std::vector<std::vector<char>> *out, std::string *err) {
  const int P = (int)space.start.size() - 1;
  plan->n_ghost = (int)garray.size();
  size_t k = 0;
  while (k < garray.size()) {
    const int64_t g = garray[k];
    if (g < 0 || g >= space.total()) { if (err) *err = "garray entry out of the column space"; return 1; }
    if (k > 0 && garray[k - 1] >= g) { if (err) *err = "garray is not strictly increasing"; return 1; }
    const int p = space.owner(g);
    if (p == my_rank) { if (err) *err = "garray lists a locally owned column"; return 1; }
    size_t e = k;
    while (e < garray.size() && garray[e] < space.start[p + 1]) {
      if (e > k && garray[e - 1] >= garray[e]) { if (err) *err = "garray is not strictly increasing"; return 1; }
      ++e;
    }
    plan->recv_off[p] = (int)k;
    plan->recv_count[p] = (int)(e - k);
    std::vector<char> &blob = (*out)[p];
    const int32_t hdr[2] = {op_id, (int32_t)(e - k)};
    blob.insert(blob.end(), (const char *)hdr, (const char *)hdr + sizeof hdr);
    for (size_t j = k; j < e; ++j) {
      const int32_t loc = (int32_t)(garray[j] - space.start[p]);
      blob.insert(blob.end(), (const char *)&loc, (const char *)&loc + 4);
    }
    k = e;
  }
  return 0;
}

// ------------------------------------------------------------------ wire format helpers
struct Writer {
  std::vector<char> buf;
  template <class T> void put(const T &v) { buf.insert(buf.end(), (const char *)&v, (const char *)&v + sizeof(T)); }
  template <class T> void put_vec(const std::vector<T> &v) {
    put<int64_t>((int64_t)v.size());
    if (!v.empty()) buf.insert(buf.end(), (const char *)v.data(), (const char *)v.data() + v.size() * sizeof(T));
  }
};
struct Reader {
  const char *p, *e;
  bool ok = true;
  Reader(const std::vector<char> &b) : p(b.data()), e(b.data() + b.size()) {}
  template <class T> T get() {
    T v{};
    if (p + sizeof(T) > e) { ok = false; return v; }
    memcpy(&v, p, sizeof(T)); p += sizeof(T);
    return v;
  }
  template <class T> void get_vec(std::vector<T> *v) {
    const int64_t n = get<int64_t>();
    if (!ok || n < 0 || p + n * (int64_t)sizeof(T) > e) { ok = false; v->clear(); return; }
    v->resize((size_t)n);
    if (n) memcpy(v->data(), p, (size_t)n * sizeof(T));
    p += n * sizeof(T);
  }
  bool done() const { return p == e; }
};

}  // namespace pfb
