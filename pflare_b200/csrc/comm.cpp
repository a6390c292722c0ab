// comm.cpp -- NCCL through dlopen (see comm.h).
#include "comm.h"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace pfb {
namespace {

struct UniqueId { char internal[128]; };
typedef int (*fn_getuid)(UniqueId *);
typedef int (*fn_init)(void **, int, UniqueId, int);
typedef int (*fn_destroy)(void *);
typedef int (*fn_sendrecv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_void)(void);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_errstr)(int);

struct Api {
  void *lib = nullptr;
  fn_getuid get_uid = nullptr;
  fn_init init = nullptr;
  fn_destroy destroy = nullptr;
  fn_sendrecv send = nullptr, recv = nullptr;
  fn_void gstart = nullptr, gend = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_errstr errstr = nullptr;
  std::string load_error;
};

Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) { a.load_error = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return; }
    a.get_uid = (fn_getuid)dlsym(a.lib, "ncclGetUniqueId");
    a.init = (fn_init)dlsym(a.lib, "ncclCommInitRank");
    a.destroy = (fn_destroy)dlsym(a.lib, "ncclCommDestroy");
    a.send = (fn_sendrecv)dlsym(a.lib, "ncclSend");
    a.recv = (fn_sendrecv)dlsym(a.lib, "ncclRecv");
    a.gstart = (fn_void)dlsym(a.lib, "ncclGroupStart");
    a.gend = (fn_void)dlsym(a.lib, "ncclGroupEnd");
    a.allreduce = (fn_allreduce)dlsym(a.lib, "ncclAllReduce");
    a.errstr = (fn_errstr)dlsym(a.lib, "ncclGetErrorString");
    if (!a.get_uid || !a.init || !a.destroy || !a.send || !a.recv || !a.gstart || !a.gend)
      a.load_error = "libnccl is missing required symbols";
  });
  return a;
}

bool ok(int rc, const char *what, std::string *err) {
  if (rc == 0) return true;
  if (err) {
    *err = std::string(what) + " failed: ";
    *err += api().errstr ? api().errstr(rc) : "nccl error";
  }
  return false;
}

int nccl_type(int bytes) { return bytes == 8 ? 8 /*ncclFloat64*/ : (bytes == 4 ? 2 /*ncclInt32*/ : 0 /*ncclInt8*/); }

}  // namespace

bool Comm::unique_id(void *id128, std::string *err) {
  Api &a = api();
  if (!a.load_error.empty()) { if (err) *err = a.load_error; return false; }
  UniqueId id;
  if (!ok(a.get_uid(&id), "ncclGetUniqueId", err)) return false;
  memcpy(id128, &id, sizeof id);
  return true;
}

Comm *Comm::create(int rank, int nranks, const void *id128, std::string *err) {
  Api &a = api();
  if (!a.load_error.empty()) { if (err) *err = a.load_error; return nullptr; }
  UniqueId id;
  memcpy(&id, id128, sizeof id);
  Comm *c = new Comm();
  c->rank_ = rank; c->nranks_ = nranks;
  if (!ok(a.init(&c->comm_, nranks, id, rank), "ncclCommInitRank", err)) { delete c; return nullptr; }
  return c;
}

Comm::~Comm() {
  if (comm_) api().destroy(comm_);
}

bool Comm::group_start(std::string *err) { return ok(api().gstart(), "ncclGroupStart", err); }
bool Comm::group_end(std::string *err) { return ok(api().gend(), "ncclGroupEnd", err); }
bool Comm::send(const void *buf, size_t count, int dtype_bytes, int peer, cudaStream_t st, std::string *err) {
  return ok(api().send((void *)buf, count, nccl_type(dtype_bytes), peer, comm_, st), "ncclSend", err);
}
bool Comm::allreduce_sum(const double *in, double *out, size_t count, cudaStream_t st, std::string *err) {
  if (!api().allreduce) { if (err) *err = "libnccl has no ncclAllReduce"; return false; }
  return ok(api().allreduce(in, out, count, 8 /*ncclFloat64*/, 0 /*ncclSum*/, comm_, st), "ncclAllReduce", err);
}
bool Comm::recv(void *buf, size_t count, int dtype_bytes, int peer, cudaStream_t st, std::string *err) {
  return ok(api().recv(buf, count, nccl_type(dtype_bytes), peer, comm_, st), "ncclRecv", err);
}

}  // namespace pfb
