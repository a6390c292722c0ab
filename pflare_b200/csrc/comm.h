// comm.h -- thin NCCL wrapper (one communicator per PC handle).  NCCL is resolved at run time
// with dlopen so that the single-GPU path has no link-time dependency and a process that
// already loaded torch's bundled libnccl shares it.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <string>

namespace pfb {

class Comm {
 public:
  static bool unique_id(void *id128, std::string *err);
  static Comm *create(int rank, int nranks, const void *id128, std::string *err);
  ~Comm();
  int rank() const { return rank_; }
  int size() const { return nranks_; }
  bool group_start(std::string *err);
  bool group_end(std::string *err);
  bool send(const void *buf, size_t count, int dtype_bytes, int peer, cudaStream_t st, std::string *err);
  bool recv(void *buf, size_t count, int dtype_bytes, int peer, cudaStream_t st, std::string *err);
  // outer Krylov method only (dot products / norms): the V-cycle itself needs no reduction
  bool allreduce_sum(const double *in, double *out, size_t count, cudaStream_t st, std::string *err);

 private:
  Comm() {}
  void *comm_ = nullptr;
  int rank_ = 0, nranks_ = 1;
};

}  // namespace pfb
