// Communicator for the row-distributed solver (SURVEY 8e "if the solver is later distributed by rows ... only an n-vector
// all-reduce per LSMR iteration remains", 8f-1 "keeping G resident and row-distributed").
// One process per GPU (torchrun); the collectives are NCCL all-reduces issued on the library's own stream, so they sit
// between the solver's kernels in stream order and are captured into the per-iteration CUDA graph with them.
// NCCL is bound at run time (dlopen): in a process that imported torch the already loaded libnccl.so.2 (torch's) is
// used, otherwise the system one; the library itself keeps no link-time dependency on NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

#include "../../include/dazim_b200.h"

namespace dzc {
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*get_id)(ncclUniqueId*) = nullptr;
  ncclResult_t (*init_rank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*all_reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*destroy)(ncclComm_t) = nullptr;
  const char* (*err)(ncclResult_t) = nullptr;
  bool ok = false;
};
static Nccl* nccl() {
  static Nccl N;
  static bool tried = false;
  if (tried) return N.ok ? &N : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    N.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (N.lib) break;
  }
  if (!N.lib) return nullptr;
  N.get_id = (decltype(N.get_id))dlsym(N.lib, "ncclGetUniqueId");
  N.init_rank = (decltype(N.init_rank))dlsym(N.lib, "ncclCommInitRank");
  N.all_reduce = (decltype(N.all_reduce))dlsym(N.lib, "ncclAllReduce");
  N.destroy = (decltype(N.destroy))dlsym(N.lib, "ncclCommDestroy");
  N.err = (decltype(N.err))dlsym(N.lib, "ncclGetErrorString");
  N.ok = N.get_id && N.init_rank && N.all_reduce && N.destroy;
  return N.ok ? &N : nullptr;
}
}  // namespace dzc

struct dazim_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, dev = 0;
};

static_assert(sizeof(ncclUniqueId) == DAZIM_COMM_ID_BYTES, "ncclUniqueId size");

extern "C" int dazim_comm_unique_id(unsigned char* id) {
  if (!id) return DAZIM_EBADARG;
  dzc::Nccl* N = dzc::nccl();
  if (!N) return DAZIM_ENCCL;
  ncclUniqueId u;
  if (N->get_id(&u) != ncclSuccess) return DAZIM_ENCCL;
  std::memcpy(id, &u, sizeof(u));
  return DAZIM_OK;
}

extern "C" int dazim_comm_create(int device, const unsigned char* id, int rank, int nranks, dazim_comm** out) {
  if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return DAZIM_EBADARG;
  dzc::Nccl* N = dzc::nccl();
  if (!N) return DAZIM_ENCCL;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return DAZIM_ECUDA + (int)e;
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  dazim_comm* c = new dazim_comm;
  c->rank = rank; c->nranks = nranks; c->dev = device;
  const ncclResult_t r = N->init_rank(&c->comm, nranks, u, rank);
  if (r != ncclSuccess) {
    std::fprintf(stderr, "dazim_comm_create: %s\n", N->err ? N->err(r) : "nccl error");
    delete c;
    return DAZIM_ENCCL;
  }
  // first collective here, not inside the first solve: NCCL connects its channels lazily (hundreds of ms)
  {
    float* w = nullptr;
    if (cudaMalloc(&w, sizeof(float)) == cudaSuccess) {
      cudaMemset(w, 0, sizeof(float));
      N->all_reduce(w, w, 1, ncclFloat32, ncclSum, c->comm, (cudaStream_t)0);
      cudaStreamSynchronize((cudaStream_t)0);
      cudaFree(w);
    }
  }
  *out = c;
  return DAZIM_OK;
}

extern "C" void dazim_comm_destroy(dazim_comm* c) {
  if (!c) return;
  dzc::Nccl* N = dzc::nccl();
  if (N && c->comm) N->destroy(c->comm);
  delete c;
}

extern "C" int dazim_comm_rank(const dazim_comm* c) { return c ? c->rank : -1; }
extern "C" int dazim_comm_size(const dazim_comm* c) { return c ? c->nranks : 0; }

namespace dzc {
// in-place sums over the ranks on stream st (capturable)
int sum_f32(void* ctx, float* buf, size_t n, cudaStream_t st) {
  dazim_comm* c = (dazim_comm*)ctx;
  Nccl* N = nccl();
  if (!N || !c || !c->comm) return DAZIM_ENCCL;
  return N->all_reduce(buf, buf, n, ncclFloat32, ncclSum, c->comm, st) == ncclSuccess ? DAZIM_OK : DAZIM_ENCCL;
}
int sum_f64(void* ctx, double* buf, size_t n, cudaStream_t st) {
  dazim_comm* c = (dazim_comm*)ctx;
  Nccl* N = nccl();
  if (!N || !c || !c->comm) return DAZIM_ENCCL;
  return N->all_reduce(buf, buf, n, ncclFloat64, ncclSum, c->comm, st) == ncclSuccess ? DAZIM_OK : DAZIM_ENCCL;
}
}  // namespace dzc
