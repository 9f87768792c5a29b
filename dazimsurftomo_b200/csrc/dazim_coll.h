// Collectives of the row-distributed LSMR (dazim_comm.cu provides them over NCCL): in-place sums over the ranks, issued
// on the solver's stream.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace dzl {
struct Coll {
  void* ctx;
  int (*sum_f32)(void*, float*, size_t, cudaStream_t);
  int (*sum_f64)(void*, double*, size_t, cudaStream_t);
  long long m_total;      // rows of the whole system
};
}  // namespace dzl
namespace dzc {
int sum_f32(void* ctx, float* buf, size_t n, cudaStream_t st);
int sum_f64(void* ctx, double* buf, size_t n, cudaStream_t st);
}  // namespace dzc
