// Sparse least-squares stage that consumes the G matrix: LSMR (lsmrModule.f90:36) with the
// reference's COO products (aprod.f90:7), single precision like the reference
// (lsmrDataModule.f90:21), for sm_100a.
//
// The two products of an iteration, y += A x and x += A^T y, are the only O(nnz) work and are
// HBM-bound (8 B per non-zero: value + index; the gathered vector lives in L2).  The COO triplets are
// turned once into CSR (for A x) and CSC (for A^T y) with a stable radix sort, so both products are
// race-free, deterministic reductions instead of atomics: one warp per row for A x; for A^T y one warp per SEGMENT of a
// column (<= 4 096 entries) and an in-order add of a column's partials (a column of G is long and there are few).
// Vector norms are two-stage deterministic reductions accumulated in double.  The scalar recurrences of LSMR
// (lsmrModule.f90:470-640) run ON THE DEVICE, in float and in the Fortran's operation order, inside the single-thread
// tails of the norm reductions; every kernel of an iteration starts with "if (state->istop) return", so the host
// enqueues iterations in batches (one CUDA graph of an iteration, replayed) and synchronises once per batch instead
// of three times per iteration.  Iterations enqueued past the stopping one are no-ops: istop / itn / x are exactly
// those of the synchronous loop.
#include "../../include/dazim_b200.h"
#include "dazim_coll.h"
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dzl {

#define LCK(x)                                                   \
  do {                                                           \
    cudaError_t e_ = (x);                                        \
    if (e_ != cudaSuccess) return DAZIM_ECUDA + (int)e_;         \
  } while (0)

template <class T>
struct Buf {
  T* p = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t alloc(size_t n, cudaStream_t s) {
    st = s;
    return cudaMallocAsync((void**)&p, std::max<size_t>(n, 1) * sizeof(T), s);
  }
  ~Buf() { if (p) cudaFreeAsync(p, st); }
};

__global__ void k_iota(int* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}
// counts[key[i]-1] += 1 (1-based keys)
__global__ void k_hist(const int* __restrict__ key, long long n, int* __restrict__ counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&counts[key[i] - 1], 1);
}
// out_idx[i] = other[perm[i]] - 1 ; out_val[i] = val[perm[i]]
__global__ void k_gather(const int* __restrict__ perm, const int* __restrict__ other, const float* __restrict__ val,
                         long long n, int* __restrict__ out_idx, float* __restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int p = perm[i];
    out_idx[i] = other[p] - 1;
    out_val[i] = val[p];
  }
}

// y[r] += sum_k val[k] * x[idx[k]], one warp per row (aprod mode 1 on CSR; mode 2 on CSC)
__global__ void __launch_bounds__(256) k_spmv_add(int nrow, const long long* __restrict__ ptr,
                                                   const int* __restrict__ idx, const float* __restrict__ val,
                                                   const float* __restrict__ x, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= nrow) return;
  const long long b = ptr[w], e = ptr[w + 1];
  float acc = 0.0f;
  long long k = b + lane;
  for (; k + 96 < e; k += 128) {
    const float v0 = val[k], v1 = val[k + 32], v2 = val[k + 64], v3 = val[k + 96];
    const int i0 = idx[k], i1 = idx[k + 32], i2 = idx[k + 64], i3 = idx[k + 96];
    acc += v0 * x[i0];
    acc += v1 * x[i1];
    acc += v2 * x[i2];
    acc += v3 * x[i3];
  }
  for (; k < e; k += 32) acc += val[k] * x[idx[k]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && e > b) y[w] = y[w] + acc;
}

// Segmented product for the TRANSPOSED matrix (CSC): a column of G holds the entries of one model cell in every ray that
// sees it -- 15 000 entries per column and only a few thousand columns at test3 size, so one warp per column leaves the
// chip idle and runs a long dependent loop (measured 632 us against 149 us for the row product on the same 94 M
// entries).  Columns are cut into segments of at most `seg` entries (every column boundary is a segment boundary),
// one warp reduces one segment, and a second kernel adds the partials of a column in segment order: same work,
// deterministic, 180 k warps instead of 2-6 k.  state == nullptr: unconditional (set-up product).
struct LsmrState;
__device__ inline bool lsmr_skip(const LsmrState* S, int need_beta);
__global__ void __launch_bounds__(256) k_spmv_seg(int nseg, const long long* __restrict__ seg_off, const int* __restrict__ idx,
                                                   const float* __restrict__ val, const float* __restrict__ x,
                                                   float* __restrict__ partial, const LsmrState* __restrict__ S, int need_beta) {
  if (S && lsmr_skip(S, need_beta)) return;
  const int lane = threadIdx.x & 31;
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= nseg) return;
  const long long b = seg_off[w], e = seg_off[w + 1];
  float acc = 0.0f;
  long long k = b + lane;
  for (; k + 96 < e; k += 128) {
    const float v0 = val[k], v1 = val[k + 32], v2 = val[k + 64], v3 = val[k + 96];
    const int i0 = idx[k], i1 = idx[k + 32], i2 = idx[k + 64], i3 = idx[k + 96];
    acc += v0 * x[i0];
    acc += v1 * x[i1];
    acc += v2 * x[i2];
    acc += v3 * x[i3];
  }
  for (; k < e; k += 32) acc += val[k] * x[idx[k]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) partial[w] = acc;
}
__global__ void k_seg_reduce_add(int ncol, const int* __restrict__ colseg, const float* __restrict__ partial,
                                 float* __restrict__ y, const LsmrState* __restrict__ S, int need_beta) {
  if (S && lsmr_skip(S, need_beta)) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const int b = colseg[c], e = colseg[c + 1];
  if (e <= b) return;
  float acc = partial[b];
  for (int s = b + 1; s < e; ++s) acc += partial[s];
  y[c] = y[c] + acc;
}

// stage 1 of a deterministic reduction: partial[b] = sum over the block's slice of a[i]*b[i] (double)
__global__ void __launch_bounds__(256) k_dot_partial(const float* __restrict__ a, const float* __restrict__ b2, int n,
                                                      double* __restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    acc += (double)a[i] * (double)b2[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += sh[i];
    partial[blockIdx.x] = s;
  }
}
// stage 2: out[0] = sum partial (fixed order); out_f = (float) of it or of its square root
__global__ void k_dot_final(const double* __restrict__ partial, int nb, int take_sqrt, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nb; ++i) s += partial[i];
    out[0] = (float)(take_sqrt ? sqrt(s) : s);
  }
}
__global__ void k_scal(float* __restrict__ x, int n, float a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = a * x[i];
}
// Local reorthogonalisation (lsmrModule.f90:740-747): for q = 1..lim, in order, d = v . V_q ; v = v - d V_q.
// The steps depend on each other, n is the model size (10^3 - 10^5) and lim reaches n/4 in the isotropic
// inversion (Main_Jt.f90:547), so three launches per step are pure launch latency: one CTA runs the whole chain,
// every thread owning the same elements of v in all steps (dot products accumulated in double, fixed tree).
__global__ void __launch_bounds__(1024) k_reorth(float* __restrict__ v, const float* __restrict__ localV, int n, int lim) {
  __shared__ double sh[32];
  __shared__ float sh_d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < lim; ++q) {
    const float* __restrict__ lq = localV + (size_t)q * n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += (double)v[i] * (double)lq[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      double t = sh[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sh_d = (float)t;
    }
    __syncthreads();
    const float d = sh_d;
    for (int i = threadIdx.x; i < n; i += 1024) v[i] = v[i] - d * lq[i];
  }
}

// hbar = h - c1*hbar ; x = x + c2*hbar ; h = v - c3*h   (lsmrModule.f90:539-541)
__global__ void k_update(float* __restrict__ hbar, float* __restrict__ h, float* __restrict__ x,
                         const float* __restrict__ v, int n, float c1, float c2, float c3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float hb = h[i] - c1 * hbar[i];
    hbar[i] = hb;
    x[i] = x[i] + c2 * hb;
    h[i] = v[i] - c3 * h[i];
  }
}

// ---- device-resident LSMR state (scalars of lsmrModule.f90:36-640) ----
struct LsmrState {
  float alpha, beta, rho, rhobar, cbar, sbar, zeta, zetabar, alphabar, betadd, betad, rhodold, tautildeold, thetatilde, d;
  float normA2, maxrbar, minrbar, normb, normA, condA, normx, normr, normAr;
  float c1, c2, c3;                      // coefficients of the hbar / x / h update
  float damp, atol, btol, ctol;
  int itn, istop, itnlim;
  int beta_pos;                          // beta > 0 in this iteration (lsmrModule.f90:486)
  int localVecs, localPointer, queueFull;
  float reorth_d;                        // v . V_q of the reorthogonalisation step in flight (large-n path)
};

__device__ inline bool lsmr_skip(const LsmrState* S, int need_beta) { return S->istop || (need_beta && !S->beta_pos); }

__device__ __host__ inline float d2norm(float a, float b) {      // lsmrModule.f90:686-711
  const float scale = fabsf(a) + fabsf(b);
  if (scale == 0.0f) return 0.0f;
  const float ra = a / scale, rb = b / scale;
  return scale * sqrtf(ra * ra + rb * rb);
}

// x = a * x with a taken from the state: mode 0: -alpha, 1: 1/beta (if beta > 0), 2: -beta (if beta > 0), 3: 1/alpha (if alpha > 0)
__global__ void k_scal_state(float* __restrict__ x, int n, const LsmrState* __restrict__ S, int mode) {
  if (S->istop) return;
  float a;
  if (mode == 0) a = -S->alpha;
  else if (mode == 1) { if (!S->beta_pos) return; a = 1.0f / S->beta; }
  else if (mode == 2) { if (!S->beta_pos) return; a = -S->beta; }
  else { if (!S->beta_pos || !(S->alpha > 0.0f)) return; a = 1.0f / S->alpha; }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = a * x[i];
}
__global__ void __launch_bounds__(256) k_spmv_add_state(int nrow, const long long* __restrict__ ptr, const int* __restrict__ idx,
                                                         const float* __restrict__ val, const float* __restrict__ x,
                                                         float* __restrict__ y, const LsmrState* __restrict__ S, int need_beta) {
  if (S->istop || (need_beta && !S->beta_pos)) return;
  const int lane = threadIdx.x & 31;
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= nrow) return;
  const long long b = ptr[w], e = ptr[w + 1];
  float acc = 0.0f;
  long long k = b + lane;
  for (; k + 96 < e; k += 128) {
    const float v0 = val[k], v1 = val[k + 32], v2 = val[k + 64], v3 = val[k + 96];
    const int i0 = idx[k], i1 = idx[k + 32], i2 = idx[k + 64], i3 = idx[k + 96];
    acc += v0 * x[i0];
    acc += v1 * x[i1];
    acc += v2 * x[i2];
    acc += v3 * x[i3];
  }
  for (; k < e; k += 32) acc += val[k] * x[idx[k]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && e > b) y[w] = y[w] + acc;
}
__global__ void __launch_bounds__(256) k_dot_partial_state(const float* __restrict__ a, int n, double* __restrict__ partial,
                                                            const LsmrState* __restrict__ S, int need_beta) {
  if (S->istop || (need_beta && !S->beta_pos)) return;
  __shared__ double sh[8];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    acc += (double)a[i] * (double)a[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += sh[i];
    partial[blockIdx.x] = s;
  }
}
// sum of the block partials by ONE WARP in a fixed order (lane l adds partials l, l+32, ...; then a fixed shuffle tree):
// deterministic, and 20 us shorter per norm than a single thread walking 592 doubles
__device__ inline float final_norm(const double* partial, int nb) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int i = lane; i < nb; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return (float)sqrt(s);
}
// beta = ||u|| (lsmrModule.f90:484) + the bookkeeping of the local reorthogonalisation queue (:489-497)
__global__ void k_tail_beta(const double* __restrict__ partial, int nb, LsmrState* S) {
  if (blockIdx.x != 0 || S->istop) return;
  const float nrm = final_norm(partial, nb);
  if (threadIdx.x != 0) return;
  S->itn = S->itn + 1;
  S->beta = nrm;
  S->beta_pos = S->beta > 0.0f ? 1 : 0;
  if (S->beta_pos && S->localVecs > 0) {
    if (S->localPointer < S->localVecs) S->localPointer = S->localPointer + 1;
    else { S->localPointer = 1; S->queueFull = 1; }
  }
}
__global__ void k_store_local(float* __restrict__ localV, const float* __restrict__ v, int n, const LsmrState* __restrict__ S) {
  if (S->istop || !S->beta_pos || S->localVecs <= 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) localV[(size_t)(S->localPointer - 1) * n + i] = v[i];
}
// alpha = ||v|| (:505) and every scalar recurrence up to the coefficients of the vector update (:514-541) and the
// norm estimates that do not need ||x|| (:548-590)
__global__ void k_tail_alpha(const double* __restrict__ partial, int nb, LsmrState* S) {
  if (blockIdx.x != 0 || S->istop) return;
  const float nrm = S->beta_pos ? final_norm(partial, nb) : 0.0f;     // beta_pos is warp-uniform
  if (threadIdx.x != 0) return;
  if (S->beta_pos) S->alpha = nrm;
  const float alpha = S->alpha, beta = S->beta, damp = S->damp;
  const float alphahat = d2norm(S->alphabar, damp);
  const float chat = S->alphabar / alphahat, shat = damp / alphahat;
  const float rhoold = S->rho;
  const float rho = d2norm(alphahat, beta);
  const float cc = alphahat / rho, s = beta / rho;
  const float thetanew = s * alpha;
  S->alphabar = cc * alpha;
  const float rhobarold = S->rhobar, zetaold = S->zeta;
  const float thetabar = S->sbar * rho, rhotemp = S->cbar * rho;
  const float rhobar = d2norm(S->cbar * rho, thetanew);
  const float cbar = S->cbar * rho / rhobar;
  const float sbar = thetanew / rhobar;
  const float zeta = cbar * S->zetabar;
  const float zetabar = -sbar * S->zetabar;
  S->c1 = thetabar * rho / (rhoold * rhobarold);
  S->c2 = zeta / (rho * rhobar);
  S->c3 = thetanew / rho;
  const float betaacute = chat * S->betadd, betacheck = -shat * S->betadd;
  const float betahat = cc * betaacute;
  const float betadd = -s * betaacute;
  const float thetatildeold = S->thetatilde;
  const float rhotildeold = d2norm(S->rhodold, thetabar);
  const float ctildeold = S->rhodold / rhotildeold, stildeold = thetabar / rhotildeold;
  const float thetatilde = stildeold * rhobar;
  const float rhodold = ctildeold * rhobar;
  const float betad = -stildeold * S->betad + ctildeold * betahat;
  const float tautildeold = (zetaold - thetatildeold * S->tautildeold) / rhotildeold;
  const float taud = (zeta - thetatilde * tautildeold) / rhodold;
  const float d = S->d + betacheck * betacheck;
  S->normr = sqrtf(d + (betad - taud) * (betad - taud) + betadd * betadd);
  float normA2 = S->normA2 + beta * beta;
  S->normA = sqrtf(normA2);
  normA2 = normA2 + alpha * alpha;
  S->normA2 = normA2;
  S->maxrbar = fmaxf(S->maxrbar, rhobarold);
  if (S->itn > 1) S->minrbar = fminf(S->minrbar, rhobarold);
  S->condA = fmaxf(S->maxrbar, rhotemp) / fminf(S->minrbar, rhotemp);
  S->normAr = fabsf(zetabar);
  S->rho = rho; S->rhobar = rhobar; S->cbar = cbar; S->sbar = sbar; S->zeta = zeta; S->zetabar = zetabar;
  S->betadd = betadd; S->thetatilde = thetatilde; S->rhodold = rhodold; S->betad = betad; S->tautildeold = tautildeold;
  S->d = d;
}
// hbar = h - c1*hbar ; x = x + c2*hbar ; h = v - c3*h   (lsmrModule.f90:539-541)
__global__ void k_update_state(float* __restrict__ hbar, float* __restrict__ h, float* __restrict__ x,
                               const float* __restrict__ v, int n, const LsmrState* __restrict__ S) {
  if (S->istop) return;
  const float c1 = S->c1, c2 = S->c2, c3 = S->c3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float hb = h[i] - c1 * hbar[i];
    hbar[i] = hb;
    x[i] = x[i] + c2 * hb;
    h[i] = v[i] - c3 * h[i];
  }
}
// normx = ||x|| (:592) and the stopping tests (:596-640)
__global__ void k_tail_normx(const double* __restrict__ partial, int nb, LsmrState* S) {
  if (blockIdx.x != 0 || S->istop) return;
  const float nrm = final_norm(partial, nb);
  if (threadIdx.x != 0) return;
  S->normx = nrm;
  const float normA = S->normA, normx = S->normx, normb = S->normb;
  const float test1 = S->normr / normb, test2 = S->normAr / (normA * S->normr), test3 = 1.0f / S->condA;
  const float t1 = test1 / (1.0f + normA * normx / normb);
  const float rtol = S->btol + S->atol * normA * normx / normb;
  int istop = 0;
  if (S->itn >= S->itnlim) istop = 7;
  if (1.0f + test3 <= 1.0f) istop = 6;
  if (1.0f + test2 <= 1.0f) istop = 5;
  if (1.0f + t1 <= 1.0f) istop = 4;
  if (test3 <= S->ctol) istop = 3;
  if (test2 <= S->atol) istop = 2;
  if (test1 <= rtol) istop = 1;
  S->istop = istop;
}
__global__ void __launch_bounds__(1024) k_reorth_state(float* __restrict__ v, const float* __restrict__ localV, int n,
                                                        const LsmrState* __restrict__ S) {
  if (S->istop || !S->beta_pos || S->localVecs <= 0) return;
  const int lim = S->queueFull ? S->localVecs : S->localPointer;
  __shared__ double sh[32];
  __shared__ float sh_d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < lim; ++q) {
    const float* __restrict__ lq = localV + (size_t)q * n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += (double)v[i] * (double)lq[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      double t = sh[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sh_d = (float)t;
    }
    __syncthreads();
    const float d = sh_d;
    for (int i = threadIdx.x; i < n; i += 1024) v[i] = v[i] - d * lq[i];
  }
}

// Large-n path of the local reorthogonalisation (n > 32 768: one CTA cannot stream 10 x n floats fast enough -- measured
// 1.9 ms of a 3.2 ms iteration at n = 960 000): the same chain, step q = dot (grid) -> tail -> axpy (grid), all guarded by
// the device state (q >= lim: no-op), no host involvement.
__device__ inline bool reorth_off(const LsmrState* S, int q) {
  if (S->istop || !S->beta_pos || S->localVecs <= 0) return true;
  const int lim = S->queueFull ? S->localVecs : S->localPointer;
  return q >= lim;
}
__global__ void __launch_bounds__(256) k_reorth_dot(const float* __restrict__ v, const float* __restrict__ lq, int n,
                                                     double* __restrict__ partial, const LsmrState* __restrict__ S, int q) {
  if (reorth_off(S, q)) return;
  __shared__ double sh[8];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += (double)v[i] * (double)lq[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void k_reorth_tail(const double* __restrict__ partial, int nb, LsmrState* S, int q) {
  if (blockIdx.x != 0 || reorth_off(S, q)) return;
  const int lane = threadIdx.x & 31;
  double t = 0.0;
  for (int i = lane; i < nb; i += 32) t += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) S->reorth_d = (float)t;
}
__global__ void k_reorth_axpy(float* __restrict__ v, const float* __restrict__ lq, int n, const LsmrState* __restrict__ S, int q) {
  if (reorth_off(S, q)) return;
  const float d = S->reorth_d;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] - d * lq[i];
}

// row-distributed solve: the block partials of a norm over the LOCAL rows fold into partial[0] (fixed order), the ranks'
// sums are added by the all-reduce, and the tails then read ONE partial
__global__ void k_fold_partial(double* __restrict__ partial, int nb) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int i = lane; i < nb; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) partial[0] = s;
}
// v = v + t (the all-reduced A^T u of the row blocks)
__global__ void k_add_state(float* __restrict__ v, const float* __restrict__ t, int n, const LsmrState* __restrict__ S, int need_beta) {
  if (S && (S->istop || (need_beta && !S->beta_pos))) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] + t[i];
}

struct Ctx {
  cudaStream_t st;
  double* partial;
  float* scal;       // device scalar
  int nb;
};

static int norm2(Ctx& c, const float* a, int n, float* out_host) {
  const int nb = std::min(c.nb, std::max(1, (n + 255) / 256));
  k_dot_partial<<<nb, 256, 0, c.st>>>(a, a, n, c.partial);
  k_dot_final<<<1, 32, 0, c.st>>>(c.partial, nb, 1, c.scal);
  LCK(cudaMemcpyAsync(out_host, c.scal, sizeof(float), cudaMemcpyDeviceToHost, c.st));
  LCK(cudaStreamSynchronize(c.st));
  return 0;
}

// events / graphs that are released on every exit path
struct Guard {
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  ~Guard() {
    for (auto e : ev) if (e) cudaEventDestroy(e);
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
};

// 1-based indices of the caller must lie in [1, lim]: a wrong index would corrupt device memory silently
__global__ void k_minmax(const int* __restrict__ a, long long n, int lim, int* __restrict__ bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (a[i] < 1 || a[i] > lim)) atomicOr(bad, 1);
}

// COO (1-based) -> compressed structure over `key` (row for CSR, col for CSC)
static int compress(cudaStream_t st, long long nnz, int nkey, const int* d_key, const int* d_other, const float* d_val,
                    long long* d_ptr, int* d_idx, float* d_v) {
  const unsigned nbk = (unsigned)((nnz + 255) / 256);
  Buf<int> keys_out, perm_in, perm_out, counts;
  LCK(keys_out.alloc(nnz, st)); LCK(perm_in.alloc(nnz, st)); LCK(perm_out.alloc(nnz, st)); LCK(counts.alloc(nkey + 1, st));
  if (nnz > 0) k_iota<<<nbk, 256, 0, st>>>(perm_in.p, nnz);
  int bits = 1;
  while ((1ll << bits) <= nkey) ++bits;
  size_t tmp_bytes = 0;
  LCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, keys_out.p, perm_in.p, perm_out.p, (long long)nnz, 0, bits, st));
  Buf<unsigned char> tmp;
  LCK(tmp.alloc(tmp_bytes, st));
  LCK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, d_key, keys_out.p, perm_in.p, perm_out.p, (long long)nnz, 0, bits, st));
  if (nnz > 0) k_gather<<<nbk, 256, 0, st>>>(perm_out.p, d_other, d_val, nnz, d_idx, d_v);
  LCK(cudaMemsetAsync(counts.p, 0, sizeof(int) * (nkey + 1), st));
  if (nnz > 0) k_hist<<<nbk, 256, 0, st>>>(d_key, nnz, counts.p);
  size_t scan_bytes = 0;
  LCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts.p, d_ptr, nkey + 1, st));
  Buf<unsigned char> tmp2;
  LCK(tmp2.alloc(scan_bytes, st));
  LCK(cub::DeviceScan::ExclusiveSum(tmp2.p, scan_bytes, counts.p, d_ptr, nkey + 1, st));
  LCK(cudaGetLastError());
  LCK(cudaStreamSynchronize(st));
  return 0;
}

// coo_on_device: row/col/rw already live in HBM (the G row block of a plan): no upload, G never leaves the GPU
int lsmr_solve(cudaStream_t st, int m, int n, long long nnz, const int* row, const int* col, const float* rw,
               const float* b, float damp, float atol, float btol, float conlim, int itnlim, int localSize, float* x,
               dazim_lsmr_info* info, bool coo_on_device, const Coll* coll) {
  if (m < 1 || n < 1 || nnz < 0 || nnz >= (1ll << 31) || !row || !col || !rw || !b || !x || !info) return DAZIM_EBADARG;
  Guard G;
  for (int i = 0; i < 3; ++i) LCK(cudaEventCreate(&G.ev[i]));
  cudaEvent_t e0 = G.ev[0], e1 = G.ev[1], e2 = G.ev[2];
  // ---- upload + CSR / CSC ----
  Buf<int> d_row, d_col, csr_idx, csc_idx, d_bad;
  Buf<float> d_val, csr_val, csc_val;
  Buf<long long> csr_ptr, csc_ptr;
  if (!coo_on_device) { LCK(d_row.alloc(nnz, st)); LCK(d_col.alloc(nnz, st)); LCK(d_val.alloc(nnz, st)); }
  LCK(csr_idx.alloc(nnz, st)); LCK(csc_idx.alloc(nnz, st)); LCK(csr_val.alloc(nnz, st)); LCK(csc_val.alloc(nnz, st));
  LCK(csr_ptr.alloc((size_t)m + 1, st)); LCK(csc_ptr.alloc((size_t)n + 1, st)); LCK(d_bad.alloc(1, st));
  LCK(cudaEventRecord(e0, st));
  if (nnz > 0 && !coo_on_device) {
    LCK(cudaMemcpyAsync(d_row.p, row, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    LCK(cudaMemcpyAsync(d_col.p, col, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    LCK(cudaMemcpyAsync(d_val.p, rw, nnz * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  const int* k_row = coo_on_device ? row : d_row.p;
  const int* k_col = coo_on_device ? col : d_col.p;
  const float* k_val = coo_on_device ? rw : d_val.p;
  if (nnz > 0) {
    // the indices come from the caller (dazim_lsmr, __lsmrmodule_MOD_lsmr): validate before they index device memory
    int bad = 0;
    LCK(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    const unsigned nbk = (unsigned)((nnz + 255) / 256);
    k_minmax<<<nbk, 256, 0, st>>>(k_row, nnz, m, d_bad.p);
    k_minmax<<<nbk, 256, 0, st>>>(k_col, nnz, n, d_bad.p);
    LCK(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LCK(cudaStreamSynchronize(st));
    if (bad) return DAZIM_EBADARG;
  }
  int rc = compress(st, nnz, m, k_row, k_col, k_val, csr_ptr.p, csr_idx.p, csr_val.p);
  if (rc) return rc;
  rc = compress(st, nnz, n, k_col, k_row, k_val, csc_ptr.p, csc_idx.p, csc_val.p);
  if (rc) return rc;
  // ---- segments of the transposed product (k_spmv_seg): at most `seg` entries each, never across a column boundary ----
  Buf<long long> seg_off;
  Buf<int> colseg;
  Buf<float> seg_partial;
  int nseg = 0;
  {
    std::vector<long long> hptr((size_t)n + 1);
    LCK(cudaMemcpyAsync(hptr.data(), csc_ptr.p, sizeof(long long) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    LCK(cudaStreamSynchronize(st));
    long long seg = (nnz / 65536 + 127) / 128 * 128;           // ~64 k segments on a big system
    seg = std::max<long long>(256, std::min<long long>(4096, seg));
    if (const char* e = getenv("DAZIM_LSMR_SEG")) seg = std::max(32, atoi(e));
    std::vector<long long> hoff;
    std::vector<int> hcs((size_t)n + 1);
    hoff.reserve((size_t)(nnz / seg) + n + 2);
    for (int c = 0; c < n; ++c) {
      hcs[c] = (int)hoff.size();
      for (long long b = hptr[c]; b < hptr[c + 1]; b += seg) hoff.push_back(b);
    }
    hcs[n] = (int)hoff.size();
    nseg = (int)hoff.size();
    hoff.push_back(hptr[n]);
    // a segment ends where the next one starts; the last segment of a column ends at the column's end = the next start
    LCK(seg_off.alloc(hoff.size(), st)); LCK(colseg.alloc((size_t)n + 1, st)); LCK(seg_partial.alloc((size_t)std::max(nseg, 1), st));
    LCK(cudaMemcpyAsync(seg_off.p, hoff.data(), sizeof(long long) * hoff.size(), cudaMemcpyHostToDevice, st));
    LCK(cudaMemcpyAsync(colseg.p, hcs.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
    LCK(cudaStreamSynchronize(st));
  }
  const unsigned gws = (unsigned)(((long long)std::max(nseg, 1) * 32 + 255) / 256);
  // ---- vectors ----
  const long long m_all = coll ? coll->m_total : (long long)m;      // the reference sizes the window by the whole system
  const int localVecs = std::max(0, (int)std::min<long long>(localSize, std::min<long long>(m_all, n)));
  Buf<float> u, v, h, hbar, dx, localV, scal, vt;
  Buf<double> partial;
  Buf<LsmrState> dS;
  LCK(u.alloc(m, st)); LCK(v.alloc(n, st)); LCK(h.alloc(n, st)); LCK(hbar.alloc(n, st)); LCK(dx.alloc(n, st));
  if (coll) LCK(vt.alloc(n, st));
  LCK(localV.alloc((size_t)localVecs * n, st)); LCK(scal.alloc(4, st)); LCK(partial.alloc(1024, st)); LCK(dS.alloc(1, st));
  Ctx c{st, partial.p, scal.p, 592};
  const unsigned gm = (unsigned)((m + 255) / 256), gn = (unsigned)((n + 255) / 256);
  const unsigned gwm = (unsigned)(((long long)m * 32 + 255) / 256);
  LCK(cudaMemcpyAsync(u.p, b, sizeof(float) * m, cudaMemcpyDefault, st));     // b may live on the host or in HBM
  LCK(cudaMemsetAsync(v.p, 0, sizeof(float) * n, st));
  LCK(cudaMemsetAsync(dx.p, 0, sizeof(float) * n, st));
  LCK(cudaMemsetAsync(hbar.p, 0, sizeof(float) * n, st));
  LCK(cudaEventRecord(e1, st));
  info->itn = 0; info->istop = 0; info->normA = 0; info->condA = 0; info->normx = 0;
  float alpha = 0.0f, beta = 0.0f;
  if (!coll) {
    if ((rc = norm2(c, u.p, m, &beta))) return rc;
  } else {
    // ||b|| over the row blocks of every rank
    const int nb0 = std::min(c.nb, std::max(1, (m + 255) / 256));
    k_dot_partial<<<nb0, 256, 0, st>>>(u.p, u.p, m, partial.p);
    k_fold_partial<<<1, 32, 0, st>>>(partial.p, nb0);
    if ((rc = coll->sum_f64(coll->ctx, partial.p, 1, st))) return rc;
    k_dot_final<<<1, 32, 0, st>>>(partial.p, 1, 1, scal.p);
    LCK(cudaMemcpyAsync(&beta, scal.p, sizeof(float), cudaMemcpyDeviceToHost, st));
    LCK(cudaStreamSynchronize(st));
  }
  if (beta > 0.0f) {
    k_scal<<<gm, 256, 0, st>>>(u.p, m, 1.0f / beta);
    k_spmv_seg<<<gws, 256, 0, st>>>(nseg, seg_off.p, csc_idx.p, csc_val.p, u.p, seg_partial.p, nullptr, 0);   // v = v + A^T u
    k_seg_reduce_add<<<gn, 256, 0, st>>>(n, colseg.p, seg_partial.p, v.p, nullptr, 0);
    if (coll && (rc = coll->sum_f32(coll->ctx, v.p, (size_t)n, st))) return rc;      // v was 0: the sum of the blocks' A^T u
    if ((rc = norm2(c, v.p, n, &alpha))) return rc;
  }
  if (alpha > 0.0f) k_scal<<<gn, 256, 0, st>>>(v.p, n, 1.0f / alpha);
  float normAr = alpha * beta;
  info->normAr = normAr; info->normr = beta;
  LsmrState hs;
  std::memset(&hs, 0, sizeof(hs));
  hs.normr = beta;
  if (normAr != 0.0f) {
    const bool localOrtho = localVecs > 0;
    if (localOrtho) LCK(cudaMemcpyAsync(localV.p, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    LCK(cudaMemcpyAsync(h.p, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    // initial scalars (lsmrModule.f90:420-460)
    hs.alpha = alpha; hs.beta = beta; hs.zetabar = alpha * beta; hs.alphabar = alpha; hs.rho = 1; hs.rhobar = 1; hs.cbar = 1;
    hs.sbar = 0; hs.betadd = beta; hs.betad = 0; hs.rhodold = 1; hs.tautildeold = 0; hs.thetatilde = 0; hs.zeta = 0; hs.d = 0;
    hs.normA2 = alpha * alpha; hs.maxrbar = 0.0f; hs.minrbar = 1e+30f; hs.normb = beta;
    hs.damp = damp; hs.atol = atol; hs.btol = btol; hs.ctol = conlim > 0.0f ? 1.0f / conlim : 0.0f;
    hs.itn = 0; hs.istop = 0; hs.itnlim = itnlim; hs.beta_pos = 1;
    hs.localVecs = localVecs; hs.localPointer = localOrtho ? 1 : 0; hs.queueFull = 0;
    LCK(cudaMemcpyAsync(dS.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
    LCK(cudaStreamSynchronize(st));
    LsmrState* S = dS.p;
    const int nbm = std::min(c.nb, std::max(1, (m + 255) / 256)), nbn = std::min(c.nb, std::max(1, (n + 255) / 256));
    int coll_rc = 0;
    auto enqueue_iteration = [&]() {
      k_scal_state<<<gm, 256, 0, st>>>(u.p, m, S, 0);                                                  // u = -alpha u
      k_spmv_add_state<<<gwm, 256, 0, st>>>(m, csr_ptr.p, csr_idx.p, csr_val.p, v.p, u.p, S, 0);       // u = u + A v
      k_dot_partial_state<<<nbm, 256, 0, st>>>(u.p, m, partial.p, S, 0);
      if (coll) {                                                                                      // ||u||^2 over all row blocks
        k_fold_partial<<<1, 32, 0, st>>>(partial.p, nbm);
        coll_rc |= coll->sum_f64(coll->ctx, partial.p, 1, st);
      }
      k_tail_beta<<<1, 32, 0, st>>>(partial.p, coll ? 1 : nbm, S);                                     // beta, queue pointer
      k_scal_state<<<gm, 256, 0, st>>>(u.p, m, S, 1);                                                  // u = u / beta
      if (localOrtho) k_store_local<<<gn, 256, 0, st>>>(localV.p, v.p, n, S);
      k_scal_state<<<gn, 256, 0, st>>>(v.p, n, S, 2);                                                  // v = -beta v
      if (!coll) {
        k_spmv_seg<<<gws, 256, 0, st>>>(nseg, seg_off.p, csc_idx.p, csc_val.p, u.p, seg_partial.p, S, 1);   // v = v + A^T u
        k_seg_reduce_add<<<gn, 256, 0, st>>>(n, colseg.p, seg_partial.p, v.p, S, 1);
      } else {
        // t = A_local^T u_local, summed over the ranks (the one n-vector exchange of the iteration), v = v + t.  The
        // all-reduce is not guarded by the device state: after a stop it adds zeros that nobody reads.
        cudaMemsetAsync(vt.p, 0, sizeof(float) * n, st);
        k_spmv_seg<<<gws, 256, 0, st>>>(nseg, seg_off.p, csc_idx.p, csc_val.p, u.p, seg_partial.p, S, 1);
        k_seg_reduce_add<<<gn, 256, 0, st>>>(n, colseg.p, seg_partial.p, vt.p, S, 1);
        coll_rc |= coll->sum_f32(coll->ctx, vt.p, (size_t)n, st);
        k_add_state<<<gn, 256, 0, st>>>(v.p, vt.p, n, S, 1);
      }
      if (localOrtho && (n <= 32768 || localVecs > 64)) k_reorth_state<<<1, 1024, 0, st>>>(v.p, localV.p, n, S);
      else if (localOrtho)
        for (int q = 0; q < localVecs; ++q) {
          k_reorth_dot<<<nbn, 256, 0, st>>>(v.p, localV.p + (size_t)q * n, n, partial.p, S, q);
          k_reorth_tail<<<1, 32, 0, st>>>(partial.p, nbn, S, q);
          k_reorth_axpy<<<gn, 256, 0, st>>>(v.p, localV.p + (size_t)q * n, n, S, q);
        }
      k_dot_partial_state<<<nbn, 256, 0, st>>>(v.p, n, partial.p, S, 1);
      k_tail_alpha<<<1, 32, 0, st>>>(partial.p, nbn, S);                                               // alpha + recurrences
      k_scal_state<<<gn, 256, 0, st>>>(v.p, n, S, 3);                                                  // v = v / alpha
      k_update_state<<<gn, 256, 0, st>>>(hbar.p, h.p, dx.p, v.p, n, S);
      k_dot_partial_state<<<nbn, 256, 0, st>>>(dx.p, n, partial.p, S, 0);
      k_tail_normx<<<1, 32, 0, st>>>(partial.p, nbn, S);                                               // ||x||, stopping tests
    };
    // one iteration captured as a CUDA graph, replayed `batch` times between host checks of the stop flag
    bool use_graph = getenv("DAZIM_LSMR_NOGRAPH") == nullptr;
    if (use_graph) {
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueue_iteration();
        if (cudaStreamEndCapture(st, &G.graph) != cudaSuccess || !G.graph ||
            cudaGraphInstantiate(&G.exec, G.graph, 0) != cudaSuccess) use_graph = false;
      } else use_graph = false;
      if (!use_graph) cudaGetLastError();
    }
    int batch = 8;
    if (const char* e = getenv("DAZIM_LSMR_BATCH")) batch = std::max(1, atoi(e));
    int ist_itn[2] = {0, 0};
    for (;;) {
      for (int k = 0; k < batch; ++k) {
        if (use_graph) LCK(cudaGraphLaunch(G.exec, st));
        else enqueue_iteration();
      }
      LCK(cudaMemcpyAsync(&ist_itn[0], &S->istop, sizeof(int), cudaMemcpyDeviceToHost, st));
      LCK(cudaMemcpyAsync(&ist_itn[1], &S->itn, sizeof(int), cudaMemcpyDeviceToHost, st));
      LCK(cudaStreamSynchronize(st));
      LCK(cudaGetLastError());
      if (coll_rc) return DAZIM_ENCCL;
      if (ist_itn[0] != 0) break;
    }
    LCK(cudaMemcpyAsync(&hs, S, sizeof(hs), cudaMemcpyDeviceToHost, st));
    LCK(cudaStreamSynchronize(st));
    if (damp > 0.0f && hs.istop == 2) hs.istop = 3;
    normAr = hs.normAr;
  }
  LCK(cudaEventRecord(e2, st));
  LCK(cudaMemcpyAsync(x, dx.p, sizeof(float) * n, cudaMemcpyDefault, st));    // so may x
  LCK(cudaStreamSynchronize(st));
  LCK(cudaGetLastError());
  info->istop = hs.istop; info->itn = hs.itn; info->normA = hs.normA; info->condA = hs.condA; info->normr = hs.normr;
  info->normAr = normAr; info->normx = hs.normx;
  cudaEventElapsedTime(&info->setup_ms, e0, e1);
  cudaEventElapsedTime(&info->solve_ms, e1, e2);
  return DAZIM_OK;
}

}  // namespace dzl
