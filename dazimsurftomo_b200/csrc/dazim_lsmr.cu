// Sparse least-squares stage that consumes the G matrix: LSMR (lsmrModule.f90:36) with the
// reference's COO products (aprod.f90:7), single precision like the reference
// (lsmrDataModule.f90:21), for sm_100a.
//
// The two products of an iteration, y += A x and x += A^T y, are the only O(nnz) work and are
// HBM-bound (8 B per non-zero: value + index; the gathered vector lives in L2).  The COO triplets are
// turned once into CSR (for A x) and CSC (for A^T y) with a stable radix sort, so both products are
// race-free, deterministic row/column reductions (one warp per row or column) instead of atomics.
// Vector norms are two-stage deterministic reductions accumulated in double; the scalar recurrences
// of LSMR run on the host in float exactly as in the Fortran.
#include "../../include/dazim_b200.h"
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>
#include <cstdio>
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

static float d2norm(float a, float b) {      // lsmrModule.f90:686-711
  const float scale = std::fabs(a) + std::fabs(b);
  if (scale == 0.0f) return 0.0f;
  const float ra = a / scale, rb = b / scale;
  return scale * std::sqrt(ra * ra + rb * rb);
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
               dazim_lsmr_info* info, bool coo_on_device) {
  if (m < 1 || n < 1 || nnz < 0 || nnz >= (1ll << 31) || !row || !col || !rw || !b || !x || !info) return DAZIM_EBADARG;
  cudaEvent_t e0, e1, e2;
  LCK(cudaEventCreate(&e0)); LCK(cudaEventCreate(&e1)); LCK(cudaEventCreate(&e2));
  // ---- upload + CSR / CSC ----
  Buf<int> d_row, d_col, csr_idx, csc_idx;
  Buf<float> d_val, csr_val, csc_val;
  Buf<long long> csr_ptr, csc_ptr;
  if (!coo_on_device) { LCK(d_row.alloc(nnz, st)); LCK(d_col.alloc(nnz, st)); LCK(d_val.alloc(nnz, st)); }
  LCK(csr_idx.alloc(nnz, st)); LCK(csc_idx.alloc(nnz, st)); LCK(csr_val.alloc(nnz, st)); LCK(csc_val.alloc(nnz, st));
  LCK(csr_ptr.alloc((size_t)m + 1, st)); LCK(csc_ptr.alloc((size_t)n + 1, st));
  LCK(cudaEventRecord(e0, st));
  if (nnz > 0 && !coo_on_device) {
    LCK(cudaMemcpyAsync(d_row.p, row, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    LCK(cudaMemcpyAsync(d_col.p, col, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    LCK(cudaMemcpyAsync(d_val.p, rw, nnz * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  const int* k_row = coo_on_device ? row : d_row.p;
  const int* k_col = coo_on_device ? col : d_col.p;
  const float* k_val = coo_on_device ? rw : d_val.p;
  int rc = compress(st, nnz, m, k_row, k_col, k_val, csr_ptr.p, csr_idx.p, csr_val.p);
  if (rc) return rc;
  rc = compress(st, nnz, n, k_col, k_row, k_val, csc_ptr.p, csc_idx.p, csc_val.p);
  if (rc) return rc;
  // ---- vectors ----
  const int localVecs = std::max(0, std::min(localSize, std::min(m, n)));
  Buf<float> u, v, h, hbar, dx, localV, scal;
  Buf<double> partial;
  LCK(u.alloc(m, st)); LCK(v.alloc(n, st)); LCK(h.alloc(n, st)); LCK(hbar.alloc(n, st)); LCK(dx.alloc(n, st));
  LCK(localV.alloc((size_t)localVecs * n, st)); LCK(scal.alloc(4, st)); LCK(partial.alloc(1024, st));
  Ctx c{st, partial.p, scal.p, 592};
  const unsigned gm = (unsigned)((m + 255) / 256), gn = (unsigned)((n + 255) / 256);
  const unsigned gwm = (unsigned)(((long long)m * 32 + 255) / 256), gwn = (unsigned)(((long long)n * 32 + 255) / 256);
  LCK(cudaMemcpyAsync(u.p, b, sizeof(float) * m, cudaMemcpyDefault, st));     // b may live on the host or in HBM
  LCK(cudaMemsetAsync(v.p, 0, sizeof(float) * n, st));
  LCK(cudaMemsetAsync(dx.p, 0, sizeof(float) * n, st));
  LCK(cudaMemsetAsync(hbar.p, 0, sizeof(float) * n, st));
  LCK(cudaEventRecord(e1, st));
  info->itn = 0; info->istop = 0; info->normA = 0; info->condA = 0; info->normx = 0;
  float alpha = 0.0f, beta = 0.0f;
  if ((rc = norm2(c, u.p, m, &beta))) return rc;
  if (beta > 0.0f) {
    k_scal<<<gm, 256, 0, st>>>(u.p, m, 1.0f / beta);
    k_spmv_add<<<gwn, 256, 0, st>>>(n, csc_ptr.p, csc_idx.p, csc_val.p, u.p, v.p);      // v = v + A^T u
    if ((rc = norm2(c, v.p, n, &alpha))) return rc;
  }
  if (alpha > 0.0f) k_scal<<<gn, 256, 0, st>>>(v.p, n, 1.0f / alpha);
  float normAr = alpha * beta;
  info->normAr = normAr; info->normr = beta;
  int itn = 0, istop = 0;
  float normA = 0, condA = 0, normx = 0, normr = beta;
  if (normAr != 0.0f) {
    bool queueFull = false;
    int localPointer = 0;
    const bool localOrtho = localVecs > 0;
    if (localOrtho) {
      localPointer = 1;
      LCK(cudaMemcpyAsync(localV.p, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    }
    float zetabar = alpha * beta, alphabar = alpha, rho = 1, rhobar = 1, cbar = 1, sbar = 0;
    LCK(cudaMemcpyAsync(h.p, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    float betadd = beta, betad = 0, rhodold = 1, tautildeold = 0, thetatilde = 0, zeta = 0, d = 0;
    float normA2 = alpha * alpha, maxrbar = 0.0f, minrbar = 1e+30f;
    const float normb = beta;
    const float ctol = conlim > 0.0f ? 1.0f / conlim : 0.0f;
    for (;;) {
      itn = itn + 1;
      k_scal<<<gm, 256, 0, st>>>(u.p, m, -alpha);
      k_spmv_add<<<gwm, 256, 0, st>>>(m, csr_ptr.p, csr_idx.p, csr_val.p, v.p, u.p);    // u = u + A v
      if ((rc = norm2(c, u.p, m, &beta))) return rc;
      if (beta > 0.0f) {
        k_scal<<<gm, 256, 0, st>>>(u.p, m, 1.0f / beta);
        if (localOrtho) {
          if (localPointer < localVecs) localPointer = localPointer + 1;
          else { localPointer = 1; queueFull = true; }
          LCK(cudaMemcpyAsync(localV.p + (size_t)(localPointer - 1) * n, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
        }
        k_scal<<<gn, 256, 0, st>>>(v.p, n, -beta);
        k_spmv_add<<<gwn, 256, 0, st>>>(n, csc_ptr.p, csc_idx.p, csc_val.p, u.p, v.p);  // v = v + A^T u
        if (localOrtho) {
          const int lim = queueFull ? localVecs : localPointer;
          k_reorth<<<1, 1024, 0, st>>>(v.p, localV.p, n, lim);
        }
        if ((rc = norm2(c, v.p, n, &alpha))) return rc;
        if (alpha > 0.0f) k_scal<<<gn, 256, 0, st>>>(v.p, n, 1.0f / alpha);
      }
      const float alphahat = d2norm(alphabar, damp);
      const float chat = alphabar / alphahat, shat = damp / alphahat;
      const float rhoold = rho;
      rho = d2norm(alphahat, beta);
      const float cc = alphahat / rho, s = beta / rho;
      const float thetanew = s * alpha;
      alphabar = cc * alpha;
      const float rhobarold = rhobar, zetaold = zeta;
      const float thetabar = sbar * rho, rhotemp = cbar * rho;
      rhobar = d2norm(cbar * rho, thetanew);
      cbar = cbar * rho / rhobar;
      sbar = thetanew / rhobar;
      zeta = cbar * zetabar;
      zetabar = -sbar * zetabar;
      k_update<<<gn, 256, 0, st>>>(hbar.p, h.p, dx.p, v.p, n, thetabar * rho / (rhoold * rhobarold), zeta / (rho * rhobar),
                                   thetanew / rho);
      const float betaacute = chat * betadd, betacheck = -shat * betadd;
      const float betahat = cc * betaacute;
      betadd = -s * betaacute;
      const float thetatildeold = thetatilde;
      const float rhotildeold = d2norm(rhodold, thetabar);
      const float ctildeold = rhodold / rhotildeold, stildeold = thetabar / rhotildeold;
      thetatilde = stildeold * rhobar;
      rhodold = ctildeold * rhobar;
      betad = -stildeold * betad + ctildeold * betahat;
      tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold;
      const float taud = (zeta - thetatilde * tautildeold) / rhodold;
      d = d + betacheck * betacheck;
      normr = std::sqrt(d + (betad - taud) * (betad - taud) + betadd * betadd);
      normA2 = normA2 + beta * beta;
      normA = std::sqrt(normA2);
      normA2 = normA2 + alpha * alpha;
      maxrbar = std::max(maxrbar, rhobarold);
      if (itn > 1) minrbar = std::min(minrbar, rhobarold);
      condA = std::max(maxrbar, rhotemp) / std::min(minrbar, rhotemp);
      normAr = std::fabs(zetabar);
      if ((rc = norm2(c, dx.p, n, &normx))) return rc;
      const float test1 = normr / normb, test2 = normAr / (normA * normr), test3 = 1.0f / condA;
      const float t1 = test1 / (1.0f + normA * normx / normb);
      const float rtol = btol + atol * normA * normx / normb;
      if (itn >= itnlim) istop = 7;
      if (1.0f + test3 <= 1.0f) istop = 6;
      if (1.0f + test2 <= 1.0f) istop = 5;
      if (1.0f + t1 <= 1.0f) istop = 4;
      if (test3 <= ctol) istop = 3;
      if (test2 <= atol) istop = 2;
      if (test1 <= rtol) istop = 1;
      if (istop != 0) break;
    }
    if (damp > 0.0f && istop == 2) istop = 3;
  }
  LCK(cudaEventRecord(e2, st));
  LCK(cudaMemcpyAsync(x, dx.p, sizeof(float) * n, cudaMemcpyDefault, st));    // so may x
  LCK(cudaStreamSynchronize(st));
  LCK(cudaGetLastError());
  info->istop = istop; info->itn = itn; info->normA = normA; info->condA = condA; info->normr = normr;
  info->normAr = normAr; info->normx = normx;
  cudaEventElapsedTime(&info->setup_ms, e0, e1);
  cudaEventElapsedTime(&info->solve_ms, e1, e2);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  return DAZIM_OK;
}

}  // namespace dzl
