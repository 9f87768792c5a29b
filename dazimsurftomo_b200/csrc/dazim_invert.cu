// The part of one outer iteration of DAzimSurfTomo (Main_Jt.f90:364-750) that sits between two G builds, for
// sm_100a, operating on the device-resident CSR block a plan produced:
//   residual + CalDdatSigma      Main_Jt.f90:425-429, CalSigamNorm.f90:2-42
//   data weighting of b and G    Main_Jt.f90:460-469        (8 B per non-zero, HBM-bound, warp per row)
//   DWS column sums (iso mode)   Main_Jt.f90:476-481
//   Tikhonov rows                TikhRegul.f90:2-105 / :108-209   (appended behind G, closed-form offsets)
//   model update + clamps        Main_Jt.f90:582-620
//   model / residual norms       CalSigamNorm.f90:45-92, 154-223, 226-352  (from the sparse rows, no dense G)
// The float32 sums that feed the computation (mean and standard deviation of |dT/T|) and the printed residual
// statistics are accumulated sequentially in the reference's order (one lane, loads batched), so sigma, the data
// weights and the weighted system are bit-identical to the Fortran arithmetic; they cost O(rows) against the
// O(non-zeros) row scaling.  2-norms (diagnostics) are deterministic two-stage reductions in double.
#include "../../include/dazim_b200.h"
#include "dazim_inv.h"
#include <cuda_runtime.h>
#include <algorithm>

namespace dzi {

// cbst = obst - dsyn ; Tdata = cbst ; deltaT = |cbst/obst|      (Main_Jt.f90:425-429, CalSigamNorm.f90:21-24)
__global__ void k_resid(int n, const float* __restrict__ obst, const float* __restrict__ dsyn, float* __restrict__ cbst,
                        float* __restrict__ tdata, float* __restrict__ deltaT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float c = obst[i] - dsyn[i];
    cbst[i] = c;
    tdata[i] = c;
    deltaT[i] = fabsf(c / obst[i]);
  }
}

// deltaT = |cbst/obst| alone (CalSigamNorm.f90:21-24), for the stand-alone CalDdatSigma entry point
__global__ void k_delta(int n, const float* __restrict__ cbst, const float* __restrict__ obst, float* __restrict__ deltaT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) deltaT[i] = fabsf(cbst[i] / obst[i]);
}

// Sequential float32 statistics of up to gridDim.x arrays (one CTA each): out[3*b+0] = sum x(i), +1 = sum |x(i)|,
// +2 = sum (x(i) - sum/n)^2, each accumulated in index order like the Fortran loops / SUM intrinsics
// (gfortran -O3 without -ffast-math does not reassociate).  The order is what makes the chain serial; the memory
// side is not: a warp streams the array through shared memory in 1024-element chunks (coalesced loads, the next
// chunk already in flight in registers) and lane 0 adds the staged values in order, so a pass costs the float add
// latency per element (~4 cycles) instead of a global-load round trip per few elements.  Warp 0 accumulates
// sum x, warp 1 sum |x| side by side; warp 0 then makes the second pass for the squared deviations.
struct SeqArrays { const float* x[6]; };
static constexpr int SEQ_CHUNK = 1024;

// kind 0: x ; 1: |x| ; 2: (x - mean)^2
template <int KIND>
__device__ __forceinline__ float seq_pass(const float* __restrict__ x, int n, float mean, float* __restrict__ stage) {
  const int lane = threadIdx.x & 31;
  float acc = 0.0f;
  float r[SEQ_CHUNK / 32];
#pragma unroll
  for (int q = 0; q < SEQ_CHUNK / 32; ++q) { const int i = q * 32 + lane; r[q] = i < n ? x[i] : 0.0f; }
  for (int base = 0; base < n; base += SEQ_CHUNK) {
#pragma unroll
    for (int q = 0; q < SEQ_CHUNK / 32; ++q) stage[q * 32 + lane] = r[q];
    const int nb = base + SEQ_CHUNK;
    if (nb < n) {
#pragma unroll
      for (int q = 0; q < SEQ_CHUNK / 32; ++q) { const int i = nb + q * 32 + lane; r[q] = i < n ? x[i] : 0.0f; }
    }
    __syncwarp();
    if (lane == 0) {
      const int cnt = min(SEQ_CHUNK, n - base);
      int i = 0;
      for (; i + 32 <= cnt; i += 32) {           // 8 x LDS.128 in flight, then 32 dependent adds
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(stage + i + 4 * q);
          v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          if (KIND == 1) v[q] = fabsf(v[q]);
          else if (KIND == 2) { const float d = v[q] - mean; v[q] = d * d; }
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) acc = acc + v[q];
      }
      for (; i < cnt; ++i) {
        const float t = stage[i];
        if (KIND == 0) acc = acc + t;
        else if (KIND == 1) acc = acc + fabsf(t);
        else { const float d = t - mean; acc = acc + d * d; }
      }
    }
    __syncwarp();
  }
  return acc;      // valid in lane 0
}

__global__ void __launch_bounds__(64) k_seq_stats(SeqArrays a, int n, float* __restrict__ out) {
  const float* __restrict__ x = a.x[blockIdx.x];
  __shared__ __align__(16) float stage[2][SEQ_CHUNK];
  __shared__ float sh_sum;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const float s = seq_pass<0>(x, n, 0.0f, stage[0]);
    if (lane == 0) { out[3 * blockIdx.x + 0] = s; sh_sum = s; }
  } else {
    const float s = seq_pass<1>(x, n, 0.0f, stage[1]);
    if (lane == 0) out[3 * blockIdx.x + 1] = s;
  }
  __syncthreads();
  if (warp == 0) {
    const float mean = sh_sum / (float)n;
    const float q2 = seq_pass<2>(x, n, mean, stage[0]);
    if (lane == 0) out[3 * blockIdx.x + 2] = q2;
  }
}

// CalSigamNorm.f90:32-41 + Main_Jt.f90:461-464.  st_dt = statistics of deltaT from k_seq_stats (sum, -, ssd).
// exp: glibc's expf is correctly rounded in all but vanishingly rare cases; the double-precision exp rounded to float
// reproduces that.
__global__ void k_sigma(int n, const float* __restrict__ deltaT, const float* __restrict__ obst,
                        const float* __restrict__ st_dt, float* __restrict__ sigmaT, float* __restrict__ datweight,
                        float* __restrict__ cbst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float stddeltaT = sqrtf(st_dt[2] / (float)n);
  const float twostdratio = fabsf(deltaT[i] / (1.5f * stddeltaT));
  float s = stddeltaT * obst[i];
  if (twostdratio > 1.0f) s = s * (float)exp((double)(twostdratio - 1.0f));
  sigmaT[i] = s;
  const float w = 1.0f / s;
  datweight[i] = w;
  cbst[i] = cbst[i] * w;
}

// rw(k) = rw(k) * datweight(row(k))   (Main_Jt.f90:467-469), one warp per CSR row: 8 B of HBM traffic per non-zero
__global__ void __launch_bounds__(256) k_scale_rows(long long nrow, const long long* __restrict__ rowptr,
                                                     const float* __restrict__ w, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrow) return;
  const long long b = rowptr[r], e = rowptr[r + 1];
  const float wr = w[r];
  for (long long k = b + lane; k < e; k += 32) val[k] = val[k] * wr;
}

// norm(col(k)) += |rw(k)|  (Main_Jt.f90:476-481); accumulated in double so that the float result does not depend on
// the order of the atomics
__global__ void k_dws(long long nnz, const int* __restrict__ col, const float* __restrict__ val, double* __restrict__ acc) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) atomicAdd(&acc[col[k] - 1], (double)fabsf(val[k]));
}
__global__ void k_d2f(int n, const double* __restrict__ a, float* __restrict__ o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)a[i];
}

// Tikhonov rows, one thread per model cell of one parameter block (TikhRegul.f90:22-56): boundary cells give one
// entry (2w), interior cells the 7-point stencil (6w, then -w at i-1, i+1, j-1, j+1, k-1, k+1).
// base = first free entry; row_base = first row id (1-based); col_off = column offset of the block.
__global__ void k_tikh(int nvx, int nvz, int nzm1, long long base, int row_base, int col_off, float weight,
                       float* __restrict__ val, int* __restrict__ col, int* __restrict__ rowid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;        // 0-based cell index, i fastest
  const int ncell = nvx * nvz * nzm1;
  if (c >= ncell) return;
  const int i = c % nvx + 1, j = (c / nvx) % nvz + 1, k = c / (nvx * nvz) + 1;
  const long long o = base + tikh_offset(i, j, k, nvx, nvz, nzm1);
  const int row = row_base + c;
  const int c0 = c + 1 + col_off;
  if (tikh_boundary(i, j, k, nvx, nvz, nzm1)) {
    val[o] = 2.0f * weight; col[o] = c0; rowid[o] = row;
  } else {
    const int cols[7] = {c0, c0 - 1, c0 + 1, c0 - nvx, c0 + nvx, c0 - nvz * nvx, c0 + nvz * nvx};
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      val[o + q] = (q == 0 ? 6.0f : -1.0f) * weight;
      col[o + q] = cols[q];
      rowid[o + q] = row;
    }
  }
}

// Main_Jt.f90:582-620, one thread per cell: clip dv, add to vsf(i+1,j+1,k), clamp; joint: gcf, gsf = dv blocks
__global__ void k_model_update(int nx, int ny, int nz, int iso_inv, float* __restrict__ dv, float* __restrict__ vsf,
                               float minvel, float maxvel, float* __restrict__ gcf, float* __restrict__ gsf) {
  const int nvx = nx - 2, nvz = ny - 2;
  const int maxvp = nvx * nvz * (nz - 1);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= maxvp) return;
  const int i = c % nvx + 1, j = (c / nvx) % nvz + 1, k = c / (nvx * nvz) + 1;
  float pertV = dv[c];
  if (pertV >= 0.5f) pertV = 0.5f;
  if (pertV <= -0.5f) pertV = -0.5f;
  if (fabsf(pertV) < 1e-5f) pertV = 0.0f;
  dv[c] = pertV;
  const size_t o = (size_t)i + (size_t)j * nx + (size_t)(k - 1) * nx * ny;
  float v = vsf[o] + pertV;
  if (v < minvel) v = minvel;
  if (v > maxvel) v = maxvel;
  vsf[o] = v;
  if (!iso_inv) {
    gcf[c] = dv[maxvp + c];
    gsf[c] = dv[2 * maxvp + c];
  }
}

// Lm(i) = rw*dv(col)/lame ; LmWeight(i) = rw*dv(col)   (CalSigamNorm.f90:262-266, :329-343) over the appended entries
__global__ void k_lm_terms(long long nre, long long nre_vs, const float* __restrict__ val, const int* __restrict__ col,
                           const float* __restrict__ dv, float lameVs, float lameGcs, float* __restrict__ lm,
                           float* __restrict__ lmw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nre) return;
  const float p = val[i] * dv[col[i] - 1];
  lm[i] = p / (i < nre_vs ? lameVs : lameGcs);
  lmw[i] = p;
}

// fwdTvs / fwdTaa / resbst of CalVsReslNorm / CalReslNormJoint from the (weighted) CSR rows: the three per-block
// sums of a row are reduced by a warp and un-weighted by 1/w(row)
__global__ void __launch_bounds__(256) k_resid_rows(long long nrow, const long long* __restrict__ rowptr,
                                                     const int* __restrict__ col, const float* __restrict__ val,
                                                     const float* __restrict__ dv, const float* __restrict__ w,
                                                     int maxvp, int nblk, const float* __restrict__ tdata,
                                                     float* __restrict__ fwdTvs, float* __restrict__ fwdTaa,
                                                     float* __restrict__ resbst, float* __restrict__ resW) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrow) return;
  const long long b = rowptr[r], e = rowptr[r + 1];
  float svs = 0.0f, sgc = 0.0f, sgs = 0.0f;
  for (long long k = b + lane; k < e; k += 32) {
    const int c = col[k] - 1;
    const float p = val[k] * dv[c];
    if (c < maxvp) svs += p;
    else if (c < 2 * maxvp) sgc += p;
    else sgs += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    svs += __shfl_xor_sync(0xffffffffu, svs, o);
    sgc += __shfl_xor_sync(0xffffffffu, sgc, o);
    sgs += __shfl_xor_sync(0xffffffffu, sgs, o);
  }
  if (lane == 0) {
    const float wr = w[r];
    const float tvs = svs / wr, taa = sgs / wr + sgc / wr;
    float res;
    if (nblk == 3) res = tdata[r] - taa - tvs;
    else res = tdata[r] - tvs;
    fwdTvs[r] = tvs;
    fwdTaa[r] = nblk == 3 ? taa : 0.0f;
    resbst[r] = res;
    resW[r] = res * wr;
  }
}

// deterministic 2-norm: partial sums of squares in double, fixed-order final sum, sqrt, float
__global__ void __launch_bounds__(256) k_sumsq_partial(const float* __restrict__ a, long long n, double* __restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
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
__global__ void k_sumsq_final(const double* __restrict__ partial, int nb, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nb; ++i) s += partial[i];
    out[0] = (float)sqrt(s);
  }
}

// ---- launchers (plain pointers; called from dazim_api.cu) ----------------------------------------------------------
static inline unsigned blocks(long long n, int t) { return (unsigned)std::max<long long>(1, (n + t - 1) / t); }

cudaError_t launch_resid(int n, const float* obst, const float* dsyn, float* cbst, float* tdata, float* deltaT,
                         cudaStream_t st) {
  if (n > 0) k_resid<<<blocks(n, 256), 256, 0, st>>>(n, obst, dsyn, cbst, tdata, deltaT);
  return cudaGetLastError();
}
cudaError_t launch_delta(int n, const float* cbst, const float* obst, float* deltaT, cudaStream_t st) {
  if (n > 0) k_delta<<<blocks(n, 256), 256, 0, st>>>(n, cbst, obst, deltaT);
  return cudaGetLastError();
}
cudaError_t launch_seq_stats(int narr, const float* const* arrs, int n, float* out, cudaStream_t st) {
  SeqArrays a;
  for (int i = 0; i < 6; ++i) a.x[i] = i < narr ? arrs[i] : nullptr;
  if (narr > 0) k_seq_stats<<<narr, 64, 0, st>>>(a, n, out);
  return cudaGetLastError();
}
cudaError_t launch_sigma(int n, const float* deltaT, const float* obst, const float* st_dt, float* sigmaT,
                         float* datweight, float* cbst, cudaStream_t st) {
  if (n > 0) k_sigma<<<blocks(n, 256), 256, 0, st>>>(n, deltaT, obst, st_dt, sigmaT, datweight, cbst);
  return cudaGetLastError();
}
cudaError_t launch_scale_rows(long long nrow, const long long* rowptr, const float* w, float* val, cudaStream_t st) {
  if (nrow > 0) k_scale_rows<<<blocks(nrow * 32, 256), 256, 0, st>>>(nrow, rowptr, w, val);
  return cudaGetLastError();
}
cudaError_t launch_dws(long long nnz, const int* col, const float* val, int ncol, double* acc, float* norm,
                       cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * (size_t)ncol, st);
  if (e != cudaSuccess) return e;
  if (nnz > 0) k_dws<<<blocks(nnz, 256), 256, 0, st>>>(nnz, col, val, acc);
  k_d2f<<<blocks(ncol, 256), 256, 0, st>>>(ncol, acc, norm);
  return cudaGetLastError();
}
// row-distributed tail: the per-column sums of the ranks' row blocks are added (double) before the conversion
cudaError_t launch_dws_finish(int ncol, const double* acc, float* norm, cudaStream_t st) {
  k_d2f<<<blocks(ncol, 256), 256, 0, st>>>(ncol, acc, norm);
  return cudaGetLastError();
}
cudaError_t launch_tikh(int nvx, int nvz, int nzm1, long long base, int row_base, int col_off, float weight, float* val,
                        int* col, int* rowid, cudaStream_t st) {
  const int ncell = nvx * nvz * nzm1;
  if (ncell > 0) k_tikh<<<blocks(ncell, 128), 128, 0, st>>>(nvx, nvz, nzm1, base, row_base, col_off, weight, val, col, rowid);
  return cudaGetLastError();
}
cudaError_t launch_model_update(int nx, int ny, int nz, int iso_inv, float* dv, float* vsf, float minvel, float maxvel,
                                float* gcf, float* gsf, cudaStream_t st) {
  const int maxvp = (nx - 2) * (ny - 2) * (nz - 1);
  if (maxvp > 0) k_model_update<<<blocks(maxvp, 128), 128, 0, st>>>(nx, ny, nz, iso_inv, dv, vsf, minvel, maxvel, gcf, gsf);
  return cudaGetLastError();
}
cudaError_t launch_lm_terms(long long nre, long long nre_vs, const float* val, const int* col, const float* dv,
                            float lameVs, float lameGcs, float* lm, float* lmw, cudaStream_t st) {
  if (nre > 0) k_lm_terms<<<blocks(nre, 256), 256, 0, st>>>(nre, nre_vs, val, col, dv, lameVs, lameGcs, lm, lmw);
  return cudaGetLastError();
}
cudaError_t launch_resid_rows(long long nrow, const long long* rowptr, const int* col, const float* val, const float* dv,
                              const float* w, int maxvp, int nblk, const float* tdata, float* fwdTvs, float* fwdTaa,
                              float* resbst, float* resW, cudaStream_t st) {
  if (nrow > 0)
    k_resid_rows<<<blocks(nrow * 32, 256), 256, 0, st>>>(nrow, rowptr, col, val, dv, w, maxvp, nblk, tdata, fwdTvs,
                                                          fwdTaa, resbst, resW);
  return cudaGetLastError();
}
// out[0] = ||a(0:n)||_2 ; partial needs >= 592 doubles
cudaError_t launch_norm2(const float* a, long long n, double* partial, float* out, cudaStream_t st) {
  const int nb = (int)std::min<long long>(592, std::max<long long>(1, (n + 255) / 256));
  k_sumsq_partial<<<nb, 256, 0, st>>>(a, n, partial);
  k_sumsq_final<<<1, 32, 0, st>>>(partial, nb, out);
  return cudaGetLastError();
}

}  // namespace dzi
