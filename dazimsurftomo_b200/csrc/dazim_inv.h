// Shared declarations of the outer-iteration stage (dazim_invert.cu <-> dazim_api.cu).
#pragma once
#include <cuda_runtime.h>

namespace dzi {

// TikhRegul.f90:22 -- cells on any face of the (nvx, nvz, nz-1) block get the one-entry row
__host__ __device__ inline bool tikh_boundary(int i, int j, int k, int nvx, int nvz, int nzm1) {
  return i == 1 || i == nvx || j == 1 || j == nvz || k == 1 || k == nzm1;
}
__host__ __device__ inline int tikh_clamp(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
// Entries the rows of all cells before (i,j,k) (1-based, i fastest) occupy inside one parameter block:
// one per cell plus six more per interior cell.  Closed form, so every row can be written independently.
__host__ __device__ inline long long tikh_offset(int i, int j, int k, int nvx, int nvz, int nzm1) {
  const int ni = nvx > 2 ? nvx - 2 : 0, nj = nvz > 2 ? nvz - 2 : 0, nk = nzm1 > 2 ? nzm1 - 2 : 0;
  long long interior = (long long)tikh_clamp(k - 2, nk) * ni * nj;       // whole interior layers below k
  if (k >= 2 && k <= nzm1 - 1) {
    interior += (long long)tikh_clamp(j - 2, nj) * ni;                   // whole interior lines of this layer
    if (j >= 2 && j <= nvz - 1) interior += tikh_clamp(i - 2, ni);       // interior cells before i on this line
  }
  const long long c = ((long long)(k - 1) * nvz + (j - 1)) * nvx + (i - 1);
  return c + 6 * interior;
}
// entries of one whole block
__host__ __device__ inline long long tikh_block_entries(int nvx, int nvz, int nzm1) {
  const int ni = nvx > 2 ? nvx - 2 : 0, nj = nvz > 2 ? nvz - 2 : 0, nk = nzm1 > 2 ? nzm1 - 2 : 0;
  return (long long)nvx * nvz * nzm1 + 6ll * ni * nj * nk;
}

cudaError_t launch_resid(int n, const float* obst, const float* dsyn, float* cbst, float* tdata, float* deltaT,
                         cudaStream_t st);
cudaError_t launch_delta(int n, const float* cbst, const float* obst, float* deltaT, cudaStream_t st);
cudaError_t launch_seq_stats(int narr, const float* const* arrs, int n, float* out, cudaStream_t st);
cudaError_t launch_sigma(int n, const float* deltaT, const float* obst, const float* st_dt, float* sigmaT,
                         float* datweight, float* cbst, cudaStream_t st);
cudaError_t launch_scale_rows(long long nrow, const long long* rowptr, const float* w, float* val, cudaStream_t st);
cudaError_t launch_dws_finish(int ncol, const double* acc, float* norm, cudaStream_t st);
cudaError_t launch_dws(long long nnz, const int* col, const float* val, int ncol, double* acc, float* norm,
                       cudaStream_t st);
cudaError_t launch_tikh(int nvx, int nvz, int nzm1, long long base, int row_base, int col_off, float weight, float* val,
                        int* col, int* rowid, cudaStream_t st);
cudaError_t launch_model_update(int nx, int ny, int nz, int iso_inv, float* dv, float* vsf, float minvel, float maxvel,
                                float* gcf, float* gsf, cudaStream_t st);
cudaError_t launch_lm_terms(long long nre, long long nre_vs, const float* val, const int* col, const float* dv,
                            float lameVs, float lameGcs, float* lm, float* lmw, cudaStream_t st);
cudaError_t launch_resid_rows(long long nrow, const long long* rowptr, const int* col, const float* val, const float* dv,
                              const float* w, int maxvp, int nblk, const float* tdata, float* fwdTvs, float* fwdTaa,
                              float* resbst, float* resW, cudaStream_t st);
cudaError_t launch_norm2(const float* a, long long n, double* partial, float* out, cudaStream_t st);

}  // namespace dzi
