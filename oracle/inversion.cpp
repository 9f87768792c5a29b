// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the pieces of the reference's outer inversion iteration that sit between the
// G build and the next G build, single precision and in the Fortran's operation order:
//   CalDdatSigma                 src/src_inv_iso_joint/CalSigamNorm.f90:2-42     (data sigma from |dT/T|)
//   data weighting of b and G    src/src_inv_iso_joint/Main_Jt.f90:460-469
//   TikhonovRegularization       src/src_inv_iso_joint/TikhRegul.f90:2-105       (iso / Gc,Gs only)
//   TikhRegul_joint              src/src_inv_iso_joint/TikhRegul.f90:108-209     (dVs block then Gc, Gs blocks)
//   model update + clamps        src/src_inv_iso_joint/Main_Jt.f90:582-620
//   Calmodel2Norm / ...Joint     src/src_inv_iso_joint/CalSigamNorm.f90:226-352
//   CalVsReslNorm / CalReslNormJoint (CalSigamNorm.f90:45-92, 154-223) with the dense MATMULs replaced by the
//       same ascending-column float32 sums over the sparse triplets (the dense arrays hold exactly those entries
//       and zeros; the reference's stale-coefficient quirk in the dense fill, SURVEY Q6, is not reproduced)
//   residual statistics          src/src_inv_iso_joint/Main_Jt.f90:432-439, 720-727
// Parity pinning: the loop built from these (oracle/pyoracle.py::invert) is compared with the reference's shipped
// inversion results example/test2_syn_iso_inv/plot_script/DSurfTomo.inv and
// example/test3_syn_joint_inv/plot_script/Gc_Gs_model.inv by scripts/pin_inversion.py (results in DESIGN.md).
#include <cmath>
#include <cstddef>
#include <vector>

namespace {

// lsmrblas.f90:247-277 (same restatement as in lsmr.cpp; kept local so the two files stay independent)
float nrm2(long n, const float* x) {
  if (n < 1) return 0.0f;
  if (n == 1) return std::fabs(x[0]);
  float scale = 0.0f, ssq = 1.0f;
  for (long i = 0; i < n; ++i) {
    if (x[i] != 0.0f) {
      const float a = std::fabs(x[i]);
      if (scale < a) {
        const float r = scale / a;
        ssq = 1.0f + ssq * (r * r);
        scale = a;
      } else {
        const float r = a / scale;
        ssq = ssq + r * r;
      }
    }
  }
  return scale * std::sqrt(ssq);
}

// one Laplacian stencil row of TikhRegul.f90 (:22-56); off = column offset of the parameter block
inline void tikh_row(int i, int j, int k, int nvx, int nvz, int nzm1, int off, float weight, int rowid, long& nar,
                     float* rw, int* iw_row, int* col) {
  const int c0 = (k - 1) * nvz * nvx + (j - 1) * nvx + i + off;
  if (i == 1 || i == nvx || j == 1 || j == nvz || k == 1 || k == nzm1) {
    col[nar] = c0; rw[nar] = 2.0f * weight; iw_row[nar] = rowid;
    nar += 1;
  } else {
    const int cols[7] = {c0, c0 - 1, c0 + 1, c0 - nvx, c0 + nvx, c0 - nvz * nvx, c0 + nvz * nvx};
    for (int q = 0; q < 7; ++q) {
      col[nar + q] = cols[q];
      rw[nar + q] = (q == 0 ? 6.0f : -1.0f) * weight;
      iw_row[nar + q] = rowid;
    }
    nar += 7;
  }
}

}  // namespace

extern "C" {

// CalSigamNorm.f90:2-42
void orc_cal_ddat_sigma(int dall, const float* obst, const float* cbst, float* sigmaT, float* meandeltaT_out) {
  std::vector<float> deltaT((size_t)dall);
  float meandeltaT = 0.0f;
  for (int i = 0; i < dall; ++i) {
    deltaT[i] = std::fabs(cbst[i] / obst[i]);
    meandeltaT = meandeltaT + deltaT[i];
  }
  meandeltaT = meandeltaT / (float)dall;
  float stddeltaT = 0.0f;
  for (int i = 0; i < dall; ++i) {
    const float d = deltaT[i] - meandeltaT;
    stddeltaT = stddeltaT + d * d;
  }
  stddeltaT = std::sqrt(stddeltaT / (float)dall);
  for (int i = 0; i < dall; ++i) {
    const float twostdratio = std::fabs(deltaT[i] / (1.5f * stddeltaT));
    if (twostdratio > 1.0f) sigmaT[i] = stddeltaT * obst[i] * std::exp(twostdratio - 1.0f);
    else sigmaT[i] = stddeltaT * obst[i];
  }
  *meandeltaT_out = meandeltaT;
}

// Main_Jt.f90:461-469: datweight = 1/sigma ; cbst *= datweight ; rw(k) *= datweight(row(k))
void orc_apply_weights(int dall, const float* sigmaT, float* datweight, float* cbst, long nar, const int* iw_row,
                       float* rw) {
  for (int i = 0; i < dall; ++i) {
    datweight[i] = 1.0f / sigmaT[i];
    cbst[i] = cbst[i] * datweight[i];
  }
  for (long k = 0; k < nar; ++k) rw[k] = rw[k] * datweight[iw_row[k] - 1];
}

// TikhRegul.f90:2-105.  rw/iw_row/col are 0-based arrays holding nar entries on entry (iw_row = iw(2:)).
void orc_tikhonov(int nx, int ny, int nz, int maxvp, int dall, long* nar_io, float* rw, int* iw_row, int* col,
                  int* count3_out, int iso_inv, float weightGcs, float weightVs) {
  const int nvz = ny - 2, nvx = nx - 2;
  long nar = *nar_io;
  int count3 = 0;
  if (iso_inv) {
    for (int k = 1; k <= nz - 1; ++k)
      for (int j = 1; j <= nvz; ++j)
        for (int i = 1; i <= nvx; ++i) {
          count3 += 1;
          tikh_row(i, j, k, nvx, nvz, nz - 1, 0, weightVs, dall + count3, nar, rw, iw_row, col);
        }
  } else {
    for (int sc = 1; sc <= 2; ++sc)
      for (int k = 1; k <= nz - 1; ++k)
        for (int j = 1; j <= nvz; ++j)
          for (int i = 1; i <= nvx; ++i) {
            count3 += 1;
            tikh_row(i, j, k, nvx, nvz, nz - 1, (sc - 1) * maxvp, weightGcs, dall + count3, nar, rw, iw_row, col);
          }
  }
  *nar_io = nar;
  *count3_out = count3;
}

// TikhRegul.f90:108-209
void orc_tikh_joint(int nx, int ny, int nz, int maxvp, int dall, long* nar_io, float* rw, int* iw_row, int* col,
                    long* narVs_out, int* count3_out, float weightGcs, float weightVs) {
  const int nvz = ny - 2, nvx = nx - 2;
  long nar = *nar_io;
  int count3 = 0;
  for (int k = 1; k <= nz - 1; ++k)
    for (int j = 1; j <= nvz; ++j)
      for (int i = 1; i <= nvx; ++i) {
        count3 += 1;
        tikh_row(i, j, k, nvx, nvz, nz - 1, 0, weightVs, dall + count3, nar, rw, iw_row, col);
      }
  *narVs_out = nar;
  for (int sc = 1; sc <= 2; ++sc)
    for (int k = 1; k <= nz - 1; ++k)
      for (int j = 1; j <= nvz; ++j)
        for (int i = 1; i <= nvx; ++i) {
          count3 += 1;
          tikh_row(i, j, k, nvx, nvz, nz - 1, sc * maxvp, weightGcs, dall + count3, nar, rw, iw_row, col);
        }
  *nar_io = nar;
  *count3_out = count3;
}

// Main_Jt.f90:582-620.  dv (maxvp or 3*maxvp) is clipped in place, vsf (nx,ny,nz) column-major updated in
// place; joint mode also fills gcf, gsf (nx-2,ny-2,nz-1).
void orc_model_update(int nx, int ny, int nz, int iso_inv, float* dv, float* vsf, float minvel, float maxvel,
                      float* gcf, float* gsf) {
  const int nvx = nx - 2, nvz = ny - 2;
  const long maxvp = (long)nvx * nvz * (nz - 1);
  for (int k = 1; k <= nz - 1; ++k)
    for (int j = 1; j <= nvz; ++j)
      for (int i = 1; i <= nvx; ++i) {
        const long c = (long)(k - 1) * nvx * nvz + (long)(j - 1) * nvx + (i - 1);
        float pertV = dv[c];
        if (pertV >= 0.5f) pertV = 0.5f;
        if (pertV <= -0.5f) pertV = -0.5f;
        if (std::fabs(pertV) < 1e-5f) pertV = 0.0f;
        dv[c] = pertV;
        float& v = vsf[(size_t)i + (size_t)j * nx + (size_t)(k - 1) * nx * ny];   // vsf(i+1,j+1,k)
        v = v + pertV;
        if (v < minvel) v = minvel;
        if (v > maxvel) v = maxvel;
        if (!iso_inv) {
          gcf[c] = dv[maxvp + c];
          gsf[c] = dv[2 * maxvp + c];
        }
      }
}

// Calmodel2Norm (CalSigamNorm.f90:226-283) when narVs < 0, Calmodel2NormJoint (:285-352) otherwise.
// out[0..5] = VsNorm2, VswNorm2, GcsNorm2, GcswNorm2, Mnorm2, MwNorm2 (iso: only the last two are set).
void orc_model_norms(long nar1, long nar, long narVs, const float* rw, const int* col, const float* dv, float lameGcs,
                     float lameVs, float* out) {
  const long Nre = nar - nar1;
  std::vector<float> Lm((size_t)Nre), LmW((size_t)Nre);
  for (int q = 0; q < 6; ++q) out[q] = 0.0f;
  if (narVs < 0) {
    for (long i = 0; i < Nre; ++i) {
      const long k = nar1 + i;
      Lm[i] = rw[k] * dv[col[k] - 1] / lameVs;
      LmW[i] = rw[k] * dv[col[k] - 1];
    }
    out[4] = nrm2(Nre, Lm.data());
    out[5] = nrm2(Nre, LmW.data());
    return;
  }
  const long NreVs = narVs - nar1;
  for (long i = 0; i < NreVs; ++i) {
    const long k = nar1 + i;
    Lm[i] = rw[k] * dv[col[k] - 1] / lameVs;
    LmW[i] = rw[k] * dv[col[k] - 1];
  }
  out[0] = nrm2(NreVs, Lm.data());
  out[1] = nrm2(NreVs, LmW.data());
  for (long i = NreVs; i < Nre; ++i) {
    const long k = nar1 + i;
    Lm[i] = rw[k] * dv[col[k] - 1] / lameGcs;
    LmW[i] = rw[k] * dv[col[k] - 1];
  }
  out[2] = nrm2(Nre - NreVs, Lm.data() + NreVs);
  out[3] = nrm2(Nre - NreVs, LmW.data() + NreVs);
  out[4] = nrm2(Nre, Lm.data());
  out[5] = nrm2(Nre, LmW.data());
}

// CalVsReslNorm (CalSigamNorm.f90:45-92) for nblk = 1, CalReslNormJoint (:154-223) for nblk = 3, on the UNWEIGHTED
// triplets of the data rows (nar1 entries, rows ascending, columns ascending inside a row):
// fwdTvs = GVs*dv(1:maxvp), fwdTaa = GGs*dv(2maxvp+1:) + GGc*dv(maxvp+1:2maxvp), resbst = Tdata - fwdTaa - fwdTvs.
// norms[0] = ||resbst||, norms[1] = ||resbst*datweight||
void orc_residuals(int dall, long maxvp, int nblk, long nar1, const float* rw, const int* iw_row, const int* col,
                   const float* dv, const float* datweight, const float* Tdata, float* fwdTvs, float* fwdTaa,
                   float* resbst, float* norms) {
  std::vector<float> tgc((size_t)dall, 0.0f), tgs((size_t)dall, 0.0f), resW((size_t)dall);
  for (int i = 0; i < dall; ++i) { fwdTvs[i] = 0.0f; fwdTaa[i] = 0.0f; }
  for (long k = 0; k < nar1; ++k) {
    const int r = iw_row[k] - 1;
    const long c = col[k] - 1;
    const float p = rw[k] * dv[c];
    if (c < maxvp) fwdTvs[r] = fwdTvs[r] + p;
    else if (c < 2 * maxvp) tgc[r] = tgc[r] + p;
    else tgs[r] = tgs[r] + p;
  }
  for (int i = 0; i < dall; ++i) {
    if (nblk == 3) {
      fwdTaa[i] = tgs[i] + tgc[i];
      resbst[i] = Tdata[i] - fwdTaa[i] - fwdTvs[i];
    } else {
      resbst[i] = Tdata[i] - fwdTvs[i];
    }
    resW[i] = resbst[i] * datweight[i];
  }
  norms[0] = nrm2(dall, resbst);
  norms[1] = nrm2(dall, resW.data());
}

// Main_Jt.f90:432-437 / :720-725: out = abs mean, std, RMS (= dnrm2/sqrt(real(dall))), mean
void orc_res_stats(int dall, const float* r, float* out) {
  float s = 0.0f;
  for (int i = 0; i < dall; ++i) s = s + r[i];
  const float mean = s / (float)dall;
  float q = 0.0f;
  for (int i = 0; i < dall; ++i) { const float d = r[i] - mean; q = q + d * d; }
  float a = 0.0f;
  for (int i = 0; i < dall; ++i) a = a + std::fabs(r[i]);
  out[0] = a / (float)dall;
  out[1] = std::sqrt(q / (float)dall);
  out[2] = nrm2(dall, r) / std::sqrt((float)dall);
  out[3] = mean;
}

}  // extern "C"
