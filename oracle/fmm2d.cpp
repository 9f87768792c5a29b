// ORACLE (test infrastructure, NOT product code). See fmm2d.hpp for the map
// of reference file:line -> function.  Compile with -ffp-contract=off and
// without -ffast-math: the reference is gfortran -O3 on x86-64/SSE2 (no FMA).
#include "fmm2d.hpp"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

static const float PI_F = 3.1415926535898f;  // CalSurfG.f90:166 (REAL*4 parameter)

static inline float cube(float x) { return x * (x * x); }  // x**3 (powi expansion)

// REAL*4 sin/cos/acos.  The reference calls libm's sinf/cosf/acosf, whose last
// bit is platform dependent (glibc vs Apple libm; neither is correctly
// rounded).  Oracle AND CUDA kernels both define them as the float rounding of
// the double-precision function, which is the correctly rounded
// result except with probability ~1e-9.
static inline float sin_r(float x) { return (float)std::sin((double)x); }
static inline float cos_r(float x) { return (float)std::cos((double)x); }
static inline float acos_r(float x) { return (float)std::acos((double)x); }

// ---------------------------------------------------------------------------
// FwdTraveltimeCPS.f90:346-409 / CalSurfGAniso_Joint.f90 Part 1
void Fmm::init(int nx, int ny, float goxdf, float gozdf, float dvxdf, float dvzdf) {
  gdx = 5; gdz = 5; asgr = 1; sgdl = 8; sgs = 8; earth = 6371.0f; fom = 1; snb = 0.5f;
  goxd = goxdf; gozd = gozdf; dvxd = dvxdf; dvzd = dvzdf;
  nvx = nx - 2; nvz = ny - 2;
  ldv = nvz + 2;
  velv.assign((size_t)(nvz + 2) * (nvx + 2), 0.0f);
  dvx = dvxd * PI_F / 180.0f;
  dvz = dvzd * PI_F / 180.0f;
  gox = (90.0f - goxd) * PI_F / 180.0f;
  goz = gozd * PI_F / 180.0f;
  nnx = (nvx - 1) * gdx + 1;
  nnz = (nvz - 1) * gdz + 1;
  dnx = dvx / (float)gdx;
  dnz = dvz / (float)gdz;
  dnxd = dvxd / (float)gdx;
  dnzd = dvzd / (float)gdz;
  nnx_c = nnx; nnz_c = nnz; dnx_c = dnx; dnz_c = dnz; gox_c = gox; goz_c = goz;
  const int nref = 2 * sgs * sgdl + 1;
  ld = std::max(nnz, nref);
  int ncol = std::max(nnx, nref);
  veln.assign((size_t)ld * ncol, 0.0f);
  velnb.assign((size_t)ld * ncol, 0.0f);
  ttn.assign((size_t)ld * ncol, 0.0f);
  nsts.assign((size_t)ld * ncol, -1);
  ldr = nref;
  ttnr.assign((size_t)nref * nref, 0.0f);
  nstsr.assign((size_t)nref * nref, -1);
  btg_px.assign((size_t)ld * ncol + 2, 0);
  btg_pz.assign((size_t)ld * ncol + 2, 0);
  rbint = 0;
}

static inline void bspl_basis(float u, float* b) {
  // CalSurfG.f90:1472-1475 (same expressions at :1546-1553 and rpathsAzim.f90:515-522)
  b[0] = cube(1.0f - u) / 6.0f;
  b[1] = (4.0f - 6.0f * (u * u) + 3.0f * cube(u)) / 6.0f;
  b[2] = (1.0f + 3.0f * u + 3.0f * (u * u) - 3.0f * cube(u)) / 6.0f;
  b[3] = cube(u) / 6.0f;
}

// CalSurfG.f90:1423-1516
void Fmm::gridder(const double* pv) {
  for (int i = 0; i <= nvz + 1; ++i)
    for (int j = 0; j <= nvx + 1; ++j) VELV(i, j) = (float)pv[i * (nvx + 2) + j];
  std::vector<float> ui((size_t)(gdx + 1) * 4), vi((size_t)(gdz + 1) * 4);
  for (int i = 1; i <= gdx + 1; ++i) {
    float u = (float)gdx;
    u = (float)(i - 1) / u;
    bspl_basis(u, &ui[(size_t)(i - 1) * 4]);
  }
  for (int i = 1; i <= gdz + 1; ++i) {
    float u = (float)gdz;
    u = (float)(i - 1) / u;
    bspl_basis(u, &vi[(size_t)(i - 1) * 4]);
  }
  for (int i = 1; i <= nvz - 1; ++i) {
    int conz = gdz;
    if (i == nvz - 1) conz = gdz + 1;
    for (int j = 1; j <= nvx - 1; ++j) {
      int conx = gdx;
      if (j == nvx - 1) conx = gdx + 1;
      for (int l = 1; l <= conz; ++l) {
        int stz = gdz * (i - 1) + l;
        for (int m = 1; m <= conx; ++m) {
          int stx = gdx * (j - 1) + m;
          float sumi = 0.0f;
          for (int i1 = 1; i1 <= 4; ++i1) {
            float sumj = 0.0f;
            for (int j1 = 1; j1 <= 4; ++j1)
              sumj = sumj + ui[(size_t)(m - 1) * 4 + (j1 - 1)] * VELV(i - 2 + i1, j - 2 + j1);
            sumi = sumi + vi[(size_t)(l - 1) * 4 + (i1 - 1)] * sumj;
          }
          VELN(stz, stx) = sumi;
        }
      }
    }
  }
}

// CalSurfG.f90:1525-1591.  Loop bounds over B-spline cells are narrowed to the
// cells that can intersect the source box (the reference visits every cell and
// CYCLEs); the values written are identical.
void Fmm::bsplrefine() {
  const int nrxr = gdx * sgdl, nrzr = gdz * sgdl;
  std::vector<float> ub((size_t)(nrxr + 1) * 4), vb((size_t)(nrzr + 1) * 4);
  for (int j = 1; j <= nrxr + 1; ++j) {
    float u = (float)nrxr;
    u = (float)(j - 1) / u;
    bspl_basis(u, &ub[(size_t)(j - 1) * 4]);
  }
  for (int i = 1; i <= nrzr + 1; ++i) {
    float v = (float)nrzr;
    v = (float)(i - 1) / v;
    bspl_basis(v, &vb[(size_t)(i - 1) * 4]);
  }
  const int origx = (vnl - 1) * sgdl + 1;
  const int origz = (vnt - 1) * sgdl + 1;
  const int ilo = std::max(1, (vnt - 1) / gdz), ihi = std::min(nvz - 1, (vnb - 1) / gdz + 1);
  const int jlo = std::max(1, (vnl - 1) / gdx), jhi = std::min(nvx - 1, (vnr - 1) / gdx + 1);
  for (int i = ilo; i <= ihi; ++i) {
    int conz = nrzr;
    if (i == nvz - 1) conz = nrzr + 1;
    for (int j = jlo; j <= jhi; ++j) {
      int conx = nrxr;
      if (j == nvx - 1) conx = nrxr + 1;
      for (int k = 1; k <= conz; ++k) {
        int st1 = gdz * (i - 1) + (k - 1) / sgdl + 1;
        if (st1 < vnt || st1 > vnb) continue;
        st1 = nrzr * (i - 1) + k;
        for (int l = 1; l <= conx; ++l) {
          int st2 = gdx * (j - 1) + (l - 1) / sgdl + 1;
          if (st2 < vnl || st2 > vnr) continue;
          st2 = nrxr * (j - 1) + l;
          float sum[4];
          for (int i1 = 1; i1 <= 4; ++i1) {
            sum[i1 - 1] = 0.0f;
            for (int j1 = 1; j1 <= 4; ++j1)
              sum[i1 - 1] = sum[i1 - 1] + ub[(size_t)(l - 1) * 4 + (j1 - 1)] * VELV(i - 2 + i1, j - 2 + j1);
            sum[i1 - 1] = vb[(size_t)(k - 1) * 4 + (i1 - 1)] * sum[i1 - 1];
          }
          int idm1 = st1 - origz + 1;
          int idm2 = st2 - origx + 1;
          if (idm1 < 1 || idm1 > nnz) continue;
          if (idm2 < 1 || idm2 > nnx) continue;
          VELN(idm1, idm2) = sum[0] + sum[1] + sum[2] + sum[3];
        }
      }
    }
  }
}

// CalSurfG.f90:2293-2314
float Fmm::bilinear(const float nv[2][2], float dsx, float dsz) {
  float biv = 0.0f;
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) {
      float produ = (1.0f - std::fabs(((float)(i - 1) * dnx - dsx) / dnx)) *
                    (1.0f - std::fabs(((float)(j - 1) * dnz - dsz) / dnz));
      biv = biv + nv[i - 1][j - 1] * produ;
    }
  return biv;
}

// CalSurfG.f90:258-457
int Fmm::travel(float scx, float scz, int urg) {
  int isx = (int)((scx - gox) / dnx) + 1;
  int isz = (int)((scz - goz) / dnz) + 1;
  int sw = 0;
  if (isx < 1 || isx > nnx) sw = 1;
  if (isz < 1 || isz > nnz) sw = 1;
  if (sw == 1) return ERR_SOURCE_OUTSIDE;
  if (isx == nnx) isx = isx - 1;
  if (isz == nnz) isz = isz - 1;
  if (urg != 2)
    for (int i = 1; i <= nnx; ++i)
      for (int j = 1; j <= nnz; ++j) NSTS(j, i) = -1;
  ntr = 0;
  if (urg == 2) {
    for (int i = 1; i <= nnx; ++i)
      for (int j = 1; j <= nnz; ++j)
        if (NSTS(j, i) > 0) addtree(j, i);
  } else {
    float vss[2][2];
    for (int i = 1; i <= 2; ++i)
      for (int j = 1; j <= 2; ++j) vss[i - 1][j - 1] = VELN(isz - 1 + j, isx - 1 + i);
    float dsx = (scx - gox) - (float)(isx - 1) * dnx;
    float dsz = (scz - goz) - (float)(isz - 1) * dnz;
    float vsrc = bilinear(vss, dsx, dsz);
    for (int i = 1; i <= 2; ++i)
      for (int j = 1; j <= 2; ++j) {
        float ax = dsx - (float)(i - 1) * dnx;
        float az = dsz - (float)(j - 1) * dnz;
        float ds = std::sqrt(ax * ax + az * az);
        TTN(isz - 1 + j, isx - 1 + i) = 2.0f * ds / (vss[i - 1][j - 1] + vsrc);
        addtree(isz - 1 + j, isx - 1 + i);
      }
  }
  while (ntr > 0) {
    int ix, iz;
    if (urg == 1) {
      ix = btg_px[1];
      iz = btg_pz[1];
      int swrg = 0;
      if (ix == 1 && vnl != 1) swrg = 1;
      if (ix == nnx && vnr != nnx) swrg = 1;   // literal: coarse index vnr vs *refined* nnx (CalSurfG.f90:369-371)
      if (iz == 1 && vnt != 1) swrg = 1;
      if (iz == nnz && vnb != nnz) swrg = 1;   // literal, same remark (:375-377)
      if (swrg == 1) {
        NSTS(iz, ix) = 0;
        break;
      }
    }
    ix = btg_px[1];
    iz = btg_pz[1];
    NSTS(iz, ix) = 0;
    ++n_accept;
    if (rec_rank) (*rec_rank)[(size_t)(ix - 1) * ld + (iz - 1)] = rec_count++;   // experiment hook (order_experiment)
    downtree();
    for (int i = ix - 1; i <= ix + 1; i += 2) {
      if (i >= 1 && i <= nnx) {
        if (NSTS(iz, i) == -1) {
          fouds2(iz, i);
          addtree(iz, i);
        } else if (NSTS(iz, i) > 0) {
          fouds2(iz, i);
          updtree(iz, i);
        }
      }
    }
    for (int i = iz - 1; i <= iz + 1; i += 2) {
      if (i >= 1 && i <= nnz) {
        if (NSTS(i, ix) == -1) {
          fouds2(i, ix);
          addtree(i, ix);
        } else if (NSTS(i, ix) > 0) {
          fouds2(i, ix);
          updtree(i, ix);
        }
      }
    }
  }
  return OK;
}

// CalSurfG.f90:557-729
void Fmm::fouds2(int iz, int ix) {
  int tsw1 = 0;
  float travm = 0.0f, trav;
  float slown = 1.0f / VELN(iz, ix);
  float ri = earth;
  float risti = ri * sin_r(gox + (float)(ix - 1) * dnx);
  for (int j = ix - 1; j <= ix + 1; j += 2) {
    if (j >= 1 && j <= nnx) {
      int swj = -1, j2;
      if (j == ix - 1) {
        j2 = j - 1;
        if (j2 >= 1) { if (NSTS(iz, j2) == 0) swj = 0; }
      } else {
        j2 = j + 1;
        if (j2 <= nnx) { if (NSTS(iz, j2) == 0) swj = 0; }
      }
      if (NSTS(iz, j) == 0 && swj == 0) {
        swj = -1;
        if (TTN(iz, j) > TTN(iz, j2)) swj = 0;
      } else {
        swj = -1;
      }
      for (int k = iz - 1; k <= iz + 1; k += 2) {
        if (k >= 1 && k <= nnz) {
          int swk = -1, k2;
          if (k == iz - 1) {
            k2 = k - 1;
            if (k2 >= 1) { if (NSTS(k2, ix) == 0) swk = 0; }
          } else {
            k2 = k + 1;
            if (k2 <= nnz) { if (NSTS(k2, ix) == 0) swk = 0; }
          }
          if (NSTS(k, ix) == 0 && swk == 0) {
            swk = -1;
            if (TTN(k, ix) > TTN(k2, ix)) swk = 0;
          } else {
            swk = -1;
          }
          int swsol = 0;
          float a = 0, b = 0, c = 0, u, v, em, tref = 0, tdiv = 1.0f;
          if (swj == 0) {
            swsol = 1;
            if (swk == 0) {
              u = 2.0f * ri * dnx;
              v = 2.0f * risti * dnz;
              em = 4.0f * TTN(iz, j) - TTN(iz, j2) - 4.0f * TTN(k, ix);
              em = em + TTN(k2, ix);
              a = v * v + u * u;
              b = 2.0f * em * (u * u);
              c = (u * u) * (em * em - (slown * slown) * (v * v));
              tref = 4.0f * TTN(iz, j) - TTN(iz, j2);
              tdiv = 3.0f;
            } else if (NSTS(k, ix) == 0) {
              u = risti * dnz;
              v = 2.0f * ri * dnx;
              em = 3.0f * TTN(k, ix) - 4.0f * TTN(iz, j) + TTN(iz, j2);
              a = v * v + 9.0f * (u * u);
              b = 6.0f * em * (u * u);
              c = (u * u) * (em * em - (slown * slown) * (v * v));
              tref = TTN(k, ix);
              tdiv = 1.0f;
            } else {
              u = 2.0f * ri * dnx;
              a = 1.0f;
              b = 0.0f;
              c = -(u * u) * (slown * slown);
              tref = 4.0f * TTN(iz, j) - TTN(iz, j2);
              tdiv = 3.0f;
            }
          } else if (NSTS(iz, j) == 0) {
            swsol = 1;
            if (swk == 0) {
              u = ri * dnx;
              v = 2.0f * risti * dnz;
              em = 3.0f * TTN(iz, j) - 4.0f * TTN(k, ix) + TTN(k2, ix);
              a = v * v + 9.0f * (u * u);
              b = 6.0f * em * (u * u);
              c = (u * u) * (em * em - (v * v) * (slown * slown));
              tref = TTN(iz, j);
              tdiv = 1.0f;
            } else if (NSTS(k, ix) == 0) {
              u = ri * dnx;
              v = risti * dnz;
              em = TTN(k, ix) - TTN(iz, j);
              a = u * u + v * v;
              b = -2.0f * (u * u) * em;
              c = (u * u) * (em * em - (v * v) * (slown * slown));
              tref = TTN(iz, j);
              tdiv = 1.0f;
            } else {
              a = 1.0f;
              b = 0.0f;
              c = -(slown * slown) * (ri * ri) * (dnx * dnx);
              tref = TTN(iz, j);
              tdiv = 1.0f;
            }
          } else {
            if (swk == 0) {
              swsol = 1;
              u = 2.0f * risti * dnz;
              a = 1.0f;
              b = 0.0f;
              c = -(u * u) * (slown * slown);
              tref = 4.0f * TTN(k, ix) - TTN(k2, ix);
              tdiv = 3.0f;
            } else if (NSTS(k, ix) == 0) {
              swsol = 1;
              a = 1.0f;
              b = 0.0f;
              c = -(slown * slown) * (risti * risti) * (dnz * dnz);
              tref = TTN(k, ix);
              tdiv = 1.0f;
            }
          }
          if (swsol == 1) {
            float rd1 = b * b - 4.0f * a * c;
            if (rd1 < 0.0f) rd1 = 0.0f;
            float tdsh = (-b + std::sqrt(rd1)) / (2.0f * a);
            trav = (tref + tdsh) / tdiv;
            if (tsw1 == 1) {
              travm = std::min(trav, travm);
            } else {
              travm = trav;
              tsw1 = 1;
            }
          }
        }
      }
    }
  }
  TTN(iz, ix) = travm;
}

// CalSurfG.f90:738-775
void Fmm::addtree(int iz, int ix) {
  ntr = ntr + 1;
  NSTS(iz, ix) = ntr;
  btg_px[ntr] = ix;
  btg_pz[ntr] = iz;
  int tpc = ntr;
  int tpp = tpc / 2;
  while (tpp > 0) {
    if (TTN(iz, ix) < TTN(btg_pz[tpp], btg_px[tpp])) {
      NSTS(iz, ix) = tpp;
      NSTS(btg_pz[tpp], btg_px[tpp]) = tpc;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
      tpc = tpp;
      tpp = tpc / 2;
    } else {
      tpp = 0;
    }
  }
}

// CalSurfG.f90:786-855
void Fmm::downtree() {
  if (ntr == 1) {
    ntr = ntr - 1;
    return;
  }
  NSTS(btg_pz[ntr], btg_px[ntr]) = 1;
  btg_px[1] = btg_px[ntr];
  btg_pz[1] = btg_pz[ntr];
  ntr = ntr - 1;
  int tpp = 1;
  int tpc = 2 * tpp;
  while (tpc < ntr) {
    float rd1 = TTN(btg_pz[tpc], btg_px[tpc]);
    float rd2 = TTN(btg_pz[tpc + 1], btg_px[tpc + 1]);
    if (rd1 > rd2) tpc = tpc + 1;
    rd1 = TTN(btg_pz[tpc], btg_px[tpc]);
    rd2 = TTN(btg_pz[tpp], btg_px[tpp]);
    if (rd1 < rd2) {
      NSTS(btg_pz[tpp], btg_px[tpp]) = tpc;
      NSTS(btg_pz[tpc], btg_px[tpc]) = tpp;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
      tpp = tpc;
      tpc = 2 * tpp;
    } else {
      tpc = ntr + 1;
    }
  }
  if (tpc == ntr) {
    float rd1 = TTN(btg_pz[tpc], btg_px[tpc]);
    float rd2 = TTN(btg_pz[tpp], btg_px[tpp]);
    if (rd1 < rd2) {
      NSTS(btg_pz[tpp], btg_px[tpp]) = tpc;
      NSTS(btg_pz[tpc], btg_px[tpc]) = tpp;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
    }
  }
}

// CalSurfG.f90:864-891
void Fmm::updtree(int iz, int ix) {
  int tpc = NSTS(iz, ix);
  int tpp = tpc / 2;
  while (tpp > 0) {
    if (TTN(iz, ix) < TTN(btg_pz[tpp], btg_px[tpp])) {
      NSTS(iz, ix) = tpp;
      NSTS(btg_pz[tpp], btg_px[tpp]) = tpc;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
      tpc = tpp;
      tpp = tpc / 2;
    } else {
      tpp = 0;
    }
  }
}

// FwdTraveltimeCPS.f90:467-645 (identical block at CalSurfGAniso_Joint.f90:491-668
// and CalSurfG.f90:1140-1316)
int Fmm::solve_source(const double* pv, float x, float z) {
  // restore coarse-grid scalars (the reference restores them at :599-604)
  nnx = nnx_c; nnz = nnz_c; dnx = dnx_c; dnz = dnz_c; gox = gox_c; goz = goz_c;
  gridder(pv);
  for (int j = 1; j <= nnx; ++j)
    for (int k = 1; k <= nnz; ++k) VELNB(k, j) = VELN(k, j);
  const int nnxb = nnx, nnzb = nnz;
  const float dnxb = dnx, dnzb = dnz, goxb = gox, gozb = goz;
  int isx = (int)((x - gox) / dnx) + 1;
  int isz = (int)((z - goz) / dnz) + 1;
  int sw = 0;
  if (isx < 1 || isx > nnx) sw = 1;
  if (isz < 1 || isz > nnz) sw = 1;
  if (sw == 1) return ERR_SOURCE_OUTSIDE;
  if (isx == nnx) isx = isx - 1;
  if (isz == nnz) isz = isz - 1;
  vnl = isx - sgs; if (vnl < 1) vnl = 1;
  vnr = isx + sgs; if (vnr > nnx) vnr = nnx;
  vnt = isz - sgs; if (vnt < 1) vnt = 1;
  vnb = isz + sgs; if (vnb > nnz) vnb = nnz;
  nrnx = (vnr - vnl) * sgdl + 1;
  nrnz = (vnb - vnt) * sgdl + 1;
  drnx = dvx / (float)(gdx * sgdl);
  drnz = dvz / (float)(gdz * sgdl);
  gorx = gox + dnx * (float)(vnl - 1);
  gorz = goz + dnz * (float)(vnt - 1);
  nnx = nrnx; nnz = nrnz; dnx = drnx; dnz = drnz; gox = gorx; goz = gorz;
  bsplrefine();
  int st = travel(x, z, 1);
  if (st != OK) return st;
  // ttnr=ttn ; nstsr=nsts  (only the refined extent is ever read back)
  for (int l = 1; l <= nnx; ++l)
    for (int k = 1; k <= nnz; ++k) {
      TTNR(k, l) = TTN(k, l);
      NSTSR(k, l) = NSTS(k, l);
    }
  const int ogx = vnl, ogz = vnt, grdfx = sgdl, grdfz = sgdl;
  {
    const int mx = std::max(nnx, nnxb), mz = std::max(nnz, nnzb);
    for (int l = 1; l <= mx; ++l)
      for (int k = 1; k <= mz; ++k) NSTS(k, l) = -1;
  }
  for (int k = 1; k <= nnz; k += grdfz) {
    int idm1 = ogz + (k - 1) / grdfz;
    for (int l = 1; l <= nnx; l += grdfx) {
      int idm2 = ogx + (l - 1) / grdfx;
      NSTS(idm1, idm2) = NSTSR(k, l);
      if (NSTS(idm1, idm2) >= 0) TTN(idm1, idm2) = TTNR(k, l);
    }
  }
  nnxr = nnx; nnzr = nnz; goxr = gox; gozr = goz; dnxr = dnx; dnzr = dnz;
  nnx = nnxb; nnz = nnzb; dnx = dnxb; dnz = dnzb; gox = goxb; goz = gozb;
  for (int j = 1; j <= nnx; ++j)
    for (int k = 1; k <= nnz; ++k) VELN(k, j) = VELNB(k, j);
  for (int k = 1; k <= nnx; ++k)
    for (int l = 1; l <= nnz; ++l) {
      if (NSTS(l, k) == 0) {
        if (l - 1 >= 1) { if (NSTS(l - 1, k) == -1) NSTS(l, k) = 1; }
        if (l + 1 <= nnz) { if (NSTS(l + 1, k) == -1) NSTS(l, k) = 1; }
        if (k - 1 >= 1) { if (NSTS(l, k - 1) == -1) NSTS(l, k) = 1; }
        if (k + 1 <= nnx) { if (NSTS(l, k + 1) == -1) NSTS(l, k) = 1; }
      }
    }
  if (rec_rank) {                          // experiment only: record the coarse march alone (fim_experiment.cpp)
    rec_rank->assign(rec_rank->size(), -1);
    rec_count = 0;
    if (rec_init_nsts) *rec_init_nsts = nsts;
    if (rec_init_ttn) *rec_init_ttn = ttn;
  }
#ifdef ORC_WITH_EXPERIMENTS
  if (fim_coarse) return travel_fim();     // experiment only (fim_experiment.cpp; liboracle_experiments.so)
#endif
  return travel(x, z, 2);
}

// CalSurfG.f90:1599-1722
int Fmm::srtimes(float scx, float scz, float rcx1, float rcz1, float* cbst1) {
  int irx = (int)((rcx1 - gox) / dnx) + 1;
  int irz = (int)((rcz1 - goz) / dnz) + 1;
  int sw = 0;
  if (irx < 1 || irx > nnx) sw = 1;
  if (irz < 1 || irz > nnz) sw = 1;
  if (sw == 1) return ERR_RECEIVER_OUTSIDE;
  if (irx == nnx) irx = irx - 1;
  if (irz == nnz) irz = irz - 1;
  int isx = (int)((scx - gox) / dnx) + 1;
  int isz = (int)((scz - goz) / dnz) + 1;
  float dpl = dnx * earth;
  float rd1 = dnz * earth * sin_r(gox);
  if (rd1 < dpl) dpl = rd1;
  rd1 = dnz * earth * sin_r(gox + (float)(nnx - 1) * dnx);
  if (rd1 < dpl) dpl = rd1;
  float t1 = (scx - rcx1) * earth;
  float sred = t1 * t1;
  float t2 = (scz - rcz1) * earth * sin_r(rcx1);
  sred = sred + t2 * t2;
  sred = std::sqrt(sred);
  if (sred < dpl) sw = 1;
  if (isx == irx) { if (isz == irz) sw = 1; }
  float trr;
  if (sw == 1) {
    float vss[2][2];
    for (int k = 1; k <= 2; ++k)
      for (int l = 1; l <= 2; ++l) vss[k - 1][l - 1] = VELN(isz - 1 + l, isx - 1 + k);
    float drx = (scx - gox) - (float)(isx - 1) * dnx;
    float drz = (scz - goz) - (float)(isz - 1) * dnz;
    float vels = bilinear(vss, drx, drz);
    for (int k = 1; k <= 2; ++k)
      for (int l = 1; l <= 2; ++l) vss[k - 1][l - 1] = VELN(irz - 1 + l, irx - 1 + k);
    drx = (rcx1 - gox) - (float)(irx - 1) * dnx;
    drz = (rcz1 - goz) - (float)(irz - 1) * dnz;
    float velr = bilinear(vss, drx, drz);
    trr = 2.0f * sred / (vels + velr);
  } else {
    float drx = (rcx1 - gox) - (float)(irx - 1) * dnx;
    float drz = (rcz1 - goz) - (float)(irz - 1) * dnz;
    trr = 0.0f;
    for (int k = 1; k <= 2; ++k)
      for (int l = 1; l <= 2; ++l) {
        float produ = (1.0f - std::fabs(((float)(l - 1) * dnz - drz) / dnz)) *
                      (1.0f - std::fabs(((float)(k - 1) * dnx - drx) / dnx));
        trr = trr + TTN(irz - 1 + l, irx - 1 + k) * produ;
      }
  }
  *cbst1 = trr;
  return OK;
}

// rpathsAzim.f90:687-793.  No IMPLICIT NONE there: stalat, stalon, evtlat,
// evtlon, delta, az, baz, piby2, predel are REAL*4; pi is a DOUBLE assigned
// from a REAL*4 literal.
void azdist(float stalat, float stalon, float evtlat, float evtlon,
            float* delta, float* az, float* baz) {
  double pi = (double)3.1415926535898f;
  float piby2 = (float)(pi / 2.0);
  double rad = 2.0 * pi / 360.0;
  double sph = (double)(1.0f / 298.257f);
  double scolat = (double)piby2 - std::atan((1.0 - sph) * (1.0 - sph) * std::tan((double)stalat * rad));
  double ecolat = (double)piby2 - std::atan((1.0 - sph) * (1.0 - sph) * std::tan((double)evtlat * rad));
  double slon = (double)stalon * rad;
  double elon = (double)evtlon * rad;
  double a = std::sin(scolat) * std::cos(slon);
  double b = std::sin(scolat) * std::sin(slon);
  double c = std::cos(scolat);
  double d = std::sin(slon);
  double e = -std::cos(slon);
  double g = -c * e;
  double h = c * d;
  double k = -std::sin(scolat);
  double aa = std::sin(ecolat) * std::cos(elon);
  double bb = std::sin(ecolat) * std::sin(elon);
  double cc = std::cos(ecolat);
  double dd = std::sin(elon);
  double ee = -std::cos(elon);
  double gg = -cc * ee;
  double hh = cc * dd;
  double kk = -std::sin(ecolat);
  float predel = (float)(a * aa + b * bb + c * cc);
  if (std::fabs(predel + 1.0f) < .000001f) predel = -1.0f;
  if (std::fabs(predel - 1.0f) < .000001f) predel = 1.0f;
  double del = (double)acos_r(predel);
  *delta = (float)(del / rad);
  double rhs1 = (aa - d) * (aa - d) + (bb - e) * (bb - e) + cc * cc - 2.0;
  double rhs2 = (aa - g) * (aa - g) + (bb - h) * (bb - h) + (cc - k) * (cc - k) - 2.0;
  double dbaz = std::atan2(rhs1, rhs2);
  if (dbaz < 0.0) dbaz = dbaz + 2 * pi;
  *baz = (float)(dbaz / rad);
  rhs1 = (a - dd) * (a - dd) + (b - ee) * (b - ee) + c * c - 2.0;
  rhs2 = (a - gg) * (a - gg) + (b - hh) * (b - hh) + (c - kk) * (c - kk) - 2.0;
  double daz = std::atan2(rhs1, rhs2);
  if (daz < 0.0) daz = daz + 2 * pi;
  *az = (float)(daz / rad);
  if (std::fabs(*baz - 360.0f) < .00001f) *baz = 0.0f;
  if (std::fabs(*az - 360.0f) < .00001f) *az = 0.0f;
}

// delsph.f90:1-28
float delsph(float flat1, float flon1, float flat2, float flon2) {
  const float R = 6371.0f;
  const float pi = 3.1415926535898f;
  float dlat = flat2 - flat1;
  float dlon = flon2 - flon1;
  float lat1 = pi / 2 - flat1;
  float lat2 = pi / 2 - flat2;
  float a = std::sin(dlat / 2) * std::sin(dlat / 2) +
            std::sin(dlon / 2) * std::sin(dlon / 2) * std::cos(lat1) * std::cos(lat2);
  float c = 2 * std::atan2(std::sqrt(a), std::sqrt(1 - a));
  return R * c;
}

// rpathsAzim.f90:16-684 (azim) and rpaths CalSurfG.f90:1735-2291 (!azim).
// fdm/fdmc/fdms are (0:nvz+1,0:nvx+1) column-major, zeroed here like the
// reference does at :169-172.  Only rgx(j), rgx(j+1) are kept (the reference
// stores the whole path for the optional ray file).
int Fmm::rpaths(float scx, float scz, float* fdm, float* fdmc, float* fdms,
                float surfrcx, float surfrcz, bool azim) {
  const int ldf = nvz + 2;
  const size_t nf = (size_t)(nvz + 2) * (nvx + 2);
  auto F = [&](float* p, int iz, int ix) -> float& { return p[(size_t)ix * ldf + iz]; };
  const int maxrp = nnx * nnz;
  int isx, isz;
  if (asgr == 1) {
    isx = (int)((scx - goxr) / dnxr) + 1;
    isz = (int)((scz - gozr) / dnzr) + 1;
  } else {
    isx = (int)((scx - gox) / dnx) + 1;
    isz = (int)((scz - goz) / dnz) + 1;
  }
  float dpl = dnx * earth;
  float rd1 = dnz * earth * sin_r(gox);
  if (rd1 < dpl) dpl = rd1;
  rd1 = dnz * earth * sin_r(gox + (float)(nnx - 1) * dnx);
  if (rd1 < dpl) dpl = rd1;
  dpl = 0.5f * dpl;
  std::memset(fdm, 0, nf * sizeof(float));
  if (azim) {
    std::memset(fdmc, 0, nf * sizeof(float));
    std::memset(fdms, 0, nf * sizeof(float));
  }
  int ipx = (int)((surfrcx - gox) / dnx) + 1;
  int ipz = (int)((surfrcz - goz) / dnz) + 1;
  int sw = 0;
  if (ipx < 1 || ipx >= nnx) sw = 1;
  if (ipz < 1 || ipz >= nnz) sw = 1;
  if (sw == 1) return ERR_RECEIVER_OUTSIDE;
  if (ipx == nnx) ipx = ipx - 1;
  if (ipz == nnz) ipz = ipz - 1;
  float rgx_j = surfrcx, rgz_j = surfrcz, rgx_j1, rgz_j1;
  float sred;
  {
    float t1 = (scx - rgx_j) * earth;
    sred = t1 * t1;
    float t2 = (scz - rgz_j) * earth * sin_r(rgx_j);
    sred = sred + t2 * t2;
    sred = std::sqrt(sred);
  }
  if (sred < 2.0f * dpl) sw = 1;
  int ipxr = 0, ipzr = 0, igref;
  if (asgr == 1) {
    ipxr = (int)((surfrcx - goxr) / dnxr) + 1;
    ipzr = (int)((surfrcz - gozr) / dnzr) + 1;
    igref = 1;
    if (ipxr < 1 || ipxr >= nnxr) igref = 0;
    if (ipzr < 1 || ipzr >= nnzr) igref = 0;
    if (igref == 1) {
      if (NSTSR(ipzr, ipxr) != 0 || NSTSR(ipzr + 1, ipxr) != 0) igref = 0;
      if (NSTSR(ipzr, ipxr + 1) != 0 || NSTSR(ipzr + 1, ipxr + 1) != 0) igref = 0;
    }
  } else {
    igref = 0;
  }
  if (sw == 0) {
    if (asgr == 1) {
      if (igref == 1 && ipxr == isx && ipzr == isz) sw = 1;
    } else {
      if (ipx == isx && ipz == isz) sw = 1;
    }
  }
  for (int j = 1; j <= maxrp; ++j) {
    if (sw == 1) break;
    ++n_steps;
    float dtx, dtz;
    if (igref == 1) {
      dtx = TTNR(ipzr, ipxr + 1) - TTNR(ipzr, ipxr);
      dtx = dtx + TTNR(ipzr + 1, ipxr + 1) - TTNR(ipzr + 1, ipxr);
      dtx = dtx / (2.0f * earth * dnxr);
      dtz = TTNR(ipzr + 1, ipxr) - TTNR(ipzr, ipxr);
      dtz = dtz + TTNR(ipzr + 1, ipxr + 1) - TTNR(ipzr, ipxr + 1);
      dtz = dtz / (2.0f * earth * sin_r(rgx_j) * dnzr);
    } else {
      dtx = TTN(ipz, ipx + 1) - TTN(ipz, ipx);
      dtx = dtx + TTN(ipz + 1, ipx + 1) - TTN(ipz + 1, ipx);
      dtx = dtx / (2.0f * earth * dnx);
      dtz = TTN(ipz + 1, ipx) - TTN(ipz, ipx);
      dtz = dtz + TTN(ipz + 1, ipx + 1) - TTN(ipz, ipx + 1);
      dtz = dtz / (2.0f * earth * sin_r(rgx_j) * dnz);
    }
    rd1 = std::sqrt(dtx * dtx + dtz * dtz);
    rgx_j1 = rgx_j - dpl * dtx / (earth * rd1);
    rgz_j1 = rgz_j - dpl * dtz / (earth * sin_r(rgx_j) * rd1);
    int ipxo = ipx, ipzo = ipz;
    if (asgr == 1) {
      ipxr = (int)((rgx_j1 - goxr) / dnxr) + 1;
      ipzr = (int)((rgz_j1 - gozr) / dnzr) + 1;
      igref = 1;
      if (ipxr < 1 || ipxr >= nnxr) igref = 0;
      if (ipzr < 1 || ipzr >= nnzr) igref = 0;
      if (igref == 1) {
        if (NSTSR(ipzr, ipxr) != 0 || NSTSR(ipzr + 1, ipxr) != 0) igref = 0;
        if (NSTSR(ipzr, ipxr + 1) != 0 || NSTSR(ipzr + 1, ipxr + 1) != 0) igref = 0;
      }
      ipx = (int)((rgx_j1 - gox) / dnx) + 1;
      ipz = (int)((rgz_j1 - goz) / dnz) + 1;
    } else {
      ipx = (int)((rgx_j1 - gox) / dnx) + 1;
      ipz = (int)((rgz_j1 - goz) / dnz) + 1;
      igref = 0;
    }
    {
      float t1 = (scx - rgx_j1) * earth;
      sred = t1 * t1;
      float t2 = (scz - rgz_j1) * earth * sin_r(rgx_j1);
      sred = sred + t2 * t2;
      sred = std::sqrt(sred);
    }
    sw = 0;
    if (sred < 2.0f * dpl) sw = 1;
    if (sw == 0) {
      if (asgr == 1) {
        if (igref == 1 && ipxr == isx && ipzr == isz) sw = 1;
      } else {
        if (ipx == isx && ipz == isz) sw = 1;
      }
    }
    if (ipx < 1) { rgx_j1 = gox; ipx = 1; rbint = 1; }
    if (ipx >= nnx) { rgx_j1 = gox + (float)(nnx - 1) * dnx; ipx = nnx - 1; rbint = 1; }
    if (ipz < 1) { rgz_j1 = goz; ipz = 1; rbint = 1; }
    if (ipz >= nnz) { rgz_j1 = goz + (float)(nnz - 1) * dnz; ipz = nnz - 1; rbint = 1; }

    float c2 = 0.0f, s2 = 0.0f;
    if (azim) {
      float rgx1 = (PI_F / 2 - rgx_j) * 180.0f / PI_F;
      float rgz1 = rgz_j * 180.0f / PI_F;
      float rgx2 = (PI_F / 2 - rgx_j1) * 180.0f / PI_F;
      float rgz2 = rgz_j1 * 180.0f / PI_F;
      float delta, az, baz;
      azdist(rgx2, rgz2, rgx1, rgz1, &delta, &az, &baz);
      float rgpsi = az / 180 * PI_F;
      c2 = cos_r(2.0f * rgpsi);
      s2 = sin_r(2.0f * rgpsi);
    }
    int ivx = (ipx - 1) / gdx + 1;
    int ivz = (ipz - 1) / gdz + 1;
    int ivxo = (ipxo - 1) / gdx + 1;
    int ivzo = (ipzo - 1) / gdz + 1;
    int nhp = 0;
    float vrat[4];
    int chp[4];
    if (ivx != ivxo) {
      nhp = nhp + 1;
      float xi;
      if (ivx > ivxo) xi = gox + (float)(ivx - 1) * dvx;
      else xi = gox + (float)ivx * dvx;
      vrat[nhp - 1] = (xi - rgx_j) / (rgx_j1 - rgx_j);
      chp[nhp - 1] = 1;
    }
    if (ivz != ivzo) {
      nhp = nhp + 1;
      float zi;
      if (ivz > ivzo) zi = goz + (float)(ivz - 1) * dvz;
      else zi = goz + (float)ivz * dvz;
      rd1 = (zi - rgz_j) / (rgz_j1 - rgz_j);
      if (nhp == 1) {
        vrat[nhp - 1] = rd1;
        chp[nhp - 1] = 2;
      } else {
        if (rd1 >= vrat[nhp - 2]) {
          vrat[nhp - 1] = rd1;
          chp[nhp - 1] = 2;
        } else {
          vrat[nhp - 1] = vrat[nhp - 2];
          chp[nhp - 1] = chp[nhp - 2];
          vrat[nhp - 2] = rd1;
          chp[nhp - 2] = 2;
        }
      }
    }
    nhp = nhp + 1;
    vrat[nhp - 1] = 1.0f;
    chp[nhp - 1] = 0;
    float drx = (rgx_j - gox) - (float)(ipxo - 1) * dnx;
    float drz = (rgz_j - goz) - (float)(ipzo - 1) * dnz;
    float vel = 0.0f;
    for (int l = 1; l <= 2; ++l)
      for (int m = 1; m <= 2; ++m) {
        float produ = (1.0f - std::fabs(((float)(m - 1) * dnz - drz) / dnz));
        produ = produ * (1.0f - std::fabs(((float)(l - 1) * dnx - drx) / dnx));
        if (ipzo - 1 + m <= nnz && ipxo - 1 + l <= nnx) vel = vel + VELN(ipzo - 1 + m, ipxo - 1 + l) * produ;
      }
    drx = (rgx_j - gox) - (float)(ivxo - 1) * dvx;
    drz = (rgz_j - goz) - (float)(ivzo - 1) * dvz;
    float v = drx / dvx;
    float w = drz / dvz;
    float vi[4], wi[4], vio[4], wio[4];
    bspl_basis(v, vi);
    bspl_basis(w, wi);
    int ivxt = ivxo, ivzt = ivzo;
    for (int k = 1; k <= nhp; ++k) {
      float velo = vel;
      for (int q = 0; q < 4; ++q) { vio[q] = vi[q]; wio[q] = wi[q]; }
      if (k > 1) {
        if (chp[k - 2] == 1) ivxt = ivx;
        else if (chp[k - 2] == 2) ivzt = ivz;
      }
      float rigz = rgz_j + vrat[k - 1] * (rgz_j1 - rgz_j);
      float rigx = rgx_j + vrat[k - 1] * (rgx_j1 - rgx_j);
      int ipxt = (int)((rigx - gox) / dnx) + 1;
      int ipzt = (int)((rigz - goz) / dnz) + 1;
      drx = (rigx - gox) - (float)(ipxt - 1) * dnx;
      drz = (rigz - goz) - (float)(ipzt - 1) * dnz;
      vel = 0.0f;
      for (int m = 1; m <= 2; ++m)
        for (int n = 1; n <= 2; ++n) {
          float produ = (1.0f - std::fabs(((float)(n - 1) * dnz - drz) / dnz));
          produ = produ * (1.0f - std::fabs(((float)(m - 1) * dnx - drx) / dnx));
          if (ipzt - 1 + n <= nnz && ipxt - 1 + m <= nnx) vel = vel + VELN(ipzt - 1 + n, ipxt - 1 + m) * produ;
        }
      drx = (rigx - gox) - (float)(ivxt - 1) * dvx;
      drz = (rigz - goz) - (float)(ivzt - 1) * dvz;
      v = drx / dvx;
      w = drz / dvz;
      bspl_basis(v, vi);
      bspl_basis(w, wi);
      float dinc;
      if (k == 1) dinc = vrat[k - 1] * dpl;
      else dinc = (vrat[k - 1] - vrat[k - 2]) * dpl;
      for (int l = 1; l <= 4; ++l)
        for (int m = 1; m <= 4; ++m) {
          float rdc1 = vi[m - 1] * wi[l - 1] / (vel * vel);
          float rdc2 = vio[m - 1] * wio[l - 1] / (velo * velo);
          float r1 = -(rdc1 + rdc2) * dinc / 2.0f;
          float r2 = F(fdm, ivzt - 2 + l, ivxt - 2 + m);
          F(fdm, ivzt - 2 + l, ivxt - 2 + m) = r1 + r2;
          if (azim) {
            r1 = -(rdc1 * c2 + rdc2 * c2) * dinc / 2.0f;
            r2 = F(fdmc, ivzt - 2 + l, ivxt - 2 + m);
            F(fdmc, ivzt - 2 + l, ivxt - 2 + m) = r1 + r2;
            r1 = -(rdc1 * s2 + rdc2 * s2) * dinc / 2.0f;
            r2 = F(fdms, ivzt - 2 + l, ivxt - 2 + m);
            F(fdms, ivzt - 2 + l, ivxt - 2 + m) = r1 + r2;
          }
        }
    }
    rgx_j = rgx_j1;
    rgz_j = rgz_j1;
  }
  return OK;
}

}  // namespace orc
