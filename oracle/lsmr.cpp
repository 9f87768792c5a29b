// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the reference's sparse least-squares solve:
//   aprod      src/src_inv_iso_joint/aprod.f90:7-60     (COO mat-vec / transposed mat-vec, sequential in k)
//   LSMR       src/src_inv_iso_joint/lsmrModule.f90:36  (Fong & Saunders LSMR, local reorthogonalisation)
//   dnrm2      src/src_inv_iso_joint/lsmrblas.f90:247   (scaled sum of squares)
// Everything is single precision like the reference (lsmrDataModule.f90:21: dp = selected_real_kind(4));
// operation order follows the Fortran.
#include <cmath>
#include <cstddef>
#include <vector>
#include <algorithm>

namespace {

// aprod.f90:35-57
void aprod(int mode, long long nnz, const int* row, const int* col, const float* rw, float* x, float* y) {
  if (mode == 1) {
    for (long long k = 0; k < nnz; ++k) y[row[k] - 1] = y[row[k] - 1] + rw[k] * x[col[k] - 1];
  } else {
    for (long long k = 0; k < nnz; ++k) x[col[k] - 1] = x[col[k] - 1] + rw[k] * y[row[k] - 1];
  }
}

// lsmrblas.f90:247-277
float nrm2(int n, const float* x) {
  if (n < 1) return 0.0f;
  if (n == 1) return std::fabs(x[0]);
  float scale = 0.0f, ssq = 1.0f;
  for (int i = 0; i < n; ++i) {
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

// lsmrModule.f90:686-711
float d2norm(float a, float b) {
  const float scale = std::fabs(a) + std::fabs(b);
  if (scale == 0.0f) return 0.0f;
  const float ra = a / scale, rb = b / scale;
  return scale * std::sqrt(ra * ra + rb * rb);
}

}  // namespace

extern "C" {

struct OrcLsmrOut { int istop, itn; float normA, condA, normr, normAr, normx; };

// lsmrModule.f90:36-750
int orc_lsmr(int m, int n, long long nnz, const int* row, const int* col, const float* rw, const float* b,
             float damp, float atol, float btol, float conlim, int itnlim, int localSize, float* x, OrcLsmrOut* out) {
  const int localVecs = std::min(localSize, std::min(m, n));
  std::vector<float> h(n), hbar(n, 0.0f), u(b, b + m), v(n, 0.0f), w(n);
  std::vector<std::vector<float>> localV(std::max(localVecs, 0), std::vector<float>(n));
  for (int i = 0; i < n; ++i) x[i] = 0.0f;
  float alpha = 0.0f, beta = nrm2(m, u.data());
  if (beta > 0.0f) {
    const float f = 1.0f / beta;
    for (int i = 0; i < m; ++i) u[i] = f * u[i];
    aprod(2, nnz, row, col, rw, v.data(), u.data());
    alpha = nrm2(n, v.data());
  }
  if (alpha > 0.0f) {
    const float f = 1.0f / alpha;
    for (int i = 0; i < n; ++i) v[i] = f * v[i];
    w = v;
  }
  out->itn = 0; out->istop = 0; out->normA = 0; out->condA = 0; out->normx = 0;
  float normAr = alpha * beta;
  out->normAr = normAr; out->normr = beta;
  if (normAr == 0.0f) return 0;
  bool localOrtho = false, queueFull = false;
  int localPointer = 0;
  if (localVecs > 0) { localPointer = 1; localOrtho = true; localV[0] = v; }
  int itn = 0, istop = 0;
  float zetabar = alpha * beta, alphabar = alpha, rho = 1, rhobar = 1, cbar = 1, sbar = 0;
  h = v;
  float betadd = beta, betad = 0, rhodold = 1, tautildeold = 0, thetatilde = 0, zeta = 0, d = 0;
  float normA2 = alpha * alpha, maxrbar = 0.0f, minrbar = 1e+30f;
  const float normb = beta;
  float ctol = 0.0f;
  if (conlim > 0.0f) ctol = 1.0f / conlim;
  float normr = beta, normA = 0, condA = 0, normx = 0;
  const bool damped = damp > 0.0f;
  for (;;) {
    itn = itn + 1;
    for (int i = 0; i < m; ++i) u[i] = (-alpha) * u[i];
    aprod(1, nnz, row, col, rw, v.data(), u.data());
    beta = nrm2(m, u.data());
    if (beta > 0.0f) {
      const float f = 1.0f / beta;
      for (int i = 0; i < m; ++i) u[i] = f * u[i];
      if (localOrtho) {                       // localVEnqueue
        if (localPointer < localVecs) localPointer = localPointer + 1;
        else { localPointer = 1; queueFull = true; }
        localV[localPointer - 1] = v;
      }
      for (int i = 0; i < n; ++i) v[i] = (-beta) * v[i];
      aprod(2, nnz, row, col, rw, v.data(), u.data());
      if (localOrtho) {                       // localVOrtho
        const int lim = queueFull ? localVecs : localPointer;
        for (int c = 0; c < lim; ++c) {
          float dd = 0.0f;
          for (int i = 0; i < n; ++i) dd = dd + v[i] * localV[c][i];
          for (int i = 0; i < n; ++i) v[i] = v[i] - dd * localV[c][i];
        }
      }
      alpha = nrm2(n, v.data());
      if (alpha > 0.0f) {
        const float f2 = 1.0f / alpha;
        for (int i = 0; i < n; ++i) v[i] = f2 * v[i];
      }
    }
    const float alphahat = d2norm(alphabar, damp);
    const float chat = alphabar / alphahat, shat = damp / alphahat;
    const float rhoold = rho;
    rho = d2norm(alphahat, beta);
    const float c = alphahat / rho, s = beta / rho;
    const float thetanew = s * alpha;
    alphabar = c * alpha;
    const float rhobarold = rhobar, zetaold = zeta;
    const float thetabar = sbar * rho, rhotemp = cbar * rho;
    rhobar = d2norm(cbar * rho, thetanew);
    cbar = cbar * rho / rhobar;
    sbar = thetanew / rhobar;
    zeta = cbar * zetabar;
    zetabar = -sbar * zetabar;
    {
      const float c1 = thetabar * rho / (rhoold * rhobarold), c2 = zeta / (rho * rhobar), c3 = thetanew / rho;
      for (int i = 0; i < n; ++i) hbar[i] = h[i] - c1 * hbar[i];
      for (int i = 0; i < n; ++i) x[i] = x[i] + c2 * hbar[i];
      for (int i = 0; i < n; ++i) h[i] = v[i] - c3 * h[i];
    }
    const float betaacute = chat * betadd, betacheck = -shat * betadd;
    const float betahat = c * betaacute;
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
    normx = nrm2(n, x);
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
  if (damped && istop == 2) istop = 3;
  out->istop = istop; out->itn = itn; out->normA = normA; out->condA = condA; out->normr = normr;
  out->normAr = normAr; out->normx = normx;
  return 0;
}

}  // extern "C"
