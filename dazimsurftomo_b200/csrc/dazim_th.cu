// K1 (Rayleigh phase-velocity root search) and K2 (TI eigenfunctions, energy
// integrals and analytic partials) for sm_100a -- the Thomson-Haskell stage
// that feeds the eikonal solver (phase-velocity maps) and the G rows (depth
// kernels).
//
// Reference semantics that shape the design (SURVEY a2-a5, Q-list):
//  * surfdisp96 (surfdisp96.f:52) is sequential in period: the bracket search
//    at period k starts from c(k-1) - 1.5*dc and its float32-rounded root feeds
//    a finite-difference kernel (H3), so the bracket grid and the Neville path
//    must be followed exactly.  Parallelism is therefore across layered
//    profiles (node x perturbation variant): one thread per profile, profiles
//    stored layer-major so that a warp's loads are coalesced.
//  * tregn96 (tregn96_subroutine.f:49) is independent per (node, period): one
//    thread per (node, period) runs the Dunkin compound-matrix sweep upward,
//    then a single downward Haskell sweep that forms the eigenfunctions, the
//    six energy integrals and the per-layer partials on the fly.
// Arithmetic is float64 (float32 where the reference stores REAL*4), compiled
// with --fmad=false; complex multiply/divide are spelled out (naive product,
// Smith quotient) like gfortran expands them.
#include "dazim_dev.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace dz {

#define NLMAX 200
#define NPMAX 60

// ----------------------------------------------------------------------------
// model preparation shared by K1/K2
__device__ __forceinline__ void brocher_dev(float vs, float& vp, float& rho) {
  // depthkernelTI.f90:53-59 (x**n expanded like gfortran -O: x2=x*x, x3=x*x2, x4=x2*x2, x5=x2*x3)
  const float s2 = vs * vs, s3 = vs * (vs * vs), s4 = (vs * vs) * (vs * vs);
  vp = 0.9409f + 2.0947f * vs - 0.8206f * s2 + 0.2683f * s3 - 0.0251f * s4;
  const float p2 = vp * vp, p3 = vp * (vp * vp), p4 = (vp * vp) * (vp * vp), p5 = (vp * vp) * (vp * (vp * vp));
  rho = 1.6612f * vp - 0.4721f * p2 + 0.0671f * p3 - 0.0043f * p4 + 0.000106f * p5;
}

// number of sub-layers of interval i (refineLayerMdl, FwdTraveltimeCPS.f90:86-91)
__device__ __forceinline__ int nsub_of(float thk, float minthk0) {
  const float minthk = thk / minthk0;
  return (int)((thk + 1.0e-4f) / minthk) + 1;
}

// ----------------------------------------------------------------------------
// K1
struct DispArgs {
  int ntask;            // profiles
  int nlayer;           // layers incl. half-space (same for all profiles)
  int kmax;
  const double* t;      // [kmax]
  // flattened (sphere-corrected) model, layer-major [layer][task]
  const float* d; const float* a; const float* b; const float* rho;
  double* cg;           // [kmax][task] float32-rounded roots (0 where no root)
  int* noroot;          // flag
};

struct Ovr { double a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz; };

// var (surfdisp96.f:868-985)
__device__ __forceinline__ void var_dev(double p, double q, double ra, double rb, double wvno, double xka, double xkb,
                                        double dpth, Ovr& o) {
  double pex = 0.0, sex = 0.0, w = 0.0, x = 0.0, cosp = 0.0, y = 0.0, z = 0.0, cosq = 0.0, fac;
  if (wvno < xka) {
    double sinp;
    sincos(p, &sinp, &cosp);
    w = sinp / ra;
    x = -ra * sinp;
  } else if (wvno == xka) {
    cosp = 1.0; w = dpth; x = 0.0;
  } else {
    pex = p;
    fac = 0.0;
    if (p < 16) fac = exp(-2.0 * p);
    cosp = (1.0 + fac) * 0.5;
    const double sinp = (1.0 - fac) * 0.5;
    w = sinp / ra;
    x = ra * sinp;
  }
  if (wvno < xkb) {
    double sinq;
    sincos(q, &sinq, &cosq);
    y = sinq / rb;
    z = -rb * sinq;
  } else if (wvno == xkb) {
    cosq = 1.0; y = dpth; z = 0.0;
  } else {
    sex = q;
    fac = 0.0;
    if (q < 16) fac = exp(-2.0 * q);
    cosq = (1.0 + fac) * 0.5;
    const double sinq = (1.0 - fac) * 0.5;
    y = sinq / rb;
    z = rb * sinq;
  }
  const double exa = pex + sex;
  o.a0 = 0.0;
  if (exa < 60.0) o.a0 = exp(-exa);
  o.cpcq = cosp * cosq; o.cpy = cosp * y; o.cpz = cosp * z; o.cqw = cosq * w; o.cqx = cosq * x;
  o.xy = x * y; o.xz = x * z; o.wy = w * y; o.wz = w * z;
}

// e <- normalise(e * ca) with ca = Dunkin's matrix of one layer (dnka surfdisp96.f:1018, normc :989)
__device__ __forceinline__ void dunkin_step(double (&e)[5], double wvno2, double gam, double gammk, double rho,
                                            const Ovr& o) {
  const double gamm1 = gam - 1.0, twgm1 = gam + gamm1, gmgmk = gam * gammk, gmgm1 = gam * gamm1;
  const double gm1sq = gamm1 * gamm1, rho2 = rho * rho, a0pq = o.a0 - o.cpcq;
  double ca[5][5];
  ca[0][0] = o.cpcq - 2.0 * gmgm1 * a0pq - gmgmk * o.xz - wvno2 * gm1sq * o.wy;
  ca[0][1] = (wvno2 * o.cpy - o.cqx) / rho;
  ca[0][2] = -(twgm1 * a0pq + gammk * o.xz + wvno2 * gamm1 * o.wy) / rho;
  ca[0][3] = (o.cpz - wvno2 * o.cqw) / rho;
  ca[0][4] = -(2.0 * wvno2 * a0pq + o.xz + wvno2 * wvno2 * o.wy) / rho2;
  ca[1][0] = (gmgmk * o.cpz - gm1sq * o.cqw) * rho;
  ca[1][1] = o.cpcq;
  ca[1][2] = gammk * o.cpz - gamm1 * o.cqw;
  ca[1][3] = -o.wz;
  ca[1][4] = ca[0][3];
  ca[3][0] = (gm1sq * o.cpy - gmgmk * o.cqx) * rho;
  ca[3][1] = -o.xy;
  ca[3][2] = gamm1 * o.cpy - gammk * o.cqx;
  ca[3][3] = ca[1][1];
  ca[3][4] = ca[0][1];
  ca[4][0] = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * o.xz + gm1sq * gm1sq * o.wy) * rho2;
  ca[4][1] = ca[3][0];
  ca[4][2] = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * o.xz + gamm1 * gm1sq * o.wy) * rho;
  ca[4][3] = ca[1][0];
  ca[4][4] = ca[0][0];
  const double t = -2.0 * wvno2;
  ca[2][0] = t * ca[4][2];
  ca[2][1] = t * ca[3][2];
  ca[2][2] = o.a0 + 2.0 * (o.cpcq - ca[0][0]);
  ca[2][3] = t * ca[1][2];
  ca[2][4] = t * ca[0][2];
  double ee[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double cr = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) cr = cr + e[j] * ca[j][i];
    ee[i] = cr;
  }
  double t1 = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i)
    if (fabs(ee[i]) > t1) t1 = fabs(ee[i]);
  if (t1 < 1.e-40) t1 = 1.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) e[i] = ee[i] / t1;
}

// dltar4 (surfdisp96.f:767-865), solid layers (llw=1)
__device__ __noinline__ double dltar4_dev(const DispArgs& A, int task, double wvno, double omga) {
  const int mmax = A.nlayer;
  const size_t n = (size_t)A.ntask;
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  const double wvno2 = wvno * wvno;
  double e[5];
  {
    const size_t o = (size_t)(mmax - 1) * n + task;
    const double am = (double)A.a[o], bm = (double)A.b[o], rho1 = (double)A.rho[o];
    const double xka = omega / am, xkb = omega / bm;
    const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    const double t = bm / omega;
    const double gammk = 2.0 * t * t, gam = gammk * wvno2, gamm1 = gam - 1.0;
    e[0] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
    e[1] = -rho1 * ra;
    e[2] = rho1 * (gamm1 - gammk * ra * rb);
    e[3] = rho1 * rb;
    e[4] = wvno2 - ra * rb;
  }
  for (int m = mmax - 2; m >= 0; --m) {
    const size_t o = (size_t)m * n + task;
    const double am = (double)A.a[o], bm = (double)A.b[o], dpth = (double)A.d[o], rho1 = (double)A.rho[o];
    const double xka = omega / am, xkb = omega / bm;
    const double t = bm / omega;
    const double gammk = 2.0 * t * t, gam = gammk * wvno2;
    const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    const double p = ra * dpth, q = rb * dpth;
    Ovr ov;
    var_dev(p, q, ra, rb, wvno, xka, xkb, dpth, ov);
    dunkin_step(e, wvno2, gam, gammk, rho1, ov);
  }
  return e[0];
}

__device__ __forceinline__ double dsign1(double x) { return signbit(x) ? -1.0 : 1.0; }

// nevill (surfdisp96.f:551-668) with half (:670)
__device__ double nevill_dev(const DispArgs& A, int task, double t, double c1, double c2, double del1, double del2) {
  const double twopi = 2.0 * 3.141592653589793;
  double x[21], y[21];
  const double omega = twopi / t;
  double c3 = 0.5 * (c1 + c2);
  double del3 = dltar4_dev(A, task, omega / c3, omega);
  int nev = 1, nctrl = 1, mm = 1;
  for (;;) {
    nctrl = nctrl + 1;
    if (nctrl >= 100) break;
    if (c3 < fmin(c1, c2) || c3 > fmax(c1, c2)) {
      nev = 0;
      c3 = 0.5 * (c1 + c2);
      del3 = dltar4_dev(A, task, omega / c3, omega);
    }
    const double s13 = del1 - del3, s32 = del3 - del2;
    if (dsign1(del3) * dsign1(del1) < 0.0) { c2 = c3; del2 = del3; }
    else { c1 = c3; del1 = del3; }
    if (fabs(c1 - c2) <= 1.e-6 * c1) break;
    if (dsign1(s13) != dsign1(s32)) nev = 0;
    const double ss1 = fabs(del1), s1 = (double)0.01f * ss1, ss2 = fabs(del2), s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || nev == 0) {
      c3 = 0.5 * (c1 + c2);
      del3 = dltar4_dev(A, task, omega / c3, omega);
      nev = 1;
      mm = 1;
    } else {
      if (nev == 2) { x[mm + 1] = c3; y[mm + 1] = del3; }
      else { x[1] = c1; y[1] = del1; x[2] = c2; y[2] = del2; mm = 1; }
      bool bad = false;
      for (int kk = 1; kk <= mm; ++kk) {
        const int j = mm - kk + 1;
        const double denom = y[mm + 1] - y[j];
        if (fabs(denom) < 1.0e-10 * fabs(y[mm + 1])) { bad = true; break; }
        x[j] = (-y[j] * x[j + 1] + y[mm + 1] * x[j]) / denom;
      }
      if (!bad) {
        c3 = x[1];
        del3 = dltar4_dev(A, task, omega / c3, omega);
        nev = 2;
        mm = mm + 1;
        if (mm > 10) mm = 10;
      } else {
        c3 = 0.5 * (c1 + c2);
        del3 = dltar4_dev(A, task, omega / c3, omega);
        nev = 1;
        mm = 1;
      }
    }
  }
  return c3;
}

// surfdisp96 main loop (:186-305) + getsol (:384-476), fundamental mode, phase velocity
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_disp(DispArgs A) {
  const int task = blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= A.ntask) return;
  const int mmax = A.nlayer;
  const size_t n = (size_t)A.ntask;
  // extremal velocities (:139-150) and the half-space Rayleigh start (gtsolh :361-382), REAL*4
  float betmx = -1.e20f, betmn = 1.e20f;
  int jmn = 0;
  for (int i = 0; i < mmax; ++i) {
    const float bi = A.b[(size_t)i * n + task];
    if (bi > 0.01f && bi < betmn) { betmn = bi; jmn = i; }
    if (bi > betmx) betmx = bi;
  }
  float cc1;
  {
    const float a = A.a[(size_t)jmn * n + task], b = A.b[(size_t)jmn * n + task];
    float c = 0.95f * b;
    for (int i = 1; i <= 5; ++i) {
      const float gamma = b / a, kappa = c / b, k2 = kappa * kappa, gk2 = (gamma * kappa) * (gamma * kappa);
      const float fac1 = sqrtf(1.0f - gk2), fac2 = sqrtf(1.0f - k2);
      const float fr = (2.0f - k2) * (2.0f - k2) - 4.0f * fac1 * fac2;
      float frp = -4.0f * (2.0f - k2) * kappa + 4.0f * fac2 * gamma * gamma * kappa / fac1 + 4.0f * fac1 * kappa / fac2;
      frp = frp / b;
      c = c - fr / frp;
    }
    cc1 = c;
  }
  cc1 = .95f * cc1;
  cc1 = .90f * cc1;
  const double cc = (double)cc1;
  const double dc = fabs((double)0.005f);
  const double onea = (double)1.5f;
  const double cm = cc;
  const double twopi = 2.0 * 3.141592653589793;
  double cprev = 0.0, del1st = 0.0;
  int k;
  for (k = 0; k < A.kmax; ++k) {
    const double t1 = A.t[k];
    double c1, clow;
    const int ifirst = (k == 0);
    if (ifirst) { c1 = cc; clow = cc; }
    else { c1 = cprev - onea * dc; clow = cm; }
    // ---- getsol ----
    const double omega = twopi / t1;
    double del1 = dltar4_dev(A, task, omega / c1, omega);
    if (ifirst) del1st = del1;
    const double plmn = dsign1(del1st) * dsign1(del1);
    int idir = +1;
    if (!ifirst && plmn < 0.0) idir = -1;
    double c2, del2;
    int iret = 0;
    for (;;) {
      if (idir > 0) c2 = c1 + dc; else c2 = c1 - dc;
      if (c2 <= clow) { idir = +1; c1 = clow; continue; }
      del2 = dltar4_dev(A, task, omega / c2, omega);
      if (dsign1(del1) != dsign1(del2)) {
        c1 = nevill_dev(A, task, t1, c1, c2, del1, del2);
        iret = (c1 > (double)betmx) ? -1 : 1;
        break;
      }
      c1 = c2;
      del1 = del2;
      if (c1 < cm) { iret = -1; break; }
      if (c1 >= ((double)betmx + dc)) { iret = -1; break; }
    }
    if (iret == -1) break;
    cprev = c1;
    A.cg[(size_t)k * n + task] = (double)(float)c1;
  }
  if (k < A.kmax) {   // :307-348: warning + zero fill
    atomicOr(A.noroot, 1);
    for (; k < A.kmax; ++k) A.cg[(size_t)k * n + task] = 0.0;
  }
}

static void launch_disp(const DispArgs& D, unsigned nblk, cudaStream_t st) {
  int minb = 6;   // measured on S200: 4 -> 967 ms, 5 -> 905, 6 -> 872, 8 -> 875 (profiles/README.md)
  if (const char* e = getenv("DAZIM_KDISP_MINB")) minb = atoi(e);
  if (minb >= 8) k_disp<8><<<nblk, 128, 0, st>>>(D);
  else if (minb >= 6) k_disp<6><<<nblk, 128, 0, st>>>(D);
  else if (minb == 5) k_disp<5><<<nblk, 128, 0, st>>>(D);
  else k_disp<4><<<nblk, 128, 0, st>>>(D);
}

// ---- profile construction -------------------------------------------------
// One thread per (node, variant).  variant 0 = unperturbed; 1+6*i+j perturbs depth node i:
// j = 0/1 Vs -/+, 2/3 Vp -/+, 4/5 rho -/+ (depthkernel, CalSurfG.f90:76-130).
struct ProfArgs {
  int nx, ny, nz, nvar;
  const float* vel;     // (nx,ny,nz)
  const float* depz;    // [nz]
  float minthk;
  int nlayer;
  float* d; float* a; float* b; float* rho;   // [layer][task], task = node*nvar + variant
  // optional raw (un-flattened) model for the TI stage, [layer][node] (variant 0 only)
  float* rthk; float* rvp; float* rvs; float* rrho;
};

__global__ void k_profiles(ProfArgs P) {
  const int task = blockIdx.x * blockDim.x + threadIdx.x;
  const int nnode = P.nx * P.ny;
  const int ntask = nnode * P.nvar;
  if (task >= ntask) return;
  const int node = task / P.nvar, var = task % P.nvar;
  const int nz = P.nz;
  const int pi_ = (var > 0) ? (var - 1) / 6 : -1, pj = (var > 0) ? (var - 1) % 6 : -1;
  const float dln = 0.01f;
  // refine (refineGrid2LayerMdl, CalSurfG.f90:2317) directly into the layer-major arrays, then flatten
  const size_t n = (size_t)ntask;
  int k = 0;
  float vs0, vp0, rh0;
  {
    vs0 = P.vel[(size_t)node];
    brocher_dev(vs0, vp0, rh0);
    if (pi_ == 0) {
      if (pj == 0) vs0 = vs0 - 0.5f * dln * vs0; else if (pj == 1) vs0 = vs0 + 0.5f * dln * vs0;
      else if (pj == 2) vp0 = vp0 - 0.5f * dln * vp0; else if (pj == 3) vp0 = vp0 + 0.5f * dln * vp0;
      else if (pj == 4) rh0 = rh0 - 0.5f * dln * rh0; else rh0 = rh0 + 0.5f * dln * rh0;
    }
  }
  // sphere(0,0) + sphere(2,1) (surfdisp96.f:480-547) fused with the refinement loop
  const double ar = 6370.0;
  double dr = 0.0, r0 = ar;
  for (int i = 1; i <= nz - 1; ++i) {
    float vs1 = P.vel[(size_t)node + (size_t)i * nnode], vp1, rh1;
    brocher_dev(vs1, vp1, rh1);
    if (pi_ == i) {
      if (pj == 0) vs1 = vs1 - 0.5f * dln * vs1; else if (pj == 1) vs1 = vs1 + 0.5f * dln * vs1;
      else if (pj == 2) vp1 = vp1 - 0.5f * dln * vp1; else if (pj == 3) vp1 = vp1 + 0.5f * dln * vp1;
      else if (pj == 4) rh1 = rh1 - 0.5f * dln * rh1; else rh1 = rh1 + 0.5f * dln * rh1;
    }
    const float thk = P.depz[i] - P.depz[i - 1];
    const int ns = nsub_of(thk, P.minthk);
    const float newthk = thk / (float)ns;
    for (int j = 1; j <= ns; ++j) {
      const float rvp = vp0 + (float)(2 * j - 1) * (vp1 - vp0) / (float)(2 * ns);
      const float rvs = vs0 + (float)(2 * j - 1) * (vs1 - vs0) / (float)(2 * ns);
      const float rrho = rh0 + (float)(2 * j - 1) * (rh1 - rh0) / (float)(2 * ns);
      if (var == 0 && P.rthk) {
        const size_t q = (size_t)k * nnode + node;
        P.rthk[q] = newthk; P.rvp[q] = rvp; P.rvs[q] = rvs; P.rrho[q] = rrho;
      }
      // flatten this layer
      dr = dr + (double)newthk;
      const double r1 = ar - dr;
      const double z0 = ar * log(ar / r0), z1 = ar * log(ar / r1);
      const double tmp = (ar + ar) / (r0 + r1);
      const float btp = (float)tmp;
      const size_t o = (size_t)k * n + task;
      P.d[o] = (float)(z1 - z0);
      P.a[o] = (float)((double)rvp * tmp);
      P.b[o] = (float)((double)rvs * tmp);
      P.rho[o] = rrho * (float)pow((double)btp, (double)-2.275f);
      r0 = r1;
      ++k;
    }
    vs0 = vs1; vp0 = vp1; rh0 = rh1;
  }
  // half-space (d = 0 after the transform; thickness 1.0 is used for the mid-point factor)
  {
    if (var == 0 && P.rthk) {
      const size_t q = (size_t)k * nnode + node;
      P.rthk[q] = 0.0f; P.rvp[q] = vp0; P.rvs[q] = vs0; P.rrho[q] = rh0;
    }
    dr = dr + (double)1.0f;
    const double r1 = ar - dr;
    const double tmp = (ar + ar) / (r0 + r1);
    const float btp = (float)tmp;
    const size_t o = (size_t)k * n + task;
    P.d[o] = 0.0f;
    P.a[o] = (float)((double)vp0 * tmp);
    P.b[o] = (float)((double)vs0 * tmp);
    P.rho[o] = rh0 * (float)pow((double)btp, (double)-2.275f);
  }
}

// sen_* = (cg2-cg1)/(dln*par) (CalSurfG.f90:90-129), pvRc = unperturbed root
__global__ void k_fdkernels(int nx, int ny, int nz, int kmax, int nvar, const float* __restrict__ vel,
                            const double* __restrict__ cg, double* __restrict__ pv, double* __restrict__ sen_vs,
                            double* __restrict__ sen_vp, double* __restrict__ sen_rho) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nnode = nx * ny;
  if (idx >= nnode * kmax) return;
  const int node = idx % nnode, k = idx / nnode;
  const size_t ntask = (size_t)nnode * nvar;
  const size_t base = (size_t)k * ntask + (size_t)node * nvar;
  pv[idx] = cg[base];
  if (nvar == 1) return;
  const float dln = 0.01f;
  for (int i = 0; i < nz; ++i) {
    const float vs = vel[(size_t)node + (size_t)i * nnode];
    float vp, rho;
    brocher_dev(vs, vp, rho);
    const size_t o = (size_t)idx + (size_t)i * nnode * kmax;
    const double* c = cg + base + 1 + 6 * i;
    sen_vs[o] = (c[1] - c[0]) / (double)(dln * vs);
    sen_vp[o] = (c[3] - c[2]) / (double)(dln * vp);
    sen_rho[o] = (c[5] - c[4]) / (double)(dln * rho);
  }
}

// ----------------------------------------------------------------------------
// K2
struct Z { double re, im; };
__device__ __forceinline__ Z mkz(double r, double i = 0.0) { Z z; z.re = r; z.im = i; return z; }
__device__ __forceinline__ Z operator+(Z a, Z b) { return mkz(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ Z operator-(Z a, Z b) { return mkz(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ Z operator-(Z a) { return mkz(-a.re, -a.im); }
__device__ __forceinline__ Z operator*(Z a, Z b) { return mkz(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ Z operator*(double r, Z a) { return mkz(r * a.re, r * a.im); }
__device__ __forceinline__ Z operator*(Z a, double r) { return mkz(a.re * r, a.im * r); }
__device__ __forceinline__ Z operator/(Z a, double r) { return mkz(a.re / r, a.im / r); }
__device__ __forceinline__ Z operator/(Z n, Z d) {
  if (fabs(d.re) < fabs(d.im)) {
    const double ratio = d.re / d.im, denom = d.re * ratio + d.im;
    return mkz((n.re * ratio + n.im) / denom, (n.im * ratio - n.re) / denom);
  }
  const double ratio = d.im / d.re, denom = d.im * ratio + d.re;
  return mkz((n.im * ratio + n.re) / denom, (n.im - n.re * ratio) / denom);
}
__device__ __forceinline__ double zabs(Z a) { return hypot(a.re, a.im); }
__device__ __forceinline__ Z zconj(Z a) { return mkz(a.re, -a.im); }
__device__ __forceinline__ Z zexp(Z a) { double s, c; sincos(a.im, &s, &c); const double e = exp(a.re); return mkz(e * c, e * s); }
__device__ __forceinline__ Z zsqrt(Z a) {
  if (a.im == 0.0) {
    if (a.re >= 0.0) return mkz(sqrt(a.re), a.im);
    return mkz(0.0, copysign(sqrt(-a.re), a.im));
  }
  const double t = sqrt((fabs(a.re) + hypot(a.re, a.im)) * 0.5);
  if (a.re >= 0.0) return mkz(t, a.im / (2.0 * t));
  return mkz(fabs(a.im) / (2.0 * t), copysign(t, a.im));
}

struct Eig { Z rp, rsv, x11, x21, x31, x41, x12, x22, x32, x42, np, nsv; };

struct EigenArgs {
  int nnode, kmax, nlayer, nz;
  const double* t;          // [kmax]
  const double* pv;         // (nnode,kmax) float32-rounded phase velocities
  // TI model after sphere_tdisp96, layer-major [layer][node]
  const double* zd; const double* zta; const double* ztc; const double* ztf; const double* ztl; const double* ztn;
  const double* zrho; const float* vtp;
  // raw refined model [layer][node] for the fold (depthkernelTI.f90:96-106)
  const float* rvp; const float* rvs; const float* rrho; const float* TA; const float* TL; const float* TF;
  const float* depz; float minthk;
  double* scratch;          // [layer][11][task]
  float* lsen;              // (nnode,kmax,nz-1)
};

// gettiegn (tregn96_subroutine.f:3163-3358), solid layers
__device__ void gettiegn_dev(double TA, double TC, double TF, double TL, double TRho, double omg, double wvn,
                             double wvno2, Eig& o) {
  const Z a = mkz(wvn * TF / (TC));
  const Z b = mkz(1.0 / (TC));
  const Z c = mkz(-TRho * omg * omg + wvn * wvn * (TA - TF * TF / (TC)));
  const Z d = mkz(-wvn);
  const Z e = mkz(1.0 / (TL));
  const Z f = mkz(-TRho * omg * omg);
  const Z ddef = mkz(wvn * wvn - TRho * omg * omg / (TL));
  const Z aabc = mkz(wvn * wvn * TA / TC - TRho * omg * omg / (TC));
  const Z bb = 2.0 * a * d + e * c + f * b;
  const Z cc = ddef * aabc;
  Z srt = zsqrt(bb * bb - 4.0 * cc);
  if (srt.im < 0.0) srt = -srt;
  Z L1, L2;
  if (bb.re < 0.0 && srt.re < 0.0) {
    L2 = (bb - srt) / 2.0;
    if (zabs(L2) > 0.0) L1 = cc / L2; else L1 = (bb + srt) / 2.0;
  } else {
    L1 = (bb + srt) / 2.0;
    if (zabs(L1) > 0.0) L2 = cc / L1; else L2 = (bb - srt) / 2.0;
  }
  const Z xka2 = mkz(wvno2) - L1, xkb2 = mkz(wvno2) - L2;
  if (zabs(xkb2) < zabs(xka2)) { const Z t = L1; L1 = L2; L2 = t; }
  o.rp = zsqrt(L1);
  o.rsv = zsqrt(L2);
  if (o.rp.re < 0.0) o.rp = -o.rp;
  if (o.rsv.re < 0.0) o.rsv = -o.rsv;
  o.x12 = (b * d - a * e);
  o.x22 = b * L2 - e * (b * c + a * a);
  o.x32 = L2 - (a * d + c * e);
  o.x42 = -a * L2 + d * (b * c + a * a);
  o.x11 = -e * L1 + b * (d * d + e * f);
  o.x21 = (b * d - a * e);
  o.x31 = d * L1 - a * (d * d + e * f);
  o.x41 = -(L1 - a * d - b * f);
  if (wvn != 0.0) {
    Z zf = mkz(wvn) / o.x11;
    o.x11 = o.x11 * zf; o.x21 = o.x21 * zf; o.x31 = o.x31 * zf; o.x41 = o.x41 * zf;
    zf = mkz(wvn) / o.x22;
    o.x12 = o.x12 * zf; o.x22 = o.x22 * zf; o.x32 = o.x32 * zf; o.x42 = o.x42 * zf;
  }
  o.np = o.x11 * o.x41 - o.x21 * o.x31;
  o.nsv = o.x12 * o.x42 - o.x22 * o.x32;
}

struct Trig { Z cosp, cosq, rsinp, rsinq, sinpr, sinqr; double pex, svex; };

// varsv (:3360-3472), solid
__device__ void varsv_dev(Z p, Z q, Z rp, Z rsv, double dm, Trig& o) {
  const double pr = p.re, pi = p.im, qr = q.re, qi = q.im;
  o.pex = pr; o.svex = qr;
  double s, c;
  sincos(pi, &s, &c);
  const Z epp = mkz(c, s) / 2.0, epm = zconj(epp);
  sincos(qi, &s, &c);
  const Z eqp = mkz(c, s) / 2.0, eqm = zconj(eqp);
  const double pfac = (pr < 15.) ? exp(-2. * pr) : 0.0;
  o.cosp = (epp + pfac * epm);
  const Z sinp = epp - pfac * epm;
  o.rsinp = (rp * sinp);
  if (fabs(pr) < (double)1.0e-5f && zabs(rp) < (double)1.0e-5f) o.sinpr = mkz(dm); else o.sinpr = (sinp / rp);
  const double svfac = (qr < 15.) ? exp(-2. * qr) : 0.0;
  o.cosq = (eqp + svfac * eqm);
  const Z sinq = eqp - svfac * eqm;
  o.rsinq = (rsv * sinq);
  if (fabs(qr) < (double)1.0e-5f && zabs(rsv) < (double)1.0e-5f) o.sinqr = mkz(dm); else o.sinqr = (sinq / rsv);
}

// ee = cd * CA(layer) with CA the TI compound matrix (dnka_tregn :1989-2984), then cnormc (:1472)
__device__ void compound_step(Z (&cd)[5], const Eig& g, const Trig& tr, double& exn) {
  const double ex = tr.pex + tr.svex;
  const double dfac = (ex > 35.0) ? 0.0 : exp(-ex);
  const Z a1 = mkz(0.5) / g.np, a2 = mkz(0.5) / g.nsv;
  const Z c1 = 2. * a1 * tr.cosp, ls1 = 2. * a1 * tr.rsinp, s1l = 2. * a1 * tr.sinpr;
  const Z c2 = 2. * a2 * tr.cosq, ls2 = 2. * a2 * tr.rsinq, s2l = 2. * a2 * tr.sinqr;
  Z x[5][3];
  x[1][1] = g.x11; x[2][1] = g.x21; x[3][1] = g.x31; x[4][1] = g.x41;
  x[1][2] = g.x12; x[2][2] = g.x22; x[3][2] = g.x32; x[4][2] = g.x42;
  Z tca11, tca12, tca13, tca15, tca16, tca21, tca22, tca23, tca25, tca31, tca32, tca33, tca51, tca52, tca61;
#include "dnka_tca_dev.inc"
  Z ca[5][5];
  ca[0][0] = tca11; ca[0][1] = tca12; ca[0][2] = tca13; ca[0][3] = tca15; ca[0][4] = tca16;
  ca[1][0] = tca21; ca[1][1] = tca22; ca[1][2] = tca23; ca[1][3] = tca25; ca[1][4] = tca15;
  ca[2][0] = 2.0 * tca31; ca[2][1] = 2.0 * tca32; ca[2][2] = 2.0 * tca33 - mkz(dfac); ca[2][3] = -2.0 * tca23; ca[2][4] = -2.0 * tca13;
  ca[3][0] = tca51; ca[3][1] = tca52; ca[3][2] = -tca32; ca[3][3] = tca22; ca[3][4] = tca12;
  ca[4][0] = tca61; ca[4][1] = tca51; ca[4][2] = -tca31; ca[4][3] = tca21; ca[4][4] = tca11;
  Z ee[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    Z cr = mkz(0.0);
#pragma unroll
    for (int j = 0; j < 5; ++j) cr = cr + cd[j] * ca[j][i];
    ee[i] = cr;
  }
  double t1 = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) { const double v = zabs(ee[i]); if (v > t1) t1 = v; }
  if (t1 < 1.e-40) t1 = 1.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) cd[i] = ee[i] / t1;
  exn = log(t1);
}

// E, E^-1 of a layer (evalg :2986-3161, solid)
__device__ void emat_dev(const Eig& g, Z (&E)[4][4], Z (&G)[4][4]) {
  const Z rp = g.rp, rsv = g.rsv, NP = g.np, NSV = g.nsv;
  G[0][0] = g.x41 * rp / (2. * rp * NP);     G[1][0] = g.x42 / (2. * rsv * NSV);
  G[2][0] = -g.x41 * rp / (-2. * rp * NP);   G[3][0] = g.x42 / (-2. * rsv * NSV);
  G[0][1] = -g.x31 / (2. * rp * NP);         G[1][1] = -g.x32 * rsv / (2. * rsv * NSV);
  G[2][1] = -g.x31 / (-2. * rp * NP);        G[3][1] = g.x32 * rsv / (-2. * rsv * NSV);
  G[0][2] = -g.x21 * rp / (2. * rp * NP);    G[1][2] = -g.x22 / (2. * rsv * NSV);
  G[2][2] = g.x21 * rp / (-2. * rp * NP);    G[3][2] = -g.x22 / (-2. * rsv * NSV);
  G[0][3] = g.x11 / (2. * rp * NP);          G[1][3] = g.x12 * rsv / (2. * rsv * NSV);
  G[2][3] = g.x11 / (-2. * rp * NP);         G[3][3] = -g.x12 * rsv / (-2. * rsv * NSV);
  E[0][0] = g.x11;        E[1][0] = g.x21 * rp;   E[2][0] = g.x31;        E[3][0] = g.x41 * rp;
  E[0][1] = g.x12 * rsv;  E[1][1] = g.x22;        E[2][1] = g.x32 * rsv;  E[3][1] = g.x42;
  E[0][2] = g.x11;        E[1][2] = -g.x21 * rp;  E[2][2] = g.x31;        E[3][2] = -g.x41 * rp;
  E[0][3] = -g.x12 * rsv; E[1][3] = g.x22;        E[2][3] = -g.x32 * rsv; E[3][3] = g.x42;
}

__device__ __forceinline__ Z ffunc_dev(Z nub, double dm) {
  if (zabs(nub) < 1.0e-08) return mkz(dm);
  const Z arg = nub * dm;
  const Z ex = (arg.re < 40.0) ? zexp(-2.0 * arg) : mkz(0.0);
  return (mkz(1.0) - ex) / (2.0 * nub);
}
__device__ __forceinline__ Z gfunc_dev(Z nub, double dm) {
  const Z arg = nub * dm;
  if (arg.re < 75) return zexp(-arg) * dm;
  return mkz(0.0);
}
__device__ __forceinline__ Z h1func_dev(Z nua, Z nub, double dm) {
  if (zabs(nub + nua) < 1.0e-08) return mkz(dm);
  const Z arg = (nua + nub) * dm;
  const Z ex = (arg.re < 40.0) ? zexp(-arg) : mkz(0.0);
  return (mkz(1.0) - ex) / (nub + nua);
}
__device__ __forceinline__ Z h2func_dev(Z nua, Z nub, double dm) {
  if (zabs(nub - nua) < 1.0e-08) return mkz(dm);
  Z arg = nua * dm;
  const Z exqp = (arg.re < 40.0) ? zexp(-arg) : mkz(0.0);
  arg = nub * dm;
  const Z exqq = (arg.re < 40.0) ? zexp(-arg) : mkz(0.0);
  return (exqq - exqp) / (nua - nub);
}

// the six integrals of energy (:3811-3822) for one layer; up = eigenfunctions at the top
// interface (ur,uz,tz,tr)(m), dn = at the bottom interface (m+1)
struct Ints { double i11, i13, i22, i24, i33, i44; };
__device__ void layer_integrals(const Z (&e)[4][4], const Z (&einv)[4][4], Z ra, Z rb, const double (&up)[4],
                                const double (&dn)[4], double dm, bool halfspace, Ints& out) {
  const Z km1pd = einv[2][0] * up[0] + einv[2][1] * up[1] + einv[2][2] * up[2] + einv[2][3] * up[3];
  const Z km1sd = einv[3][0] * up[0] + einv[3][1] * up[1] + einv[3][2] * up[2] + einv[3][3] * up[3];
  const int II[6] = {0, 0, 1, 1, 2, 3}, JJ[6] = {0, 2, 1, 3, 2, 3};
  double r[6];
  if (!halfspace) {
    const Z kmpu = einv[0][0] * dn[0] + einv[0][1] * dn[1] + einv[0][2] * dn[2] + einv[0][3] * dn[3];
    const Z kmsu = einv[1][0] * dn[0] + einv[1][1] * dn[1] + einv[1][2] * dn[2] + einv[1][3] * dn[3];
    const Z FA = ffunc_dev(ra, dm), GA = gfunc_dev(ra, dm), FB = ffunc_dev(rb, dm), GB = gfunc_dev(rb, dm);
    const Z H1 = h1func_dev(ra, rb, dm), H2 = h2func_dev(ra, rb, dm);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int I = II[q], J = JJ[q];
      const Z c = e[I][0] * e[J][0] * kmpu * kmpu * FA
                + e[I][2] * e[J][2] * km1pd * km1pd * FA
                + e[I][1] * e[J][1] * kmsu * kmsu * FB
                + e[I][3] * e[J][3] * km1sd * km1sd * FB
                + H1 * ((e[I][0] * e[J][1] + e[I][1] * e[J][0]) * kmpu * kmsu +
                        (e[I][2] * e[J][3] + e[I][3] * e[J][2]) * km1pd * km1sd)
                + H2 * ((e[I][0] * e[J][3] + e[I][3] * e[J][0]) * kmpu * km1sd +
                        (e[I][1] * e[J][2] + e[I][2] * e[J][1]) * km1pd * kmsu)
                + GA * (e[I][0] * e[J][2] + e[I][2] * e[J][0]) * kmpu * km1pd
                + GB * (e[I][1] * e[J][3] + e[I][3] * e[J][1]) * kmsu * km1sd;
      r[q] = c.re;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int I = II[q], J = JJ[q];
      const Z c = e[I][2] * e[J][2] * km1pd * km1pd / (2.0 * ra)
                + (e[I][2] * e[J][3] + e[I][3] * e[J][2]) * km1pd * km1sd / (ra + rb)
                + e[I][3] * e[J][3] * km1sd * km1sd / (2.0 * rb);
      r[q] = c.re;
    }
  }
  out.i11 = r[0]; out.i13 = r[1]; out.i22 = r[2]; out.i24 = r[3]; out.i33 = r[4]; out.i44 = r[5];
}

#define NSLOT 11
__global__ void __launch_bounds__(64) k_eigen(EigenArgs A) {
  const int task = blockIdx.x * blockDim.x + threadIdx.x;
  const int ntask = A.nnode * A.kmax;
  if (task >= ntask) return;
  const int node = task % A.nnode, per = task / A.nnode;
  const int mmax = A.nlayer;
  const size_t nn = (size_t)A.nnode;
  const size_t nt = (size_t)ntask;
#define SCR(m, s) A.scratch[((size_t)(m) * NSLOT + (s)) * nt + task]
#define MOD(arr, m) arr[(size_t)(m) * nn + node]
  const float twopi = 2.f * 3.141592654f;                  // REAL*4, tregn96_subroutine.f:435
  const double t = (double)(float)A.t[per];
  const double omega = (double)twopi / t;
  double cph = (double)(float)A.pv[(size_t)node + (size_t)per * nn];
  double wvno = omega / cph;
  if (!(cph > 0.0)) {   // no root at this period: the reference would divide by zero; emit zeros
    for (int j = 0; j < A.nz - 1; ++j) A.lsen[(size_t)node + (size_t)per * nn + (size_t)j * nn * A.kmax] = 0.0f;
    return;
  }
  const double om2 = omega * omega, wvno2 = wvno * wvno;
  Eig g;
  // ---- up (:1831-1987): compound vector from the half-space to the surface ----
  Z cd[5];
  {
    gettiegn_dev(MOD(A.zta, mmax - 1), MOD(A.ztc, mmax - 1), MOD(A.ztf, mmax - 1), MOD(A.ztl, mmax - 1),
                 MOD(A.zrho, mmax - 1), omega, wvno, wvno2, g);
    Z E[4][4], G[4][4];
    emat_dev(g, E, G);
    cd[0] = mkz((G[0][0] * G[1][1] - G[0][1] * G[1][0]).re);
    cd[1] = mkz((G[0][0] * G[1][2] - G[0][2] * G[1][0]).re);
    cd[2] = mkz((G[0][0] * G[1][3] - G[0][3] * G[1][0]).re);
    cd[3] = mkz((G[0][1] * G[1][3] - G[0][3] * G[1][1]).re);
    cd[4] = mkz((G[0][2] * G[1][3] - G[0][3] * G[1][2]).re);
  }
  double exsum = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) { SCR(mmax - 1, 2 * i) = cd[i].re; }
  SCR(mmax - 1, 10) = 0.0;
  for (int m = mmax - 2; m >= 0; --m) {
    const double dm = MOD(A.zd, m);
    gettiegn_dev(MOD(A.zta, m), MOD(A.ztc, m), MOD(A.ztf, m), MOD(A.ztl, m), MOD(A.zrho, m), omega, wvno, wvno2, g);
    Trig tr;
    varsv_dev(g.rp * dm, g.rsv * dm, g.rp, g.rsv, dm, tr);
    double exn;
    compound_step(cd, g, tr, exn);
    exsum = exsum + tr.pex + tr.svex + exn;
#pragma unroll
    for (int i = 0; i < 5; ++i) SCR(m, 2 * i) = cd[i].re;   // only the real parts are used downstream (:1705-1710)
    SCR(m, 10) = exsum;
  }
  const double exe1 = exsum;
  const double f1213 = -cd[1].re;
  // ---- down (:3558-3706) fused with svfunc (:1616-1740) and energy (:3774-3991) ----
  double up[4], dn[4];
  up[0] = (cd[2] / cd[1]).re; up[1] = 1.0; up[2] = 0.0; up[3] = 0.0;   // ur, uz, tz, tr at the surface
  double vv[4] = {1.0, 0.0, 0.0, 0.0};
  double exa = 0.0;
  double sumi0 = 0.0, sumi1 = 0.0, sumi2 = 0.0;
  for (int m = 0; m < mmax; ++m) {
    const double TA = MOD(A.zta, m), TC = MOD(A.ztc, m), TF = MOD(A.ztf, m), TL = MOD(A.ztl, m), rho = MOD(A.zrho, m);
    const double dm = MOD(A.zd, m);
    gettiegn_dev(TA, TC, TF, TL, rho, omega, wvno, wvno2, g);
    const bool half = (m == mmax - 1);
    if (!half) {
      Trig tr;
      varsv_dev(g.rp * dm, g.rsv * dm, g.rp, g.rsv, dm, tr);
      // hska (:3474-3556) with the larger exponent factored out (:3654-3679)
      double dfac, cpex;
      Z tc = tr.cosp, trs = tr.rsinp, tsr = tr.sinpr, qc = tr.cosq, qrs = tr.rsinq, qsr = tr.sinqr;
      if (tr.pex > tr.svex) {
        dfac = ((tr.pex - tr.svex) > 40.0) ? 0.0 : exp(-(tr.pex - tr.svex));
        cpex = tr.pex;
        qc = dfac * qc; qrs = dfac * qrs; qsr = dfac * qsr;
      } else {
        dfac = ((tr.svex - tr.pex) > 40.0) ? 0.0 : exp(-(tr.svex - tr.pex));
        cpex = tr.svex;
        tc = dfac * tc; trs = dfac * trs; tsr = dfac * tsr;
      }
      const Z cosp = tc / g.np, sinpr = tsr / g.np, rsinp = trs / g.np;
      const Z cossv = qc / g.nsv, sinsvr = qsr / g.nsv, rsinsv = qrs / g.nsv;
      double AA[4][4];
      AA[0][0] = (g.x11 * g.x41 * cosp + g.x12 * g.x42 * cossv).re;
      AA[0][1] = (-g.x11 * g.x31 * sinpr - g.x12 * g.x32 * rsinsv).re;
      AA[0][2] = (-g.x11 * g.x21 * cosp - g.x12 * g.x22 * cossv).re;
      AA[0][3] = (g.x11 * g.x11 * sinpr + g.x12 * g.x12 * rsinsv).re;
      AA[1][0] = (g.x21 * g.x41 * rsinp + g.x22 * g.x42 * sinsvr).re;
      AA[1][1] = (-g.x21 * g.x31 * cosp - g.x22 * g.x32 * cossv).re;
      AA[1][2] = (-g.x21 * g.x21 * rsinp - g.x22 * g.x22 * sinsvr).re;
      AA[2][0] = (g.x31 * g.x41 * cosp + g.x32 * g.x42 * cossv).re;
      AA[2][1] = (-g.x31 * g.x31 * sinpr - g.x32 * g.x32 * rsinsv).re;
      AA[3][0] = (g.x41 * g.x41 * rsinp + g.x42 * g.x42 * sinsvr).re;
      AA[1][3] = -AA[0][2]; AA[2][2] = AA[1][1]; AA[2][3] = -AA[0][1]; AA[3][1] = -AA[2][0]; AA[3][2] = -AA[1][0];
      AA[3][3] = AA[0][0];
      double aa0[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double cc = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) cc = cc + AA[i][j] * vv[j];
        aa0[i] = cc;
      }
      double t1 = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) if (fabs(aa0[i]) > t1) t1 = fabs(aa0[i]);
      if (t1 < 1.e-40) t1 = 1.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) vv[i] = aa0[i] / t1;
      exa = exa + cpex + log(t1);
      // eigenfunctions at interface m+1 (:1704-1740)
      const double cd1 = SCR(m + 1, 0), cd2 = SCR(m + 1, 2), cd3 = SCR(m + 1, 4), cd4 = -cd3, cd5 = SCR(m + 1, 6),
                   cd6 = SCR(m + 1, 8);
      const double tz1 = -vv[3], tz2 = -vv[2], tz3 = vv[1], tz4 = vv[0];
      const double uu1 = tz2 * cd6 - tz3 * cd5 + tz4 * cd4;
      const double uu2 = -tz1 * cd6 + tz3 * cd3 - tz4 * cd2;
      const double uu3 = tz1 * cd5 - tz2 * cd3 + tz4 * cd1;
      const double uu4 = -tz1 * cd4 + tz2 * cd2 - tz3 * cd1;
      const double ext = exa + SCR(m + 1, 10) - exe1;
      if (ext > -80.0 && ext < 80.0) {
        const double fact = exp(ext);
        dn[0] = uu1 * fact / f1213; dn[1] = uu2 * fact / f1213; dn[2] = uu3 * fact / f1213; dn[3] = uu4 * fact / f1213;
      } else {
        dn[0] = 0.0; dn[1] = 0.0; dn[2] = 0.0; dn[3] = 0.0;
      }
    }
    // energy integrals of layer m
    Z E[4][4], G[4][4];
    emat_dev(g, E, G);
    Ints I;
    layer_integrals(E, G, g.rp, g.rsv, up, dn, dm, half, I);
    // getmat (:3993-4068)
    const double ah = sqrt(TA / rho), av = sqrt(TC / rho), bv = sqrt(TL / rho);
    const double eta = TF / (TA - 2. * TL);
    const double a12 = -wvno, a14 = 1.0 / TL, a21 = wvno * TF / TC, a23 = 1.0 / TC;
    const double URUR = I.i11, UZUZ = I.i22;
    const double DURDUR = a12 * a12 * I.i22 + 2. * a12 * a14 * I.i24 + a14 * a14 * I.i44;
    const double DUZDUZ = a21 * a21 * I.i11 + 2. * a21 * a23 * I.i13 + a23 * a23 * I.i33;
    const double URDUZ = a21 * I.i11 + a23 * I.i13;
    const double UZDUR = a12 * I.i22 + a14 * I.i24;
    sumi0 = sumi0 + rho * (URUR + UZUZ);
    sumi1 = sumi1 + TL * UZUZ + TA * URUR;
    sumi2 = sumi2 + TL * UZDUR - TF * URDUZ;
    const double facah = rho * ah * (URUR - 2. * eta * URDUZ / wvno);
    const double facav = rho * av * DUZDUZ / wvno2;
    const double facbv = rho * bv * (UZUZ + 2. * UZDUR / wvno + DURDUR / wvno2 + 4. * eta * URDUZ / wvno);
    const double facn = -TF * URDUZ / (wvno * eta);
    // reuse the scratch of this layer (its compound vector has been consumed)
    SCR(m, 1) = facah; SCR(m, 3) = facav; SCR(m, 5) = facbv; SCR(m, 7) = facn;
#pragma unroll
    for (int i = 0; i < 4; ++i) up[i] = dn[i];
  }
  const double ugr = (wvno * sumi1 + sumi2) / (omega * sumi0);
  const double nrm = ugr * sumi0;
  // ---- gammap (:3708-3772): causal-Q dispersion correction of c (dogam = .true., Qp=150, Qs=50) ----
  {
    const double pi = 3.141592653589493;   // sic
    const double qa = (double)(1.0f / 150.0f), qb = (double)(1.0f / 50.0f);
    double dc = 0.0;
    for (int m = 0; m < mmax; ++m) {
      const double TA = MOD(A.zta, m), TC = MOD(A.ztc, m), TL = MOD(A.ztl, m), TN = MOD(A.ztn, m), rho = MOD(A.zrho, m);
      const double ah = sqrt(TA / rho), av = sqrt(TC / rho), bh = sqrt(TN / rho), bv = sqrt(TL / rho);
      const double dcdah = SCR(m, 1) / nrm, dcdav = SCR(m, 3) / nrm, dcdbv = SCR(m, 5) / nrm, dcdbh = 0.0 / nrm;
      double x = dcdbh * bh * qb + dcdbv * bv * qb;
      double omgref = 2.0 * pi * 1.0;
      dc = dc + log(omega / omgref) * x / pi;
      x = dcdav * av * qa + dcdah * ah * qa;
      omgref = 2.0 * pi * 1.0;
      dc = dc + log(omega / omgref) * x / pi;
    }
    cph = omega / wvno;
    cph = cph + dc;
    wvno = omega / cph;
    cph = omega / wvno;
  }
  // ---- sprayl (:1232-1296) + fold into Lsen_Gsc (depthkernelTI.f90:96-106) ----
  const double ar = 6370.0;
  const double qq = cph / (2. * ar * omega);
  const double tm = sqrt(1. + qq * qq);
  const double tm3 = tm * (tm * tm);
  int k = 0;
  for (int j = 0; j < A.nz - 1; ++j) {
    const float thk = A.depz[j + 1] - A.depz[j];
    const int ns = nsub_of(thk, A.minthk);
    float L = 0.0f;
    for (int s = 0; s < ns; ++s, ++k) {
      const double vtp = (double)MOD(A.vtp, k);
      const double d_ah = (SCR(k, 1) / nrm) * vtp / tm3;
      const double d_bv = (SCR(k, 5) / nrm) * vtp / tm3;
      const double d_n = SCR(k, 7) / nrm;
      const float dcdah = (fabs(d_ah) < 1.0e-36) ? 0.0f : (float)d_ah;      // chksiz :1216
      const float dcdbv = (fabs(d_bv) < 1.0e-36) ? 0.0f : (float)d_bv;
      const float dcdn = (fabs(d_n) < 1.0e-36) ? 0.0f : (float)d_n;
      const float rrho = MOD(A.rrho, k), rvp = MOD(A.rvp, k), rvs = MOD(A.rvs, k);
      const float TAk = MOD(A.TA, k), TLk = MOD(A.TL, k), TFk = MOD(A.TF, k);
      float den = (TAk - 2.0f * TLk);
      den = den * den;
      const float dcR_dA = 0.5f / (rrho * rvp) * dcdah - TFk / den * dcdn;
      const float dcR_dL = 0.5f / (rrho * rvs) * dcdbv + 2.0f * TFk / den * dcdn;
      L = L + dcR_dA * TAk + dcR_dL * TLk;
    }
    A.lsen[(size_t)node + (size_t)per * nn + (size_t)j * nn * A.kmax] = L;
  }
#undef SCR
#undef MOD
}

// TI model of a node for tregn96: TA.. (depthkernelTI.f90:71-78), sphere_tdisp96 (:771-838), bldsph (:1298-1380)
struct TiArgs {
  int nnode, nlayer;
  const float* rthk; const float* rvp; const float* rvs; const float* rrho;   // [layer][node]
  float* TA; float* TL; float* TF;                                            // raw moduli (float)
  double* zd; double* zta; double* ztc; double* ztf; double* ztl; double* ztn; double* zrho; float* vtp;
};

__global__ void k_timodel(TiArgs A) {
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= A.nnode) return;
  const size_t nn = (size_t)A.nnode;
  const int mmax = A.nlayer;
  const double ar = (double)6371.f;
  double r0 = ar + (double)0.0f;
  const double arb = 6370.0;
  double r0b = arb;
  for (int i = 0; i < mmax; ++i) {
    const size_t o = (size_t)i * nn + node;
    const float rho = A.rrho[o], vp = A.rvp[o], vs = A.rvs[o];
    const float TA = rho * (vp * vp), TL = rho * (vs * vs);
    const float TF = 1.0f * (TA - 2 * TL);
    A.TA[o] = TA; A.TL[o] = TL; A.TF[o] = TF;
    float d = (i == mmax - 1) ? 1.0f : A.rthk[o];
    const double r1 = r0 - (double)d;
    const double z0 = ar * log(ar / r0), z1 = ar * log(ar / r1);
    float dflat = (float)(z1 - z0);
    const double tmp = (ar + ar) / (r0 + r1);
    const float frho = (float)((double)rho * pow(tmp, (double)(-2.275f)));
    const double pw = pow(tmp, (double)(-0.2750f));
    const float fTA = (float)((double)TA * pw);
    const float fTL = (float)((double)TL * pw);
    r0 = r1;
    if (i == mmax - 1) dflat = 0.0f;
    // bldsph uses zd (flattened thickness; 1.0 for the half-space)
    const double zdb = (i == mmax - 1) ? 1.0 : (double)dflat;
    const double r1b = r0b * exp(-zdb / arb);
    const double tmpb = (arb + arb) / (r0b + r1b);
    A.vtp[o] = (float)tmpb;
    r0b = r1b;
    A.zd[o] = zdb; A.zta[o] = (double)fTA; A.ztc[o] = (double)fTA; A.ztf[o] = (double)TF; A.ztl[o] = (double)fTL;
    A.ztn[o] = (double)fTL; A.zrho[o] = (double)frho;
  }
}

// ----------------------------------------------------------------------------
// host drivers
// stream-ordered allocations (recycled by the default pool between calls)
static thread_local cudaStream_t g_th_stream = nullptr;
template <class T>
struct TBuf {
  T* p = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t alloc(size_t n) { st = g_th_stream; return cudaMallocAsync((void**)&p, std::max<size_t>(n, 1) * sizeof(T), st); }
  ~TBuf() { if (p) cudaFreeAsync(p, st); }
};

#define TCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return 100 + (int)e_; } while (0)

static int count_layers(int nz, const float* depz, float minthk) {
  int k = 0;
  for (int i = 1; i <= nz - 1; ++i) {
    const float thk = depz[i] - depz[i - 1];
    const float m = thk / minthk;
    k += (int)((thk + 1.0e-4f) / m) + 1;
  }
  return k + 1;
}

// shared: profiles -> roots.  nvar = 1 (pvRc only) or 1+6*nz (finite-difference kernels)
static int run_disp(cudaStream_t st, int nx, int ny, int nz, int nvar, const float* d_vel, const float* d_depz,
                    float minthk, int nlayer, int kmax, const double* d_t, double* d_cg, float* rthk, float* rvp,
                    float* rvs, float* rrho, int* noroot_host, cudaEvent_t ev_start) {
  const size_t nnode = (size_t)nx * ny, ntask = nnode * nvar;
  TBuf<float> d, a, b, rho;
  TBuf<int> flag;
  TCK(d.alloc(ntask * nlayer)); TCK(a.alloc(ntask * nlayer)); TCK(b.alloc(ntask * nlayer)); TCK(rho.alloc(ntask * nlayer));
  TCK(flag.alloc(1));
  TCK(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
  TCK(cudaEventRecord(ev_start, st));     // timed region = kernels only (allocations are outside)
  ProfArgs P;
  P.nx = nx; P.ny = ny; P.nz = nz; P.nvar = nvar; P.vel = d_vel; P.depz = d_depz; P.minthk = minthk; P.nlayer = nlayer;
  P.d = d.p; P.a = a.p; P.b = b.p; P.rho = rho.p; P.rthk = rthk; P.rvp = rvp; P.rvs = rvs; P.rrho = rrho;
  k_profiles<<<(unsigned)((ntask + 127) / 128), 128, 0, st>>>(P);
  TCK(cudaGetLastError());
  DispArgs D;
  D.ntask = (int)ntask; D.nlayer = nlayer; D.kmax = kmax; D.t = d_t; D.d = d.p; D.a = a.p; D.b = b.p; D.rho = rho.p;
  D.cg = d_cg; D.noroot = flag.p;
  launch_disp(D, (unsigned)((ntask + 127) / 128), st);
  TCK(cudaGetLastError());
  TCK(cudaMemcpyAsync(noroot_host, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TCK(cudaStreamSynchronize(st));
  return 0;
}

int th_depthkernel(cudaStream_t st, int nx, int ny, int nz, const float* vel, double* pvRc, double* sen_vs,
                   double* sen_vp, double* sen_rho, int kmaxRc, const double* tRc, const float* depz, float minthk,
                   float* ms, long long* nlaunch) {
  g_th_stream = st;
  if (nz < 2 || nz > NLMAX || kmaxRc > NPMAX || kmaxRc < 1) return 5;
  const int nlayer = count_layers(nz, depz, minthk);
  if (nlayer > NLMAX) return 5;
  const size_t nnode = (size_t)nx * ny;
  const int nvar = 1 + 6 * nz;
  TBuf<float> d_vel, d_depz;
  TBuf<double> d_t, d_cg, d_pv, d_s[3];
  TCK(d_vel.alloc(nnode * nz)); TCK(d_depz.alloc(nz)); TCK(d_t.alloc(kmaxRc));
  TCK(d_cg.alloc(nnode * nvar * kmaxRc)); TCK(d_pv.alloc(nnode * kmaxRc));
  for (int i = 0; i < 3; ++i) TCK(d_s[i].alloc(nnode * kmaxRc * nz));
  cudaEvent_t e0, e1;
  TCK(cudaEventCreate(&e0)); TCK(cudaEventCreate(&e1));
  TCK(cudaMemcpyAsync(d_vel.p, vel, nnode * nz * 4, cudaMemcpyHostToDevice, st));
  TCK(cudaMemcpyAsync(d_depz.p, depz, nz * 4, cudaMemcpyHostToDevice, st));
  TCK(cudaMemcpyAsync(d_t.p, tRc, kmaxRc * 8, cudaMemcpyHostToDevice, st));
  int noroot = 0;
  int rc = run_disp(st, nx, ny, nz, nvar, d_vel.p, d_depz.p, minthk, nlayer, kmaxRc, d_t.p, d_cg.p, nullptr, nullptr,
                    nullptr, nullptr, &noroot, e0);
  if (rc) return rc;
  k_fdkernels<<<(unsigned)((nnode * kmaxRc + 127) / 128), 128, 0, st>>>(nx, ny, nz, kmaxRc, nvar, d_vel.p, d_cg.p, d_pv.p,
                                                                        d_s[0].p, d_s[1].p, d_s[2].p);
  TCK(cudaGetLastError());
  TCK(cudaEventRecord(e1, st));
  TCK(cudaMemcpyAsync(pvRc, d_pv.p, nnode * kmaxRc * 8, cudaMemcpyDeviceToHost, st));
  TCK(cudaMemcpyAsync(sen_vs, d_s[0].p, nnode * kmaxRc * nz * 8, cudaMemcpyDeviceToHost, st));
  TCK(cudaMemcpyAsync(sen_vp, d_s[1].p, nnode * kmaxRc * nz * 8, cudaMemcpyDeviceToHost, st));
  TCK(cudaMemcpyAsync(sen_rho, d_s[2].p, nnode * kmaxRc * nz * 8, cudaMemcpyDeviceToHost, st));
  TCK(cudaStreamSynchronize(st));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  if (ms) *ms = t;
  if (nlaunch) *nlaunch += 3;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (noroot)
    fprintf(stderr, " WARNING:improper initial value in disper - no zero found (depthkernel; zero-filled like the reference)\n");
  return 0;
}

int th_depthkernel_ti(cudaStream_t st, int nx, int ny, int nz, const float* vel, double* pvRc, int kmaxRc,
                      const double* tRc, const float* depz, float minthk, float* Lsen_Gsc, float* ms,
                      long long* nlaunch) {
  g_th_stream = st;
  if (nz < 2 || nz > NLMAX || kmaxRc > NPMAX || kmaxRc < 1) return 5;
  const int nlayer = count_layers(nz, depz, minthk);
  if (nlayer > NLMAX) return 5;
  const size_t nnode = (size_t)nx * ny, nl = (size_t)nlayer;
  TBuf<float> d_vel, d_depz, rthk, rvp, rvs, rrho, TA, TL, TF, vtp, d_lsen;
  TBuf<double> d_t, d_cg, zd, zta, ztc, ztf, ztl, ztn, zrho, scratch;
  TCK(d_vel.alloc(nnode * nz)); TCK(d_depz.alloc(nz)); TCK(d_t.alloc(kmaxRc)); TCK(d_cg.alloc(nnode * kmaxRc));
  TCK(rthk.alloc(nnode * nl)); TCK(rvp.alloc(nnode * nl)); TCK(rvs.alloc(nnode * nl)); TCK(rrho.alloc(nnode * nl));
  TCK(TA.alloc(nnode * nl)); TCK(TL.alloc(nnode * nl)); TCK(TF.alloc(nnode * nl)); TCK(vtp.alloc(nnode * nl));
  TCK(zd.alloc(nnode * nl)); TCK(zta.alloc(nnode * nl)); TCK(ztc.alloc(nnode * nl)); TCK(ztf.alloc(nnode * nl));
  TCK(ztl.alloc(nnode * nl)); TCK(ztn.alloc(nnode * nl)); TCK(zrho.alloc(nnode * nl));
  TCK(scratch.alloc(nnode * kmaxRc * nl * NSLOT));
  TCK(d_lsen.alloc(nnode * kmaxRc * (nz - 1)));
  cudaEvent_t e0, e1;
  TCK(cudaEventCreate(&e0)); TCK(cudaEventCreate(&e1));
  TCK(cudaMemcpyAsync(d_vel.p, vel, nnode * nz * 4, cudaMemcpyHostToDevice, st));
  TCK(cudaMemcpyAsync(d_depz.p, depz, nz * 4, cudaMemcpyHostToDevice, st));
  TCK(cudaMemcpyAsync(d_t.p, tRc, kmaxRc * 8, cudaMemcpyHostToDevice, st));
  int noroot = 0;
  int rc = run_disp(st, nx, ny, nz, 1, d_vel.p, d_depz.p, minthk, nlayer, kmaxRc, d_t.p, d_cg.p, rthk.p, rvp.p, rvs.p,
                    rrho.p, &noroot, e0);
  if (rc) return rc;
  TiArgs T;
  T.nnode = (int)nnode; T.nlayer = nlayer; T.rthk = rthk.p; T.rvp = rvp.p; T.rvs = rvs.p; T.rrho = rrho.p;
  T.TA = TA.p; T.TL = TL.p; T.TF = TF.p; T.zd = zd.p; T.zta = zta.p; T.ztc = ztc.p; T.ztf = ztf.p; T.ztl = ztl.p;
  T.ztn = ztn.p; T.zrho = zrho.p; T.vtp = vtp.p;
  k_timodel<<<(unsigned)((nnode + 127) / 128), 128, 0, st>>>(T);
  TCK(cudaGetLastError());
  EigenArgs E;
  E.nnode = (int)nnode; E.kmax = kmaxRc; E.nlayer = nlayer; E.nz = nz; E.t = d_t.p; E.pv = d_cg.p;
  E.zd = zd.p; E.zta = zta.p; E.ztc = ztc.p; E.ztf = ztf.p; E.ztl = ztl.p; E.ztn = ztn.p; E.zrho = zrho.p; E.vtp = vtp.p;
  E.rvp = rvp.p; E.rvs = rvs.p; E.rrho = rrho.p; E.TA = TA.p; E.TL = TL.p; E.TF = TF.p; E.depz = d_depz.p;
  E.minthk = minthk; E.scratch = scratch.p; E.lsen = d_lsen.p;
  const size_t ntask = nnode * kmaxRc;
  k_eigen<<<(unsigned)((ntask + 63) / 64), 64, 0, st>>>(E);
  TCK(cudaGetLastError());
  TCK(cudaEventRecord(e1, st));
  TCK(cudaMemcpyAsync(pvRc, d_cg.p, nnode * kmaxRc * 8, cudaMemcpyDeviceToHost, st));
  TCK(cudaMemcpyAsync(Lsen_Gsc, d_lsen.p, nnode * kmaxRc * (nz - 1) * 4, cudaMemcpyDeviceToHost, st));
  TCK(cudaStreamSynchronize(st));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  if (ms) *ms = t;
  if (nlaunch) *nlaunch += 4;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (noroot)
    fprintf(stderr, " WARNING:improper initial value in disper - no zero found (depthkernelTI; zero-filled like the reference)\n");
  return 0;
}

// nprof independent layered profiles given explicitly (test seam for surfdisp96)
__global__ void k_flatten_given(int ntask, int nlayer, const float* thk, const float* vp, const float* vs,
                                const float* rho, float* d, float* a, float* b, float* r) {
  const int task = blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= ntask) return;
  const double ar = 6370.0;
  double dr = 0.0, r0 = ar;
  for (int i = 0; i < nlayer; ++i) {
    const size_t in = (size_t)task * nlayer + i;
    const float th = (i == nlayer - 1) ? 1.0f : thk[in];
    dr = dr + (double)th;
    const double r1 = ar - dr;
    const double z0 = ar * log(ar / r0), z1 = ar * log(ar / r1);
    const double tmp = (ar + ar) / (r0 + r1);
    const float btp = (float)tmp;
    const size_t o = (size_t)i * ntask + task;
    d[o] = (i == nlayer - 1) ? 0.0f : (float)(z1 - z0);
    a[o] = (float)((double)vp[in] * tmp);
    b[o] = (float)((double)vs[in] * tmp);
    r[o] = rho[in] * (float)pow((double)btp, (double)-2.275f);
    r0 = r1;
  }
}

int th_surfdisp96(cudaStream_t st, int nprof, int nlayer, const float* thk, const float* vp, const float* vs,
                  const float* rho, int kmax, const double* t, double* cg) {
  g_th_stream = st;
  if (nlayer < 2 || nlayer > NLMAX || kmax > NPMAX || kmax < 1 || nprof < 1) return 5;
  const size_t n = (size_t)nprof * nlayer;
  TBuf<float> in[4], fl[4];
  TBuf<double> d_t, d_cg;
  TBuf<int> flag;
  const float* src[4] = {thk, vp, vs, rho};
  for (int i = 0; i < 4; ++i) {
    TCK(in[i].alloc(n)); TCK(fl[i].alloc(n));
    TCK(cudaMemcpyAsync(in[i].p, src[i], n * 4, cudaMemcpyHostToDevice, st));
  }
  TCK(d_t.alloc(kmax)); TCK(d_cg.alloc((size_t)nprof * kmax)); TCK(flag.alloc(1));
  TCK(cudaMemsetAsync(flag.p, 0, 4, st));
  TCK(cudaMemcpyAsync(d_t.p, t, kmax * 8, cudaMemcpyHostToDevice, st));
  k_flatten_given<<<(nprof + 127) / 128, 128, 0, st>>>(nprof, nlayer, in[0].p, in[1].p, in[2].p, in[3].p, fl[0].p, fl[1].p,
                                                       fl[2].p, fl[3].p);
  TCK(cudaGetLastError());
  DispArgs D;
  D.ntask = nprof; D.nlayer = nlayer; D.kmax = kmax; D.t = d_t.p; D.d = fl[0].p; D.a = fl[1].p; D.b = fl[2].p;
  D.rho = fl[3].p; D.cg = d_cg.p; D.noroot = flag.p;
  launch_disp(D, (unsigned)((nprof + 127) / 128), st);
  TCK(cudaGetLastError());
  // cg comes back as (kmax, nprof) column-major == [prof][k]; device layout is [k][prof]
  std::vector<double> tmp((size_t)nprof * kmax);
  TCK(cudaMemcpyAsync(tmp.data(), d_cg.p, tmp.size() * 8, cudaMemcpyDeviceToHost, st));
  TCK(cudaStreamSynchronize(st));
  for (int p = 0; p < nprof; ++p)
    for (int k = 0; k < kmax; ++k) cg[(size_t)p * kmax + k] = tmp[(size_t)k * nprof + p];
  return 0;
}

}  // namespace dz
