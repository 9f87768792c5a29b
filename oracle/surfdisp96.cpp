// ORACLE (test infrastructure, NOT product code).
// CPU restatement of src/src_forward/surfdisp96.f (byte-identical copy in
// src/src_inv_iso_joint/), Rayleigh-wave / phase-velocity branch only
// (iwave=2, igr=0), which is the only branch the reference drivers use
// (depthkernelTI.f90:66, CalSurfG.f90:58).  Implicit F77 typing is honoured:
// betmx, betmn, cc1 and gtsolh's locals are REAL*4; everything under
// "implicit double precision (a-h,o-z)" is double.  The SAVEd del1st/dhalf
// (surfdisp96.f:409,509) become members of a per-call context (serial
// semantics).
#include "oracle.h"
#include <cmath>
#include <algorithm>

namespace orc {

namespace {
const int NL = 200;

struct Ctx {
  float d[NL], a[NL], b[NL], rho[NL], rtp[NL], dtp[NL], btp[NL];
  int mmax, llw;
  double del1st;
  float dhalf;
  long neval;
};

inline double dsign1(double x) { return std::signbit(x) ? -1.0 : 1.0; }  // dsign(1.0d0,x)

// surfdisp96.f:480-547
void sphere(Ctx& m, int ifunc, int iflag) {
  double ar = 6370.0, dr = 0.0, r0 = ar, r1, z0, z1, tmp;
  const int mmax = m.mmax;
  m.d[mmax - 1] = 1.0f;
  if (iflag == 0) {
    for (int i = 0; i < mmax; ++i) {
      m.dtp[i] = m.d[i];
      m.rtp[i] = m.rho[i];
    }
    for (int i = 0; i < mmax; ++i) {
      dr = dr + (double)m.d[i];
      r1 = ar - dr;
      z0 = ar * std::log(ar / r0);
      z1 = ar * std::log(ar / r1);
      m.d[i] = (float)(z1 - z0);
      tmp = (ar + ar) / (r0 + r1);
      m.a[i] = (float)((double)m.a[i] * tmp);
      m.b[i] = (float)((double)m.b[i] * tmp);
      m.btp[i] = (float)tmp;
      r0 = r1;
    }
    m.dhalf = m.d[mmax - 1];
  } else {
    m.d[mmax - 1] = m.dhalf;
    for (int i = 0; i < mmax; ++i) {
      if (ifunc == 1) {
        float p = m.btp[i];
        // btp**(-5): integer power through __powisf2 then reciprocal
        float y = p; float x2 = p * p; float x4 = x2 * x2; y = y * x4;
        m.rho[i] = m.rtp[i] * (1.0f / y);
      } else if (ifunc == 2) {
        m.rho[i] = m.rtp[i] * (float)std::pow((double)m.btp[i], (double)-2.275f);  // REAL*4 powf -> rounded double pow
      }
    }
  }
  m.d[mmax - 1] = 0.0f;
}

// surfdisp96.f:361-382 (all REAL*4)
void gtsolh(float a, float b, float* cout) {
  float c = 0.95f * b;
  for (int i = 1; i <= 5; ++i) {
    float gamma = b / a;
    float kappa = c / b;
    float k2 = kappa * kappa;
    float gk2 = (gamma * kappa) * (gamma * kappa);
    float fac1 = std::sqrt(1.0f - gk2);
    float fac2 = std::sqrt(1.0f - k2);
    float fr = (2.0f - k2) * (2.0f - k2) - 4.0f * fac1 * fac2;
    float frp = -4.0f * (2.0f - k2) * kappa + 4.0f * fac2 * gamma * gamma * kappa / fac1 +
                4.0f * fac1 * kappa / fac2;
    frp = frp / b;
    c = c - fr / frp;
  }
  *cout = c;
}

struct Ovr { double a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz; };

// surfdisp96.f:868-985
void var(double p, double q, double ra, double rb, double wvno, double xka, double xkb,
         double dpth, double& w, double& cosp, double& exa, Ovr& o) {
  exa = 0.0;
  o.a0 = 1.0;
  double pex = 0.0, sex = 0.0;
  double sinp, x = 0.0, fac, sinq, y = 0.0, z = 0.0, cosq = 0.0;
  w = 0.0; cosp = 0.0;
  if (wvno < xka) {
    sinp = std::sin(p);
    w = sinp / ra;
    x = -ra * sinp;
    cosp = std::cos(p);
  } else if (wvno == xka) {
    cosp = 1.0;
    w = dpth;
    x = 0.0;
  } else if (wvno > xka) {
    pex = p;
    fac = 0.0;
    if (p < 16) fac = std::exp(-2.0 * p);
    cosp = (1.0 + fac) * 0.5;
    sinp = (1.0 - fac) * 0.5;
    w = sinp / ra;
    x = ra * sinp;
  }
  if (wvno < xkb) {
    sinq = std::sin(q);
    y = sinq / rb;
    z = -rb * sinq;
    cosq = std::cos(q);
  } else if (wvno == xkb) {
    cosq = 1.0;
    y = dpth;
    z = 0.0;
  } else if (wvno > xkb) {
    sex = q;
    fac = 0.0;
    if (q < 16) fac = std::exp(-2.0 * q);
    cosq = (1.0 + fac) * 0.5;
    sinq = (1.0 - fac) * 0.5;
    y = sinq / rb;
    z = rb * sinq;
  }
  exa = pex + sex;
  o.a0 = 0.0;
  if (exa < 60.0) o.a0 = std::exp(-exa);
  o.cpcq = cosp * cosq;
  o.cpy = cosp * y;
  o.cpz = cosp * z;
  o.cqw = cosq * w;
  o.cqx = cosq * x;
  o.xy = x * y;
  o.xz = x * z;
  o.wy = w * y;
  o.wz = w * z;
  // (the rescaled cosq,y,z at :978-983 are locals that are never used again)
}

// surfdisp96.f:989-1014
void normc(double* ee, double& ex) {
  ex = 0.0;
  double t1 = 0.0;
  for (int i = 0; i < 5; ++i)
    if (std::fabs(ee[i]) > t1) t1 = std::fabs(ee[i]);
  if (t1 < 1.e-40) t1 = 1.0;
  for (int i = 0; i < 5; ++i) {
    double t2 = ee[i];
    t2 = t2 / t1;
    ee[i] = t2;
  }
  ex = std::log(t1);
}

// surfdisp96.f:1018-1062; ca(i,j) stored as ca[i-1][j-1]
void dnka(double ca[5][5], double wvno2, double gam, double gammk, double rho, const Ovr& o) {
  const double one = 1.0, two = 2.0;
  double gamm1 = gam - one;
  double twgm1 = gam + gamm1;
  double gmgmk = gam * gammk;
  double gmgm1 = gam * gamm1;
  double gm1sq = gamm1 * gamm1;
  double rho2 = rho * rho;
  double a0pq = o.a0 - o.cpcq;
  ca[0][0] = o.cpcq - two * gmgm1 * a0pq - gmgmk * o.xz - wvno2 * gm1sq * o.wy;
  ca[0][1] = (wvno2 * o.cpy - o.cqx) / rho;
  ca[0][2] = -(twgm1 * a0pq + gammk * o.xz + wvno2 * gamm1 * o.wy) / rho;
  ca[0][3] = (o.cpz - wvno2 * o.cqw) / rho;
  ca[0][4] = -(two * wvno2 * a0pq + o.xz + wvno2 * wvno2 * o.wy) / rho2;
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
  ca[4][0] = -(two * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * o.xz + gm1sq * gm1sq * o.wy) * rho2;
  ca[4][1] = ca[3][0];
  ca[4][2] = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * o.xz + gamm1 * gm1sq * o.wy) * rho;
  ca[4][3] = ca[1][0];
  ca[4][4] = ca[0][0];
  double t = -two * wvno2;
  ca[2][0] = t * ca[4][2];
  ca[2][1] = t * ca[3][2];
  ca[2][2] = o.a0 + two * (o.cpcq - ca[0][0]);
  ca[2][3] = t * ca[1][2];
  ca[2][4] = t * ca[0][2];
}

// surfdisp96.f:767-865
double dltar4(Ctx& m, double wvno, double omga) {
  ++m.neval;
  double e[5], ee[5], ca[5][5];
  Ovr o;
  const int mmax = m.mmax;
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  double wvno2 = wvno * wvno;
  double xka = omega / (double)m.a[mmax - 1];
  double xkb = omega / (double)m.b[mmax - 1];
  double wvnop = wvno + xka;
  double wvnom = std::fabs(wvno - xka);
  double ra = std::sqrt(wvnop * wvnom);
  wvnop = wvno + xkb;
  wvnom = std::fabs(wvno - xkb);
  double rb = std::sqrt(wvnop * wvnom);
  double t = (double)m.b[mmax - 1] / omega;
  double gammk = 2.0 * t * t;
  double gam = gammk * wvno2;
  double gamm1 = gam - 1.0;
  double rho1 = (double)m.rho[mmax - 1];
  e[0] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
  e[1] = -rho1 * ra;
  e[2] = rho1 * (gamm1 - gammk * ra * rb);
  e[3] = rho1 * rb;
  e[4] = wvno2 - ra * rb;
  double w, cosp, exa;
  for (int l = mmax - 1; l >= m.llw; --l) {
    const int i0 = l - 1;
    xka = omega / (double)m.a[i0];
    xkb = omega / (double)m.b[i0];
    t = (double)m.b[i0] / omega;
    gammk = 2.0 * t * t;
    gam = gammk * wvno2;
    wvnop = wvno + xka;
    wvnom = std::fabs(wvno - xka);
    ra = std::sqrt(wvnop * wvnom);
    wvnop = wvno + xkb;
    wvnom = std::fabs(wvno - xkb);
    rb = std::sqrt(wvnop * wvnom);
    double dpth = (double)m.d[i0];
    rho1 = (double)m.rho[i0];
    double p = ra * dpth;
    double q = rb * dpth;
    var(p, q, ra, rb, wvno, xka, xkb, dpth, w, cosp, exa, o);
    dnka(ca, wvno2, gam, gammk, rho1, o);
    for (int i = 0; i < 5; ++i) {
      double cr = 0.0;
      for (int j = 0; j < 5; ++j) cr = cr + e[j] * ca[j][i];
      ee[i] = cr;
    }
    normc(ee, exa);
    for (int i = 0; i < 5; ++i) e[i] = ee[i];
  }
  if (m.llw != 1) {
    xka = omega / (double)m.a[0];
    wvnop = wvno + xka;
    wvnom = std::fabs(wvno - xka);
    ra = std::sqrt(wvnop * wvnom);
    double dpth = (double)m.d[0];
    rho1 = (double)m.rho[0];
    double p = ra * dpth;
    double znul = 1.0e-05;
    var(p, znul, ra, znul, wvno, xka, znul, dpth, w, cosp, exa, o);
    double w0 = -rho1 * w;
    return cosp * e[0] + w0 * e[1];
  }
  return e[0];
}

// surfdisp96.f:670-680
void half(Ctx& m, double c1, double c2, double& c3, double& del3, double omega) {
  c3 = 0.5 * (c1 + c2);
  double wvno = omega / c3;
  del3 = dltar4(m, wvno, omega);
}

// surfdisp96.f:551-668
void nevill(Ctx& m, double t, double c1, double c2, double del1, double del2, double& cc) {
  const double twopi = 2.0 * 3.141592653589793;
  double x[21], y[21];
  double c3, del3;
  double omega = twopi / t;
  half(m, c1, c2, c3, del3, omega);
  int nev = 1;
  int nctrl = 1;
  int mm = 1;
  for (;;) {
    nctrl = nctrl + 1;
    if (nctrl >= 100) break;
    if (c3 < std::min(c1, c2) || c3 > std::max(c1, c2)) {
      nev = 0;
      half(m, c1, c2, c3, del3, omega);
    }
    double s13 = del1 - del3;
    double s32 = del3 - del2;
    if (dsign1(del3) * dsign1(del1) < 0.0) {
      c2 = c3;
      del2 = del3;
    } else {
      c1 = c3;
      del1 = del3;
    }
    if (std::fabs(c1 - c2) <= 1.e-6 * c1) break;
    if (dsign1(s13) != dsign1(s32)) nev = 0;
    double ss1 = std::fabs(del1);
    double s1 = (double)0.01f * ss1;
    double ss2 = std::fabs(del2);
    double s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || nev == 0) {
      half(m, c1, c2, c3, del3, omega);
      nev = 1;
      mm = 1;
    } else {
      if (nev == 2) {
        x[mm + 1] = c3;
        y[mm + 1] = del3;
      } else {
        x[1] = c1;
        y[1] = del1;
        x[2] = c2;
        y[2] = del2;
        mm = 1;
      }
      bool bad = false;
      for (int kk = 1; kk <= mm; ++kk) {
        int j = mm - kk + 1;
        double denom = y[mm + 1] - y[j];
        if (std::fabs(denom) < 1.0e-10 * std::fabs(y[mm + 1])) { bad = true; break; }
        x[j] = (-y[j] * x[j + 1] + y[mm + 1] * x[j]) / denom;
      }
      if (!bad) {
        c3 = x[1];
        double wvno = omega / c3;
        del3 = dltar4(m, wvno, omega);
        nev = 2;
        mm = mm + 1;
        if (mm > 10) mm = 10;
      } else {
        half(m, c1, c2, c3, del3, omega);
        nev = 1;
        mm = 1;
      }
    }
  }
  cc = c3;
}

// surfdisp96.f:384-476
void getsol(Ctx& m, double t1, double& c1, double clow, double dc, double cm, float betmx,
            int& iret, int ifirst) {
  const double twopi = 2.0 * 3.141592653589793;
  double omega = twopi / t1;
  double wvno = omega / c1;
  double del1 = dltar4(m, wvno, omega);
  if (ifirst == 1) m.del1st = del1;
  double plmn = dsign1(m.del1st) * dsign1(del1);
  int idir = +1;
  if (ifirst == 1) idir = +1;
  else if (ifirst != 1 && plmn >= 0.0) idir = +1;
  else if (ifirst != 1 && plmn < 0.0) idir = -1;
  double c2, del2, cn;
  for (;;) {
    if (idir > 0) c2 = c1 + dc;
    else c2 = c1 - dc;
    if (c2 <= clow) {
      idir = +1;
      c1 = clow;
    }
    if (c2 <= clow) continue;
    omega = twopi / t1;
    wvno = omega / c2;
    del2 = dltar4(m, wvno, omega);
    if (dsign1(del1) != dsign1(del2)) {
      nevill(m, t1, c1, c2, del1, del2, cn);
      c1 = cn;
      if (c1 > (double)betmx) { iret = -1; return; }
      iret = 1;
      return;
    }
    c1 = c2;
    del1 = del2;
    if (c1 < cm) { iret = -1; return; }
    if (c1 >= ((double)betmx + dc)) { iret = -1; return; }
  }
}

}  // namespace

// surfdisp96.f:52-354
int surfdisp96(const float* thkm, const float* vpm, const float* vsm, const float* rhom,
               int nlayer, int iflsph, int iwave, int mode, int igr, int kmax,
               const double* t, double* cg, long* neval) {
  if (iwave != 2 || igr != 0) return ERR_BAD_ARG;  // only the branch the reference drivers use
  if (nlayer > NL || nlayer < 2 || kmax > 60) return ERR_LAYERS;
  Ctx m;
  m.neval = 0;
  m.del1st = 0.0;
  m.dhalf = 0.0f;
  const int mmax = nlayer;
  m.mmax = mmax;
  for (int i = 0; i < mmax; ++i) {
    m.b[i] = vsm[i];
    m.a[i] = vpm[i];
    m.d[i] = thkm[i];
    m.rho[i] = rhom[i];
  }
  const float sone0 = 1.500f, ddc0 = 0.005f;
  m.llw = 1;
  if (m.b[0] <= 0.0f) m.llw = 2;
  const double one = 1.0e-2;
  if (iflsph == 1) sphere(m, 0, 0);
  int jmn = 1, jsol = 1;
  float betmx = -1.e20f, betmn = 1.e20f;
  for (int i = 0; i < mmax; ++i) {
    if (m.b[i] > 0.01f && m.b[i] < betmn) {
      betmn = m.b[i];
      jmn = i + 1;
      jsol = 1;
    } else if (m.b[i] <= 0.01f && m.a[i] < betmn) {
      betmn = m.a[i];
      jmn = i + 1;
      jsol = 0;
    }
    if (m.b[i] > betmx) betmx = m.b[i];
  }
  double c[61], cb[61];
  (void)cb;
  const int ifunc = 2;
  if (iflsph == 1) sphere(m, ifunc, 1);
  float ddc = ddc0, sone = sone0;
  if (sone < 0.01f) sone = 2.0f;
  double onea = (double)sone;
  float cc1;
  if (jsol == 0) cc1 = betmn;
  else gtsolh(m.a[jmn - 1], m.b[jmn - 1], &cc1);
  cc1 = .95f * cc1;
  cc1 = .90f * cc1;
  double cc = (double)cc1;
  double dc = (double)ddc;
  dc = std::fabs(dc);
  double c1 = cc;
  double cm = cc;
  double clow = cc;
  for (int i = 1; i <= kmax; ++i) { cb[i] = 0.0; c[i] = 0.0; }
  int ift = 999;
  int status = OK;
  for (int iq = 1; iq <= mode; ++iq) {
    const int is = 1, ie = kmax;
    int k;
    bool fail = false;
    for (k = is; k <= ie; ++k) {
      if (k >= ift) { fail = true; break; }
      double t1 = t[k - 1];
      int ifirst;
      if (k == is && iq == 1) {
        c1 = cc; clow = cc; ifirst = 1;
      } else if (k == is && iq > 1) {
        c1 = c[is] + one * dc; clow = c1; ifirst = 1;
      } else if (k > is && iq > 1) {
        ifirst = 0;
        clow = c[k] + one * dc;
        c1 = c[k - 1];
        if (c1 < clow) c1 = clow;
      } else {
        ifirst = 0;
        c1 = c[k - 1] - onea * dc;
        clow = cm;
      }
      int iret;
      getsol(m, t1, c1, clow, dc, cm, betmx, iret, ifirst);
      if (iret == -1) { fail = true; break; }
      c[k] = c1;
      float cc0 = (float)c[k];
      cg[k - 1] = (double)cc0;
    }
    if (fail) {
      // surfdisp96.f:307-348: "improper initial value in disper - no zero found";
      // remaining periods are zero-filled
      if (iq == 1) status = OK;  // the reference only prints a warning
      ift = k;
      for (int i = k; i <= ie; ++i) cg[i - 1] = 0.0;
    }
  }
  if (neval) *neval = m.neval;
  return status;
}

}  // namespace orc
