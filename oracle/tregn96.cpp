// ORACLE (test infrastructure, NOT product code).
// CPU restatement of src/src_forward/tregn96_subroutine.f (tregn96.f in
// src_inv_iso_joint differs by 3 comment lines): TI-medium Rayleigh-wave
// eigenfunctions, energy integrals and the analytic partials dc/dA_h, dc/dbeta_v,
// dc/deta per layer, with the causal-Q and sphericity corrections the reference
// applies.  Reference line numbers below are for tregn96_subroutine.f.
//
// Scope: solid layers only (iwat=0 everywhere), fundamental mode, hs=hr=0
// (the reference reads them uninitialised; 0 means no layer is
// inserted and lss=lrr=1).  dc/dh (getdcdh :4244) is not restated: the drivers
// discard it.  COMMON-block state becomes a local context, so the routine is
// re-entrant (the reference is not).
//
// Complex arithmetic is spelled out the way gfortran (-fcx-fortran-rules)
// expands it: naive multiplication, Smith division.
#include "oracle.h"
#include <cmath>
#include <complex>

namespace orc {
namespace {

const int NL = 200;
const int NP = 60;

struct Z {
  double re, im;
  Z() : re(0), im(0) {}
  Z(double r) : re(r), im(0) {}
  Z(double r, double i) : re(r), im(i) {}
};
inline Z operator+(Z a, Z b) { return Z(a.re + b.re, a.im + b.im); }
inline Z operator-(Z a, Z b) { return Z(a.re - b.re, a.im - b.im); }
inline Z operator-(Z a) { return Z(-a.re, -a.im); }
inline Z operator*(Z a, Z b) { return Z(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline Z operator*(double r, Z a) { return Z(r * a.re, r * a.im); }
inline Z operator*(Z a, double r) { return Z(a.re * r, a.im * r); }
inline Z operator/(Z a, double r) { return Z(a.re / r, a.im / r); }
inline Z operator/(Z n, Z d) {  // Smith's algorithm as GCC expands complex division
  if (std::fabs(d.re) < std::fabs(d.im)) {
    const double ratio = d.re / d.im, denom = d.re * ratio + d.im;
    return Z((n.re * ratio + n.im) / denom, (n.im * ratio - n.re) / denom);
  }
  const double ratio = d.im / d.re, denom = d.im * ratio + d.re;
  return Z((n.im * ratio + n.re) / denom, (n.im - n.re * ratio) / denom);
}
inline double zabs(Z a) { return std::hypot(a.re, a.im); }   // cdabs
inline Z zsqrt(Z a) { std::complex<double> r = std::sqrt(std::complex<double>(a.re, a.im)); return Z(r.real(), r.imag()); }
inline Z zexp(Z a) { const double e = std::exp(a.re); return Z(e * std::cos(a.im), e * std::sin(a.im)); }
inline Z zconj(Z a) { return Z(a.re, -a.im); }

struct Eig {  // gettiegn outputs
  Z rp, rsv, x11, x21, x31, x41, x12, x22, x32, x42, np, nsv;
};

struct Ctx {
  int mmax;
  // common/timod/ (double copies of the flattened model)
  double zd[NL], zta[NL], ztc[NL], ztf[NL], ztl[NL], ztn[NL], zrho[NL], zqai[NL], zqbi[NL], zfrefp[NL], zfrefs[NL];
  // common/sphereL/
  float vtp[NL], dtp[NL], rtp[NL];
  // common/eigfun/
  double ur[NL], uz[NL], tz[NL], tr[NL], uu0[4];
  double dcdah[NL], dcdav[NL], dcdbh[NL], dcdbv[NL], dcdn[NL], dcdr[NL];
  // common/dunk/, common/hask/
  Z cd[NL][5];
  double exe[NL], exa[NL];
  double vv[NL][4];
  // common/sumi/
  double sumi0, sumi1, sumi2, sumi3, flagr, are, ugr;
  // common/emat/
  Z e[4][4], einv[4][4], ra, rb;
};

// gettiegn :3163-3358 (solid branch)
void gettiegn(const Ctx& c, int m, double omg, double wvn, double omega2, double wvno2, Eig& o) {
  const double TA = c.zta[m], TC = c.ztc[m], TF = c.ztf[m], TL = c.ztl[m], TRho = c.zrho[m];
  const Z a = wvn * TF / (TC);
  const Z b = 1.0 / (TC);
  const Z cc_ = -TRho * omg * omg + wvn * wvn * (TA - TF * TF / (TC));
  const Z d = -wvn;
  const Z e = 1.0 / (TL);
  const Z f = -TRho * omg * omg;
  const Z ddef = wvn * wvn - TRho * omg * omg / (TL);
  const Z aabc = wvn * wvn * TA / TC - TRho * omg * omg / (TC);
  const Z bb = 2.0 * a * d + e * cc_ + f * b;
  const Z cc = ddef * aabc;
  Z srt = zsqrt(bb * bb - 4.0 * cc);
  if (srt.im < 0.0) srt = -srt;
  Z L2[2];
  if (bb.re < 0.0 && srt.re < 0.0) {
    L2[1] = (bb - srt) / 2.0;
    if (zabs(L2[1]) > 0.0) L2[0] = cc / L2[1];
    else L2[0] = (bb + srt) / 2.0;
  } else {
    L2[0] = (bb + srt) / 2.0;
    if (zabs(L2[0]) > 0.0) L2[1] = cc / L2[0];
    else L2[1] = (bb - srt) / 2.0;
  }
  const Z xka2 = Z(wvno2) - L2[0];
  const Z xkb2 = Z(wvno2) - L2[1];
  if (zabs(xkb2) < zabs(xka2)) { const Z t = L2[0]; L2[0] = L2[1]; L2[1] = t; }
  o.rp = zsqrt(L2[0]);
  o.rsv = zsqrt(L2[1]);
  if (o.rp.re < 0.0) o.rp = -o.rp;
  if (o.rsv.re < 0.0) o.rsv = -o.rsv;
  o.x12 = (b * d - a * e);
  o.x22 = b * L2[1] - e * (b * cc_ + a * a);
  o.x32 = L2[1] - (a * d + cc_ * e);
  o.x42 = -a * L2[1] + d * (b * cc_ + a * a);
  o.x11 = -e * L2[0] + b * (d * d + e * f);
  o.x21 = (b * d - a * e);
  o.x31 = d * L2[0] - a * (d * d + e * f);
  o.x41 = -(L2[0] - a * d - b * f);
  if (wvn != 0.0) {
    Z zfac = Z(wvn) / o.x11;
    o.x11 = o.x11 * zfac; o.x21 = o.x21 * zfac; o.x31 = o.x31 * zfac; o.x41 = o.x41 * zfac;
    zfac = Z(wvn) / o.x22;
    o.x12 = o.x12 * zfac; o.x22 = o.x22 * zfac; o.x32 = o.x32 * zfac; o.x42 = o.x42 * zfac;
  }
  o.np = o.x11 * o.x41 - o.x21 * o.x31;
  o.nsv = o.x12 * o.x42 - o.x22 * o.x32;
}

// evalg :2986-3161 (solid branch): E, E^-1 of layer m and the 5 compound minors of E^-1
void evalg(Ctx& c, int m, double wvno, double om, double om2, double wvno2, Z gbr[5]) {
  Eig g;
  gettiegn(c, m, om, wvno, om2, wvno2, g);
  const Z rp = g.rp, rsv = g.rsv, NPz = g.np, NSV = g.nsv;
  c.ra = rp; c.rb = rsv;
  Z G[4][4];
  G[0][0] = g.x41 * rp / (2. * rp * NPz);
  G[1][0] = g.x42 / (2. * rsv * NSV);
  G[2][0] = -g.x41 * rp / (-2. * rp * NPz);
  G[3][0] = g.x42 / (-2. * rsv * NSV);
  G[0][1] = -g.x31 / (2. * rp * NPz);
  G[1][1] = -g.x32 * rsv / (2. * rsv * NSV);
  G[2][1] = -g.x31 / (-2. * rp * NPz);
  G[3][1] = g.x32 * rsv / (-2. * rsv * NSV);
  G[0][2] = -g.x21 * rp / (2. * rp * NPz);
  G[1][2] = -g.x22 / (2. * rsv * NSV);
  G[2][2] = g.x21 * rp / (-2. * rp * NPz);
  G[3][2] = -g.x22 / (-2. * rsv * NSV);
  G[0][3] = g.x11 / (2. * rp * NPz);
  G[1][3] = g.x12 * rsv / (2. * rsv * NSV);
  G[2][3] = g.x11 / (-2. * rp * NPz);
  G[3][3] = -g.x12 * rsv / (-2. * rsv * NSV);
  Z (*E)[4] = c.e;
  E[0][0] = g.x11;        E[1][0] = g.x21 * rp;   E[2][0] = g.x31;        E[3][0] = g.x41 * rp;
  E[0][1] = g.x12 * rsv;  E[1][1] = g.x22;        E[2][1] = g.x32 * rsv;  E[3][1] = g.x42;
  E[0][2] = g.x11;        E[1][2] = -g.x21 * rp;  E[2][2] = g.x31;        E[3][2] = -g.x41 * rp;
  E[0][3] = -g.x12 * rsv; E[1][3] = g.x22;        E[2][3] = -g.x32 * rsv; E[3][3] = g.x42;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) c.einv[i][j] = G[i][j];
  const Z CG1 = G[0][0] * G[1][1] - G[0][1] * G[1][0];
  const Z CG2 = G[0][0] * G[1][2] - G[0][2] * G[1][0];
  const Z CG3 = G[0][0] * G[1][3] - G[0][3] * G[1][0];
  const Z CG5 = G[0][1] * G[1][3] - G[0][3] * G[1][1];
  const Z CG6 = G[0][2] * G[1][3] - G[0][3] * G[1][2];
  gbr[0] = CG1; gbr[1] = CG2; gbr[2] = CG3; gbr[3] = CG5; gbr[4] = CG6;
}

// varsv :3360-3472 (solid branch)
void varsv(Z p, Z q, Z rp, Z rsv, Z& cosp, Z& cosq, Z& rsinp, Z& rsinq, Z& sinpr, Z& sinqr, double& pex,
           double& svex, double dm) {
  const double pr = p.re, pi = p.im, qr = q.re, qi = q.im;
  pex = pr;
  svex = qr;
  const Z epp = Z(std::cos(pi), std::sin(pi)) / 2.0;
  const Z epm = zconj(epp);
  const Z eqp = Z(std::cos(qi), std::sin(qi)) / 2.0;
  const Z eqm = zconj(eqp);
  double pfac, svfac;
  if (pr < 15.) pfac = std::exp(-2. * pr); else pfac = 0.0;
  cosp = (epp + pfac * epm);
  const Z sinp = epp - pfac * epm;
  rsinp = (rp * sinp);
  if (std::fabs(pr) < (double)1.0e-5f && zabs(rp) < (double)1.0e-5f) sinpr = dm;
  else sinpr = (sinp / rp);
  if (qr < 15.) svfac = std::exp(-2. * qr); else svfac = 0.0;
  cosq = (eqp + svfac * eqm);
  const Z sinq = eqp - svfac * eqm;
  rsinq = (rsv * sinq);
  if (std::fabs(qr) < (double)1.0e-5f && zabs(rsv) < (double)1.0e-5f) sinqr = dm;
  else sinqr = (sinq / rsv);
}

// dnka_tregn :1989-2984 (solid branch); CA(i,j) -> ca[i-1][j-1]
void dnka_tregn(Z ca[5][5], Z cosp, Z rsinp, Z sinpr, Z cossv, Z rsinsv, Z sinsvr, const Eig& g, double ex) {
  double dfac;
  if (ex > 35.0) dfac = 0.0; else dfac = std::exp(-ex);
  const Z a1 = Z(0.5) / g.np;
  const Z a2 = Z(0.5) / g.nsv;
  const Z c1 = 2. * a1 * cosp;
  const Z ls1 = 2. * a1 * rsinp;
  const Z s1l = 2. * a1 * sinpr;
  const Z c2 = 2. * a2 * cossv;
  const Z ls2 = 2. * a2 * rsinsv;
  const Z s2l = 2. * a2 * sinsvr;
  Z x[5][3];
  x[1][1] = g.x11; x[2][1] = g.x21; x[3][1] = g.x31; x[4][1] = g.x41;
  x[1][2] = g.x12; x[2][2] = g.x22; x[3][2] = g.x32; x[4][2] = g.x42;
  Z tca11, tca12, tca13, tca15, tca16, tca21, tca22, tca23, tca25, tca31, tca32, tca33, tca51, tca52, tca61;
#include "dnka_tca.inc"
  ca[0][0] = tca11; ca[0][1] = tca12; ca[0][2] = tca13; ca[0][3] = tca15; ca[0][4] = tca16;
  ca[1][0] = tca21; ca[1][1] = tca22; ca[1][2] = tca23; ca[1][3] = tca25; ca[1][4] = tca15;
  ca[2][0] = 2 * tca31; ca[2][1] = 2 * tca32; ca[2][2] = 2 * tca33 - Z(dfac); ca[2][3] = -2 * tca23; ca[2][4] = -2 * tca13;
  ca[3][0] = tca51; ca[3][1] = tca52; ca[3][2] = -tca32; ca[3][3] = tca22; ca[3][4] = tca12;
  ca[4][0] = tca61; ca[4][1] = tca51; ca[4][2] = -tca31; ca[4][3] = tca21; ca[4][4] = tca11;
}

// hska :3474-3556 (solid branch)
void hska(double AA[4][4], Z tcosp, Z trsinp, Z tsinpr, Z tcossv, Z trsinsv, Z tsinsvr, const Eig& g) {
  const Z cosp = tcosp / g.np, sinpr = tsinpr / g.np, rsinp = trsinp / g.np;
  const Z cossv = tcossv / g.nsv, sinsvr = tsinsvr / g.nsv, rsinsv = trsinsv / g.nsv;
  const Z &x11 = g.x11, &x21 = g.x21, &x31 = g.x31, &x41 = g.x41, &x12 = g.x12, &x22 = g.x22, &x32 = g.x32, &x42 = g.x42;
  AA[0][0] = (x11 * x41 * cosp + x12 * x42 * cossv).re;
  AA[0][1] = (-x11 * x31 * sinpr - x12 * x32 * rsinsv).re;
  AA[0][2] = (-x11 * x21 * cosp - x12 * x22 * cossv).re;
  AA[0][3] = (x11 * x11 * sinpr + x12 * x12 * rsinsv).re;
  AA[1][0] = (x21 * x41 * rsinp + x22 * x42 * sinsvr).re;
  AA[1][1] = (-x21 * x31 * cosp - x22 * x32 * cossv).re;
  AA[1][2] = (-x21 * x21 * rsinp - x22 * x22 * sinsvr).re;
  AA[2][0] = (x31 * x41 * cosp + x32 * x42 * cossv).re;
  AA[2][1] = (-x31 * x31 * sinpr - x32 * x32 * rsinsv).re;
  AA[3][0] = (x41 * x41 * rsinp + x42 * x42 * sinsvr).re;
  AA[1][3] = -AA[0][2];
  AA[2][2] = AA[1][1];
  AA[2][3] = -AA[0][1];
  AA[3][1] = -AA[2][0];
  AA[3][2] = -AA[1][0];
  AA[3][3] = AA[0][0];
}

// up :1831-1987
void up(Ctx& c, double omega, double wvno, double& fr) {
  const int mmax = c.mmax;
  const double wvno2 = wvno * wvno, om2 = omega * omega;
  Z gbr[5];
  evalg(c, mmax - 1, wvno, omega, om2, wvno2, gbr);
  for (int i = 0; i < 5; ++i) c.cd[mmax - 1][i] = Z(gbr[i].re);
  c.exe[mmax - 1] = 0.0;
  double exsum = 0.0;
  for (int m = mmax - 2; m >= 0; --m) {
    Eig g;
    gettiegn(c, m, omega, wvno, om2, wvno2, g);
    const Z p = g.rp * c.zd[m], q = g.rsv * c.zd[m];
    Z cosp, cossv, rsinp, rsinsv, sinpr, sinsvr;
    double pex, svex;
    varsv(p, q, g.rp, g.rsv, cosp, cossv, rsinp, rsinsv, sinpr, sinsvr, pex, svex, c.zd[m]);
    Z ca[5][5];
    dnka_tregn(ca, cosp, rsinp, sinpr, cossv, rsinsv, sinsvr, g, pex + svex);
    Z ee[5];
    for (int i = 0; i < 5; ++i) {
      Z cr(0.0);
      for (int j = 0; j < 5; ++j) cr = cr + c.cd[m + 1][j] * ca[j][i];
      ee[i] = cr;
    }
    // cnormc :1472-1508
    double t1 = 0.0;
    for (int i = 0; i < 5; ++i)
      if (zabs(ee[i]) > t1) t1 = zabs(ee[i]);
    if (t1 < 1.e-40) t1 = 1.0;
    for (int i = 0; i < 5; ++i) ee[i] = ee[i] / t1;
    const double exn = std::log(t1);
    exsum = exsum + pex + svex + exn;
    c.exe[m] = exsum;
    for (int i = 0; i < 5; ++i) c.cd[m][i] = ee[i];
  }
  fr = c.cd[0][0].re;
}

// down :3558-3706
void down(Ctx& c, double omega, double wvno) {
  const int mmax = c.mmax;
  const double om2 = omega * omega, wvno2 = wvno * wvno;
  c.vv[0][0] = 1.0; c.vv[0][1] = 0.0; c.vv[0][2] = 0.0; c.vv[0][3] = 0.0;
  c.exa[0] = 0.0;
  double exsum = 0.0;
  for (int m = 0; m < mmax - 1; ++m) {
    Eig g;
    gettiegn(c, m, omega, wvno, om2, wvno2, g);
    const Z p = g.rp * c.zd[m], q = g.rsv * c.zd[m];
    Z cosp, cossv, rsinp, rsinsv, sinpr, sinsvr;
    double pex, svex, dfac, cpex;
    varsv(p, q, g.rp, g.rsv, cosp, cossv, rsinp, rsinsv, sinpr, sinsvr, pex, svex, c.zd[m]);
    double AA[4][4];
    if (pex > svex) {
      if ((pex - svex) > 40.0) dfac = 0.0; else dfac = std::exp(-(pex - svex));
      cpex = pex;
      hska(AA, cosp, rsinp, sinpr, dfac * cossv, dfac * rsinsv, dfac * sinsvr, g);
    } else {
      if ((svex - pex) > 40.0) dfac = 0.0; else dfac = std::exp(-(svex - pex));
      cpex = svex;
      hska(AA, dfac * cosp, dfac * rsinp, dfac * sinpr, cossv, rsinsv, sinsvr, g);
    }
    double aa0[4];
    for (int i = 0; i < 4; ++i) {
      double cc = 0.0;
      for (int j = 0; j < 4; ++j) cc = cc + AA[i][j] * c.vv[m][j];
      aa0[i] = cc;
    }
    // rnormc :1510-1546
    double t1 = 0.0;
    for (int i = 0; i < 4; ++i)
      if (std::fabs(aa0[i]) > t1) t1 = std::fabs(aa0[i]);
    if (t1 < 1.e-40) t1 = 1.0;
    for (int i = 0; i < 4; ++i) aa0[i] = aa0[i] / t1;
    const double ex2 = std::log(t1);
    exsum = exsum + cpex + ex2;
    c.exa[m + 1] = exsum;
    for (int i = 0; i < 4; ++i) c.vv[m + 1][i] = aa0[i];
  }
}

// svfunc :1556-1829
void svfunc(Ctx& c, double omega, double wvno) {
  double fr;
  up(c, omega, wvno, fr);
  down(c, omega, wvno);
  const double f1213 = -c.cd[0][1].re;
  c.ur[0] = (c.cd[0][2] / c.cd[0][1]).re;
  c.uz[0] = 1.0;
  c.tz[0] = 0.0;
  c.tr[0] = 0.0;
  c.uu0[0] = c.ur[0]; c.uu0[1] = 1.0; c.uu0[2] = fr; c.uu0[3] = fr;
  for (int i = 1; i < c.mmax; ++i) {
    const double cd1 = c.cd[i][0].re, cd2 = c.cd[i][1].re, cd3 = c.cd[i][2].re, cd4 = -c.cd[i][2].re,
                 cd5 = c.cd[i][3].re, cd6 = c.cd[i][4].re;
    const double tz1 = -c.vv[i][3], tz2 = -c.vv[i][2], tz3 = c.vv[i][1], tz4 = c.vv[i][0];
    const double uu1 = tz2 * cd6 - tz3 * cd5 + tz4 * cd4;
    const double uu2 = -tz1 * cd6 + tz3 * cd3 - tz4 * cd2;
    const double uu3 = tz1 * cd5 - tz2 * cd3 + tz4 * cd1;
    const double uu4 = -tz1 * cd4 + tz2 * cd2 - tz3 * cd1;
    const double ext = c.exa[i] + c.exe[i] - c.exe[0];
    if (ext > -80.0 && ext < 80.0) {
      const double fact = std::exp(ext);
      c.ur[i] = uu1 * fact / f1213;
      c.uz[i] = uu2 * fact / f1213;
      c.tz[i] = uu3 * fact / f1213;
      c.tr[i] = uu4 * fact / f1213;
    } else {
      c.ur[i] = 0.0; c.uz[i] = 0.0; c.tz[i] = 0.0; c.tr[i] = 0.0;
    }
  }
}

struct Mat { double a12, a14, a21, a23, ah, av, bh, bv, eta, rho, TA, TC, TF, TL, TN; };

// getmat :3993-4068 (solid branch)
void getmat(const Ctx& c, int m, double wvno, Mat& o) {
  o.ah = std::sqrt(c.zta[m] / c.zrho[m]);
  o.av = std::sqrt(c.ztc[m] / c.zrho[m]);
  o.bh = std::sqrt(c.ztn[m] / c.zrho[m]);
  o.bv = std::sqrt(c.ztl[m] / c.zrho[m]);
  o.rho = c.zrho[m];
  o.TL = c.ztl[m]; o.TN = c.ztn[m]; o.TC = c.ztc[m]; o.TA = c.zta[m]; o.TF = c.ztf[m];
  o.eta = o.TF / (o.TA - 2. * o.TL);
  o.a12 = -wvno;
  o.a14 = 1.0 / o.TL;
  o.a21 = wvno * o.TF / o.TC;
  o.a23 = 1.0 / o.TC;
}

// ffunc/gfunc/h1func/h2func :1382-1470
Z ffunc(Z nub, double dm) {
  if (zabs(nub) < 1.0e-08) return Z(dm);
  const Z argcd = nub * dm;
  Z exqq;
  if (argcd.re < 40.0) exqq = zexp(-2.0 * argcd); else exqq = Z(0.0);
  return (Z(1.0) - exqq) / (2.0 * nub);
}
Z gfunc(Z nub, double dm) {
  const Z argcd = nub * dm;
  if (argcd.re < 75) return zexp(-argcd) * dm;
  return Z(0.0, 0.0);
}
Z h1func(Z nua, Z nub, double dm) {
  if (zabs(nub + nua) < 1.0e-08) return Z(dm);
  const Z argcd = (nua + nub) * dm;
  Z exqq;
  if (argcd.re < 40.0) exqq = zexp(-argcd); else exqq = Z(0.0);
  return (Z(1.0) - exqq) / (nub + nua);
}
Z h2func(Z nua, Z nub, double dm) {
  if (zabs(nub - nua) < 1.0e-08) return Z(dm);
  Z argcd = nua * dm, exqp, exqq;
  if (argcd.re < 40.0) exqp = zexp(-argcd); else exqp = Z(0.0);
  argcd = nub * dm;
  if (argcd.re < 40.0) exqq = zexp(-argcd); else exqq = Z(0.0);
  return (exqq - exqp) / (nua - nub);
}

// intijr :4070-4242 (solid branch); i,j are 1-based like the reference
double intijr(Ctx& c, int i, int j, int m, int typelyr, double om, double om2, double wvno, double wvno2) {
  Z gbr[5];
  evalg(c, m, wvno, om, om2, wvno2, gbr);
  const Z ra = c.ra, rb = c.rb;
  Z (*e)[4] = c.e;
  Z (*einv)[4] = c.einv;
  const int I = i - 1, J = j - 1;
  Z cint;
  if (typelyr == 0) {
    const Z km1pd = einv[2][0] * c.ur[m] + einv[2][1] * c.uz[m] + einv[2][2] * c.tz[m] + einv[2][3] * c.tr[m];
    const Z km1sd = einv[3][0] * c.ur[m] + einv[3][1] * c.uz[m] + einv[3][2] * c.tz[m] + einv[3][3] * c.tr[m];
    const Z kmpu = einv[0][0] * c.ur[m + 1] + einv[0][1] * c.uz[m + 1] + einv[0][2] * c.tz[m + 1] + einv[0][3] * c.tr[m + 1];
    const Z kmsu = einv[1][0] * c.ur[m + 1] + einv[1][1] * c.uz[m + 1] + einv[1][2] * c.tz[m + 1] + einv[1][3] * c.tr[m + 1];
    const double dm = c.zd[m];
    const Z FA = ffunc(ra, dm), GA = gfunc(ra, dm), FB = ffunc(rb, dm), GB = gfunc(rb, dm);
    const Z H1 = h1func(ra, rb, dm), H2 = h2func(ra, rb, dm);
    cint = e[I][0] * e[J][0] * kmpu * kmpu * FA
         + e[I][2] * e[J][2] * km1pd * km1pd * FA
         + e[I][1] * e[J][1] * kmsu * kmsu * FB
         + e[I][3] * e[J][3] * km1sd * km1sd * FB
         + H1 * ((e[I][0] * e[J][1] + e[I][1] * e[J][0]) * kmpu * kmsu +
                 (e[I][2] * e[J][3] + e[I][3] * e[J][2]) * km1pd * km1sd)
         + H2 * ((e[I][0] * e[J][3] + e[I][3] * e[J][0]) * kmpu * km1sd +
                 (e[I][1] * e[J][2] + e[I][2] * e[J][1]) * km1pd * kmsu)
         + GA * (e[I][0] * e[J][2] + e[I][2] * e[J][0]) * kmpu * km1pd
         + GB * (e[I][1] * e[J][3] + e[I][3] * e[J][1]) * kmsu * km1sd;
  } else {
    const Z km1pd = einv[2][0] * c.ur[m] + einv[2][1] * c.uz[m] + einv[2][2] * c.tz[m] + einv[2][3] * c.tr[m];
    const Z km1sd = einv[3][0] * c.ur[m] + einv[3][1] * c.uz[m] + einv[3][2] * c.tz[m] + einv[3][3] * c.tr[m];
    cint = e[I][2] * e[J][2] * km1pd * km1pd / (2.0 * ra)
         + (e[I][2] * e[J][3] + e[I][3] * e[J][2]) * km1pd * km1sd / (ra + rb)
         + e[I][3] * e[J][3] * km1sd * km1sd / (2.0 * rb);
  }
  return cint.re;
}

// energy :3774-3991 (solid branch, without getdcdh)
void energy(Ctx& c, double om, double wvno) {
  const int mmax = c.mmax;
  c.sumi0 = c.sumi1 = c.sumi2 = c.sumi3 = 0.0;
  const double cph = om / wvno, om2 = om * om, wvno2 = wvno * wvno;
  for (int m = 0; m < mmax; ++m) {
    Mat g;
    getmat(c, m, wvno, g);
    const int typelyr = (m == mmax - 1) ? 1 : 0;
    const double INT11 = intijr(c, 1, 1, m, typelyr, om, om2, wvno, wvno2);
    const double INT13 = intijr(c, 1, 3, m, typelyr, om, om2, wvno, wvno2);
    const double INT22 = intijr(c, 2, 2, m, typelyr, om, om2, wvno, wvno2);
    const double INT24 = intijr(c, 2, 4, m, typelyr, om, om2, wvno, wvno2);
    const double INT33 = intijr(c, 3, 3, m, typelyr, om, om2, wvno, wvno2);
    const double INT44 = intijr(c, 4, 4, m, typelyr, om, om2, wvno, wvno2);
    const double URUR = INT11, UZUZ = INT22;
    const double DURDUR = g.a12 * g.a12 * INT22 + 2. * g.a12 * g.a14 * INT24 + g.a14 * g.a14 * INT44;
    const double DUZDUZ = g.a21 * g.a21 * INT11 + 2. * g.a21 * g.a23 * INT13 + g.a23 * g.a23 * INT33;
    const double URDUZ = g.a21 * INT11 + g.a23 * INT13;
    const double UZDUR = g.a12 * INT22 + g.a14 * INT24;
    c.sumi0 = c.sumi0 + g.rho * (URUR + UZUZ);
    c.sumi1 = c.sumi1 + g.TL * UZUZ + g.TA * URUR;
    c.sumi2 = c.sumi2 + g.TL * UZDUR - g.TF * URDUZ;
    c.sumi3 = c.sumi3 + g.TL * DURDUR + g.TC * DUZDUZ;
    const double facah = g.rho * g.ah * (URUR - 2. * g.eta * URDUZ / wvno);
    const double facav = g.rho * g.av * DUZDUZ / wvno2;
    const double facbh = 0.0;
    const double facbv = g.rho * g.bv * (UZUZ + 2. * UZDUR / wvno + DURDUR / wvno2 + 4. * g.eta * URDUZ / wvno);
    const double facn = -g.TF * URDUZ / (wvno * g.eta);
    c.dcdah[m] = facah; c.dcdav[m] = facav; c.dcdbh[m] = facbh; c.dcdbv[m] = facbv; c.dcdn[m] = facn;
    const double facr = -0.5 * cph * cph * (URUR + UZUZ);
    c.dcdr[m] = 0.5 * (g.av * facav + g.ah * facah + g.bv * facbv) / g.rho + facr;
  }
  c.flagr = om2 * c.sumi0 - wvno2 * c.sumi1 - 2.0 * wvno * c.sumi2 - c.sumi3;
  c.ugr = (wvno * c.sumi1 + c.sumi2) / (om * c.sumi0);
  c.are = wvno / (2.0 * om * c.ugr * c.sumi0);
  for (int m = 0; m < mmax; ++m) {
    c.dcdah[m] = c.dcdah[m] / (c.ugr * c.sumi0);
    c.dcdav[m] = c.dcdav[m] / (c.ugr * c.sumi0);
    c.dcdbh[m] = c.dcdbh[m] / (c.ugr * c.sumi0);
    c.dcdbv[m] = c.dcdbv[m] / (c.ugr * c.sumi0);
    c.dcdr[m] = c.dcdr[m] / (c.ugr * c.sumi0);
    c.dcdn[m] = c.dcdn[m] / (c.ugr * c.sumi0);
  }
}

// gammap :3708-3772 (note the reference's pi = 3.141592653589493)
void gammap(Ctx& c, double omega, double& wvno, double& gammar) {
  gammar = 0.0;
  double dc = 0.0;
  const double pi = 3.141592653589493;
  for (int i = 0; i < c.mmax; ++i) {
    Mat g;
    getmat(c, i, wvno, g);
    double x = c.dcdbh[i] * g.bh * c.zqbi[i] + c.dcdbv[i] * g.bv * c.zqbi[i];
    gammar = gammar + x;
    double omgref = 2.0 * pi * c.zfrefs[i];
    dc = dc + std::log(omega / omgref) * x / pi;
    x = c.dcdav[i] * g.av * c.zqai[i] + c.dcdah[i] * g.ah * c.zqai[i];
    gammar = gammar + x;
    omgref = 2.0 * pi * c.zfrefp[i];
    dc = dc + std::log(omega / omgref) * x / pi;
  }
  double cph = omega / wvno;
  gammar = 0.5 * wvno * gammar / cph;
  cph = cph + dc;
  wvno = omega / cph;
}

}  // namespace

// tregn96 :49-719
int tregn96(int nl_in, const float* d_in, const float* TA_in, const float* TC_in, const float* TF_in,
            const float* TL_in, const float* TN_in, const float* TRho_in, const float* qai_in,
            const float* qbi_in, const float* etapi_in, const float* etasi_in, const float* frefpi_in,
            const float* frefsi_in, int Nt_in, const float* t_in, const float* cp_in, float* dcdah_out,
            float* dcdbv_out, float* dcdn_out) {
  (void)etapi_in; (void)etasi_in;
  if (nl_in > NL || nl_in < 2 || Nt_in > NP) return ERR_LAYERS;
  static thread_local Ctx ctx;
  Ctx& c = ctx;
  const int mmax = nl_in;
  c.mmax = mmax;
  float d[NL], TA[NL], TC[NL], TF[NL], TL[NL], TN[NL], TRho[NL], qai[NL], qbi[NL], frefpi[NL], frefsi[NL];
  for (int i = 0; i < mmax; ++i) {
    d[i] = d_in[i]; TA[i] = TA_in[i]; TC[i] = TC_in[i]; TF[i] = TF_in[i]; TN[i] = TN_in[i]; TL[i] = TL_in[i];
    TRho[i] = TRho_in[i]; qai[i] = qai_in[i]; qbi[i] = qbi_in[i]; frefpi[i] = frefpi_in[i]; frefsi[i] = frefsi_in[i];
    if (TN[i] <= 1.0e-4f * TA[i]) return ERR_BAD_ARG;   // fluid layers are outside the restated scope
  }
  for (int i = 0; i < mmax; ++i) {   // :283-298 (dogam = .true.)
    if (qai[i] > 1.0f) qai[i] = 1.0f / qai[i];
    if (qbi[i] > 1.0f) qbi[i] = 1.0f / qbi[i];
    c.zqai[i] = qai[i];
    c.zqbi[i] = qbi[i];
    if (frefpi[i] <= 0.0f) frefpi[i] = 1.0f;
    if (frefsi[i] <= 0.0f) frefsi[i] = 1.0f;
    c.zfrefp[i] = frefpi[i];
    c.zfrefs[i] = frefsi[i];
  }
  // sphere_tdisp96 :771-838 (radius 6371, TF left untouched)
  {
    const double ar = (double)6371.f;
    double r0 = ar + (double)0.0f, r1, z0, z1, tmp;
    d[mmax - 1] = 1.0f;
    for (int i = 0; i < mmax; ++i) {
      r1 = r0 - (double)d[i];
      z0 = ar * std::log(ar / r0);
      z1 = ar * std::log(ar / r1);
      d[i] = (float)(z1 - z0);
      tmp = (ar + ar) / (r0 + r1);
      const float rhosph = TRho[i];
      TRho[i] = (float)((double)rhosph * std::pow(tmp, (double)(-2.275f)));
      TA[i] = (float)((double)TA[i] * std::pow(tmp, (double)(-0.2750f)));
      TC[i] = (float)((double)TC[i] * std::pow(tmp, (double)(-0.2750f)));
      const float elsph = TL[i];
      TL[i] = (float)((double)elsph * std::pow(tmp, (double)(-0.2750f)));
      const float ensph = TN[i];
      TN[i] = (float)((double)ensph * std::pow(tmp, (double)(-0.2750f)));
      r0 = r1;
    }
    d[mmax - 1] = 0.0f;
  }
  for (int i = 0; i < mmax; ++i) {   // :374-397
    c.zd[i] = d[i]; c.zta[i] = TA[i]; c.ztc[i] = TC[i]; c.ztl[i] = TL[i]; c.ztn[i] = TN[i]; c.ztf[i] = TF[i];
    c.zrho[i] = TRho[i];
  }
  // bldsph :1298-1380 (leaves zd(mmax)=1.0; only layers < mmax use zd afterwards)
  {
    const double ar = 6370.0;
    double r0 = ar, r1, tmp;
    c.zd[mmax - 1] = 1.0;
    for (int i = 0; i < mmax; ++i) {
      r1 = r0 * std::exp(-c.zd[i] / ar);
      tmp = (ar + ar) / (r0 + r1);
      c.vtp[i] = (float)tmp;
      c.rtp[i] = (float)std::pow(tmp, (double)(-2.275f));
      c.dtp[i] = (float)(ar / r0);
      r0 = r1;
    }
  }
  // insert/srclyr with depth 0: no layer added, lss = lrr = 1 (:419-422)
  const float twopi = 2.f * 3.141592654f;   // REAL*4, :435
  for (int idisp = 0; idisp < Nt_in; ++idisp) {
    const double t = (double)t_in[idisp];
    const double omega = (double)twopi / t;
    double cph = (double)cp_in[idisp];
    double wvno = omega / cph;
    svfunc(c, omega, wvno);
    energy(c, omega, wvno);
    double gammar;
    gammap(c, omega, wvno, gammar);
    cph = omega / wvno;
    // sprayl :1232-1296 (dcdn is NOT rescaled)
    {
      const double ar = 6370.0;
      const double q = cph / (2. * ar * omega);
      const double tm = std::sqrt(1. + q * q);
      const double tm3 = tm * (tm * tm);
      for (int i = 0; i < mmax; ++i) {
        c.dcdah[i] = c.dcdah[i] * (double)c.vtp[i] / tm3;
        c.dcdav[i] = c.dcdav[i] * (double)c.vtp[i] / tm3;
        c.dcdbh[i] = c.dcdbh[i] * (double)c.vtp[i] / tm3;
        c.dcdbv[i] = c.dcdbv[i] * (double)c.vtp[i] / tm3;
        c.dcdr[i] = c.dcdr[i] * (double)c.rtp[i] / tm3;
      }
    }
    // chksiz :1216-1230 and the output copy :668-680; (NP,NL) column-major
    for (int i = 0; i < mmax; ++i) {
      auto sp = [](double v) { return std::fabs(v) < 1.0e-36 ? 0.0f : (float)v; };
      dcdah_out[(size_t)idisp + (size_t)i * NP] = sp(c.dcdah[i]);
      dcdbv_out[(size_t)idisp + (size_t)i * NP] = sp(c.dcdbv[i]);
      dcdn_out[(size_t)idisp + (size_t)i * NP] = sp(c.dcdn[i]);
    }
  }
  return OK;
}

}  // namespace orc
