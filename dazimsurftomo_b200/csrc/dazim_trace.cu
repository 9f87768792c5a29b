// K4 (receiver travel time + ray back-trace + Frechet footprint) and
// K5 (G-row assembly) for sm_100a.
//
// K4 follows srtimes (CalSurfG.f90:1599) and rpathsAzim (rpathsAzim.f90:16;
// rpaths CalSurfG.f90:1735 when AZIM=false) arithmetic step for step in
// float32 (azdist in float64), one THREAD per ray so that every footprint
// element is accumulated sequentially in the reference's order (bit-exact
// float sums).  What is new is the data structure: instead of a
// dense (nvz+2)x(nvx+2) map per ray, the 4x4 block of control points under
// the current B-spline cell lives in registers and is spilled to a per-thread
// sparse slot list (cell -> slot through a 16-bit index map) only when the ray
// changes cell.  When the 32 rays of a warp are done, the warp walks each
// ray's bounding box cooperatively, emitting the |fdm|>=1e-4 entries in the
// reference's (jj,kk) order with ballot compaction and clearing the index map.
//
// K5 turns footprints into CSR rows (two passes: count, scan, fill) with the
// reference's value formulae and its second 1e-4 threshold
// (CalSurfGAniso_Joint.f90:714-752, CalSurfG.f90:1339-1364) or, in forward
// mode, into T_aa = G_c.Gc + G_s.Gs accumulated in the reference's column
// order (FwdTraveltimeCPS.f90:710-761).
#include "dazim_dev.h"
#include <cub/device/device_scan.cuh>

namespace dz {

__device__ __forceinline__ float sin_r(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cube(float x) { return x * (x * x); }
__device__ __forceinline__ void bspl_basis(float u, float* b) {
  b[0] = cube(1.0f - u) / 6.0f;
  b[1] = (4.0f - 6.0f * (u * u) + 3.0f * cube(u)) / 6.0f;
  b[2] = (1.0f + 3.0f * u + 3.0f * (u * u) - 3.0f * cube(u)) / 6.0f;
  b[3] = cube(u) / 6.0f;
}

#define PI_F 3.1415926535898f



struct Foot {
  unsigned short* map; int* skey; float* v0; float* v1; float* v2;
  int nslot, cap, ldc;
  int zmin, zmax, xmin, xmax;
  int overflow;
};

template <bool AZIM>
__device__ __forceinline__ void win_flush(Foot& f, int wz, int wx, const float (&acc)[3][16]) {
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int cell = (wz - 1 + l) * f.ldc + (wx - 1 + m);
      int s = f.map[cell];
      if (s == 0) {
        if (f.nslot >= f.cap) { f.overflow = 1; continue; }
        s = ++f.nslot;
        f.map[cell] = (unsigned short)s;
        f.skey[s - 1] = cell;
      }
      f.v0[s - 1] = acc[0][l * 4 + m];
      if (AZIM) { f.v1[s - 1] = acc[1][l * 4 + m]; f.v2[s - 1] = acc[2][l * 4 + m]; }
    }
}

template <bool AZIM>
__device__ __forceinline__ void win_load(Foot& f, int wz, int wx, float (&acc)[3][16]) {
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int cell = (wz - 1 + l) * f.ldc + (wx - 1 + m);
      const int s = f.map[cell];
      acc[0][l * 4 + m] = s ? f.v0[s - 1] : 0.0f;
      if (AZIM) {
        acc[1][l * 4 + m] = s ? f.v1[s - 1] : 0.0f;
        acc[2][l * 4 + m] = s ? f.v2[s - 1] : 0.0f;
      }
    }
  f.zmin = min(f.zmin, wz); f.zmax = max(f.zmax, wz);
  f.xmin = min(f.xmin, wx); f.xmax = max(f.xmax, wx);
}

// azimuth of the segment, azdist(rgx2,rgz2,rgx1,rgz1,...) -> az (rpathsAzim.f90:429-436, :687-793).
// Only the az branch is evaluated (delta and baz are never used by the caller).
__device__ __forceinline__ float azimuth_deg(float stalat, float stalon, float evtlat, float evtlon) {
  const double pi = (double)3.1415926535898f;
  const float piby2 = (float)(pi / 2.0);
  const double rad = 2.0 * pi / 360.0;
  const double sph = (double)(1.0f / 298.257f);
  const double scolat = (double)piby2 - atan((1.0 - sph) * (1.0 - sph) * tan((double)stalat * rad));
  const double ecolat = (double)piby2 - atan((1.0 - sph) * (1.0 - sph) * tan((double)evtlat * rad));
  const double slon = (double)stalon * rad;
  const double elon = (double)evtlon * rad;
  double ss, cs, sl, cl, se, ce, sel, cel;
  sincos(scolat, &ss, &cs);
  sincos(slon, &sl, &cl);
  sincos(ecolat, &se, &ce);
  sincos(elon, &sel, &cel);
  const double a = ss * cl;
  const double b = ss * sl;
  const double c = cs;
  const double cc = ce;
  const double dd = sel;
  const double ee = -cel;
  const double gg = -cc * ee;
  const double hh = cc * dd;
  const double kk = -se;
  const double rhs1 = (a - dd) * (a - dd) + (b - ee) * (b - ee) + c * c - 2.0;
  const double rhs2 = (a - gg) * (a - gg) + (b - hh) * (b - hh) + (c - kk) * (c - kk) - 2.0;
  double daz = atan2(rhs1, rhs2);
  if (daz < 0.0) daz = daz + 2 * pi;
  float az = (float)(daz / rad);
  if (fabsf(az - 360.0f) < .00001f) az = 0.0f;
  return az;
}

__device__ __forceinline__ float bilin_vel(const float* __restrict__ veln, const GridC& g, int ipz, int ipx,
                                           float drx, float drz, bool xfirst) {
  // vel = sum veln(ipz-1+m, ipx-1+l)*produ, with the bound guard of rpathsAzim.f90:503/:555
  float vel = 0.0f;
#pragma unroll
  for (int l = 1; l <= 2; ++l)
#pragma unroll
    for (int m = 1; m <= 2; ++m) {
      float produ = (1.0f - fabsf(((float)(m - 1) * g.dnz - drz) / g.dnz));
      produ = produ * (1.0f - fabsf(((float)(l - 1) * g.dnx - drx) / g.dnx));
      if (ipz - 1 + m <= g.nnz && ipx - 1 + l <= g.nnx && ipz - 1 + m >= 1 && ipx - 1 + l >= 1)
        vel = vel + veln[(size_t)(ipx - 1 + l - 1) * g.nnz + (ipz - 1 + m - 1)] * produ;
    }
  return vel;
}

template <bool AZIM>
__device__ void trace_one(const TraceArgs& A, const RayRec rr, Foot& f, int& flags, unsigned long long& nsteps) {
  const GridC& g = A.g;
  const SrcRec sr = A.src[rr.src];
  const size_t ncoarse = (size_t)g.nnx * g.nnz;
  const float* __restrict__ veln = A.veln_c + (size_t)sr.period * ncoarse;
  const float* __restrict__ ttn = A.ttn_c + (size_t)rr.src * coarse_field_size(g.nnx, g.nnz);
  const float* __restrict__ ttnr = A.ttn_r + (size_t)rr.src * REF_N;
  const float scx = sr.scx, scz = sr.scz, earth = g.earth;
  const float rcx = rr.rcx, rcz = rr.rcz;
  const int nnx = g.nnx, nnz = g.nnz;
#define TTN(iz, ix) ttn[cidx((ix)-1, (iz)-1, nnz)]
#define TTNR(iz, ix) ttnr[((ix)-1) * REF_LD + ((iz)-1)]
// nstsr(iz,ix) /= 0  <=>  not alive  <=>  sign bit of the encoded refined field
#define NSTSR(iz, ix) (__float_as_int(ttnr[((ix)-1) * REF_LD + ((iz)-1)]) >> 31)

  // ------------------------- srtimes (CalSurfG.f90:1644-1717) -------------------------
  {
    int irx = (int)((rcx - g.gox) / g.dnx) + 1;
    int irz = (int)((rcz - g.goz) / g.dnz) + 1;
    if (irx < 1 || irx > nnx || irz < 1 || irz > nnz) { flags |= 2; A.dsurf[rr.row] = 0.0f; return; }
    if (irx == nnx) irx = irx - 1;
    if (irz == nnz) irz = irz - 1;
    const float t1 = (scx - rcx) * earth;
    float sred = t1 * t1;
    const float t2 = (scz - rcz) * earth * sin_r(rcx);
    sred = sred + t2 * t2;
    sred = sqrtf(sred);
    int sw = 0;
    if (sred < g.dpl_full) sw = 1;
    if (sr.isx_c == irx && sr.isz_c == irz) sw = 1;
    float trr;
    if (sw == 1) {
      const int isx = min(max(sr.isx_c, 1), nnx - 1), isz = min(max(sr.isz_c, 1), nnz - 1);
      float drx = (scx - g.gox) - (float)(sr.isx_c - 1) * g.dnx;
      float drz = (scz - g.goz) - (float)(sr.isz_c - 1) * g.dnz;
      float vels = 0.0f, velr = 0.0f;
#pragma unroll
      for (int k = 1; k <= 2; ++k)
#pragma unroll
        for (int l = 1; l <= 2; ++l) {
          const float produ = (1.0f - fabsf(((float)(k - 1) * g.dnx - drx) / g.dnx)) *
                              (1.0f - fabsf(((float)(l - 1) * g.dnz - drz) / g.dnz));
          vels = vels + veln[(size_t)(isx - 1 + k - 1) * nnz + (isz - 1 + l - 1)] * produ;
        }
      drx = (rcx - g.gox) - (float)(irx - 1) * g.dnx;
      drz = (rcz - g.goz) - (float)(irz - 1) * g.dnz;
#pragma unroll
      for (int k = 1; k <= 2; ++k)
#pragma unroll
        for (int l = 1; l <= 2; ++l) {
          const float produ = (1.0f - fabsf(((float)(k - 1) * g.dnx - drx) / g.dnx)) *
                              (1.0f - fabsf(((float)(l - 1) * g.dnz - drz) / g.dnz));
          velr = velr + veln[(size_t)(irx - 1 + k - 1) * nnz + (irz - 1 + l - 1)] * produ;
        }
      trr = 2.0f * sred / (vels + velr);
    } else {
      const float drx = (rcx - g.gox) - (float)(irx - 1) * g.dnx;
      const float drz = (rcz - g.goz) - (float)(irz - 1) * g.dnz;
      trr = 0.0f;
#pragma unroll
      for (int k = 1; k <= 2; ++k)
#pragma unroll
        for (int l = 1; l <= 2; ++l) {
          const float produ = (1.0f - fabsf(((float)(l - 1) * g.dnz - drz) / g.dnz)) *
                              (1.0f - fabsf(((float)(k - 1) * g.dnx - drx) / g.dnx));
          trr = trr + TTN(irz - 1 + l, irx - 1 + k) * produ;
        }
    }
    A.dsurf[rr.row] = trr;
  }

  // ------------------------- rpathsAzim (rpathsAzim.f90:145-612) -------------------------
  const int isx = sr.isx_t, isz = sr.isz_t;
  const float dpl = g.dpl_half;
  int ipx = (int)((rcx - g.gox) / g.dnx) + 1;
  int ipz = (int)((rcz - g.goz) / g.dnz) + 1;
  if (ipx < 1 || ipx >= nnx || ipz < 1 || ipz >= nnz) { flags |= 2; return; }
  float rgx_j = rcx, rgz_j = rcz;
  int sw = 0;
  {
    const float t1 = (scx - rgx_j) * earth;
    float sred = t1 * t1;
    const float t2 = (scz - rgz_j) * earth * sin_r(rgx_j);
    sred = sred + t2 * t2;
    sred = sqrtf(sred);
    if (sred < 2.0f * dpl) sw = 1;
  }
  int ipxr = (int)((rcx - sr.goxr) / sr.dnxr) + 1;
  int ipzr = (int)((rcz - sr.gozr) / sr.dnzr) + 1;
  int igref = 1;
  if (ipxr < 1 || ipxr >= sr.nnxr) igref = 0;
  if (ipzr < 1 || ipzr >= sr.nnzr) igref = 0;
  if (igref == 1) {
    if (NSTSR(ipzr, ipxr) != 0 || NSTSR(ipzr + 1, ipxr) != 0) igref = 0;
    if (NSTSR(ipzr, ipxr + 1) != 0 || NSTSR(ipzr + 1, ipxr + 1) != 0) igref = 0;
  }
  if (sw == 0 && igref == 1 && ipxr == isx && ipzr == isz) sw = 1;

  float acc[3][16];
  int wz = -1000, wx = -1000;   // cached window origin (ivzt, ivxt); none yet
  const long maxrp = (long)nnx * nnz;
  for (long j = 1; j <= maxrp; ++j) {
    if (sw == 1) break;
    ++nsteps;
    const float sinj = sin_r(rgx_j);
    float dtx, dtz;
    if (igref == 1) {
      dtx = TTNR(ipzr, ipxr + 1) - TTNR(ipzr, ipxr);
      dtx = dtx + TTNR(ipzr + 1, ipxr + 1) - TTNR(ipzr + 1, ipxr);
      dtx = dtx / (2.0f * earth * sr.dnxr);
      dtz = TTNR(ipzr + 1, ipxr) - TTNR(ipzr, ipxr);
      dtz = dtz + TTNR(ipzr + 1, ipxr + 1) - TTNR(ipzr, ipxr + 1);
      dtz = dtz / (2.0f * earth * sinj * sr.dnzr);
    } else {
      dtx = TTN(ipz, ipx + 1) - TTN(ipz, ipx);
      dtx = dtx + TTN(ipz + 1, ipx + 1) - TTN(ipz + 1, ipx);
      dtx = dtx / (2.0f * earth * g.dnx);
      dtz = TTN(ipz + 1, ipx) - TTN(ipz, ipx);
      dtz = dtz + TTN(ipz + 1, ipx + 1) - TTN(ipz, ipx + 1);
      dtz = dtz / (2.0f * earth * sinj * g.dnz);
    }
    float rd1 = sqrtf(dtx * dtx + dtz * dtz);
    float rgx_j1 = rgx_j - dpl * dtx / (earth * rd1);
    float rgz_j1 = rgz_j - dpl * dtz / (earth * sinj * rd1);
    const int ipxo = ipx, ipzo = ipz;
    ipxr = (int)((rgx_j1 - sr.goxr) / sr.dnxr) + 1;
    ipzr = (int)((rgz_j1 - sr.gozr) / sr.dnzr) + 1;
    igref = 1;
    if (ipxr < 1 || ipxr >= sr.nnxr) igref = 0;
    if (ipzr < 1 || ipzr >= sr.nnzr) igref = 0;
    if (igref == 1) {
      if (NSTSR(ipzr, ipxr) != 0 || NSTSR(ipzr + 1, ipxr) != 0) igref = 0;
      if (NSTSR(ipzr, ipxr + 1) != 0 || NSTSR(ipzr + 1, ipxr + 1) != 0) igref = 0;
    }
    ipx = (int)((rgx_j1 - g.gox) / g.dnx) + 1;
    ipz = (int)((rgz_j1 - g.goz) / g.dnz) + 1;
    {
      const float t1 = (scx - rgx_j1) * earth;
      float sred = t1 * t1;
      const float t2 = (scz - rgz_j1) * earth * sin_r(rgx_j1);
      sred = sred + t2 * t2;
      sred = sqrtf(sred);
      sw = 0;
      if (sred < 2.0f * dpl) sw = 1;
    }
    if (sw == 0 && igref == 1 && ipxr == isx && ipzr == isz) sw = 1;
    if (ipx < 1) { rgx_j1 = g.gox; ipx = 1; flags |= 1; }
    if (ipx >= nnx) { rgx_j1 = g.gox + (float)(nnx - 1) * g.dnx; ipx = nnx - 1; flags |= 1; }
    if (ipz < 1) { rgz_j1 = g.goz; ipz = 1; flags |= 1; }
    if (ipz >= nnz) { rgz_j1 = g.goz + (float)(nnz - 1) * g.dnz; ipz = nnz - 1; flags |= 1; }

    float c2 = 0.0f, s2 = 0.0f;
    if (AZIM) {
      const float rgx1 = (PI_F / 2 - rgx_j) * 180.0f / PI_F;
      const float rgz1 = rgz_j * 180.0f / PI_F;
      const float rgx2 = (PI_F / 2 - rgx_j1) * 180.0f / PI_F;
      const float rgz2 = rgz_j1 * 180.0f / PI_F;
      const float az = azimuth_deg(rgx2, rgz2, rgx1, rgz1);
      const float rgpsi = az / 180 * PI_F;
      double ds2, dc2;
      sincos((double)(2.0f * rgpsi), &ds2, &dc2);
      c2 = (float)dc2;
      s2 = (float)ds2;
    }
    const int ivx = (ipx - 1) / g.gdx + 1;
    const int ivz = (ipz - 1) / g.gdz + 1;
    const int ivxo = (ipxo - 1) / g.gdx + 1;
    const int ivzo = (ipzo - 1) / g.gdz + 1;
    int nhp = 0;
    float vrat[3];
    int chp[3];
    if (ivx != ivxo) {
      nhp = nhp + 1;
      float xi;
      if (ivx > ivxo) xi = g.gox + (float)(ivx - 1) * g.dvx;
      else xi = g.gox + (float)ivx * g.dvx;
      vrat[0] = (xi - rgx_j) / (rgx_j1 - rgx_j);
      chp[0] = 1;
    }
    if (ivz != ivzo) {
      nhp = nhp + 1;
      float zi;
      if (ivz > ivzo) zi = g.goz + (float)(ivz - 1) * g.dvz;
      else zi = g.goz + (float)ivz * g.dvz;
      rd1 = (zi - rgz_j) / (rgz_j1 - rgz_j);
      if (nhp == 1) {
        vrat[0] = rd1;
        chp[0] = 2;
      } else {
        if (rd1 >= vrat[0]) {
          vrat[1] = rd1;
          chp[1] = 2;
        } else {
          vrat[1] = vrat[0];
          chp[1] = chp[0];
          vrat[0] = rd1;
          chp[0] = 2;
        }
      }
    }
    nhp = nhp + 1;
    vrat[nhp - 1] = 1.0f;
    chp[nhp - 1] = 0;
    float drx = (rgx_j - g.gox) - (float)(ipxo - 1) * g.dnx;
    float drz = (rgz_j - g.goz) - (float)(ipzo - 1) * g.dnz;
    float vel = bilin_vel(veln, g, ipzo, ipxo, drx, drz, true);
    drx = (rgx_j - g.gox) - (float)(ivxo - 1) * g.dvx;
    drz = (rgz_j - g.goz) - (float)(ivzo - 1) * g.dvz;
    float vi[4], wi[4], vio[4], wio[4];
    bspl_basis(drx / g.dvx, vi);
    bspl_basis(drz / g.dvz, wi);
    int ivxt = ivxo, ivzt = ivzo;
    float vprev = 0.0f;
    for (int k = 1; k <= nhp; ++k) {
      const float velo = vel;
#pragma unroll
      for (int q = 0; q < 4; ++q) { vio[q] = vi[q]; wio[q] = wi[q]; }
      if (k > 1) {
        if (chp[k - 2] == 1) ivxt = ivx;
        else if (chp[k - 2] == 2) ivzt = ivz;
      }
      const float vr = (k == 1) ? vrat[0] : (k == 2 ? vrat[1] : vrat[2]);
      const float rigz = rgz_j + vr * (rgz_j1 - rgz_j);
      const float rigx = rgx_j + vr * (rgx_j1 - rgx_j);
      const int ipxt = (int)((rigx - g.gox) / g.dnx) + 1;
      const int ipzt = (int)((rigz - g.goz) / g.dnz) + 1;
      drx = (rigx - g.gox) - (float)(ipxt - 1) * g.dnx;
      drz = (rigz - g.goz) - (float)(ipzt - 1) * g.dnz;
      // note the m/n loop nest of rpathsAzim.f90:551-559 (x outer, z inner): same sum order as bilin_vel
      vel = bilin_vel(veln, g, ipzt, ipxt, drx, drz, true);
      drx = (rigx - g.gox) - (float)(ivxt - 1) * g.dvx;
      drz = (rigz - g.goz) - (float)(ivzt - 1) * g.dvz;
      bspl_basis(drx / g.dvx, vi);
      bspl_basis(drz / g.dvz, wi);
      float dinc;
      if (k == 1) dinc = vr * dpl;
      else dinc = (vr - vprev) * dpl;
      vprev = vr;
      if (ivzt != wz || ivxt != wx) {
        if (ivzt < 1 || ivzt > g.nvz - 1 || ivxt < 1 || ivxt > g.nvx - 1) { f.overflow = 1; sw = 1; break; }
        if (wz != -1000) win_flush<AZIM>(f, wz, wx, acc);
        wz = ivzt; wx = ivxt;
        win_load<AZIM>(f, wz, wx, acc);
      }
      const float vel2 = vel * vel, velo2 = velo * velo;
#pragma unroll
      for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float rdc1 = vi[m] * wi[l] / vel2;
          const float rdc2 = vio[m] * wio[l] / velo2;
          float r1 = -(rdc1 + rdc2) * dinc / 2.0f;
          acc[0][l * 4 + m] = r1 + acc[0][l * 4 + m];
          if (AZIM) {
            r1 = -(rdc1 * c2 + rdc2 * c2) * dinc / 2.0f;
            acc[1][l * 4 + m] = r1 + acc[1][l * 4 + m];
            r1 = -(rdc1 * s2 + rdc2 * s2) * dinc / 2.0f;
            acc[2][l * 4 + m] = r1 + acc[2][l * 4 + m];
          }
        }
    }
    rgx_j = rgx_j1;
    rgz_j = rgz_j1;
  }
  if (wz != -1000) win_flush<AZIM>(f, wz, wx, acc);
#undef TTN
#undef TTNR
#undef NSTSR
}

template <bool AZIM>
__global__ void __launch_bounds__(128) k_trace(TraceArgs A) {
  const int lane = threadIdx.x & 31;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const GridC& g = A.g;
  const int ldc = g.nvx + 2;
  const int ncell = (g.nvz + 2) * ldc;
  const float ftol = 1e-4f;
  Foot f;
  f.map = A.map + (size_t)tid * ncell;
  f.skey = A.skey + (size_t)tid * A.cap;
  f.v0 = A.sval + (size_t)tid * 3 * A.cap;
  f.v1 = f.v0 + A.cap;
  f.v2 = f.v1 + A.cap;
  f.cap = A.cap;
  f.ldc = ldc;
  int flags = 0;
  unsigned long long nsteps = 0;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(A.counter, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= A.nray) break;
    const int ri = base + lane;
    f.nslot = 0; f.overflow = 0;
    f.zmin = 1 << 30; f.zmax = -1; f.xmin = 1 << 30; f.xmax = -1;
    RayRec rr;
    rr.row = -1;
    if (ri < A.nray) {
      rr = A.ray[ri];
      trace_one<AZIM>(A, rr, f, flags, nsteps);
      if (f.overflow) flags |= 4;
    }
    __syncwarp();
    // ---- cooperative ordered emission, one ray of the warp at a time ----
    for (int r = 0; r < 32; ++r) {
      if (base + r >= A.nray) break;
      const int nslot = __shfl_sync(0xffffffffu, f.nslot, r);
      const int row = __shfl_sync(0xffffffffu, rr.row, r);
      int zlo = __shfl_sync(0xffffffffu, f.zmin, r) - 1, zhi = __shfl_sync(0xffffffffu, f.zmax, r) + 2;
      int xlo = __shfl_sync(0xffffffffu, f.xmin, r) - 1, xhi = __shfl_sync(0xffffffffu, f.xmax, r) + 2;
      if (nslot == 0) {
        if (lane == 0) { A.fp_off[row] = 0; A.fp_cnt[row] = 0; }
        continue;
      }
      unsigned long long off = 0;
      if (lane == 0) off = atomicAdd(A.pool_used, (unsigned long long)nslot);
      off = __shfl_sync(0xffffffffu, off, 0);
      const bool room = (off + (unsigned long long)nslot <= A.pool_cap);
      if (!room) flags |= 8;
      const size_t tbase = (size_t)(tid - lane + r);
      unsigned short* map = A.map + tbase * ncell;
      const float* v0 = A.sval + tbase * 3 * A.cap;
      const float* v1 = v0 + A.cap;
      const float* v2 = v1 + A.cap;
      const int wx_n = xhi - xlo + 1;
      const int ntot = (zhi - zlo + 1) * wx_n;
      int cnt = 0;
      for (int e0 = 0; e0 < ntot; e0 += 32) {
        const int e = e0 + lane;
        int cell = -1, s = 0;
        float a0 = 0.0f;
        bool keep = false;
        if (e < ntot) {
          const int z = zlo + e / wx_n, x = xlo + e % wx_n;
          cell = z * ldc + x;
          s = map[cell];
          if (s) {
            map[cell] = 0;
            a0 = v0[s - 1];
            keep = A.emit_all || (z >= 1 && z <= g.nvz && x >= 1 && x <= g.nvx && fabsf(a0) >= ftol);
          }
        }
        const unsigned msk = __ballot_sync(0xffffffffu, keep);
        if (keep && room) {
          const size_t p = off + cnt + __popc(msk & ((1u << lane) - 1));
          A.fp_cell[p] = cell;
          A.fp_fdm[p] = a0;
          if (AZIM) { A.fp_fdmc[p] = v1[s - 1]; A.fp_fdms[p] = v2[s - 1]; }
        }
        cnt += __popc(msk);
      }
      if (lane == 0) { A.fp_off[row] = (int)off; A.fp_cnt[row] = room ? cnt : 0; }
    }
    __syncwarp();
  }
  if (flags) atomicOr(A.flags, flags);
  if (nsteps) atomicAdd(A.n_steps, nsteps);
}

cudaError_t launch_trace(const TraceArgs& A, bool azim, int nblocks, cudaStream_t st) {
  if (azim) k_trace<true><<<nblocks, 128, 0, st>>>(A);
  else k_trace<false><<<nblocks, 128, 0, st>>>(A);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K5


__global__ void k_coef(int nx, int ny, int nz, const float* __restrict__ vels, float* __restrict__ coe_a,
                       float* __restrict__ coe_rho) {
  // coe_a / coe_rho of CalSurfGAniso_Joint.f90:720-726 for every (node, layer)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nxy = nx * ny;
  if (idx >= nxy * (nz - 1)) return;
  const float v = vels[idx];
  const float v2 = v * v, v3 = v * (v * v), v4 = (v * v) * (v * v);
  const float ca = (2.0947f - (0.8206f * 2) * v + (0.2683f * 3) * v2 - (0.0251f * 4) * v3);
  const float vp = 0.9409f + 2.0947f * v - 0.8206f * v2 + 0.2683f * v3 - 0.0251f * v4;
  const float p2 = vp * vp, p3 = vp * (vp * vp), p4 = (vp * vp) * (vp * vp);
  const float cr = ca * (1.6612f - (0.4721f * 2) * vp + (0.0671f * 3) * p2 - (0.0043f * 4) * p3 +
                         (0.000106f * 5) * p4);
  coe_a[idx] = ca;
  coe_rho[idx] = cr;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_assemble(AsmArgs A) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= A.nrow) return;
  const int row = A.row0 + w;
  const int off = A.fp_off[row], nf = A.fp_cnt[row];
  const int knumi = A.row_knumi[row];
  const int nlay = A.nz - 1;
  const int nblk = (A.mode == 2) ? 3 : 1;
  const size_t nxy = (size_t)A.nx * A.ny;
  const long long nparpi = (long long)A.nvx * A.nvz * nlay;
  const int ldc = A.nvx + 2;
  const float ftol = 1e-4f;
  const int ntot = nblk * nlay * nf;
  long long base = 0;
  if (FILL) base = A.rowptr[w];
  int cnt = 0;
  for (int e0 = 0; e0 < ntot; e0 += 32) {
    const int e = e0 + lane;
    bool keep = false;
    float val = 0.0f;
    int colid = 0;
    if (e < ntot) {
      const int c = e % nf, kb = e / nf;
      const int k = kb % nlay, blk = kb / nlay;     // k 0-based layer
      const int cell = A.fp_cell[off + c];
      const int jj = cell / ldc, kk = cell % ldc;
      const size_t node = (size_t)jj * (A.nvx + 2) + kk;   // 0-based of jj*(nvx+2)+kk+1
      if (blk == 0) {
        const size_t q = node + (size_t)knumi * nxy + (size_t)k * nxy * A.kmax;
        const size_t qc = node + (size_t)k * nxy;
        const double r = (A.sen_vp[q] * (double)A.coe_a[qc] + A.sen_rho[q] * (double)A.coe_rho[qc] + A.sen_vs[q]) *
                         (double)A.fp_fdm[off + c];
        val = (float)r;
      } else {
        const float L = A.lsen[node + (size_t)knumi * nxy + (size_t)k * nxy * A.kmax];
        val = L * (blk == 1 ? A.fp_fdmc[off + c] : A.fp_fdms[off + c]);
      }
      keep = fabsf(val) > ftol;
      colid = (int)((long long)blk * nparpi + (long long)k * A.nvx * A.nvz + (long long)(jj - 1) * A.nvx + kk);  // 1-based nn
    }
    const unsigned msk = __ballot_sync(0xffffffffu, keep);
    if (FILL && keep) {
      const long long p = base + cnt + __popc(msk & ((1u << lane) - 1));
      A.val[p] = val;
      A.col[p] = colid;
      if (A.rowid) A.rowid[p] = row + 1;
    }
    cnt += __popc(msk);
  }
  if (!FILL && lane == 0) A.nnz_row[w] = cnt;
}



__global__ void k_taa(TaaArgs A) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= A.nrow) return;
  const int row = A.row0 + w;
  const int off = A.fp_off[row], nf = A.fp_cnt[row];
  const int knumi = A.row_knumi[row];
  const size_t nxy = (size_t)A.nx * A.ny;
  const int ldc = A.nvx + 2;
  float sgc = 0.0f, sgs = 0.0f;
  for (int k = 0; k < A.nz - 1; ++k)
    for (int c = 0; c < nf; ++c) {
      const int cell = A.fp_cell[off + c];
      const int jj = cell / ldc, kk = cell % ldc;
      const size_t node = (size_t)jj * (A.nvx + 2) + kk;
      const float L = A.lsen[node + (size_t)knumi * nxy + (size_t)k * nxy * A.kmax];
      const size_t gi = (size_t)(kk - 1) + (size_t)(jj - 1) * A.nvx + (size_t)k * A.nvx * A.nvz;
      sgc = sgc + (L * A.fp_fdmc[off + c]) * A.gc[gi];
      sgs = sgs + (L * A.fp_fdms[off + c]) * A.gs[gi];
    }
  A.taa[row] = sgc + sgs;
}

cudaError_t launch_coef(int nx, int ny, int nz, const float* vels, float* ca, float* cr, cudaStream_t st) {
  const int n = nx * ny * (nz - 1);
  k_coef<<<(n + 255) / 256, 256, 0, st>>>(nx, ny, nz, vels, ca, cr);
  return cudaGetLastError();
}

cudaError_t launch_assemble(const AsmArgs& A, bool fill, cudaStream_t st) {
  if (A.nrow <= 0) return cudaSuccess;
  const int nb = (A.nrow * 32 + 127) / 128;
  if (fill) k_assemble<true><<<nb, 128, 0, st>>>(A);
  else k_assemble<false><<<nb, 128, 0, st>>>(A);
  return cudaGetLastError();
}

cudaError_t launch_taa(const TaaArgs& A, cudaStream_t st) {
  if (A.nrow <= 0) return cudaSuccess;
  k_taa<<<(A.nrow + 127) / 128, 128, 0, st>>>(A);
  return cudaGetLastError();
}

// exclusive scan of 64-bit counts -> row pointers.  (An int input would make cub accumulate in int: a joint G
// of the Yunnan-shaped survey has 4.8e9 non-zeros.)
cudaError_t scan_rowptr(const long long* counts, long long* rowptr, int n, void* tmp, size_t* tmp_bytes, cudaStream_t st) {
  // rowptr[0..n-1] = exclusive sum; the caller adds the last element
  return cub::DeviceScan::ExclusiveSum(tmp, *tmp_bytes, counts, rowptr, n, st);
}

}  // namespace dz
