// K0 (B-spline dicing) and K3 (eikonal solve) for sm_100a.
//
// K3 reproduces the reference's narrow-band fast-marching solve
// (module traveltime, CalSurfG.f90:234-893) step for step: the accepted value
// of a node depends on the order in which the binary heap pops nodes
// (SURVEY H1), so the heap discipline (addtree/downtree/updtree) is kept
// exactly.  The parallelism is (a) across (period, source) solves -- one warp
// per solve, up to 32 solves resident per SM, the whole batch in flight at
// once -- and (b) inside one accept step: the four neighbour updates are
// mutually independent (a node being updated is never "alive", and only alive
// nodes are read), so the 32 lanes fetch the 4 x 8 stencil nodes in one
// coalesced-by-line gather, 16 lanes solve the 4 x 4 quadrant quadratics, and
// a shuffle-min collapses them.  The heap lives in shared memory as
// (key, node) pairs so sift operations never touch global memory for keys;
// levels beyond the shared capacity spill to a per-solve global array.
#include "dazim_dev.h"

namespace dz {

__constant__ float c_ubasis[41 * 4];   // refined B-spline basis, u=(l-1)/40, l=1..41 (CalSurfG.f90:1540-1555)
__constant__ float c_cbasis[6 * 4];    // coarse basis, u=(m-1)/5, m=1..6 (CalSurfG.f90:1469-1488)

cudaError_t upload_basis(const float* ub, const float* cb) {
  cudaError_t e = cudaMemcpyToSymbol(c_ubasis, ub, sizeof(float) * 41 * 4);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_cbasis, cb, sizeof(float) * 6 * 4);
}

// ---------------------------------------------------------------------------
// K0: coarse dicing, one thread per propagation node, one grid.y per period.
// velv: [nper][(nvz+2)*(nvx+2)] float, velv(i,j) at i*(nvx+2)+j  (gridder, CalSurfG.f90:1450-1457)
// veln: [nper][nnx*nnz] column-major (z fastest)
__global__ void k_dice_coarse(GridC g, const float* __restrict__ velv, float* __restrict__ veln) {
  const int per = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nnx * g.nnz) return;
  const int stz = idx % g.nnz + 1, stx = idx / g.nnz + 1;
  // cell (i,j) and local index (l,m): the last cell also owns its far edge
  int i = (stz - 1) / g.gdz + 1, l = (stz - 1) % g.gdz + 1;
  if (i > g.nvz - 1) { i = g.nvz - 1; l = g.gdz + 1; }
  int j = (stx - 1) / g.gdx + 1, m = (stx - 1) % g.gdx + 1;
  if (j > g.nvx - 1) { j = g.nvx - 1; m = g.gdx + 1; }
  const float* vv = velv + (size_t)per * (g.nvz + 2) * (g.nvx + 2);
  const int ldv = g.nvx + 2;
  float sumi = 0.0f;
#pragma unroll
  for (int i1 = 1; i1 <= 4; ++i1) {
    float sumj = 0.0f;
#pragma unroll
    for (int j1 = 1; j1 <= 4; ++j1)
      sumj = sumj + c_cbasis[(m - 1) * 4 + (j1 - 1)] * vv[(i - 2 + i1) * ldv + (j - 2 + j1)];
    sumi = sumi + c_cbasis[(l - 1) * 4 + (i1 - 1)] * sumj;
  }
  veln[(size_t)per * g.nnx * g.nnz + idx] = sumi;
}

// ---------------------------------------------------------------------------
// K3


struct Heap {
  float* sk; int* sn;          // shared part
  float* gk; int* gn;          // spill part
  int hcap, hspill;
  int ntr;
  __device__ __forceinline__ float key(int p) const { return p < hcap ? sk[p] : gk[p - hcap]; }
  __device__ __forceinline__ int node(int p) const { return p < hcap ? sn[p] : gn[p - hcap]; }
  __device__ __forceinline__ void set(int p, float k, int n) {
    if (p < hcap) { sk[p] = k; sn[p] = n; } else { gk[p - hcap] = k; gn[p - hcap] = n; }
  }
};

// addtree / updtree share the sift-up (CalSurfG.f90:760-774, :876-890)
__device__ __forceinline__ void sift_up(Heap& h, int* __restrict__ nsts, int tpc, float k, int n) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const float kp = h.key(tpp);
    if (k < kp) {
      const int np = h.node(tpp);
      h.set(tpc, kp, np);
      nsts[np] = tpc;
      tpc = tpp;
      tpp = tpc >> 1;
    } else {
      break;
    }
  }
  h.set(tpc, k, n);
  nsts[n] = tpc;
}

// downtree (CalSurfG.f90:786-855)
__device__ __forceinline__ void pop_root(Heap& h, int* __restrict__ nsts) {
  if (h.ntr == 1) { h.ntr = 0; return; }
  const float k = h.key(h.ntr);
  const int n = h.node(h.ntr);
  h.ntr -= 1;
  const int ntr = h.ntr;
  int tpp = 1, tpc = 2;
  while (tpc < ntr) {
    float rd1 = h.key(tpc);
    const float rd2 = h.key(tpc + 1);
    if (rd1 > rd2) { tpc = tpc + 1; rd1 = rd2; }
    if (rd1 < k) {
      const int nc = h.node(tpc);
      h.set(tpp, rd1, nc);
      nsts[nc] = tpp;
      tpp = tpc;
      tpc = 2 * tpp;
    } else {
      tpc = ntr + 1;
    }
  }
  if (tpc == ntr) {
    const float rd1 = h.key(tpc);
    if (rd1 < k) {
      const int nc = h.node(tpc);
      h.set(tpp, rd1, nc);
      nsts[nc] = tpp;
      tpp = tpc;
    }
  }
  h.set(tpp, k, n);
  nsts[n] = tpp;
}

// One quadrant of fouds2 (CalSurfG.f90:634-723): returns trial time, valid flag through ok.
__device__ __forceinline__ float quadrant(int sj, int sj2, float tj, float tj2, int sk, int sk2, float tk,
                                          float tk2, float slown, float ri, float risti, float dnx,
                                          float dnz, bool& ok) {
  // sj/sk: status of first neighbours (-2 = outside grid); sj2/sk2: second neighbours
  int swj = -1, swk = -1;
  if (sj2 == 0 && sj == 0 && tj > tj2) swj = 0;
  if (sk2 == 0 && sk == 0 && tk > tk2) swk = 0;
  float a = 1.0f, b = 0.0f, c = 0.0f, tref = 0.0f, tdiv = 1.0f, u, v, em;
  bool sol = false;
  if (swj == 0) {
    sol = true;
    if (swk == 0) {
      u = 2.0f * ri * dnx;
      v = 2.0f * risti * dnz;
      em = 4.0f * tj - tj2 - 4.0f * tk;
      em = em + tk2;
      a = v * v + u * u;
      b = 2.0f * em * (u * u);
      c = (u * u) * (em * em - (slown * slown) * (v * v));
      tref = 4.0f * tj - tj2;
      tdiv = 3.0f;
    } else if (sk == 0) {
      u = risti * dnz;
      v = 2.0f * ri * dnx;
      em = 3.0f * tk - 4.0f * tj + tj2;
      a = v * v + 9.0f * (u * u);
      b = 6.0f * em * (u * u);
      c = (u * u) * (em * em - (slown * slown) * (v * v));
      tref = tk;
      tdiv = 1.0f;
    } else {
      u = 2.0f * ri * dnx;
      a = 1.0f;
      b = 0.0f;
      c = -(u * u) * (slown * slown);
      tref = 4.0f * tj - tj2;
      tdiv = 3.0f;
    }
  } else if (sj == 0) {
    sol = true;
    if (swk == 0) {
      u = ri * dnx;
      v = 2.0f * risti * dnz;
      em = 3.0f * tj - 4.0f * tk + tk2;
      a = v * v + 9.0f * (u * u);
      b = 6.0f * em * (u * u);
      c = (u * u) * (em * em - (v * v) * (slown * slown));
      tref = tj;
      tdiv = 1.0f;
    } else if (sk == 0) {
      u = ri * dnx;
      v = risti * dnz;
      em = tk - tj;
      a = u * u + v * v;
      b = -2.0f * (u * u) * em;
      c = (u * u) * (em * em - (v * v) * (slown * slown));
      tref = tj;
      tdiv = 1.0f;
    } else {
      a = 1.0f;
      b = 0.0f;
      c = -(slown * slown) * (ri * ri) * (dnx * dnx);
      tref = tj;
      tdiv = 1.0f;
    }
  } else {
    if (swk == 0) {
      sol = true;
      u = 2.0f * risti * dnz;
      a = 1.0f;
      b = 0.0f;
      c = -(u * u) * (slown * slown);
      tref = 4.0f * tk - tk2;
      tdiv = 3.0f;
    } else if (sk == 0) {
      sol = true;
      a = 1.0f;
      b = 0.0f;
      c = -(slown * slown) * (risti * risti) * (dnz * dnz);
      tref = tk;
      tdiv = 1.0f;
    }
  }
  ok = sol && sj != -2 && sk != -2;
  float rd1 = b * b - 4.0f * a * c;
  if (rd1 < 0.0f) rd1 = 0.0f;
  const float tdsh = (-b + sqrtf(rd1)) / (2.0f * a);
  return (tref + tdsh) / tdiv;
}

// The narrow-band march (travel's DO WHILE, CalSurfG.f90:356-456) on one grid.
// urg==1: refined grid with the early exit of :362-382.
__device__ void march(Heap& h, const int nnx, const int nnz, const int ld, const float dnx, const float dnz,
                      const float earth, const float* __restrict__ veln, const float* __restrict__ risti_tab,
                      float* __restrict__ ttn, int* __restrict__ nsts, const int urg, const bool ex_l,
                      const bool ex_r, const bool ex_t, const bool ex_b, const int lane,
                      unsigned long long& nacc, int& overflow) {
  const int nb = lane >> 3;          // neighbour 0..3: (iz,ix-1),(iz,ix+1),(iz-1,ix),(iz+1,ix)
  const int m = lane & 7;            // stencil slot around that neighbour
  // stencil slot offsets: 0:(0,-1) 1:(0,-2) 2:(0,+1) 3:(0,+2) 4:(-1,0) 5:(-2,0) 6:(+1,0) 7:(+2,0)
  const int mdx = (m < 4) ? ((m & 2) ? 1 : -1) * ((m & 1) ? 2 : 1) : 0;
  const int mdz = (m >= 4) ? ((m & 2) ? 1 : -1) * ((m & 1) ? 2 : 1) : 0;
  const int ndx = (nb == 0) ? -1 : (nb == 1 ? 1 : 0);
  const int ndz = (nb == 2) ? -1 : (nb == 3 ? 1 : 0);
  while (h.ntr > 0) {
    const int pn = h.node(1);
    const int ix = pn / ld + 1, iz = pn % ld + 1;
    if (urg == 1) {
      if ((ix == 1 && ex_l) || (ix == nnx && ex_r) || (iz == 1 && ex_t) || (iz == nnz && ex_b)) {
        nsts[pn] = 0;
        break;
      }
    }
    nsts[pn] = 0;
    ++nacc;
    pop_root(h, nsts);
    // ---- gather the 4 x 8 stencil ----
    const int cx = ix + ndx, cz = iz + ndz;            // neighbour handled by this lane group
    const bool cin = (cx >= 1 && cx <= nnx && cz >= 1 && cz <= nnz);
    const int sx = cx + mdx, sz = cz + mdz;
    int st = -2;
    float tt = 0.0f;
    if (cin && sx >= 1 && sx <= nnx && sz >= 1 && sz <= nnz) {
      const int o = (sx - 1) * ld + (sz - 1);
      st = nsts[o];
      tt = ttn[o];
    }
    int cst = -2;
    float slown = 0.0f, risti = 0.0f;
    if (cin) {
      const int o = (cx - 1) * ld + (cz - 1);
      cst = nsts[o];
      slown = 1.0f / veln[o];
      risti = risti_tab[cx - 1];
    }
    // ---- 16 quadrant solves: lane q = nb*8 + (jside*2+kside) uses slots {2*jside, 2*jside+1, 4+2*kside, 5+2*kside}
    const int base = lane & 24;
    const int js = (lane >> 1) & 1, ks = lane & 1;
    const int sj = __shfl_sync(0xffffffffu, st, base + 2 * js);
    const int sj2 = __shfl_sync(0xffffffffu, st, base + 2 * js + 1);
    const int sk = __shfl_sync(0xffffffffu, st, base + 4 + 2 * ks);
    const int sk2 = __shfl_sync(0xffffffffu, st, base + 5 + 2 * ks);
    const float tj = __shfl_sync(0xffffffffu, tt, base + 2 * js);
    const float tj2 = __shfl_sync(0xffffffffu, tt, base + 2 * js + 1);
    const float tk = __shfl_sync(0xffffffffu, tt, base + 4 + 2 * ks);
    const float tk2 = __shfl_sync(0xffffffffu, tt, base + 5 + 2 * ks);
    bool ok = false;
    float trav = quadrant(sj, sj2, tj, tj2, sk, sk2, tk, tk2, slown, earth, risti, dnx, dnz, ok);
    if (!ok || (lane & 4)) trav = __int_as_float(0x7f800000);  // +inf: lanes 4..7 of each group idle
    trav = fminf(trav, __shfl_xor_sync(0xffffffffu, trav, 1));
    trav = fminf(trav, __shfl_xor_sync(0xffffffffu, trav, 2));
    // ---- apply in the reference order: x-1, x+1, z-1, z+1 ----
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int qst = __shfl_sync(0xffffffffu, cst, q * 8);
      const float qt = __shfl_sync(0xffffffffu, trav, q * 8);
      if (qst == -2 || qst == 0) continue;
      const int qx = ix + ((q == 0) ? -1 : (q == 1 ? 1 : 0));
      const int qz = iz + ((q == 2) ? -1 : (q == 3 ? 1 : 0));
      const int o = (qx - 1) * ld + (qz - 1);
      ttn[o] = qt;
      if (qst == -1) {
        if (h.ntr + 1 >= h.hcap + h.hspill) { overflow = 1; h.ntr = 0; return; }
        h.ntr += 1;
        sift_up(h, nsts, h.ntr, qt, o);
      } else {
        // position may have moved while earlier neighbours sifted: re-read it
        sift_up(h, nsts, nsts[o], qt, o);
      }
    }
  }
}

__global__ void __launch_bounds__(32) k_fmm(FmmArgs A) {
  extern __shared__ unsigned char smem_raw[];
  const int s = blockIdx.x;
  if (s >= A.nsrc) return;
  const int lane = threadIdx.x;
  const GridC& g = A.g;
  const SrcRec sr = A.src[s];
  Heap h;
  h.sk = reinterpret_cast<float*>(smem_raw);
  h.sn = reinterpret_cast<int*>(smem_raw + sizeof(float) * A.hcap);
  h.gk = A.hspill_k + (size_t)s * A.hspill;
  h.gn = A.hspill_n + (size_t)s * A.hspill;
  h.hcap = A.hcap;
  h.hspill = A.hspill;
  h.ntr = 0;
  unsigned long long nacc = 0;
  int overflow = 0;

  float* veln_r = A.veln_r + (size_t)s * REF_N;
  float* ttn_r = A.ttn_r + (size_t)s * REF_N;
  int* nsts_r = A.nsts_r + (size_t)s * REF_N;
  const size_t ncoarse = (size_t)g.nnx * g.nnz;
  float* ttn_c = A.ttn_c + (size_t)s * ncoarse;
  int* nsts_c = A.nsts_c + (size_t)s * ncoarse;
  const float* veln_c = A.veln_c + (size_t)sr.period * ncoarse;
  const float* vv = A.velv + (size_t)sr.period * (g.nvz + 2) * (g.nvx + 2);
  const int ldv = g.nvx + 2;

  // ---- refined velocity nodes (bsplrefine, CalSurfG.f90:1559-1590) + status reset ----
  const int nrxr = g.gdx * g.sgdl, nrzr = g.gdz * g.sgdl;
  const int origx = (sr.vnl - 1) * g.sgdl + 1, origz = (sr.vnt - 1) * g.sgdl + 1;
  for (int e = lane; e < sr.nnxr * sr.nnzr; e += 32) {
    const int idm1 = e % sr.nnzr + 1, idm2 = e / sr.nnzr + 1;
    const int st1 = idm1 + origz - 1, st2 = idm2 + origx - 1;
    int i = (st1 - 1) / nrzr + 1, k = (st1 - 1) % nrzr + 1;
    if (i > g.nvz - 1) { i = g.nvz - 1; k = nrzr + 1; }
    int j = (st2 - 1) / nrxr + 1, l = (st2 - 1) % nrxr + 1;
    if (j > g.nvx - 1) { j = g.nvx - 1; l = nrxr + 1; }
    float sum[4];
#pragma unroll
    for (int i1 = 1; i1 <= 4; ++i1) {
      float sacc = 0.0f;
#pragma unroll
      for (int j1 = 1; j1 <= 4; ++j1)
        sacc = sacc + c_ubasis[(l - 1) * 4 + (j1 - 1)] * vv[(i - 2 + i1) * ldv + (j - 2 + j1)];
      sum[i1 - 1] = c_ubasis[(k - 1) * 4 + (i1 - 1)] * sacc;
    }
    const int o = (idm2 - 1) * REF_LD + (idm1 - 1);
    veln_r[o] = sum[0] + sum[1] + sum[2] + sum[3];
    nsts_r[o] = -1;
  }
  __syncwarp();

  // ---- source cell initialisation (travel, CalSurfG.f90:324-345) ----
  {
    const int isx = sr.isx_r, isz = sr.isz_r;
    float vss[2][2];
#pragma unroll
    for (int i = 1; i <= 2; ++i)
#pragma unroll
      for (int j = 1; j <= 2; ++j) vss[i - 1][j - 1] = veln_r[(isx - 1 + i - 1) * REF_LD + (isz - 1 + j - 1)];
    const float dsx = sr.dsx_r, dsz = sr.dsz_r;
    float vsrc = 0.0f;
#pragma unroll
    for (int i = 1; i <= 2; ++i)
#pragma unroll
      for (int j = 1; j <= 2; ++j) {
        const float produ = (1.0f - fabsf(((float)(i - 1) * sr.dnxr - dsx) / sr.dnxr)) *
                            (1.0f - fabsf(((float)(j - 1) * sr.dnzr - dsz) / sr.dnzr));
        vsrc = vsrc + vss[i - 1][j - 1] * produ;
      }
#pragma unroll
    for (int i = 1; i <= 2; ++i)
#pragma unroll
      for (int j = 1; j <= 2; ++j) {
        const float ax = dsx - (float)(i - 1) * sr.dnxr;
        const float az = dsz - (float)(j - 1) * sr.dnzr;
        const float ds = sqrtf(ax * ax + az * az);
        const float t0 = 2.0f * ds / (vss[i - 1][j - 1] + vsrc);
        const int o = (isx - 1 + i - 1) * REF_LD + (isz - 1 + j - 1);
        ttn_r[o] = t0;
        h.ntr += 1;
        sift_up(h, nsts_r, h.ntr, t0, o);
      }
  }
  // ---- refined march; exit tests literal to CalSurfG.f90:366-377 (vnr/vnb vs *refined* nnx/nnz) ----
  march(h, sr.nnxr, sr.nnzr, REF_LD, sr.dnxr, sr.dnzr, g.earth, veln_r, A.risti_r + (size_t)s * REF_LD, ttn_r,
        nsts_r, 1, sr.vnl != 1, sr.vnr != sr.nnxr, sr.vnt != 1, sr.vnb != sr.nnzr, lane, nacc, overflow);
  __syncwarp();

  // ---- hand-off to the coarse grid (FwdTraveltimeCPS.f90:576-632) ----
  for (size_t e = lane; e < ncoarse; e += 32) nsts_c[e] = -1;
  __syncwarp();
  {
    const int nkx = (sr.nnxr - 1) / g.sgdl + 1, nkz = (sr.nnzr - 1) / g.sgdl + 1;
    for (int e = lane; e < nkx * nkz; e += 32) {
      const int kz = e % nkz, kx = e / nkz;
      const int k = 1 + kz * g.sgdl, l = 1 + kx * g.sgdl;
      const int idm1 = sr.vnt + kz, idm2 = sr.vnl + kx;
      const int orf = (l - 1) * REF_LD + (k - 1);
      const int oc = (idm2 - 1) * g.nnz + (idm1 - 1);
      const int v = nsts_r[orf];
      nsts_c[oc] = v;
      if (v >= 0) ttn_c[oc] = ttn_r[orf];
    }
  }
  __syncwarp();
  // alive nodes with a far neighbour become close (:615-632).  Restricted to the
  // injected box: everything outside it is far.
  for (int e = lane; e < (sr.vnr - sr.vnl + 1) * (sr.vnb - sr.vnt + 1); e += 32) {
    const int nz_b = sr.vnb - sr.vnt + 1;
    const int l = sr.vnt + e % nz_b, k = sr.vnl + e / nz_b;
    const int o = (k - 1) * g.nnz + (l - 1);
    if (nsts_c[o] == 0) {
      bool mk = false;
      if (l - 1 >= 1 && nsts_c[o - 1] == -1) mk = true;
      if (l + 1 <= g.nnz && nsts_c[o + 1] == -1) mk = true;
      if (k - 1 >= 1 && nsts_c[o - g.nnz] == -1) mk = true;
      if (k + 1 <= g.nnx && nsts_c[o + g.nnz] == -1) mk = true;
      if (mk) nsts_c[o] = 1;
    }
  }
  __syncwarp();
  // ---- heap build in scan order i=1..nnx, j=1..nnz (travel urg=2, CalSurfG.f90:311-317) ----
  h.ntr = 0;
  for (int k = sr.vnl; k <= sr.vnr && !overflow; ++k) {
    for (int l0 = sr.vnt; l0 <= sr.vnb; l0 += 32) {
      const int l = l0 + lane;
      int v = -1;
      float t = 0.0f;
      const int o = (k - 1) * g.nnz + (l - 1);
      if (l <= sr.vnb) { v = nsts_c[o]; t = ttn_c[o]; }
      unsigned msk = __ballot_sync(0xffffffffu, v > 0);
      while (msk) {
        const int b = __ffs(msk) - 1;
        msk &= msk - 1;
        const float tb = __shfl_sync(0xffffffffu, t, b);
        const int ob = (k - 1) * g.nnz + (l0 + b - 1);
        if (h.ntr + 1 >= h.hcap + h.hspill) { overflow = 1; break; }
        h.ntr += 1;
        sift_up(h, nsts_c, h.ntr, tb, ob);
      }
    }
  }
  __syncwarp();
  if (!overflow)
    march(h, g.nnx, g.nnz, g.nnz, g.dnx, g.dnz, g.earth, veln_c, A.risti_c, ttn_c, nsts_c, 2, false, false, false,
          false, lane, nacc, overflow);
  if (lane == 0) {
    if (overflow) atomicOr(A.flags, 16);
    atomicAdd(A.n_accept, nacc);
  }
}

cudaError_t launch_dice_coarse(const GridC& g, int nper, const float* velv, float* veln, cudaStream_t st) {
  dim3 grid((g.nnx * g.nnz + 255) / 256, nper);
  k_dice_coarse<<<grid, 256, 0, st>>>(g, velv, veln);
  return cudaGetLastError();
}

cudaError_t launch_fmm(const FmmArgs& A, cudaStream_t st) {
  const size_t smem = (size_t)A.hcap * 8;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_fmm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  k_fmm<<<A.nsrc, 32, smem, st>>>(A);
  return cudaGetLastError();
}

}  // namespace dz
