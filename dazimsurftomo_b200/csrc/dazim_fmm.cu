// K0 (B-spline dicing) and K3 (eikonal solve) for sm_100a.
//
// K3 reproduces the reference's narrow-band fast-marching solve
// (module traveltime, CalSurfG.f90:234-893) step for step: the accepted value
// of a node depends on the order in which the binary heap pops nodes
// (SURVEY H1), so the heap discipline (addtree/downtree/updtree) is kept
// exactly.  The accept chain of one solve is inherently serial; the design
// therefore minimises the LATENCY of one accept step and takes its throughput
// from the (period x source) axis:
//
//  * one HALF-WARP (16 lanes) per solve, two solves per warp, persistent CTAs
//    pulling solve pairs from a queue.  16 lanes is exactly the width of one
//    accept step: 4 neighbours x 4 quadrant quadratics (fouds2), 4 x 4 stencil
//    directions, 4 x 4 per-neighbour scalars.
//  * the binary heap lives in shared memory as (key, node) pairs (one LDS.128
//    fetches both children of a sift-down level); only positions beyond the
//    shared capacity spill to a per-slot global array.
//  * status and travel time share one 32-bit word per node ("E"): alive = +t,
//    close (in the heap) = -t (sign bit), far = 0xFFFFFFFF.  A stencil gather is
//    then 32 loads per accept instead of 64, and the final coarse field IS the
//    travel-time array the ray tracer reads.
//  * the stencil gather for the four neighbours is ISSUED BEFORE the heap pop,
//    so its global-memory latency hides behind the shared-memory sift-down.
//    Heap positions of the four neighbours are loaded at the same time and
//    patched in registers while the pop / earlier sift-ups move entries, so no
//    global read-after-write sits on the critical path.
#include "dazim_dev.h"
#include "dazim_tps.h"
#include <cstdio>
#include <cstdlib>
#include <algorithm>

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
// veln: [nper][nnx*nnz] column-major (z fastest); slow = 1/veln (fouds2's slown, CalSurfG.f90:583)
__global__ void k_dice_coarse(GridC g, const float* __restrict__ velv, float* __restrict__ veln,
                              float* __restrict__ slow) {
  const int per = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nnx * g.nnz) return;
  const float sumi = dice_coarse_node(g, velv + (size_t)per * (g.nvz + 2) * (g.nvx + 2), c_cbasis, idx);
  veln[(size_t)per * g.nnx * g.nnz + idx] = sumi;
  slow[(size_t)per * g.nnx * g.nnz + idx] = 1.0f / sumi;
}

// ---------------------------------------------------------------------------
// K3

// E_FAR / E_OUT / E_SIGN, quadrant(), e_status(), nidx(), ndecode(): dazim_tps.h (shared with the thread-per-solve kernel).
// In THIS file's kernels a close node keeps its trial value under the sign bit (nsts > 0 is hpos).

// Heap entries are (key bits, node offset) pairs.  Positions 1..hcap-1 live in shared memory,
// positions >= hcap in a per-slot global array ("spill").  hpos[node] is the reference's nsts
// back pointer (heap position of a close node); every move of an entry updates it.
struct Heap {
  int2* sm;            // shared part (entry 0 unused)
  int2* gl;            // spill part
  int hcap, hspill;
  int ntr;
};
#define HKEY(e) __int_as_float((e).x)

__device__ __forceinline__ int2 hget(const Heap& h, int p) { return p < h.hcap ? h.sm[p] : h.gl[p - h.hcap]; }
__device__ __forceinline__ void hput(Heap& h, int* __restrict__ hpos, int p, int2 e) {
  if (p < h.hcap) h.sm[p] = e; else h.gl[p - h.hcap] = e;
  hpos[e.y] = p;
}
__device__ __forceinline__ int2 shfl2(unsigned hm, int2 v, int src) {
  return make_int2(__shfl_sync(hm, v.x, src, 16), __shfl_sync(hm, v.y, src, 16));
}

// addtree / updtree share the sift-up (CalSurfG.f90:760-774, :876-890).  Plain version
// (source-cell initialisation, coarse heap build, fall-back of the accept step).
// Returns the final position; *moved is set when at least one parent was moved down.
__device__ __forceinline__ int sift_up(Heap& h, int* __restrict__ hpos, int tpc, float k, int n, bool* moved) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const int2 par = hget(h, tpp);
    if (k < HKEY(par)) {
      hput(h, hpos, tpc, par);
      if (moved) *moved = true;
      tpc = tpp;
      tpp = tpc >> 1;
    } else {
      break;
    }
  }
  hput(h, hpos, tpc, make_int2(__float_as_int(k), n));
  return tpc;
}

// downtree (CalSurfG.f90:786-855).  `last` = heap[ntr] (fetched early by the caller).
// Shared levels: one LDS.128 fetches both children; the loop is kept to the fewest instructions
// (the accept chain is bound by dependent-issue latency, not by the shared-memory latency).
// Spill levels: the 14 entries of the next three levels under the hole are fetched by
// 14 lanes in ONE round trip and resolved with shuffles.
template <unsigned CM>
__device__ __forceinline__ void pop_root(Heap& h, int* __restrict__ hpos, const int2 last, const int sl,
                                         const unsigned hm_rt) {
  const unsigned hm = CM ? CM : hm_rt;   // compile-time mask: no convergence check (MATCH.ANY) per shuffle
  if (h.ntr == 1) { h.ntr = 0; return; }
  const float k = HKEY(last);
  h.ntr -= 1;
  const int ntr = h.ntr;
  int tpp = 1, tpc = 2;
  bool placed = false;
  const int lim = min(ntr, h.hcap - 1);      // tpc < lim  <=>  both children exist and live in shared memory
  while (tpc < lim) {
    const int4 pr = *reinterpret_cast<const int4*>(h.sm + tpc);
    const bool right = __int_as_float(pr.x) > __int_as_float(pr.z);
    const int2 c = right ? make_int2(pr.z, pr.w) : make_int2(pr.x, pr.y);
    tpc += right ? 1 : 0;
    if (!(HKEY(c) < k)) { placed = true; break; }
    h.sm[tpp] = c;
    hpos[c.y] = tpp;
    tpp = tpc;
    tpc = 2 * tpp;
  }
  if (!placed && tpc <= ntr) {
    if (tpc < h.hcap) {
      // tpc == ntr: a single child, in shared memory
      const int2 c = h.sm[tpc];
      if (HKEY(c) < k) {
        h.sm[tpp] = c;
        hpos[c.y] = tpp;
        tpp = tpc;
      }
    } else {
      // children live in the spill part
      const int j = (sl < 2) ? 1 : (sl < 6 ? 2 : 3);
      const int i = sl - ((1 << j) - 2);
      for (;;) {
        const int p = (tpp << j) + i;
        int2 ent = make_int2(0x7f800000, -1);
        if (sl < 14 && p <= ntr) ent = h.gl[p - h.hcap];
        int rel = 0, lvbase = 0;
        bool done = false;
#pragma unroll
        for (int lv = 1; lv <= 3; ++lv) {
          if (done || tpc > ntr) { done = true; continue; }
          const int l0 = lvbase + 2 * rel;
          const int2 c0 = shfl2(hm, ent, l0), c1 = shfl2(hm, ent, l0 + 1);
          const bool right = (tpc < ntr) && (HKEY(c0) > HKEY(c1));
          const int2 c = right ? c1 : c0;
          tpc += right ? 1 : 0;
          if (!(HKEY(c) < k)) { done = true; continue; }
          hput(h, hpos, tpp, c);
          tpp = tpc;
          tpc = 2 * tpp;
          rel = 2 * rel + (right ? 1 : 0);
          lvbase = (2 << lv) - 2;
        }
        if (done || tpc > ntr) break;
      }
    }
  }
  hput(h, hpos, tpp, last);
}

// Software prefetch of the words an accept step of node pn will gather (hint only).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int URG>
__device__ __forceinline__ void stencil_prefetch(const int pn, const int nnx, const int nnz, const int ld, const float inv_ld,
                                             const float* slow, const unsigned* E, const int* hpos, const int sl) {
  if (pn < 0) return;
  const int nb = sl >> 2, d = sl & 3;
  int ix, iz;
  ndecode<URG>(pn, ld, inv_ld, ix, iz);
  const int cx = ix + ((nb == 0) ? -1 : (nb == 1 ? 1 : 0)), cz = iz + ((nb == 2) ? -1 : (nb == 3 ? 1 : 0));
  if (cx < 0 || cx >= nnx || cz < 0 || cz >= nnz) return;
  const int co = nidx<URG>(cx, cz, ld);
  const int ddx = (d == 0) ? -1 : (d == 1 ? 1 : 0), ddz = (d == 2) ? -1 : (d == 3 ? 1 : 0);
  const int s1x = cx + ddx, s1z = cz + ddz, s2x = s1x + ddx, s2z = s1z + ddz;
  if (s1x >= 0 && s1x < nnx && s1z >= 0 && s1z < nnz) prefetch_l2(E + nidx<URG>(s1x, s1z, ld));
  if (s2x >= 0 && s2x < nnx && s2z >= 0 && s2z < nnz) prefetch_l2(E + nidx<URG>(s2x, s2z, ld));
  if (d == 0) prefetch_l2(E + co);
  else if (d == 1) prefetch_l2(hpos + co);
  else if (d == 2) prefetch_l2(slow + cx * ld + cz);
}

// The narrow-band march (travel's DO WHILE, CalSurfG.f90:356-456) on one grid,
// executed by one half-warp: sl = lane within the half, hm = its shuffle mask.
// URG==1: refined grid with the early exit of :362-382.
template <int URG, unsigned CM, bool PF>
__device__ void march(Heap& h, const int nnx, const int nnz, const int ld, const float dnx,
                      const float dnz, const float earth, const float* __restrict__ slow,
                      const float* __restrict__ risti_tab, unsigned* __restrict__ E, int* __restrict__ hpos,
                      const bool ex_l, const bool ex_r, const bool ex_t, const bool ex_b, const int sl,
                      const unsigned hm_rt, unsigned long long& nacc, int& overflow, const bool act0) {
  const unsigned hm = CM ? CM : hm_rt;
  const int nb = sl >> 2;            // neighbour 0..3: (iz,ix-1),(iz,ix+1),(iz-1,ix),(iz+1,ix)
  const int d = sl & 3;              // stencil direction of this lane: x-1, x+1, z-1, z+1
  const int ndx = (nb == 0) ? -1 : (nb == 1 ? 1 : 0);
  const int ndz = (nb == 2) ? -1 : (nb == 3 ? 1 : 0);
  const int ddx = (d == 0) ? -1 : (d == 1 ? 1 : 0);
  const int ddz = (d == 2) ? -1 : (d == 3 ? 1 : 0);
  const int base = sl & 12;
  const float inv_ld = 1.0f / (float)ld;
  // CM == 0 (two solves per warp): the call is warp-collective and both halves re-converge at the
  // vote below once per accept step, so the instruction stream they have in common is issued once.
  bool done = !act0;
  for (;;) {
    const bool act = !done && h.ntr > 0;
    if (CM == 0) { if (!__any_sync(0xffffffffu, act)) break; }
    else {
      if (!act) break;
      __syncwarp(CM);        // the lanes of the solve run the scalar heap code redundantly: re-align them every step
    }
    if (!act) continue;
    const int2 root = h.sm[1];
    const int pn = root.y;
    // the element that will be sifted down from the root (may live in the spill part: fetch it now)
    const int2 last = hget(h, h.ntr);
    if (PF) {
      // the root after this pop (first level of downtree, decided now) is almost always the next node to be
      // accepted: start pulling its stencil lines towards the L2 one full accept step ahead
      const int n1 = h.ntr - 1;
      int pred = -1;
      if (n1 == 1) pred = last.y;
      else if (n1 >= 2) {
        int2 c = h.sm[2];
        if (n1 >= 3) {
          const int2 c3 = h.sm[3];
          if (HKEY(c) > HKEY(c3)) c = c3;
        }
        pred = (HKEY(c) < HKEY(last)) ? c.y : last.y;
      }
      stencil_prefetch<URG>(pred, nnx, nnz, ld, inv_ld, slow, E, hpos, sl);
    }
    int ix, iz;                                        // 0-based grid coordinates of the node
    ndecode<URG>(pn, ld, inv_ld, ix, iz);
    // the popped node becomes alive with its trial value (= its heap key)
    E[pn] = (unsigned)root.x & ~E_SIGN;
    if (URG == 1) {
      if ((ix == 0 && ex_l) || (ix == nnx - 1 && ex_r) || (iz == 0 && ex_t) || (iz == nnz - 1 && ex_b)) {
        done = true;
        continue;
      }
    }
    ++nacc;
    // ---- issue the gather: 4 neighbours x 4 directions x (first, second) stencil node ----
    const int cx = ix + ndx, cz = iz + ndz;            // neighbour handled by this lane group
    const bool cin = (cx >= 0 && cx < nnx && cz >= 0 && cz < nnz);
    const int co = nidx<URG>(cx, cz, ld);
    unsigned e1 = E_OUT, e2 = E_OUT;
    {
      const int s1x = cx + ddx, s1z = cz + ddz, s2x = s1x + ddx, s2z = s1z + ddz;
      if (cin && s1x >= 0 && s1x < nnx && s1z >= 0 && s1z < nnz) e1 = E[nidx<URG>(s1x, s1z, ld)];
      if (cin && s2x >= 0 && s2x < nnx && s2z >= 0 && s2z < nnz) e2 = E[nidx<URG>(s2x, s2z, ld)];
    }
    unsigned cval = 0;                                 // d=0: E[c], 1: hpos[c], 2: slowness, 3: R sin(theta)
    if (cin) {
      if (d == 0) cval = E[co];
      else if (d == 1) cval = (unsigned)hpos[co];
      else if (d == 2) cval = __float_as_uint(slow[cx * ld + cz]);
      else cval = __float_as_uint(risti_tab[cx]);
    }
    // ---- pop the root while the loads are in flight ----
    pop_root<CM>(h, hpos, last, sl, hm);
    // ---- neighbour scalars ----
    const unsigned cE = __shfl_sync(hm, cval, base + 0, 16);
    const float slown = __uint_as_float(__shfl_sync(hm, cval, base + 2, 16));
    const float risti = __uint_as_float(__shfl_sync(hm, cval, base + 3, 16));
    const int cst = !cin ? -2 : (cE == E_FAR ? -1 : ((int)cE >= 0 ? 0 : 1));
    // ---- 16 quadrant solves: lane (nb, js, ks) pairs x-direction js with z-direction 2+ks ----
    const int js = (sl >> 1) & 1, ks = sl & 1;
    const unsigned ej1 = __shfl_sync(hm, e1, base + js, 16), ej2 = __shfl_sync(hm, e2, base + js, 16);
    const unsigned ek1 = __shfl_sync(hm, e1, base + 2 + ks, 16), ek2 = __shfl_sync(hm, e2, base + 2 + ks, 16);
    bool ok = false;
    float trav = quadrant(e_status(ej1), e_status(ej2), __uint_as_float(ej1), __uint_as_float(ej2), e_status(ek1),
                          e_status(ek2), __uint_as_float(ek1), __uint_as_float(ek2), slown, earth, risti, dnx, dnz, ok);
    if (!ok) trav = __int_as_float(0x7f800000);
    trav = fminf(trav, __shfl_xor_sync(hm, trav, 1, 16));
    trav = fminf(trav, __shfl_xor_sync(hm, trav, 2, 16));
    // ---- per-neighbour uniform scalars + start positions of the four sift-ups ----
    int qst[4], qo[4], spos[4];
    float qt[4];
    int nins = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      qst[q] = __shfl_sync(hm, cst, q * 4, 16);
      qt[q] = __shfl_sync(hm, trav, q * 4, 16);
      const int hl = (int)__shfl_sync(hm, cval, q * 4 + 1, 16);      // hpos[neighbour] as of before the pop
      qo[q] = nidx<URG>(ix + ((q == 0) ? -1 : (q == 1 ? 1 : 0)), iz + ((q == 2) ? -1 : (q == 3 ? 1 : 0)), ld);
      spos[q] = 0;
      if (qst[q] == -1) spos[q] = h.ntr + (++nins);
      else if (qst[q] == 1) spos[q] = hl;
    }
    if (h.ntr + nins >= h.hcap + h.hspill) { overflow = 1; h.ntr = 0; done = true; continue; }
    if (h.ntr + nins < h.hcap) {
      // ---- fast path: every heap position this step can touch lives in shared memory ----
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (qst[q] != 1) continue;
        // a position read before the pop is still right iff that heap slot holds the neighbour
        int vn = -1;
        if (spos[q] >= 1 && spos[q] <= h.ntr) vn = h.sm[spos[q]].y;
        if (vn != qo[q]) spos[q] = hpos[qo[q]];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (qst[q] == -2 || qst[q] == 0) continue;
        E[qo[q]] = __float_as_uint(qt[q]) | E_SIGN;
        if (qst[q] == -1) h.ntr += 1;
        const float k = qt[q];
        int tpc = spos[q];
        for (int tpp = tpc >> 1; tpp > 0; tpp >>= 1) {
          const int2 par = h.sm[tpp];
          if (!(k < HKEY(par))) break;
          h.sm[tpc] = par;
          hpos[par.y] = tpc;
          // a later neighbour that sits on this path moves down with its parent slot: keep its
          // position in registers (a read-back of hpos would put a global round trip on the chain)
#pragma unroll
          for (int r = q + 1; r < 4; ++r)
            if (par.y == qo[r]) spos[r] = tpc;
          tpc = tpp;
        }
        h.sm[tpc] = make_int2(__float_as_int(k), qo[q]);
        hpos[qo[q]] = tpc;
      }
      continue;
    }
    // ---- deep path.  One round trip for every spilled entry the sift-ups can need: lane (q, j)
    //      fetches the entry at spos[q] >> j (j = 0: the entry itself, to verify a close
    //      neighbour's position) ----
    int myp = 0;
    {
      const int sp = (nb == 0) ? spos[0] : (nb == 1 ? spos[1] : (nb == 2 ? spos[2] : spos[3]));
      myp = sp >> d;
    }
    int2 cent = make_int2(0, -1);
    if (myp >= h.hcap && myp <= h.ntr) cent = h.gl[myp - h.hcap];
    bool chain_ok = true;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (qst[q] != 1) continue;
      int2 v = make_int2(0, -1);
      if (spos[q] >= 1 && spos[q] <= h.ntr) v = (spos[q] < h.hcap) ? h.sm[spos[q]] : shfl2(hm, cent, q * 4);
      if (v.y != qo[q]) {                  // moved by the pop: take the fresh back pointer
        spos[q] = hpos[qo[q]];
        chain_ok = false;
      }
    }
    // ---- apply in the reference order: x-1, x+1, z-1, z+1 ----
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (qst[q] == -2 || qst[q] == 0) continue;
      E[qo[q]] = __float_as_uint(qt[q]) | E_SIGN;
      if (qst[q] == -1) h.ntr += 1;
      const float k = qt[q];
      int tpc = spos[q];
      bool moved = false;
      for (int j = 1;; ++j) {
        const int tpp = tpc >> 1;
        if (tpp == 0) break;
        int2 par;
        if (tpp < h.hcap) par = h.sm[tpp];
        else if (chain_ok && j <= 3) par = shfl2(hm, cent, q * 4 + j);
        else par = h.gl[tpp - h.hcap];
        if (!(k < HKEY(par))) break;
        hput(h, hpos, tpc, par);
        moved = true;
#pragma unroll
        for (int r = q + 1; r < 4; ++r)
          if (par.y == qo[r]) spos[r] = tpc;
        tpc = tpp;
      }
      hput(h, hpos, tpc, make_int2(__float_as_int(k), qo[q]));
      if (q < 3) {
        // later sift-ups: prefetched entries are stale if this one changed a spilled position they use
        if (moved) chain_ok = false;
        else if (tpc >= h.hcap && __ballot_sync(hm, myp == tpc && nb > q) != 0u) chain_ok = false;
      }
    }
  }
}

// refined velocity node (bsplrefine, CalSurfG.f90:1559-1590); idm1/idm2 1-based refined indices
__device__ __forceinline__ float refined_vel(const GridC& g, const SrcRec& sr, const float* __restrict__ vv,
                                             int idm1, int idm2) {
  const int ldv = g.nvx + 2;
  const int nrxr = g.gdx * g.sgdl, nrzr = g.gdz * g.sgdl;
  const int origx = (sr.vnl - 1) * g.sgdl + 1, origz = (sr.vnt - 1) * g.sgdl + 1;
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
  return sum[0] + sum[1] + sum[2] + sum[3];
}

template <int SPC>
__global__ void __launch_bounds__(32, 20) k_fmm(FmmArgs A) {
  constexpr unsigned CM = (SPC == 1) ? 0xffffu : 0u;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int half = lane >> 4, sl = lane & 15;
  const unsigned hm = CM ? CM : (0xffffu << (half * 16));
  const GridC& g = A.g;
  const size_t ncoarse = (size_t)g.nnx * g.nnz;          // slow_c / veln_c (plain column-major)
  const size_t ncf = coarse_field_size(g.nnx, g.nnz);    // E_c / hpos_c (interleaved layout)
  const int slot = blockIdx.x * 2 + half;
  Heap h;
  h.sm = reinterpret_cast<int2*>(smem_raw) + (size_t)half * A.hcap;
  h.gl = A.hspill + (size_t)slot * A.hspill_n;
  h.hcap = A.hcap;
  h.hspill = A.hspill_n;
  int* hpos_c = A.hpos_c + (size_t)slot * ncf;
  int* hpos_r = A.hpos_r + (size_t)slot * REF_N;
  float* slow_r = A.slow_r + (size_t)slot * REF_N;
  unsigned long long nacc = 0;
  int overflow = 0;

  // first pair = own CTA index, later pairs from the queue (so a grid that covers all pairs is one pair per CTA)
  for (int pair = blockIdx.x;;) {
    __syncwarp();
    if (pair < 0) {
      if (lane == 0) pair = gridDim.x + atomicAdd(A.queue, 1);
      pair = __shfl_sync(0xffffffffu, pair, 0);
    }
    if (pair * SPC >= A.nsrc) break;
    const int s = pair * SPC + half;
    const bool active = (half < SPC) && (s < A.nsrc) && !overflow;
    const int sc = min(s, A.nsrc - 1);                 // inactive half: valid addresses, no side effects
    const SrcRec sr = A.src[sc];
    unsigned* E_r = A.E_r + (size_t)sc * REF_N;
    unsigned* E_c = A.E_c + (size_t)sc * ncf;
    const float* slow_c = A.slow_c + (size_t)sr.period * ncoarse;
    const float* vv = A.velv + (size_t)sr.period * (g.nvz + 2) * (g.nvx + 2);
    if (active) {
      if (A.slot_of && sl == 0) A.slot_of[s] = slot;
      h.ntr = 0;

      // ---- refined slowness nodes (bsplrefine) + status reset ----
      for (int e = sl; e < sr.nnxr * sr.nnzr; e += 16) {
        const int idm1 = e % sr.nnzr + 1, idm2 = e / sr.nnzr + 1;
        const int o = (idm2 - 1) * REF_LD + (idm1 - 1);
        slow_r[o] = 1.0f / refined_vel(g, sr, vv, idm1, idm2);
        E_r[o] = E_FAR;
      }
      __syncwarp(hm);

      // ---- source cell initialisation (travel, CalSurfG.f90:324-345) ----
      {
        const int isx = sr.isx_r, isz = sr.isz_r;
        float vss[2][2];
#pragma unroll
        for (int i = 1; i <= 2; ++i)
#pragma unroll
          for (int j = 1; j <= 2; ++j) vss[i - 1][j - 1] = refined_vel(g, sr, vv, isz - 1 + j, isx - 1 + i);
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
            E_r[o] = __float_as_uint(t0) | E_SIGN;
            h.ntr += 1;
            sift_up(h, hpos_r, h.ntr, t0, o, nullptr);
          }
      }
    }
    // ---- refined march; exit tests literal to CalSurfG.f90:366-377 (vnr/vnb vs *refined* nnx/nnz) ----
    march<1, CM, false>(h, sr.nnxr, sr.nnzr, REF_LD, sr.dnxr, sr.dnzr, g.earth, slow_r,
                 A.risti_r + (size_t)sc * REF_LD, E_r, hpos_r, sr.vnl != 1, sr.vnr != sr.nnxr, sr.vnt != 1,
                 sr.vnb != sr.nnzr, sl, hm, nacc, overflow, active);
    if (active) {
      __syncwarp(hm);

      // ---- hand-off to the coarse grid (FwdTraveltimeCPS.f90:576-632); E_c was preset to FAR ----
      {
        const int nkx = (sr.nnxr - 1) / g.sgdl + 1, nkz = (sr.nnzr - 1) / g.sgdl + 1;
        for (int e = sl; e < nkx * nkz; e += 16) {
          const int kz = e % nkz, kx = e / nkz;
          const int orf = (kx * g.sgdl) * REF_LD + (kz * g.sgdl);
          const int oc = cidx(sr.vnl + kx - 1, sr.vnt + kz - 1, g.nnz);
          E_c[oc] = E_r[orf];
        }
      }
      __syncwarp(hm);
      // alive nodes with a far neighbour become close (:615-632).  Restricted to the
      // injected box: everything outside it is far.  Two passes (decide, then mark) keep
      // the test on the *original* alive set exactly like a scan that only turns 0 into 1.
      {
        const int nz_b = sr.vnb - sr.vnt + 1, nb_tot = (sr.vnr - sr.vnl + 1) * nz_b;
        for (int e0 = 0; e0 < nb_tot; e0 += 16) {
          const int e = e0 + sl;
          bool mk = false;
          int o = 0;
          if (e < nb_tot) {
            const int l = sr.vnt + e % nz_b, k = sr.vnl + e / nz_b;
            o = cidx(k - 1, l - 1, g.nnz);
            if ((int)E_c[o] >= 0) {
              if (l - 1 >= 1 && E_c[cidx(k - 1, l - 2, g.nnz)] == E_FAR) mk = true;
              if (l + 1 <= g.nnz && E_c[cidx(k - 1, l, g.nnz)] == E_FAR) mk = true;
              if (k - 1 >= 1 && E_c[cidx(k - 2, l - 1, g.nnz)] == E_FAR) mk = true;
              if (k + 1 <= g.nnx && E_c[cidx(k, l - 1, g.nnz)] == E_FAR) mk = true;
            }
          }
          __syncwarp(hm);
          if (mk) E_c[o] |= E_SIGN;
        }
      }
      __syncwarp(hm);
      // ---- heap build in scan order i=1..nnx, j=1..nnz (travel urg=2, CalSurfG.f90:311-317) ----
      h.ntr = 0;
      {
        for (int k = sr.vnl; k <= sr.vnr && !overflow; ++k) {
          for (int l0 = sr.vnt; l0 <= sr.vnb; l0 += 16) {
            const int l = l0 + sl;
            unsigned ev = E_FAR;
            if (l <= sr.vnb) ev = E_c[cidx(k - 1, l - 1, g.nnz)];
            unsigned msk = __ballot_sync(hm, (int)ev < 0 && ev != E_FAR) >> (half * 16);
            while (msk) {
              const int b = __ffs(msk) - 1;
              msk &= msk - 1;
              const float tb = __uint_as_float(__shfl_sync(hm, ev, b, 16) & ~E_SIGN);
              if (h.ntr + 1 >= h.hcap + h.hspill) { overflow = 1; break; }
              h.ntr += 1;
              sift_up(h, hpos_c, h.ntr, tb, cidx(k - 1, l0 + b - 1, g.nnz), nullptr);
            }
          }
        }
      }
      __syncwarp(hm);
    }
    march<2, CM, (SPC == 2)>(h, g.nnx, g.nnz, g.nnz, g.dnx, g.dnz, g.earth, slow_c, A.risti_c, E_c, hpos_c,
                 false, false, false, false, sl, hm, nacc, overflow, active && !overflow);
    pair = -1;
  }
  if (sl == 0) {
    if (overflow) atomicOr(A.flags, 16);
    if (nacc) atomicAdd(A.n_accept, nacc);
  }
}

// ---------------------------------------------------------------------------
// K3, latency mode ("duo"): one CTA of TWO warps per solve.  Used when every solve of the launch
// can be resident at once, i.e. when the run time is the length of one solve's serial accept chain.
// Warp H owns the heap (pop, verification, sift-ups); warp Q owns the stencil work of the same
// accept step (gather of the 4 neighbours' stencils, the 16 quadrant quadratics, the status
// stores).  Q's ~350 instructions run while H sifts the root down, so the serial chain of one
// accept shrinks to  pop + apply  instead of  pop + gather + quadrants + apply.  Two block
// barriers per accept hand the node id to Q and the four trial times back to H.  The arithmetic
// and the heap discipline are the same functions as in k_fmm: results are bit-identical.
#define DUO_FULL 0xffffffffu
// Producer/consumer barriers between the two warps (PTX named barriers, 64 threads):
//   DUO_X: H arrives after posting the node it is about to accept (+ the predicted next one), Q waits
//   DUO_Y: Q arrives after posting the four trial times of that node, H waits before its sift-ups
#define DUO_X 1
#define DUO_Y 2
__device__ __forceinline__ void duo_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void duo_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}

template <int URG>
__device__ void march_duo_H(Heap& h, int* comm, const int nnx, const int nnz, const int ld,
                            unsigned* __restrict__ E, int* __restrict__ hpos, const bool ex_l, const bool ex_r,
                            const bool ex_t, const bool ex_b, const int sl, unsigned long long& nacc, int& overflow) {
  const int nb = sl >> 2, d = sl & 3;
  const float inv_ld = 1.0f / (float)ld;
#ifdef DAZIM_DUO_PROF
  long long p_top = 0, p_pop = 0, p_wait = 0, p_apply = 0;
  int p_open = 0;
#endif
  for (;;) {
#ifdef DAZIM_DUO_PROF
    const long long t0 = clock64();
    if (p_open) { p_apply += t0; p_open = 0; }
#endif
    __syncwarp();            // the 32 lanes run the scalar heap code redundantly: re-align them every step
    bool stop = (h.ntr == 0) || overflow;
    int2 root = make_int2(0, 0), last = make_int2(0, 0);
    int pred = -1, cand0 = -1, cand1 = -1, cand2 = -1;
    if (!stop) {
      root = h.sm[1];
      last = hget(h, h.ntr);
      E[root.y] = (unsigned)root.x & ~E_SIGN;       // the popped node becomes alive with its heap key
      if (URG == 1) {
        const int pn = root.y;
        int ix = (int)((float)pn * inv_ld);
        int iz = pn - ix * ld;
        if (iz < 0) { ix -= 1; iz += ld; } else if (iz >= ld) { ix += 1; iz -= ld; }
        if ((ix == 0 && ex_l) || (ix == nnx - 1 && ex_r) || (iz == 0 && ex_t) || (iz == nnz - 1 && ex_b)) stop = true;
      }
      // Which node will be the root after this pop?  (first level of downtree, decided now.)  Unless one of
      // the four sift-ups puts a smaller key on top it is the next node to be accepted, and Q works on it
      // speculatively while this warp pops and sifts.
      const int n1 = h.ntr - 1;
      if (n1 == 1) pred = last.y;
      else if (n1 >= 2) {
        int2 c = h.sm[2];
        int wpos = 2;
        if (n1 >= 3) {
          const int2 c3 = h.sm[3];
          if (HKEY(c) > HKEY(c3)) { c = c3; wpos = 3; }
        }
        pred = (HKEY(c) < HKEY(last)) ? c.y : last.y;
        // candidates for the step after that (the loser of the two children and the winner's children):
        // Q only pulls their stencil lines into the L2, one step ahead of the speculation that will need them
        if (n1 >= 3) cand0 = h.sm[5 - wpos].y;
        if (2 * wpos <= n1 && 2 * wpos < h.hcap) cand1 = h.sm[2 * wpos].y;
        if (2 * wpos + 1 <= n1 && 2 * wpos + 1 < h.hcap) cand2 = h.sm[2 * wpos + 1].y;
      }
    }
    if (threadIdx.x == 0) {
      comm[0] = stop ? -1 : root.y;
      comm[2] = pred;
      comm[3] = cand0;
      comm[4] = cand1;
      comm[5] = cand2;
    }
    duo_arrive(DUO_X);
#ifdef DAZIM_DUO_PROF
    if (stop && URG == 2 && blockIdx.x == 0 && threadIdx.x == 0)
      printf("[duo prof] H: per accept cycles: top %.0f pop %.0f waitY+read %.0f apply %.0f (accepts %llu)\n",
             (double)p_top / nacc, (double)p_pop / nacc, (double)p_wait / nacc, (double)p_apply / nacc, nacc);
    const long long t1 = clock64();
    p_top += t1 - t0;
#endif
    if (stop) break;
    ++nacc;
    pop_root<DUO_FULL>(h, hpos, last, sl, DUO_FULL);
#ifdef DAZIM_DUO_PROF
    const long long t2 = clock64();
    p_pop += t2 - t1;
#endif
    duo_sync(DUO_Y);                                   // Q's results for this node are in comm[16..31]
    int qst[4], qo[4], spos[4];
    float qt[4];
    int nins = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int4 r = reinterpret_cast<const int4*>(comm + 16)[q];    // (status, hpos read earlier, trial time, offset)
      qst[q] = r.x;
      qt[q] = __int_as_float(r.z);
      qo[q] = r.w;
      spos[q] = 0;
      if (qst[q] == -1) spos[q] = h.ntr + (++nins);
      else if (qst[q] == 1) spos[q] = r.y;
    }
#ifdef DAZIM_DUO_PROF
    const long long t3 = clock64();
    p_wait += t3 - t2;
    p_apply -= t3;
    p_open = 1;
#endif
    if (h.ntr + nins >= h.hcap + h.hspill) { overflow = 1; h.ntr = 0; continue; }
    if (h.ntr + nins < h.hcap) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (qst[q] != 1) continue;
        // Q read this position while the heap was moving: it is right iff that slot holds the neighbour
        int vn = -1;
        if (spos[q] >= 1 && spos[q] <= h.ntr) vn = h.sm[spos[q]].y;
        if (vn != qo[q]) spos[q] = hpos[qo[q]];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (qst[q] == -2 || qst[q] == 0) continue;
        if (qst[q] == -1) h.ntr += 1;
        const float k = qt[q];
        int tpc = spos[q];
        for (int tpp = tpc >> 1; tpp > 0; tpp >>= 1) {
          const int2 par = h.sm[tpp];
          if (!(k < HKEY(par))) break;
          h.sm[tpc] = par;
          hpos[par.y] = tpc;
#pragma unroll
          for (int r = q + 1; r < 4; ++r)
            if (par.y == qo[r]) spos[r] = tpc;
          tpc = tpp;
        }
        h.sm[tpc] = make_int2(__float_as_int(k), qo[q]);
        hpos[qo[q]] = tpc;
      }
      continue;
    }
    // deep path (rare in this mode): same scheme as k_fmm
    int myp = 0;
    {
      const int sp = (nb == 0) ? spos[0] : (nb == 1 ? spos[1] : (nb == 2 ? spos[2] : spos[3]));
      myp = sp >> d;
    }
    int2 cent = make_int2(0, -1);
    if (myp >= h.hcap && myp <= h.ntr) cent = h.gl[myp - h.hcap];
    bool chain_ok = true;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (qst[q] != 1) continue;
      int2 v = make_int2(0, -1);
      if (spos[q] >= 1 && spos[q] <= h.ntr) v = (spos[q] < h.hcap) ? h.sm[spos[q]] : shfl2(DUO_FULL, cent, q * 4);
      if (v.y != qo[q]) {
        spos[q] = hpos[qo[q]];
        chain_ok = false;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (qst[q] == -2 || qst[q] == 0) continue;
      if (qst[q] == -1) h.ntr += 1;
      const float k = qt[q];
      int tpc = spos[q];
      bool moved = false;
      for (int j = 1;; ++j) {
        const int tpp = tpc >> 1;
        if (tpp == 0) break;
        int2 par;
        if (tpp < h.hcap) par = h.sm[tpp];
        else if (chain_ok && j <= 3) par = shfl2(DUO_FULL, cent, q * 4 + j);
        else par = h.gl[tpp - h.hcap];
        if (!(k < HKEY(par))) break;
        hput(h, hpos, tpc, par);
        moved = true;
#pragma unroll
        for (int r = q + 1; r < 4; ++r)
          if (par.y == qo[r]) spos[r] = tpc;
        tpc = tpp;
      }
      hput(h, hpos, tpc, make_int2(__float_as_int(k), qo[q]));
      if (q < 3) {
        if (moved) chain_ok = false;
        else if (tpc >= h.hcap && __ballot_sync(DUO_FULL, myp == tpc && nb > q) != 0u) chain_ok = false;
      }
    }
  }
}

// stencil work of one accept step for node pn (warp Q).  patch >= 0: the node itself is still marked
// close in E (speculative call): read it as alive with its trial value, which is what H will have
// written by the time the result is used.
struct QRes { int cst; unsigned cval; float trav; int co; };
#ifdef DAZIM_DUO_PROF
__device__ long long g_q_t[6];
#define QT(i) do { if (blockIdx.x == 0 && threadIdx.x == 32) { const long long t_ = clock64(); g_q_t[i] += t_ - tq_; tq_ = t_; } } while (0)
#else
#define QT(i)
#endif
template <int URG>
__device__ __forceinline__ QRes duo_stencil(const int pn, const int patch, const int nnx, const int nnz, const int ld,
                                            const float inv_ld, const float dnx, const float dnz, const float earth,
                                            const float* __restrict__ slow, const float* __restrict__ risti_tab,
                                            const unsigned* E, const int* hpos, const int sl) {
  const int nb = sl >> 2, d = sl & 3;
  const int ndx = (nb == 0) ? -1 : (nb == 1 ? 1 : 0);
  const int ndz = (nb == 2) ? -1 : (nb == 3 ? 1 : 0);
  const int ddx = (d == 0) ? -1 : (d == 1 ? 1 : 0);
  const int ddz = (d == 2) ? -1 : (d == 3 ? 1 : 0);
  const int base = sl & 12;
#ifdef DAZIM_DUO_PROF
  long long tq_ = clock64();
#endif
  int ix, iz;
  ndecode<URG>(pn, ld, inv_ld, ix, iz);
  const int cx = ix + ndx, cz = iz + ndz;
  const bool cin = (cx >= 0 && cx < nnx && cz >= 0 && cz < nnz);
  const int co = nidx<URG>(cx, cz, ld);
  unsigned e1 = E_OUT, e2 = E_OUT;
  {
    const int s1x = cx + ddx, s1z = cz + ddz, s2x = s1x + ddx, s2z = s1z + ddz;
    const int o1 = nidx<URG>(s1x, s1z, ld), o2 = nidx<URG>(s2x, s2z, ld);
    if (cin && s1x >= 0 && s1x < nnx && s1z >= 0 && s1z < nnz) e1 = E[o1];
    if (cin && s2x >= 0 && s2x < nnx && s2z >= 0 && s2z < nnz) e2 = E[o2];
    if (o1 == patch && e1 != E_OUT) e1 &= ~E_SIGN;
    if (o2 == patch && e2 != E_OUT) e2 &= ~E_SIGN;
  }
  unsigned cval = 0;                                 // d=0: E[c], 1: hpos[c], 2: slowness, 3: R sin(theta)
  if (cin) {
    if (d == 0) cval = E[co];
    else if (d == 1) cval = (unsigned)hpos[co];
    else if (d == 2) cval = __float_as_uint(slow[cx * ld + cz]);
    else cval = __float_as_uint(risti_tab[cx]);
  }
  QT(0);
  const unsigned cE = __shfl_sync(DUO_FULL, cval, base + 0, 16);
  QT(1);
  const float slown = __uint_as_float(__shfl_sync(DUO_FULL, cval, base + 2, 16));
  const float risti = __uint_as_float(__shfl_sync(DUO_FULL, cval, base + 3, 16));
  QRes r;
  r.cst = !cin ? -2 : (cE == E_FAR ? -1 : ((int)cE >= 0 ? 0 : 1));
  const int js = (sl >> 1) & 1, ks = sl & 1;
  const unsigned ej1 = __shfl_sync(DUO_FULL, e1, base + js, 16), ej2 = __shfl_sync(DUO_FULL, e2, base + js, 16);
  const unsigned ek1 = __shfl_sync(DUO_FULL, e1, base + 2 + ks, 16), ek2 = __shfl_sync(DUO_FULL, e2, base + 2 + ks, 16);
  QT(2);
  bool ok = false;
  float trav = quadrant(e_status(ej1), e_status(ej2), __uint_as_float(ej1), __uint_as_float(ej2), e_status(ek1),
                        e_status(ek2), __uint_as_float(ek1), __uint_as_float(ek2), slown, earth, risti, dnx, dnz, ok);
  if (!ok) trav = __int_as_float(0x7f800000);
  QT(3);
  trav = fminf(trav, __shfl_xor_sync(DUO_FULL, trav, 1, 16));
  trav = fminf(trav, __shfl_xor_sync(DUO_FULL, trav, 2, 16));
  QT(4);
  r.cval = cval;
  r.trav = trav;
  r.co = co;
  return r;
}

template <int URG>
__device__ void march_duo_Q(int* comm, const int nnx, const int nnz, const int ld, const float dnx, const float dnz,
                            const float earth, const float* __restrict__ slow, const float* __restrict__ risti_tab,
                            unsigned* E, const int* hpos, const int sl) {
  const int d = sl & 3;
  const float inv_ld = 1.0f / (float)ld;
  int spec = -2;                                       // node the registers below were computed for
#ifdef DAZIM_DUO_PROF
  long long q_hit = 0, q_miss = 0;
#endif
  QRes r;
  r.cst = -2; r.cval = 0; r.trav = 0.0f; r.co = 0;
  for (;;) {
    duo_sync(DUO_X);
    const int pn = comm[0], pred = comm[2];
#ifdef DAZIM_DUO_PROF
    if (pn < 0 && ld != REF_LD && blockIdx.x == 0 && threadIdx.x == 32)
      printf("[duo prof] Q: speculation hits %lld misses %lld; per call cycles: issue %.0f loadwait %.0f shfl %.0f quadrant %.0f reduce %.0f\n",
             q_hit, q_miss, (double)g_q_t[0] / (q_hit + q_miss), (double)g_q_t[1] / (q_hit + q_miss), (double)g_q_t[2] / (q_hit + q_miss),
             (double)g_q_t[3] / (q_hit + q_miss), (double)g_q_t[4] / (q_hit + q_miss));
#endif
    if (pn < 0) break;
#ifdef DAZIM_DUO_PROF
    if (pn != spec) ++q_miss; else ++q_hit;
#endif
    if (pn != spec)                                    // first step, or a sift-up put another node on top
      r = duo_stencil<URG>(pn, -1, nnx, nnz, ld, inv_ld, dnx, dnz, earth, slow, risti_tab, E, hpos, sl);
    // hand the four (status, heap position as read, trial time, offset) records to H; mark the
    // neighbours close with their new trial time (only this warp reads or writes E from here on)
    int out = r.cst;
    if (d == 1) out = (int)r.cval;
    else if (d == 2) out = __float_as_int(r.trav);
    else if (d == 3) out = r.co;
    if (threadIdx.x - 32 < 16) comm[16 + sl] = out;
    if (d == 0 && (r.cst == -1 || r.cst == 1) && threadIdx.x - 32 < 16) E[r.co] = __float_as_uint(r.trav) | E_SIGN;
    duo_arrive(DUO_Y);
    // pull the stencil lines of the nodes that can be accepted two steps from now towards the L2
    // (lanes 0-15 take one candidate, the mirror lanes 16-31 another)
#ifdef DAZIM_DUO_CANDPF
    stencil_prefetch<URG>(comm[3 + ((threadIdx.x >> 4) & 1)], nnx, nnz, ld, inv_ld, slow, E, hpos, sl);
    stencil_prefetch<URG>(((threadIdx.x >> 4) & 1) ? -1 : comm[5], nnx, nnz, ld, inv_ld, slow, E, hpos, sl);
#endif
    // speculate on the next node while H pops this one and sifts its neighbours
    spec = -2;
    if (pred >= 0) {
      r = duo_stencil<URG>(pred, pred, nnx, nnz, ld, inv_ld, dnx, dnz, earth, slow, risti_tab, E, hpos, sl);
      spec = pred;
    }
  }
}

template <int MINB>
__global__ void __launch_bounds__(64, MINB) k_fmm_duo(FmmArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sl = lane & 15;
  const GridC& g = A.g;
  const size_t ncoarse = (size_t)g.nnx * g.nnz;
  const size_t ncf = coarse_field_size(g.nnx, g.nnz);
  const int slot = blockIdx.x * 2;
  Heap h;
  h.sm = reinterpret_cast<int2*>(smem_raw);
  h.gl = A.hspill + (size_t)slot * A.hspill_n;
  h.hcap = A.hcap;
  h.hspill = A.hspill_n;
  h.ntr = 0;
  int* comm = reinterpret_cast<int*>(smem_raw + (size_t)A.hcap * 8);       // 32 ints
  int* hpos_c = A.hpos_c + (size_t)slot * ncf;
  int* hpos_r = A.hpos_r + (size_t)slot * REF_N;
  float* slow_r = A.slow_r + (size_t)slot * REF_N;
  unsigned long long nacc = 0;
  int overflow = 0;
  for (int s = blockIdx.x; s < A.nsrc;) {
    const SrcRec sr = A.src[s];
    if (A.slot_of && tid == 0) A.slot_of[s] = slot;
    unsigned* E_r = A.E_r + (size_t)s * REF_N;
    unsigned* E_c = A.E_c + (size_t)s * ncf;
    const float* slow_c = A.slow_c + (size_t)sr.period * ncoarse;
    const float* vv = A.velv + (size_t)sr.period * (g.nvz + 2) * (g.nvx + 2);
    // ---- refined slowness nodes (bsplrefine) + status reset: all 64 threads ----
    for (int e = tid; e < sr.nnxr * sr.nnzr; e += 64) {
      const int idm1 = e % sr.nnzr + 1, idm2 = e / sr.nnzr + 1;
      const int o = (idm2 - 1) * REF_LD + (idm1 - 1);
      slow_r[o] = 1.0f / refined_vel(g, sr, vv, idm1, idm2);
      E_r[o] = E_FAR;
    }
    __syncthreads();
    if (warp == 0) {
      // ---- source cell initialisation (travel, CalSurfG.f90:324-345) ----
      h.ntr = 0;
      const int isx = sr.isx_r, isz = sr.isz_r;
      float vss[2][2];
#pragma unroll
      for (int i = 1; i <= 2; ++i)
#pragma unroll
        for (int j = 1; j <= 2; ++j) vss[i - 1][j - 1] = refined_vel(g, sr, vv, isz - 1 + j, isx - 1 + i);
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
          E_r[o] = __float_as_uint(t0) | E_SIGN;
          h.ntr += 1;
          sift_up(h, hpos_r, h.ntr, t0, o, nullptr);
        }
      // ---- refined march; exit tests literal to CalSurfG.f90:366-377 ----
      march_duo_H<1>(h, comm, sr.nnxr, sr.nnzr, REF_LD, E_r, hpos_r, sr.vnl != 1, sr.vnr != sr.nnxr, sr.vnt != 1,
                     sr.vnb != sr.nnzr, sl, nacc, overflow);
    } else {
      march_duo_Q<1>(comm, sr.nnxr, sr.nnzr, REF_LD, sr.dnxr, sr.dnzr, g.earth, slow_r, A.risti_r + (size_t)s * REF_LD, E_r,
                  hpos_r, sl);
    }
    __syncthreads();
    // ---- hand-off to the coarse grid (FwdTraveltimeCPS.f90:576-632); E_c was preset to FAR ----
    {
      const int nkx = (sr.nnxr - 1) / g.sgdl + 1, nkz = (sr.nnzr - 1) / g.sgdl + 1;
      for (int e = tid; e < nkx * nkz; e += 64) {
        const int kz = e % nkz, kx = e / nkz;
        E_c[cidx(sr.vnl + kx - 1, sr.vnt + kz - 1, g.nnz)] = E_r[(kx * g.sgdl) * REF_LD + (kz * g.sgdl)];
      }
    }
    __syncthreads();
    {
      // alive nodes with a far neighbour become close (:615-632): decide on the original field, then mark
      const int nz_b = sr.vnb - sr.vnt + 1, nb_tot = (sr.vnr - sr.vnl + 1) * nz_b;
      for (int e0 = 0; e0 < nb_tot; e0 += 64) {
        const int e = e0 + tid;
        bool mk = false;
        int o = 0;
        if (e < nb_tot) {
          const int l = sr.vnt + e % nz_b, k = sr.vnl + e / nz_b;
          o = cidx(k - 1, l - 1, g.nnz);
          if ((int)E_c[o] >= 0) {
            if (l - 1 >= 1 && E_c[cidx(k - 1, l - 2, g.nnz)] == E_FAR) mk = true;
            if (l + 1 <= g.nnz && E_c[cidx(k - 1, l, g.nnz)] == E_FAR) mk = true;
            if (k - 1 >= 1 && E_c[cidx(k - 2, l - 1, g.nnz)] == E_FAR) mk = true;
            if (k + 1 <= g.nnx && E_c[cidx(k, l - 1, g.nnz)] == E_FAR) mk = true;
          }
        }
        __syncthreads();
        if (mk) E_c[o] |= E_SIGN;
      }
    }
    __syncthreads();
    if (warp == 0) {
      // ---- heap build in scan order i=1..nnx, j=1..nnz (travel urg=2, CalSurfG.f90:311-317) ----
      h.ntr = 0;
      for (int k = sr.vnl; k <= sr.vnr && !overflow; ++k) {
        for (int l0 = sr.vnt; l0 <= sr.vnb; l0 += 16) {
          const int l = l0 + sl;
          unsigned ev = E_FAR;
          if (l <= sr.vnb) ev = E_c[cidx(k - 1, l - 1, g.nnz)];
          unsigned msk = __ballot_sync(DUO_FULL, (int)ev < 0 && ev != E_FAR) & 0xffffu;
          while (msk) {
            const int b = __ffs(msk) - 1;
            msk &= msk - 1;
            const float tb = __uint_as_float(__shfl_sync(DUO_FULL, ev, b, 16) & ~E_SIGN);
            if (h.ntr + 1 >= h.hcap + h.hspill) { overflow = 1; break; }
            h.ntr += 1;
            sift_up(h, hpos_c, h.ntr, tb, cidx(k - 1, l0 + b - 1, g.nnz), nullptr);
          }
        }
      }
      march_duo_H<2>(h, comm, g.nnx, g.nnz, g.nnz, E_c, hpos_c, false, false, false, false, sl, nacc, overflow);
    } else {
      march_duo_Q<2>(comm, g.nnx, g.nnz, g.nnz, g.dnx, g.dnz, g.earth, slow_c, A.risti_c, E_c, hpos_c, sl);
    }
    __syncthreads();
    // next solve: first one = own CTA index, later ones from the queue
    if (tid == 0) comm[1] = gridDim.x + atomicAdd(A.queue, 1);
    __syncthreads();
    s = comm[1];
    __syncthreads();
  }
  if (tid == 0) {
    if (overflow) atomicOr(A.flags, 16);
    if (nacc) atomicAdd(A.n_accept, nacc);
  }
}

// MINB = 10: 96 registers (fastest single solve, 1 480 solves per chip); MINB = 16: 64 registers, 2 368 solves per chip
cudaError_t fmm_duo_max_ctas(int hcap, int nsm, int minb, int* nctas) {
  const size_t smem = (size_t)hcap * 8 + 128;
  const void* fn = minb >= 16 ? (const void*)k_fmm_duo<16> : (const void*)k_fmm_duo<10>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 64, smem);
  if (e != cudaSuccess) return e;
  *nctas = per_sm * nsm;
  return cudaSuccess;
}

cudaError_t launch_fmm_duo(const FmmArgs& A, int nctas, int minb, cudaStream_t st) {
  const size_t smem = (size_t)A.hcap * 8 + 128;
  const bool wide = minb >= 16;
  const void* fn = wide ? (const void*)k_fmm_duo<16> : (const void*)k_fmm_duo<10>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (wide) k_fmm_duo<16><<<nctas, 64, smem, st>>>(A);
  else k_fmm_duo<10><<<nctas, 64, smem, st>>>(A);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K3, thread per solve (dazim_tps.h): the default eikonal kernel.
// k_tps_init: refined slowness nodes (bsplrefine) + status reset, one thread per refined node.
__global__ void k_tps_init(TpsArgs A) {
  const GridC& g = A.g;
  for (int s = blockIdx.y; s < A.nsrc; s += gridDim.y) {
    const SrcRec sr = A.src[s];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= sr.nnxr * sr.nnzr) continue;
    const float* vv = A.velv + (size_t)sr.period * (g.nvz + 2) * (g.nvx + 2);
    const int idm1 = e % sr.nnzr + 1, idm2 = e / sr.nnzr + 1;
    const size_t o = (size_t)s * REF_N + (size_t)(idm2 - 1) * REF_LD + (idm1 - 1);
    A.slow_r[o] = 1.0f / refined_vel_t(g, sr, vv, c_ubasis, idm1, idm2);
    A.E_r[o] = E_FAR;
  }
}

__global__ void __launch_bounds__(32, 2) k_fmm_tps(TpsArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const GridC& g = A.g;
  const size_t ncoarse = (size_t)g.nnx * g.nnz;          // slow_c (plain column-major)
  const size_t ncf = coarse_field_size(g.nnx, g.nnz);    // E_c (interleaved layout)
  TpsState S;
  S.sm = reinterpret_cast<int2*>(smem_raw) + lane;
  S.stride = 32;
  S.hcap = A.hcap;
  S.htot = A.hcap + A.hspill_n - 2;     // two slots of slack: tps_pop_root reads sibling pairs as one int4
  S.E = nullptr;
  S.overflow = 0;
  S.prof = nullptr; S.pt0 = 0;
  tps_reset(S);
  unsigned long long nacc = 0;
  for (int s0 = blockIdx.x * 32; s0 < A.nsrc; s0 += gridDim.x * 32) {
    const int s = s0 + lane;
    const bool act = (s < A.nsrc) && !S.overflow;
    const int sc = min(s, A.nsrc - 1);                   // inactive lane: valid addresses, no side effects
    const SrcRec sr = A.src[sc];
    unsigned* E_r = A.E_r + (size_t)sc * REF_N;
    unsigned* E_c = A.E_c + (size_t)sc * ncf;
    const float* slow_r = A.slow_r + (size_t)sc * REF_N;
    const float* slow_c = A.slow_c + (size_t)sr.period * ncoarse;
    const float* vv = A.velv + (size_t)sr.period * (g.nvz + 2) * (g.nvx + 2);
    S.gl = A.hspill + (size_t)sc * A.hspill_n;
    if (act) tps_source_init(S, g, sr, vv, c_ubasis, E_r);
    __syncwarp();
    {
      const TpsGrid G = tps_grid_refined(g, sr, slow_r, A.risti_r + (size_t)sc * REF_LD, E_r);
      bool run = act;
      for (;;) {
        if (run) run = tps_step<1>(S, G, nacc);
        if (!__any_sync(0xffffffffu, run)) break;
      }
    }
    if (act && !S.overflow) {
      tps_refined_finish(S, E_r, A.hpos_r_out ? A.hpos_r_out + (size_t)sc * REF_N : nullptr);
      tps_handoff(S, g, sr, E_r, E_c);
    }
    __syncwarp();
    {
      const TpsGrid G = tps_grid_coarse(g, slow_c, A.risti_c, E_c);
      bool run = act && !S.overflow;
      for (;;) {
        if (run) run = tps_step<2>(S, G, nacc);
        if (!__any_sync(0xffffffffu, run)) break;
      }
    }
  }
  if (S.overflow) atomicOr(A.flags, 16);
  if (nacc) atomicAdd(A.n_accept, nacc);
}

// ---------------------------------------------------------------------------
// K3, cohort kernel: one CTA per LANES solves (LANES = 8, 16 or 32).  Warp 0 ("heap warp") runs, one lane per solve,
// the scalar heap code of dazim_tps.h (tps_pre, tps_pop, tps_apply); the 4 x LANES threads of the "stencil warps" each
// own ONE of the four neighbours of the node one solve is accepting: the gather of that neighbour's stencil and its
// four quadrant quadratics (tps_neighbour) run while the heap warp sifts the root down.  No shuffle and no shared
// state between solves; the two roles hand over through a few KB of shared memory and two producer/consumer named
// barriers per accept round.  Arithmetic and heap discipline are exactly those of the one-thread kernel (and of its
// host twin).  Fewer lanes per heap warp = less SIMT divergence in the heap code (every lane's solve takes its own
// path through pop / insert / update: measured 37 % of the round at 32 lanes) at the price of more warps.
#define COH_X 1      // heap warp arrives after posting the nodes being accepted, stencil threads wait
#define COH_Y 2      // stencil threads arrive after posting the neighbour records, heap warp waits
template <int NT>
__device__ __forceinline__ void coh_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }
template <int NT>
__device__ __forceinline__ void coh_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NT) : "memory");
}
// exchange area: head[lane] = int4 (node or -1, ix, iz, tself) posted by the heap warp; flag (one int, "any lane still
// running"); res[q * L + lane] = int4 (status, heap position read, trial time, offset) posted by the stencil thread of
// neighbour q.  One 16-byte shared access per record; the named barriers order the accesses (no volatile needed).
__host__ __device__ constexpr int coh_xch_ints(int L) { return 4 * L + 32 + 4 * L + 32 * L; }   // head, flag, predicted (node, key, node after), 2 x records
__host__ __device__ constexpr int coh_threads(int L, int QS) { return 32 + 4 * L * QS; }

template <int URG, int LANES, int QS>
__device__ void coh_march_heap(TpsState& S, const TpsGrid& G, int* xch, const int lane, const bool act,
                               unsigned long long& nacc, const bool prof) {
  constexpr int NT = coh_threads(LANES, QS);
  int4* head = reinterpret_cast<int4*>(xch);
  volatile int* flag = xch + 4 * LANES;
  int4* pred = reinterpret_cast<int4*>(xch + 4 * LANES + 32);
  const int4* res = reinterpret_cast<const int4*>(xch + 8 * LANES + 32);
  bool run = act;
  int ins[4] = {-1, -1, -1, -1};      // nodes the previous round inserted (were far): see the note on records computed ahead
  int cur = 0;                         // record buffer of this round
  long long c_pre = 0, c_pop = 0, c_wait = 0, c_apply = 0, c_nread = 0, t0 = 0, t1 = 0;
  unsigned long long rounds = 0;
  for (;;) {
    if (prof) t0 = clock64();
    TpsPre P;
    P.pn = -1; P.ix = 0; P.iz = 0; P.tself = 0; P.pred = -1; P.predk = 0; P.pred2 = -1; P.last = make_int2(0, 0);
    if (run) run = tps_pre<URG>(S, G, nacc, P);
    if (lane < LANES) { head[lane] = make_int4(run ? P.pn : -1, P.ix, P.iz, (int)P.tself); pred[lane] = make_int4(run ? P.pred : -1, P.predk & 0x7fffffff, run ? P.pred2 : -1, 0); }
    const bool any = __any_sync(0xffffffffu, run);
    if (lane == 0) *flag = any ? 1 : 0;
    coh_arrive<NT>(COH_X);
    if (!any) {
      // end of this march: wait until every stencil thread has READ the stop flag before the exchange area is
      // written again (the next march posts its first round right after the hand-off; found by racecheck)
      coh_sync<NT>(COH_Y);
      break;
    }
    if (prof) { t1 = clock64(); c_pre += t1 - t0; t0 = t1; }
    tps_pop<true>(S, P, run);                        // sift the root down while the stencil threads work (all lanes enter:
                                                     // the routine holds the warp's reconvergence points)
    if (prof) { t1 = clock64(); c_pop += t1 - t0; t0 = t1; }
    coh_sync<NT>(COH_Y);
    if (prof) { t1 = clock64(); c_wait += t1 - t0; t0 = t1; }
    {
      TpsNb N[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int4 r = make_int4(0, 0, 0, 0);
        if (run) r = res[(cur * 4 + q) * LANES + lane];
        N[q].qst = r.x; N[q].qid = r.y; N[q].qt = __int_as_float(r.z); N[q].co = r.w;
      }
      // Records may have been computed AHEAD (coh_march_stencil), while the previous round's updates were still being
      // applied.  What can differ from a gather made now: (1) a neighbour that round inserted was still far when it was
      // read -> it is close now: take its position from E (qid = 0 fails the slot check in tps_apply); (2) heap
      // positions: verified in tps_apply anyway.  Alive values are final when written, so the trial times are exact.
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (N[q].qst == -1 && (N[q].co == ins[0] || N[q].co == ins[1] || N[q].co == ins[2] || N[q].co == ins[3])) { N[q].qst = 1; N[q].qid = 0; }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) ins[q] = (run && N[q].qst == -1) ? N[q].co : -1;
      cur ^= 1;
      if (prof) { __syncwarp(); t1 = clock64(); c_nread += t1 - t0; t0 = t1; }
      run = tps_apply<URG, true>(S, G, N, run);
    }
    if (prof) { __syncwarp(); t1 = clock64(); c_apply += t1 - t0; ++rounds; }
  }
  if (prof && lane == 0 && rounds) {
    printf("[coh prof] lanes %d urg %d: rounds %llu, cycles per round: pre %.0f pop %.0f waitY %.0f read-records %.0f apply %.0f (heap size at end %d)\n",
           LANES, URG, rounds, (double)c_pre / rounds, (double)c_pop / rounds, (double)c_wait / rounds, (double)c_nread / rounds, (double)c_apply / rounds, S.ntr);
    if (S.prof) {
      printf("[coh prof] urg %d lane 0 split: pop{last %.0f shared %.0f spilled+place %.0f} apply{setup+issue %.0f wait+verify %.0f",
             URG, (double)S.prof[0] / rounds, (double)S.prof[1] / rounds, (double)S.prof[2] / rounds, (double)S.prof[3] / rounds,
             (double)S.prof[4] / rounds);
      printf(" neighbour set-up %.0f move %.0f place %.0f}\n", (double)S.prof[6] / rounds, (double)S.prof[7] / rounds,
             (double)S.prof[8] / rounds);
      for (int i = 0; i < 16; ++i) S.prof[i] = 0;
    }
  }
}

// Stencil thread t serves solve l = t / (4 QS), neighbour q = (t / QS) % 4, part = t % QS of that neighbour's four
// quadrant solves: the QS parts of one neighbour are adjacent lanes and reduce with shuffles.
// WORKING AHEAD: the gather + quadratics of one neighbour take ~8 000 cycles (a DRAM round trip with 8 000 fronts in
// flight, then four IEEE quadratics in a row) against ~3 000 for the heap warp's pop: done after the node is posted they
// are the critical path.  So right after delivering the records of round r the stencil threads compute the records of
// the node the heap warp PREDICTS for round r + 1 (the root after this pop, with its key), into the other record buffer,
// while the heap warp applies round r.  In round r + 1 they compare the posted (node, key) with what they worked on:
// equal -> the records are already there; different (an update of round r put another node, or a smaller key, on top:
// 172 of 991 910 rounds on S200) -> computed now, as before.  What a record computed ahead can miss is patched by the
// heap warp (coh_march_heap).  pf2: the stencil lines of the node expected AFTER the predicted one are pulled towards
// the L2 one round earlier still (hint only).
template <int URG, int LANES, int QS>
__device__ void coh_march_stencil(const TpsGrid& G, int* xch, const int l, const int q, const int part, const bool pf2) {
  constexpr int NT = coh_threads(LANES, QS);
  const int4* head = reinterpret_cast<const int4*>(xch);
  volatile int* flag = xch + 4 * LANES;
  const int4* pred = reinterpret_cast<const int4*>(xch + 4 * LANES + 32);
  int4* res = reinterpret_cast<int4*>(xch + 8 * LANES + 32);
  const float inv_ld = G.inv_ld;
  int spec_node = -1, spec_key = 0, cur = 0;
  const unsigned gm = QS == 1 ? 0u : (((1u << QS) - 1u) << ((threadIdx.x & 31) & ~(QS - 1)));   // the parts of my neighbour
  for (;;) {
    coh_sync<NT>(COH_X);
    if (!*flag) { coh_arrive<NT>(COH_Y); break; }      // acknowledge the stop flag (see coh_march_heap)
    const int4 hd = head[l];
    const int4 pd = pred[l];
    if (hd.x >= 0 && !(hd.x == spec_node && hd.w == spec_key)) {
      TpsNb R = tps_neighbour_part<URG, QS>(G, hd.y, hd.z, (unsigned)hd.w, q, part);
      if (QS >= 2) R.qt = fminf(R.qt, __shfl_xor_sync(gm, R.qt, 1));
      if (QS >= 4) R.qt = fminf(R.qt, __shfl_xor_sync(gm, R.qt, 2));
      if (part == 0) res[(cur * 4 + q) * LANES + l] = make_int4(R.qst, R.qid, __float_as_int(R.qt), R.co);
    }
    coh_arrive<NT>(COH_Y);
    cur ^= 1;
    spec_node = pd.x; spec_key = pd.y;
    if (pf2 && pd.z >= 0 && part == 0) {
      int px, pz;
      ndecode<URG>(pd.z, G.ld, inv_ld, px, pz);
      tps_neighbour_prefetch<URG>(G, px, pz, q);
    }
    if (pd.x >= 0) {
      int px, pz;
      ndecode<URG>(pd.x, G.ld, inv_ld, px, pz);
      TpsNb R = tps_neighbour_part<URG, QS>(G, px, pz, (unsigned)pd.y, q, part);
      if (QS >= 2) R.qt = fminf(R.qt, __shfl_xor_sync(gm, R.qt, 1));
      if (QS >= 4) R.qt = fminf(R.qt, __shfl_xor_sync(gm, R.qt, 2));
      if (part == 0) res[(cur * 4 + q) * LANES + l] = make_int4(R.qst, R.qid, __float_as_int(R.qt), R.co);
    }
  }
}

template <int LANES, int QS>
__global__ void __launch_bounds__(32 + 4 * LANES * QS, LANES == 32 ? 2 : (LANES == 16 ? 4 : 7)) k_fmm_coh(TpsArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // solve served by this thread: heap warp lane l < LANES; stencil thread t = tid - 32: solve t / (4 QS), neighbour
  // (t / QS) % 4, part t % QS
  const int l = warp == 0 ? lane : (tid - 32) / (4 * QS);
  const int q = warp == 0 ? 0 : ((tid - 32) / QS) % 4;
  const int part = warp == 0 ? 0 : (tid - 32) % QS;
  const GridC& g = A.g;
  const size_t ncoarse = (size_t)g.nnx * g.nnz;          // slow_c (plain column-major)
  const size_t ncf = coarse_field_size(g.nnx, g.nnz);    // E_c (interleaved layout)
  int* xch = reinterpret_cast<int*>(smem_raw + (size_t)A.hcap * LANES * 8);    // 16-byte aligned: hcap even, LANES >= 8
  TpsState S;
  S.sm = reinterpret_cast<int2*>(smem_raw) + l;
  S.stride = LANES;
  S.hcap = A.hcap;
  S.htot = A.hcap + A.hspill_n - 2;     // two slots of slack: tps_pop_root reads sibling pairs as one int4
  S.E = nullptr;
  S.overflow = 0;
  __shared__ long long s_prof[16];
  if (tid < 16) s_prof[tid] = 0;
  __syncthreads();
  S.prof = (A.prof && blockIdx.x == 0 && warp == 0 && lane == 0) ? s_prof : nullptr;
  S.pt0 = 0;
  tps_reset(S);
  unsigned long long nacc = 0;
  for (int s0 = blockIdx.x * LANES; s0 < A.nsrc; s0 += gridDim.x * LANES) {
    const int s = (warp != 0 || lane < LANES) ? s0 + l : A.nsrc;
    const int sc = min(s, A.nsrc - 1);                   // inactive lane: valid addresses, no side effects
    const SrcRec sr = A.src[sc];
    unsigned* E_r = A.E_r + (size_t)sc * REF_N;
    unsigned* E_c = A.E_c + (size_t)sc * ncf;
    const float* slow_r = A.slow_r + (size_t)sc * REF_N;
    const float* slow_c = A.slow_c + (size_t)sr.period * ncoarse;
    const TpsGrid Gr = tps_grid_refined(g, sr, slow_r, A.risti_r + (size_t)sc * REF_LD, E_r);
    const TpsGrid Gc = tps_grid_coarse(g, slow_c, A.risti_c, E_c);
    if (warp == 0) {
      const bool act = (s < A.nsrc) && !S.overflow;
      const float* vv = A.velv + (size_t)sr.period * (g.nvz + 2) * (g.nvx + 2);
      S.gl = A.hspill + (size_t)sc * A.hspill_n;
      if (act) tps_source_init(S, g, sr, vv, c_ubasis, E_r);
      coh_march_heap<1, LANES, QS>(S, Gr, xch, lane, act, nacc, A.prof && blockIdx.x == 0);
      if (act && !S.overflow) {
        tps_refined_finish(S, E_r, A.hpos_r_out ? A.hpos_r_out + (size_t)sc * REF_N : nullptr);
        tps_handoff(S, g, sr, E_r, E_c);
      }
      __syncwarp();
      coh_march_heap<2, LANES, QS>(S, Gc, xch, lane, act && !S.overflow, nacc, A.prof && blockIdx.x == 0);
    } else {
      coh_march_stencil<1, LANES, QS>(Gr, xch, l, q, part, A.pf2 != 0);
      coh_march_stencil<2, LANES, QS>(Gc, xch, l, q, part, A.pf2 != 0);
    }
  }
  if (warp == 0) {
    if (S.overflow) atomicOr(A.flags, 16);
    if (nacc) atomicAdd(A.n_accept, nacc);
  }
}

// (lanes, qs) variants built: 8 x {1, 2, 4}, 16 x {1, 2}, 32 x 1
static const void* coh_fn(int lanes, int qs) {
  if (lanes == 8) return qs == 4 ? (const void*)k_fmm_coh<8, 4> : (qs == 2 ? (const void*)k_fmm_coh<8, 2> : (const void*)k_fmm_coh<8, 1>);
  if (lanes == 16) return qs >= 2 ? (const void*)k_fmm_coh<16, 2> : (const void*)k_fmm_coh<16, 1>;
  return (const void*)k_fmm_coh<32, 1>;
}
int coh_norm_qs(int lanes, int qs) { return lanes == 8 ? (qs >= 4 ? 4 : (qs >= 2 ? 2 : 1)) : (lanes == 16 ? (qs >= 2 ? 2 : 1) : 1); }
size_t fmm_coh_smem(int hcap, int lanes) { return (size_t)hcap * lanes * 8 + (size_t)coh_xch_ints(lanes) * 4; }

// resident warps (= CTAs of 32 solves) per SM for a given shared heap capacity
// resident CTAs for a given shared heap capacity; coh = lanes per heap warp (8 / 16 / 32) or 0 for the one-thread kernel;
// qs = stencil threads per neighbour
cudaError_t fmm_tps_max_ctas(int hcap, int nsm, int coh, int qs, int* nctas) {
  const size_t smem = coh ? fmm_coh_smem(hcap, coh) : (size_t)hcap * 32 * 8;
  qs = coh ? coh_norm_qs(coh, qs) : 1;
  const void* fn = coh ? coh_fn(coh, qs) : (const void*)k_fmm_tps;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, coh ? coh_threads(coh, qs) : 32, smem);
  if (e != cudaSuccess) return e;
  *nctas = per_sm * nsm;
  return cudaSuccess;
}

cudaError_t launch_fmm_tps(const TpsArgs& A, int nctas, int coh, int qs, cudaStream_t st) {
  if (A.nsrc <= 0) return cudaSuccess;
  dim3 gi((REF_N + 255) / 256, (unsigned)std::min(A.nsrc, 65535));
  k_tps_init<<<gi, 256, 0, st>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t smem = coh ? fmm_coh_smem(A.hcap, coh) : (size_t)A.hcap * 32 * 8;
  qs = coh ? coh_norm_qs(coh, qs) : 1;
  e = cudaFuncSetAttribute(coh ? coh_fn(coh, qs) : (const void*)k_fmm_tps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (coh == 8 && qs == 4) k_fmm_coh<8, 4><<<nctas, coh_threads(8, 4), smem, st>>>(A);
  else if (coh == 8 && qs == 2) k_fmm_coh<8, 2><<<nctas, coh_threads(8, 2), smem, st>>>(A);
  else if (coh == 8) k_fmm_coh<8, 1><<<nctas, coh_threads(8, 1), smem, st>>>(A);
  else if (coh == 16 && qs == 2) k_fmm_coh<16, 2><<<nctas, coh_threads(16, 2), smem, st>>>(A);
  else if (coh == 16) k_fmm_coh<16, 1><<<nctas, coh_threads(16, 1), smem, st>>>(A);
  else if (coh == 32) k_fmm_coh<32, 1><<<nctas, coh_threads(32, 1), smem, st>>>(A);
  else k_fmm_tps<<<nctas, 32, smem, st>>>(A);
  return cudaGetLastError();
}

// test seam: (E, hpos) -> the reference's (ttn, nsts) pair
__global__ void k_decode_status(const unsigned* __restrict__ E, const int* __restrict__ hpos, size_t n, int nnz_tiled,
                                float* __restrict__ ttn, int* __restrict__ nsts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // output index: plain column-major
  if (i >= n) return;
  // nnz_tiled > 0: the input is a coarse field in the interleaved layout
  const size_t j = nnz_tiled > 0 ? (size_t)cidx((int)(i / nnz_tiled), (int)(i % nnz_tiled), nnz_tiled) : i;
  const unsigned e = E[j];
  int st;
  float t;
  if (e == E_FAR) { st = -1; t = 0.0f; }
  else if ((int)e >= 0) { st = 0; t = __uint_as_float(e); }
  else { st = hpos ? hpos[j] : 1; t = __uint_as_float(e & ~E_SIGN); }
  if (ttn) ttn[i] = t;
  if (nsts) nsts[i] = st;
}

cudaError_t launch_decode_status(const unsigned* E, const int* hpos, size_t n, int nnz_tiled, float* ttn, int* nsts,
                                 cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_decode_status<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, hpos, n, nnz_tiled, ttn, nsts);
  return cudaGetLastError();
}

cudaError_t launch_dice_coarse(const GridC& g, int nper, const float* velv, float* veln, float* slow, cudaStream_t st) {
  dim3 grid((g.nnx * g.nnz + 255) / 256, nper);
  k_dice_coarse<<<grid, 256, 0, st>>>(g, velv, veln, slow);
  return cudaGetLastError();
}

// number of CTAs (= solve pairs in flight) the device can hold for a given shared heap capacity
cudaError_t fmm_max_ctas(int hcap, int spc, int nsm, int* nctas) {
  const size_t smem = (size_t)hcap * 8 * spc;
  cudaError_t e = cudaFuncSetAttribute(spc == 1 ? k_fmm<1> : k_fmm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spc == 1 ? k_fmm<1> : k_fmm<2>, 32, smem);
  if (e != cudaSuccess) return e;
  *nctas = per_sm * nsm;
  return cudaSuccess;
}

cudaError_t launch_fmm(const FmmArgs& A, int nctas, cudaStream_t st) {
  const size_t smem = (size_t)A.hcap * 8 * A.spc;
  cudaError_t e = cudaFuncSetAttribute(A.spc == 1 ? k_fmm<1> : k_fmm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (A.spc == 1) k_fmm<1><<<nctas, 32, smem, st>>>(A);
  else k_fmm<2><<<nctas, 32, smem, st>>>(A);
  return cudaGetLastError();
}

}  // namespace dz
