// K3 "thread per solve" (TPS): the reference's narrow-band fast-marching solve (module traveltime,
// CalSurfG.f90:234-893) with ONE THREAD PER (period, source) SOLVE, 32 solves per warp.
//
// Why: the accept chain of one solve is strictly serial (the result depends on the binary heap's pop
// order, ties included), so the chip's parallelism has to come from the (period x source) axis.  The half-warp and
// two-warp kernels of dazim_fmm.cu spend ~600 warp instructions per accept, almost all of them scalar heap code
// executed redundantly by 16-32 lanes, and can keep at most 1 480 - 5 920 solves resident.  Here every lane runs the
// scalar code of its OWN solve: ~1 500 thread instructions per accept = ~47 warp instructions per accept, 9 472 solves
// resident on 148 SMs (2 warps per SM), the 16 quadrant quadratics of an accept step are independent instructions
// of one thread (ILP instead of lanes), and no shuffle, vote-free inner loops, no barrier and no shared state exist
// between solves.
//
// Data layout (per solve):
//   heap      (key bits, node offset) pairs; positions 1..hcap-1 in shared memory, INTERLEAVED across the lanes of
//             the heap warp (entry p of lane l at sm[p * 32 + l]: every lane always hits its own two banks),
//             positions >= hcap in a per-solve global array.
//   E         one 32-bit word per node: alive = +t, far = 0xFFFFFFFF, close = 0x80000000 | heap position.  A close
//             node's trial time lives only in the heap (fouds2 never reads it, CalSurfG.f90:592-629), so the word is
//             free to hold the reference's nsts back pointer: the gather that tells a stencil warp "this neighbour
//             is close" hands the heap warp its position in the same word, and the separate 4 MB hpos field per
//             resident solve of the round-1 kernels is gone.  (A first version kept positions in a per-solve
//             id -> position table; measured on the B200 its id allocator and the extra dependent load cost the heap
//             warp ~3 000 cycles per accept: profiles/r2_k3_cohort_cycle_split.txt.)
//
// The code below is __host__ __device__: dazim_fmm.cu instantiates it in the kernel k_fmm_tps AND in a host twin
// (dazim_debug_fmm_host_twin, test seam only) so that the exact logic is checked against the oracle on machines
// without a GPU.  No product entry point calls the host twin.
#pragma once
#include "dazim_dev.h"
#include <math.h>

#if defined(__CUDACC__)
#define TPS_HD __host__ __device__ __forceinline__
#else
#define TPS_HD inline
#endif

namespace dz {

#define E_FAR 0xFFFFFFFFu     // never touched (nsts = -1)
#define E_OUT 0xFFFFFFFEu     // outside the grid (register-only sentinel)
#define E_SIGN 0x80000000u    // close (nsts > 0)

TPS_HD int e_status(unsigned e) { return e == E_OUT ? -2 : ((int)e >= 0 ? 0 : 1); }

// One quadrant of fouds2 (CalSurfG.f90:634-723): returns trial time, valid flag through ok.
// sj/sk: status of the first neighbours (0 alive, -2 = outside grid); sj2/sk2: second neighbours.
// The reference selects one of nine stencils (first/second order in x and z, or one-sided) by
// nested IFs.  The nest is evaluated here WITHOUT branches: every case's operands are chosen with selects and one
// common expression tree is evaluated.  Each case keeps the reference's operation order (x2 and the commutations
// used are exact in IEEE arithmetic), so the result is bit-identical to the branchy form:
//   A swj&swk  : u=2Rdx v=2Rsdz em=4tj-tj2-4tk+tk2  a=v2+u2   b=(2em)u2   c=u2(em2-s2v2) tref=4tj-tj2 /3
//   B swj&k1   : u=Rsdz v=2Rdx  em=3tk-4tj+tj2      a=v2+9u2  b=(6em)u2   c=u2(em2-s2v2) tref=tk
//   C swj only : u=2Rdx                             a=1 b=0   c=-(u2 s2)                tref=4tj-tj2 /3
//   D j1&swk   : u=Rdx  v=2Rsdz em=3tj-4tk+tk2      a=v2+9u2  b=(6em)u2   c=u2(em2-v2s2) tref=tj
//   E j1&k1    : u=Rdx  v=Rsdz  em=tk-tj            a=u2+v2   b=(-2u2)em  c=u2(em2-v2s2) tref=tj
//   F j1 only  :                                    a=1 b=0   c=-(s2 R2)dx2             tref=tj
//   G swk only : u=2Rsdz                            a=1 b=0   c=-(u2 s2)                tref=4tk-tk2 /3
//   H k1 only  :                                    a=1 b=0   c=-(s2 Rs2)dz2            tref=tk
TPS_HD float quadrant(int sj, int sj2, float tj, float tj2, int sk, int sk2, float tk, float tk2, float slown, float ri,
                      float risti, float dnx, float dnz, bool& ok) {
  // A stencil word that is not alive is not a time: it is NaN (far, outside) or a tiny negative number (close: heap
  // position under the sign bit).  Every case below uses only ALIVE times, so replacing the others by 0 changes no
  // result -- but it keeps NaN / denormal operands out of the common expression tree, whose IEEE square root and two
  // divisions otherwise leave their fast paths for the ~60-instruction special-operand subroutines in every quadrant
  // that has no valid stencil (most of them; measured: 14 % of the cohort kernel's issue samples sat in those routines).
  tj = (sj == 0) ? tj : 0.0f;
  tj2 = (sj2 == 0) ? tj2 : 0.0f;
  tk = (sk == 0) ? tk : 0.0f;
  tk2 = (sk2 == 0) ? tk2 : 0.0f;
  const bool j1 = (sj == 0), k1 = (sk == 0);
  const bool swj = j1 && (sj2 == 0) && (tj > tj2);
  const bool swk = k1 && (sk2 == 0) && (tk > tk2);
  const bool cA = swj && swk, cB = swj && !swk && k1, cC = swj && !k1;
  const bool cD = !swj && j1 && swk, cE = !swj && j1 && !swk && k1, cF = !swj && j1 && !k1;
  const bool cG = !j1 && swk, cH = !j1 && !swk && k1;
  ok = (cA || cB || cC || cD || cE || cF || cG || cH) && sj != -2 && sk != -2;
  const float ux1 = ri * dnx, ux2 = 2.0f * ri * dnx;          // 2.0f*ri*dnx == 2*(ri*dnx) exactly
  const float vz1 = risti * dnz, vz2 = 2.0f * risti * dnz;
  const float s2 = slown * slown;
  const float u = (cA || cC) ? ux2 : (cB ? vz1 : (cG ? vz2 : ux1));
  const float v = (cA || cD) ? vz2 : (cB ? ux2 : vz1);
  const float fj = 4.0f * tj - tj2, fk = 4.0f * tk - tk2;      // second-order one-sided values
  float emA = fj - 4.0f * tk;
  emA = emA + tk2;
  const float emB = 3.0f * tk - 4.0f * tj + tj2;
  const float emD = 3.0f * tj - 4.0f * tk + tk2;
  const float emE = tk - tj;
  const float em = cA ? emA : (cB ? emB : (cD ? emD : emE));
  const float uu = u * u, vv = v * v;
  float a = 1.0f;
  if (cA || cE) a = vv + uu;
  if (cB || cD) a = vv + 9.0f * uu;
  float b = 0.0f;
  if (cA) b = 2.0f * em * uu;
  if (cB || cD) b = 6.0f * em * uu;
  if (cE) b = -2.0f * uu * em;
  float c = uu * (em * em - s2 * vv);                          // A, B, D, E
  if (cC || cG) c = -uu * s2;
  if (cF) c = -s2 * (ri * ri) * (dnx * dnx);
  if (cH) c = -s2 * (risti * risti) * (dnz * dnz);
  const float tref = (cA || cC) ? fj : (cG ? fk : ((cB || cH) ? tk : tj));
  const float tdiv = (cA || cC || cG) ? 3.0f : 1.0f;
  float rd1 = b * b - 4.0f * a * c;
  if (rd1 < 0.0f) rd1 = 0.0f;
  const float tdsh = (-b + sqrtf(rd1)) / (2.0f * a);
  return (tref + tdsh) / tdiv;
}

// Node offset of (ix, iz) in a per-solve field: the refined box (URG == 1) is plain column-major with
// leading dimension ld; the coarse grid (URG == 2, ld == nnz) uses the interleaved layout of dazim_dev.h.
template <int URG>
TPS_HD int nidx(int ix, int iz, int ld) { return URG == 2 ? cidx(ix, iz, ld) : ix * ld + iz; }
// inverse: float quotient, exact after one correction (offsets < 2^30, ld <= 32767)
template <int URG>
TPS_HD void ndecode(int o, int ld, float inv_ld, int& ix, int& iz) {
  const int b = (URG == 2) ? (o >> 3) : o;
  int q = (int)((float)b * inv_ld);
  int r = b - q * ld;
  if (r < 0) { q -= 1; r += ld; } else if (r >= ld) { q += 1; r -= ld; }
  ix = (URG == 2) ? (q * 8 + (o & 7)) : q;
  iz = r;
}

// coarse dicing of one propagation node (gridder, CalSurfG.f90:1450-1488): idx = (stx-1)*nnz + (stz-1);
// vv = control values of the period, velv(i,j) at i*(nvx+2)+j; cb = coarse basis table [6][4]
TPS_HD float dice_coarse_node(const GridC& g, const float* vv, const float* cb, int idx) {
  const int stz = idx % g.nnz + 1, stx = idx / g.nnz + 1;
  // cell (i,j) and local index (l,m): the last cell also owns its far edge
  int i = (stz - 1) / g.gdz + 1, l = (stz - 1) % g.gdz + 1;
  if (i > g.nvz - 1) { i = g.nvz - 1; l = g.gdz + 1; }
  int j = (stx - 1) / g.gdx + 1, m = (stx - 1) % g.gdx + 1;
  if (j > g.nvx - 1) { j = g.nvx - 1; m = g.gdx + 1; }
  const int ldv = g.nvx + 2;
  float sumi = 0.0f;
  for (int i1 = 1; i1 <= 4; ++i1) {
    float sumj = 0.0f;
    for (int j1 = 1; j1 <= 4; ++j1)
      sumj = sumj + cb[(m - 1) * 4 + (j1 - 1)] * vv[(i - 2 + i1) * ldv + (j - 2 + j1)];
    sumi = sumi + cb[(l - 1) * 4 + (i1 - 1)] * sumj;
  }
  return sumi;
}

// refined velocity node (bsplrefine, CalSurfG.f90:1559-1590); idm1/idm2 1-based refined indices;
// ub = refined B-spline basis table [41][4] (constant memory on the device)
TPS_HD float refined_vel_t(const GridC& g, const SrcRec& sr, const float* vv, const float* ub, int idm1, int idm2) {
  const int ldv = g.nvx + 2;
  const int nrxr = g.gdx * g.sgdl, nrzr = g.gdz * g.sgdl;
  const int origx = (sr.vnl - 1) * g.sgdl + 1, origz = (sr.vnt - 1) * g.sgdl + 1;
  const int st1 = idm1 + origz - 1, st2 = idm2 + origx - 1;
  int i = (st1 - 1) / nrzr + 1, k = (st1 - 1) % nrzr + 1;
  if (i > g.nvz - 1) { i = g.nvz - 1; k = nrzr + 1; }
  int j = (st2 - 1) / nrxr + 1, l = (st2 - 1) % nrxr + 1;
  if (j > g.nvx - 1) { j = g.nvx - 1; l = nrxr + 1; }
  float sum[4];
  for (int i1 = 1; i1 <= 4; ++i1) {
    float sacc = 0.0f;
    for (int j1 = 1; j1 <= 4; ++j1)
      sacc = sacc + ub[(l - 1) * 4 + (j1 - 1)] * vv[(i - 2 + i1) * ldv + (j - 2 + j1)];
    sum[i1 - 1] = ub[(k - 1) * 4 + (i1 - 1)] * sacc;
  }
  return sum[0] + sum[1] + sum[2] + sum[3];
}

// ---------------------------------------------------------------------------------------------------------------
struct TpsArgs {
  GridC g;
  const SrcRec* src;          // [nsrc]
  int nsrc;
  const float* velv;          // [nper][(nvz+2)(nvx+2)]
  const float* slow_c;        // [nper][nnx*nnz]  1/veln (fouds2's slown), plain column-major
  const float* risti_c;       // [nnx]   earth*sin(gox+(ix-1)*dnx), host computed
  const float* risti_r;       // [nsrc][REF_LD]
  unsigned* E_c;              // [nsrc][coarse_field_size]  (preset to E_FAR by the host)
  unsigned* E_r;              // [nsrc][REF_N]              (preset + slow_r by k_tps_init)
  float* slow_r;              // [nsrc][REF_N]   refined slowness
  int2* hspill;               // [nsrc][hspill_n] heap entries beyond the shared capacity
  int hspill_n;
  int hcap;                   // heap positions 1..hcap-1 live in shared memory
  int* hpos_r_out;            // optional test seam [nsrc][REF_N]: heap slots of the close nodes after the refined march
  int* flags;                 // bit4 (16): heap / id overflow
  unsigned long long* n_accept;
  int pf2;                    // cohort kernel: prefetch the stencil of the node after the predicted one (DAZIM_COH_PF2, default on)
  int prof;                   // DAZIM_COH_PROF=1: lane 0 of CTA 0 prints its cycle split per accept (cohort kernel)
};

struct TpsState {
  int2* sm;                   // this solve's shared heap part; position p at sm[p * stride]
  int stride;
  int2* gl;                   // spill part: position p at gl[p - hcap]
  int hcap, htot;             // htot = hcap + hspill_n - 2
  int ntr;
  int2 last_e;                // copy of heap[last_p], kept by tps_apply (valid while last_p == ntr): the next pop's `last`
  int last_p;
  unsigned* E;                // status / time / back-pointer words of the grid being marched
  int overflow;
  int stopped_at_root;        // refined march left through the exit test: the root is alive and stays in the heap
  long long* prof;            // cycle accumulators of the profiled lane (DAZIM_COH_PROF), else nullptr
  long long pt0;
};

// Reconvergence point for the lanes of a heap warp (cohort kernel).  The heap code is full of data-dependent branches and
// early exits; without explicit convergence points the lanes of a warp, once diverged, tend to run the REST of the
// routine one group after the other (measured: one lane's own path through tps_apply is 3 500 cycles, the warp needed
// 9 200).  Every lane of the warp must reach these points: the routines below take a `run` flag instead of being
// called conditionally.  No-op on the host and in the one-thread kernel (SYNC = false).
#if defined(__CUDA_ARCH__)
#define TPS_SYNCWARP(on) do { if (on) __syncwarp(); } while (0)
#else
#define TPS_SYNCWARP(on) do { } while (0)
#endif

#if defined(__CUDA_ARCH__)
#define TPS_TICK0(S) do { if ((S).prof) (S).pt0 = clock64(); } while (0)
#define TPS_TICK(S, i) do { if ((S).prof) { const long long t_ = clock64(); (S).prof[i] += t_ - (S).pt0; (S).pt0 = t_; } } while (0)
#else
#define TPS_TICK0(S) do { } while (0)
#define TPS_TICK(S, i) do { } while (0)
#endif

#define TKEY(e) tps_as_float((e).x)
TPS_HD float tps_as_float(int b) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(b);
#else
  union { int i; float f; } u; u.i = b; return u.f;
#endif
}
TPS_HD int tps_as_int(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  union { int i; float f; } u; u.f = f; return u.i;
#endif
}
TPS_HD float tps_inf() { return tps_as_float(0x7f800000); }

TPS_HD int2 tps_hget(const TpsState& S, int p) { return p < S.hcap ? S.sm[(size_t)p * S.stride] : S.gl[p - S.hcap]; }
// every placement of an entry updates the node's back pointer (the reference's nsts(node) = heap position)
TPS_HD void tps_hput(TpsState& S, int p, int2 e) {
  if (p < S.hcap) S.sm[(size_t)p * S.stride] = e; else S.gl[p - S.hcap] = e;
  S.E[e.y] = E_SIGN | (unsigned)p;
}
TPS_HD void tps_reset(TpsState& S) { S.ntr = 0; S.stopped_at_root = 0; S.last_p = -1; S.last_e = make_int2(0, 0); }

// addtree / updtree share the sift-up (CalSurfG.f90:760-774, :876-890).
// plain version (source-cell initialisation, coarse heap build)
TPS_HD void tps_sift_up_plain(TpsState& S, int tpc, float k, int node) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const int2 par = tps_hget(S, tpp);
    if (!(k < TKEY(par))) break;
    tps_hput(S, tpc, par);
    tpc = tpp;
    tpp = tpc >> 1;
  }
  tps_hput(S, tpc, make_int2(tps_as_int(k), node));
}

// downtree (CalSurfG.f90:786-855).  `last` = heap[ntr] (fetched early by the caller).
// Shared levels one by one (two loads from the lane's own banks).  Spilled levels: the 14 entries of the next THREE
// levels under the hole are contiguous per level (2p..2p+1, 4p..4p+3, 8p..8p+7) and are fetched with seven 16-byte
// loads in ONE round trip, then the path is resolved from registers with selects -- a dependent global load per level
// was the largest single term of the accept chain.  hcap and the spill stride are even, so a sibling pair never
// straddles the shared / spilled boundary and every pair is one aligned int4.
template <bool SYNC>
TPS_HD void tps_pop_root(TpsState& S, const int2 last, const bool run) {
  const bool single = (S.ntr == 1);
  const bool go = run && !single;
  if (run && single) S.ntr = 0;
  TPS_TICK0(S);
  const float k = TKEY(last);
  if (go) S.ntr -= 1;
  const int ntr = S.ntr;
  int tpp = 1, tpc = 2;
  TPS_TICK(S, 0);      // 0: wait for `last`
  // shared levels: both children exist and live in shared memory
  const int lim = ntr < S.hcap - 1 ? ntr : S.hcap - 1;
  bool placed = !go;
  while (!placed && tpc < lim) {
    const int2 c0 = S.sm[(size_t)tpc * S.stride], c1 = S.sm[(size_t)(tpc + 1) * S.stride];
    const bool right = TKEY(c0) > TKEY(c1);
    const int2 c = right ? c1 : c0;
    tpc += right ? 1 : 0;
    if (!(TKEY(c) < k)) { placed = true; break; }
    S.sm[(size_t)tpp * S.stride] = c;
    S.E[c.y] = E_SIGN | (unsigned)tpp;
    tpp = tpc;
    tpc = 2 * tpp;
  }
  TPS_SYNCWARP(SYNC);
  if (go && !placed && tpc <= ntr && tpc < S.hcap) {
    // tpc == ntr: a single child, in shared memory
    const int2 c = S.sm[(size_t)tpc * S.stride];
    if (TKEY(c) < k) { tps_hput(S, tpp, c); tpp = tpc; }
    placed = true;
  }
  TPS_TICK(S, 1);      // 1: shared levels
  // spilled levels, three at a time
  while (go && !placed && tpc <= ntr) {
    const int4* g4 = reinterpret_cast<const int4*>(S.gl);
    const int b1 = (2 * tpp - S.hcap) >> 1, b2 = (4 * tpp - S.hcap) >> 1, b3 = (8 * tpp - S.hcap) >> 1;   // int4 indices
    const int4 z = make_int4(0x7f800000, 0, 0x7f800000, 0);
    int4 l1 = z, l2a = z, l2b = z, l3a = z, l3b = z, l3c = z, l3d = z;
    l1 = g4[b1];                                   // 2p <= ntr here; the slot after ntr exists (spill slack)
    if (4 * tpp <= ntr) l2a = g4[b2];
    if (4 * tpp + 2 <= ntr) l2b = g4[b2 + 1];
    if (8 * tpp <= ntr) l3a = g4[b3];
    if (8 * tpp + 2 <= ntr) l3b = g4[b3 + 1];
    if (8 * tpp + 4 <= ntr) l3c = g4[b3 + 2];
    if (8 * tpp + 6 <= ntr) l3d = g4[b3 + 3];
    // level 1
    {
      const int2 c0 = make_int2(l1.x, l1.y), c1 = make_int2(l1.z, l1.w);
      const bool right = (tpc < ntr) && (TKEY(c0) > TKEY(c1));
      const int2 c = right ? c1 : c0;
      tpc += right ? 1 : 0;
      if (!(TKEY(c) < k)) break;
      tps_hput(S, tpp, c);
      tpp = tpc; tpc = 2 * tpp;
      if (tpc > ntr) break;
      // level 2
      const int4 m = right ? l2b : l2a;
      const int2 d0 = make_int2(m.x, m.y), d1 = make_int2(m.z, m.w);
      const bool right2 = (tpc < ntr) && (TKEY(d0) > TKEY(d1));
      const int2 d = right2 ? d1 : d0;
      tpc += right2 ? 1 : 0;
      if (!(TKEY(d) < k)) break;
      tps_hput(S, tpp, d);
      tpp = tpc; tpc = 2 * tpp;
      if (tpc > ntr) break;
      // level 3
      const int4 n = right ? (right2 ? l3d : l3c) : (right2 ? l3b : l3a);
      const int2 f0 = make_int2(n.x, n.y), f1 = make_int2(n.z, n.w);
      const bool right3 = (tpc < ntr) && (TKEY(f0) > TKEY(f1));
      const int2 f = right3 ? f1 : f0;
      tpc += right3 ? 1 : 0;
      if (!(TKEY(f) < k)) break;
      tps_hput(S, tpp, f);
      tpp = tpc; tpc = 2 * tpp;
    }
  }
  if (go) tps_hput(S, tpp, last);
  TPS_SYNCWARP(SYNC);
  TPS_TICK(S, 2);      // 2: spilled levels + placement
}

// Grid view of one march
struct TpsGrid {
  int nnx, nnz, ld;
  float inv_ld;               // 1 / ld (ndecode)
  float dnx, dnz, earth;
  const float* slow;          // [ix * ld + iz] plain
  const float* risti_tab;     // [ix]
  unsigned* E;
  bool ex_l, ex_r, ex_t, ex_b;
};

// One accept step of travel's DO WHILE (CalSurfG.f90:356-456) in three pieces, so that the same code serves the
// one-thread-per-solve kernel (pre + 4 x neighbour + post in one thread), the cohort kernel (pre/post on the heap
// warp, one neighbour per stencil warp) and the host twin.
struct TpsPre { int pn, ix, iz, pred, predk, pred2; unsigned tself; int2 last; };   // pred / predk: node (and its key) that will be on top after this pop; pred2: a guess for the one after (hints)
struct TpsNb { int qst, qid, co; float qt; };     // neighbour status (-2 outside, -1 far, 0 alive, 1 close), heap position read, offset, trial

// (1) the node on top of the heap becomes alive.  Returns false when the march is over (heap empty, overflow, or the
//     refined march reached the edge of the source box: CalSurfG.f90:362-382).
template <int URG>
TPS_HD bool tps_pre(TpsState& S, const TpsGrid& G, unsigned long long& nacc, TpsPre& P) {
  if (S.ntr <= 0 || S.overflow) return false;
  const int2 root = tps_hget(S, 1);
  P.pn = root.y;
  P.last = (S.last_p == S.ntr) ? S.last_e : tps_hget(S, S.ntr);     // usually kept in registers by tps_apply (a spilled slot is a global load)
  ndecode<URG>(P.pn, G.ld, G.inv_ld, P.ix, P.iz);
  P.tself = (unsigned)root.x & ~E_SIGN;
  G.E[P.pn] = P.tself;                             // the popped node becomes alive with its trial value (= its heap key)
  // which node will be the root after this pop?  (first level of downtree, decided now; unless one of the four updates
  // puts a smaller key on top it is the next node to be accepted: the stencil threads prefetch its stencil lines)
  P.pred = -1;
  P.predk = 0;
  P.pred2 = -1;
  {
    const int n1 = S.ntr - 1;
    if (n1 == 1) { P.pred = P.last.y; P.predk = P.last.x; }
    else if (n1 >= 2) {
      const int2 c2 = tps_hget(S, 2);
      int2 c = c2, o = make_int2(0x7f800000, -1);
      int pc = 2;
      if (n1 >= 3) { const int2 c3 = tps_hget(S, 3); if (TKEY(c2) > TKEY(c3)) { c = c3; o = c2; pc = 3; } else o = c3; }
      if (!(TKEY(c) < TKEY(P.last))) { o = c; c = P.last; pc = 0; }
      P.pred = c.y; P.predk = c.x;
      // the one after: the other child of the root or a child of the predicted node (prefetch hint only)
      if (pc && 2 * pc + 1 <= n1) {
        const int2 g0 = tps_hget(S, 2 * pc), g1 = tps_hget(S, 2 * pc + 1);
        if (TKEY(g0) < TKEY(o)) o = g0;
        if (TKEY(g1) < TKEY(o)) o = g1;
      }
      P.pred2 = o.y;
    }
  }
  if (URG == 1) {
    if ((P.ix == 0 && G.ex_l) || (P.ix == G.nnx - 1 && G.ex_r) || (P.iz == 0 && G.ex_t) || (P.iz == G.nnz - 1 && G.ex_b)) {
      S.stopped_at_root = 1;
      return false;
    }
  }
  ++nacc;
  return true;
}

// (2) neighbour q of the accepted node (ix, iz): its status and, if it is far or close, its trial time = the minimum
//     of fouds2's four quadrant solves over the nodes that are alive now.  Reads E only.  The four quadrants can be
//     split over QS = 1, 2 or 4 threads (part = 0..QS-1): thread `part` evaluates quadrants part, part + QS, ... and
//     returns the minimum over ITS quadrants (the caller reduces over the parts; min is exact and order-free).
//     quadrant index = js * 2 + ks (js: x-1 / x+1 side, ks: z-1 / z+1 side).
template <int URG, int QS>
TPS_HD TpsNb tps_neighbour_part(const TpsGrid& G, const int ix, const int iz, const unsigned tself, const int q, const int part) {
  const int ndx = (q == 0) ? -1 : (q == 1 ? 1 : 0), ndz = (q == 2) ? -1 : (q == 3 ? 1 : 0);
  const int cx = ix + ndx, cz = iz + ndz, ld = G.ld;
  TpsNb R;
  R.co = nidx<URG>(cx, cz, ld);
  R.qid = 0;
  R.qt = tps_inf();
  if (!(cx >= 0 && cx < G.nnx && cz >= 0 && cz < G.nnz)) { R.qst = -2; return R; }
  const unsigned* E = G.E;
  // ALL loads of the neighbour are issued together -- its own word, the first / second stencil nodes in the four
  // directions x-1, x+1, z-1, z+1, slowness, R sin(theta) -- before the status is looked at: with 8 000 fronts in flight
  // every gather is a DRAM round trip, and "status first, stencil only if the neighbour is not alive" made it two in a
  // row (measured: the heap warp waited 5 300 cycles per accept for its stencil threads).  The first node back towards
  // the accepted node is that node itself (alive with tself).  A thread only loads the directions its quadrants use.
  const unsigned cE = E[R.co];
  unsigned e1[4], e2[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int d = 0; d < 4; ++d) {
    e1[d] = E_OUT; e2[d] = E_OUT;
    const bool need = (QS == 1) || (QS == 2 && (d >= 2 || d == part)) || (QS == 4 && (d == (part >> 1) || d == 2 + (part & 1)));
    if (need) {
      const int ddx = (d == 0) ? -1 : (d == 1 ? 1 : 0), ddz = (d == 2) ? -1 : (d == 3 ? 1 : 0);
      const int s1x = cx + ddx, s1z = cz + ddz, s2x = s1x + ddx, s2z = s1z + ddz;
      if (ddx == -ndx && ddz == -ndz) e1[d] = tself;
      else if (s1x >= 0 && s1x < G.nnx && s1z >= 0 && s1z < G.nnz) e1[d] = E[nidx<URG>(s1x, s1z, ld)];
      if (s2x >= 0 && s2x < G.nnx && s2z >= 0 && s2z < G.nnz) e2[d] = E[nidx<URG>(s2x, s2z, ld)];
    }
  }
  const float slown = G.slow[cx * ld + cz], risti = G.risti_tab[cx];
  R.qst = (cE == E_FAR ? -1 : ((int)cE >= 0 ? 0 : 1));
  R.qid = (int)(cE & ~E_SIGN);
  if (R.qst == 0) return R;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 4 / QS; ++i) {
    const int qd = (QS == 2) ? part * 2 + i : part + i * QS;      // QS = 2: part = x side, i = z side
    const int js = qd >> 1, ks = qd & 1;
    bool ok = false;
    float trav = quadrant(e_status(e1[js]), e_status(e2[js]), tps_as_float((int)e1[js]), tps_as_float((int)e2[js]),
                          e_status(e1[2 + ks]), e_status(e2[2 + ks]), tps_as_float((int)e1[2 + ks]),
                          tps_as_float((int)e2[2 + ks]), slown, G.earth, risti, G.dnx, G.dnz, ok);
    if (!ok) trav = tps_inf();
    R.qt = fminf(R.qt, trav);
  }
  return R;
}
// hint only: pull the lines the gather of neighbour q of node (ix, iz) will touch towards the L2
template <int URG>
TPS_HD void tps_neighbour_prefetch(const TpsGrid& G, const int ix, const int iz, const int q) {
#if defined(__CUDA_ARCH__)
  const int ndx = (q == 0) ? -1 : (q == 1 ? 1 : 0), ndz = (q == 2) ? -1 : (q == 3 ? 1 : 0);
  const int cx = ix + ndx, cz = iz + ndz, ld = G.ld;
  if (!(cx >= 0 && cx < G.nnx && cz >= 0 && cz < G.nnz)) return;
  const unsigned* E = G.E;
  asm volatile("prefetch.global.L2 [%0];" ::"l"(E + nidx<URG>(cx, cz, ld)));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(G.slow + cx * ld + cz));
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int ddx = (d == 0) ? -1 : (d == 1 ? 1 : 0), ddz = (d == 2) ? -1 : (d == 3 ? 1 : 0);
    const int s2x = cx + 2 * ddx, s2z = cz + 2 * ddz;      // the second stencil node; the first shares its sector or the centre's
    if (s2x >= 0 && s2x < G.nnx && s2z >= 0 && s2z < G.nnz) asm volatile("prefetch.global.L2 [%0];" ::"l"(E + nidx<URG>(s2x, s2z, ld)));
    const int s1x = cx + ddx, s1z = cz + ddz;
    if (ddz != 0 && s1x >= 0 && s1x < G.nnx && s1z >= 0 && s1z < G.nnz) asm volatile("prefetch.global.L2 [%0];" ::"l"(E + nidx<URG>(s1x, s1z, ld)));
  }
#endif
}
template <int URG>
TPS_HD TpsNb tps_neighbour(const TpsGrid& G, const int ix, const int iz, const unsigned tself, const int q) {
  return tps_neighbour_part<URG, 1>(G, ix, iz, tself, q, 0);
}

// (3) pop the root ...
template <bool SYNC>
TPS_HD void tps_pop(TpsState& S, const TpsPre& P, const bool run) { tps_pop_root<SYNC>(S, P.last, run); }
// (4) ... then insert / update the four neighbours in the reference order x-1, x+1, z-1, z+1 (addtree / updtree,
//     CalSurfG.f90:738-774, :864-890).  A close neighbour's position was read by the stencil warp BEFORE the pop, so
//     it is verified against the heap ("does that slot hold the neighbour?") together with the fetch of its parent:
//     one round trip for the four neighbours; only a position the pop really moved is re-read from E.  The common
//     case "the new key is not smaller than its parent" is a single store.  (Fetching three ancestor levels per
//     neighbour up front was tried and measured slower: 11 600 against 9 200 cycles per round on S200 -- the extra
//     loads and register patching cost more than the rare multi-level move saves; profiles/r2_k3_cohort_cycle_split.txt.)
template <int URG, bool SYNC>
TPS_HD bool tps_apply(TpsState& S, const TpsGrid& G, const TpsNb (&N)[4], bool run) {
  int qst[4], spos[4], ppos[4];
  int2 pent[4], sent[4];
  int nins = 0;
  TPS_TICK0(S);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    qst[q] = run ? N[q].qst : 0;                    // a lane that is not running sees four alive neighbours: nothing to do
    if (qst[q] == -1) ++nins;
  }
  if (run && S.ntr + nins >= S.htot) {
    S.overflow = 1; S.ntr = 0; run = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) qst[q] = 0;
    nins = 0;
  }
  // The entry of the LAST heap slot as it will stand after this step is the next pop's `last`: with an insert that slot
  // is written below (every write to it is seen here); without one it is read now, under the latency of the updates.
  const int nfinal = S.ntr + nins;
  int2 laste = make_int2(0, 0);
  if (run && nins == 0 && S.ntr >= 1) laste = tps_hget(S, S.ntr);
  // start positions: close = back pointer as read before the pop, far = next free heap slots in order
  int nt = S.ntr;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    spos[q] = 0;
    if (qst[q] == 1) spos[q] = N[q].qid;
    else if (qst[q] == -1) spos[q] = ++nt;
    ppos[q] = spos[q] >> 1;
    sent[q] = make_int2(0, -1);
    pent[q] = make_int2(0, 0);
    if (qst[q] == 1 && spos[q] >= 1 && spos[q] <= S.ntr) sent[q] = tps_hget(S, spos[q]);
    if ((qst[q] == 1 || qst[q] == -1) && ppos[q] > 0) pent[q] = tps_hget(S, ppos[q]);
  }
  TPS_SYNCWARP(SYNC);
  TPS_TICK(S, 3);      // 3: statuses, issue of the slot + parent loads
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    if (qst[q] == 1 && sent[q].y != N[q].co) {       // moved by the pop: take the fresh back pointer
      spos[q] = (int)(S.E[N[q].co] & ~E_SIGN);
      ppos[q] = -1;                                   // parent is fetched below
    }
  }
  TPS_SYNCWARP(SYNC);
  TPS_TICK(S, 4);      // 4: wait for the loads + verification
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    if (qst[q] == 1 || qst[q] == -1) {
      if (qst[q] == -1) S.ntr += 1;
      const float k = N[q].qt;
      int tpc = spos[q];
      const int tpc0 = tpc;
      if ((tpc >> 1) != ppos[q]) { ppos[q] = tpc >> 1; if (ppos[q] > 0) pent[q] = tps_hget(S, ppos[q]); }   // moved earlier in this step
      if (ppos[q] > 0 && k < TKEY(pent[q])) {
        // the key moves up: generic loop (addtree / updtree sift-up), patching what later neighbours hold in registers
        int tpp = ppos[q];
        int2 par = pent[q];
        for (;;) {
          tps_hput(S, tpc, par);
          if (tpc == nfinal) laste = par;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
          for (int r = 0; r < 4; ++r) {
            if (r > q && qst[r] == 1 && N[r].co == par.y) spos[r] = tpc;    // a later close neighbour moved down
            if (r > q && ppos[r] == tpc) pent[r] = par;                     // a later parent slot changed content
          }
          tpc = tpp;
          tpp = tpc >> 1;
          if (tpp == 0) break;
          par = tps_hget(S, tpp);
          if (!(k < TKEY(par))) break;
        }
      }
      const int2 e = make_int2(tps_as_int(k), N[q].co);
      // a close neighbour that stays where it is keeps its back pointer: only the key changes (no write to E)
      if (qst[q] == 1 && tpc == tpc0) { if (tpc < S.hcap) S.sm[(size_t)tpc * S.stride] = e; else S.gl[tpc - S.hcap] = e; }
      else tps_hput(S, tpc, e);
      if (tpc == nfinal) laste = e;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int r = 0; r < 4; ++r)
        if (r > q && ppos[r] == tpc) pent[r] = e;
    }
    TPS_SYNCWARP(SYNC);
    TPS_TICK(S, 6 + (q & 1));    // 6 / 7: neighbours
  }
  S.last_e = laste;
  S.last_p = (run && nfinal >= 1) ? nfinal : -1;
  return run;
}

template <int URG>
TPS_HD bool tps_step(TpsState& S, const TpsGrid& G, unsigned long long& nacc) {
  TpsPre P;
  if (!tps_pre<URG>(S, G, nacc, P)) return false;
  TpsNb N[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) N[q] = tps_neighbour<URG>(G, P.ix, P.iz, P.tself, q);
  tps_pop<false>(S, P, true);
  return tps_apply<URG, false>(S, G, N, true);
}

// ---- source cell initialisation (travel, CalSurfG.f90:324-345) on the refined grid ----
TPS_HD void tps_source_init(TpsState& S, const GridC& g, const SrcRec& sr, const float* vv, const float* ub, unsigned* E_r) {
  tps_reset(S);
  S.E = E_r;
  const int isx = sr.isx_r, isz = sr.isz_r;
  float vss[2][2];
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) vss[i - 1][j - 1] = refined_vel_t(g, sr, vv, ub, isz - 1 + j, isx - 1 + i);
  const float dsx = sr.dsx_r, dsz = sr.dsz_r;
  float vsrc = 0.0f;
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) {
      const float produ = (1.0f - fabsf(((float)(i - 1) * sr.dnxr - dsx) / sr.dnxr)) *
                          (1.0f - fabsf(((float)(j - 1) * sr.dnzr - dsz) / sr.dnzr));
      vsrc = vsrc + vss[i - 1][j - 1] * produ;
    }
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) {
      const float ax = dsx - (float)(i - 1) * sr.dnxr;
      const float az = dsz - (float)(j - 1) * sr.dnzr;
      const float ds = sqrtf(ax * ax + az * az);
      const float t0 = 2.0f * ds / (vss[i - 1][j - 1] + vsrc);
      const int o = (isx - 1 + i - 1) * REF_LD + (isz - 1 + j - 1);
      S.ntr += 1;
      tps_sift_up_plain(S, S.ntr, t0, o);
    }
}

// ---- after the refined march: the close nodes get their trial value back (the hand-off and the ray tracer read
//      E_r as "alive = +t, close = t | sign, far"), optionally their heap slots (test seam = the reference's nstsr) ----
TPS_HD void tps_refined_finish(TpsState& S, unsigned* E_r, int* hpos_out) {
  for (int p = 1; p <= S.ntr; ++p) {
    if (p == 1 && S.stopped_at_root) { if (hpos_out) hpos_out[tps_hget(S, 1).y] = 1; continue; }
    const int2 e = tps_hget(S, p);
    const int n = e.y;
    E_r[n] = (unsigned)e.x | E_SIGN;
    if (hpos_out) hpos_out[n] = p;
  }
}

// ---- hand-off to the coarse grid (FwdTraveltimeCPS.f90:576-632) + heap build (travel urg=2, CalSurfG.f90:311-317).
//      E_c was preset to FAR. ----
TPS_HD void tps_handoff(TpsState& S, const GridC& g, const SrcRec& sr, const unsigned* E_r, unsigned* E_c) {
  const int nkx = (sr.nnxr - 1) / g.sgdl + 1, nkz = (sr.nnzr - 1) / g.sgdl + 1;
  for (int kx = 0; kx < nkx; ++kx)
    for (int kz = 0; kz < nkz; ++kz)
      E_c[cidx(sr.vnl + kx - 1, sr.vnt + kz - 1, g.nnz)] = E_r[(kx * g.sgdl) * REF_LD + (kz * g.sgdl)];
  // alive nodes with a far neighbour become close (:615-632).  Restricted to the injected box: everything outside it
  // is far.  The test runs on the ORIGINAL alive set: a node marked close here keeps its time under the sign bit, and
  // "far" is an exact bit pattern, so marking in place cannot change a later test.
  for (int k = sr.vnl; k <= sr.vnr; ++k)
    for (int l = sr.vnt; l <= sr.vnb; ++l) {
      const int o = cidx(k - 1, l - 1, g.nnz);
      if ((int)E_c[o] >= 0) {
        bool mk = false;
        if (l - 1 >= 1 && E_c[cidx(k - 1, l - 2, g.nnz)] == E_FAR) mk = true;
        if (l + 1 <= g.nnz && E_c[cidx(k - 1, l, g.nnz)] == E_FAR) mk = true;
        if (k - 1 >= 1 && E_c[cidx(k - 2, l - 1, g.nnz)] == E_FAR) mk = true;
        if (k + 1 <= g.nnx && E_c[cidx(k, l - 1, g.nnz)] == E_FAR) mk = true;
        if (mk) E_c[o] |= E_SIGN;
      }
    }
  // heap build in scan order i=1..nnx, j=1..nnz; the node's word switches from "t | sign" to "position | sign"
  tps_reset(S);
  S.E = E_c;
  for (int k = sr.vnl; k <= sr.vnr && !S.overflow; ++k)
    for (int l = sr.vnt; l <= sr.vnb; ++l) {
      const int o = cidx(k - 1, l - 1, g.nnz);
      const unsigned ev = E_c[o];
      if ((int)ev < 0 && ev != E_FAR) {
        if (S.ntr + 1 >= S.htot) { S.overflow = 1; break; }
        S.ntr += 1;
        tps_sift_up_plain(S, S.ntr, tps_as_float((int)(ev & ~E_SIGN)), o);
      }
    }
}

TPS_HD TpsGrid tps_grid_refined(const GridC& g, const SrcRec& sr, const float* slow_r, const float* risti_r, unsigned* E_r) {
  TpsGrid G;
  G.nnx = sr.nnxr; G.nnz = sr.nnzr; G.ld = REF_LD; G.inv_ld = 1.0f / (float)REF_LD; G.dnx = sr.dnxr; G.dnz = sr.dnzr; G.earth = g.earth;
  G.slow = slow_r; G.risti_tab = risti_r; G.E = E_r;
  // exit tests literal to CalSurfG.f90:366-377 (vnr/vnb vs *refined* nnx/nnz)
  G.ex_l = sr.vnl != 1; G.ex_r = sr.vnr != sr.nnxr; G.ex_t = sr.vnt != 1; G.ex_b = sr.vnb != sr.nnzr;
  return G;
}
TPS_HD TpsGrid tps_grid_coarse(const GridC& g, const float* slow_c, const float* risti_c, unsigned* E_c) {
  TpsGrid G;
  G.nnx = g.nnx; G.nnz = g.nnz; G.ld = g.nnz; G.inv_ld = 1.0f / (float)g.nnz; G.dnx = g.dnx; G.dnz = g.dnz; G.earth = g.earth;
  G.slow = slow_c; G.risti_tab = risti_c; G.E = E_c;
  G.ex_l = G.ex_r = G.ex_t = G.ex_b = false;
  return G;
}

}  // namespace dz
