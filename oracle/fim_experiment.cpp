// ORACLE-side EXPERIMENT (test/analysis infrastructure, NOT product code, NOT the reference's algorithm).
//
// Question it answers ("why K3 is an exact heap march"): how far from the reference's
// heap fast-marching field does an order-free fixed point of the SAME local solver land?  BASELINE.json's north star
// names a fast-iterative-method (FIM) kernel; a FIM converges to the fixed point
//     T(X) = F(X ; { neighbours N : T(N) < T(X) })
// where F is the reference's quadrant solver fouds2 (CalSurfG.f90:557-729) with "N is alive" replaced by the causal
// test "T(N) < T(X)" -- what "alive when X was last updated" means in a heap march whose pops are monotone.  The
// reference's result differs from that fixed point wherever its pops are not monotone, wherever equal keys were broken
// by heap position, and wherever a trial value was OVERWRITTEN by a later, larger one (CalSurfG.f90:728 is an
// assignment, not a minimum).  This file computes the fixed point for the coarse-grid continuation (the refined source
// box and the hand-off stay the reference's, FwdTraveltimeCPS.f90:576-632) by Gauss-Seidel passes in four orderings --
// any order-free scheme (FIM active lists, fast sweeping, Jacobi) reaches the same fixed point -- so that
// scripts/fim_vs_fmm.py can measure the difference in travel times, ray footprints and the G sparsity pattern.
#include "fmm2d.hpp"
#include <algorithm>
#include <cmath>

namespace orc {

static inline float sin_rf(float x) { return (float)std::sin((double)x); }
static const float FIM_INF = 1.0e30f;

// fouds2 with the alive tests replaced by the causal test against tcur (the node's current value); returns the minimum
// over the quadrants that have a usable neighbour, or FIM_INF.  Arithmetic is the reference's, operation for operation.
float Fmm::fouds2_values(int iz, int ix, float tcur, bool second_order) {
  int tsw1 = 0;
  float travm = FIM_INF, trav;
  const float slown = 1.0f / VELN(iz, ix);
  const float ri = earth;
  const float risti = ri * sin_rf(gox + (float)(ix - 1) * dnx);
  auto usable = [&](int z, int x) { return TTN(z, x) < tcur; };
  for (int j = ix - 1; j <= ix + 1; j += 2) {
    if (j < 1 || j > nnx) continue;
    int swj = -1, j2 = (j == ix - 1) ? j - 1 : j + 1;
    if (second_order && j2 >= 1 && j2 <= nnx && usable(iz, j2)) swj = 0;
    const bool aj = usable(iz, j);
    if (aj && swj == 0) {
      swj = -1;
      if (TTN(iz, j) > TTN(iz, j2)) swj = 0;
    } else {
      swj = -1;
    }
    for (int k = iz - 1; k <= iz + 1; k += 2) {
      if (k < 1 || k > nnz) continue;
      int swk = -1, k2 = (k == iz - 1) ? k - 1 : k + 1;
      if (second_order && k2 >= 1 && k2 <= nnz && usable(k2, ix)) swk = 0;
      const bool ak = usable(k, ix);
      if (ak && swk == 0) {
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
        } else if (ak) {
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
      } else if (aj) {
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
        } else if (ak) {
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
        } else if (ak) {
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
        const float tdsh = (-b + std::sqrt(rd1)) / (2.0f * a);
        trav = (tref + tdsh) / tdiv;
        if (tsw1 == 1) travm = std::min(trav, travm);
        else { travm = trav; tsw1 = 1; }
      }
    }
  }
  return travm;
}

// Fixed point on the coarse grid after the refined -> coarse hand-off.  Entry state = the state travel(urg = 2) gets:
// NSTS == 0 interior alive nodes (final), NSTS > 0 the initial narrow band (alive-with-a-far-neighbour nodes and
// injected close nodes, with their injected values), NSTS == -1 far.  Interior alive nodes stay fixed (the reference
// never touches them again); band values are upper bounds that an update may lower; far nodes start at +infinity.
int Fmm::travel_fim() {
  std::vector<char> fixed((size_t)nnx * nnz, 0);     // 1: interior alive (final), 2: initial narrow band
  std::vector<float> injected((size_t)nnx * nnz, 0.0f);
  auto id = [&](int iz, int ix) { return (size_t)(ix - 1) * nnz + (iz - 1); };
  for (int ix = 1; ix <= nnx; ++ix)
    for (int iz = 1; iz <= nnz; ++iz) {
      if (NSTS(iz, ix) == 0) fixed[id(iz, ix)] = 1;
      else if (NSTS(iz, ix) > 0) { fixed[id(iz, ix)] = 2; injected[id(iz, ix)] = TTN(iz, ix); }
      else TTN(iz, ix) = FIM_INF;
    }
  // A band node keeps its injected (refined-grid) value in the reference unless a non-interior direct neighbour is
  // accepted before it -- only then does fouds2 overwrite it with a coarse-stencil value (CalSurfG.f90:394-417).
  auto band_is_recomputed = [&](int iz, int ix, float tx) {
    const int dz[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
    for (int q = 0; q < 4; ++q) {
      const int z = iz + dz[q], x = ix + dx[q];
      if (z < 1 || z > nnz || x < 1 || x > nnx) continue;
      if (fixed[id(z, x)] != 1 && TTN(z, x) < tx) return true;
    }
    return false;
  };
  fim_sweeps = 0;
  fim_converged = 0;
  // Phase 1: first-order scheme, monotone (minimum) updates: an upper bound with the right causal structure.
  // Phase 2: the reference's mixed first/second-order solver, values OVERWRITTEN (a second-order extrapolation from
  // not-yet-converged neighbours can undershoot; with minimum-only updates such a transient would be frozen in),
  // repeated until a whole pass changes nothing bit for bit.
  for (int phase = 1; phase <= 2; ++phase) {
    for (int pass = 0; pass < 2000; ++pass) {
      long changed = 0;
      for (int ord = 0; ord < 4; ++ord) {
        const int x0 = (ord & 1) ? nnx : 1, x1 = (ord & 1) ? 0 : nnx + 1, dx = (ord & 1) ? -1 : 1;
        const int z0 = (ord & 2) ? nnz : 1, z1 = (ord & 2) ? 0 : nnz + 1, dz = (ord & 2) ? -1 : 1;
        for (int ix = x0; ix != x1; ix += dx)
          for (int iz = z0; iz != z1; iz += dz) {
            const char kind = fixed[id(iz, ix)];
            if (kind == 1) continue;
            const float cur = TTN(iz, ix);
            if (kind == 2 && !band_is_recomputed(iz, ix, injected[id(iz, ix)])) {
              if (cur != injected[id(iz, ix)]) { TTN(iz, ix) = injected[id(iz, ix)]; ++changed; }
              continue;
            }
            if (phase == 1) {
              const float t = fouds2_values(iz, ix, cur, false);
              if (t < cur) { TTN(iz, ix) = t; ++changed; }
            } else {
              // neighbours count as upwind when they are below the value they produce: test against +infinity first,
              // then keep only those below the candidate (one refinement is enough at a fixed point)
              float t = fouds2_values(iz, ix, FIM_INF, true);
              t = fouds2_values(iz, ix, t, true);
              if (t >= FIM_INF) t = cur;
              if (t != cur) { TTN(iz, ix) = t; ++changed; }
            }
          }
      }
      ++fim_sweeps;
      if (changed == 0) { if (phase == 2) fim_converged = 1; break; }
    }
  }
  for (int ix = 1; ix <= nnx; ++ix)
    for (int iz = 1; iz <= nnz; ++iz) NSTS(iz, ix) = 0;
  n_accept += (long)nnx * nnz;
  return OK;
}

}  // namespace orc

extern "C" {
// coarse travel-time field of one source from the fixed-point experiment: ttn (nnz,nnx) column-major; returns the
// number of Gauss-Seidel passes in *sweeps
int orc_fmm_source_fim(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv, float scx,
                       float scz, float* ttn, long* sweeps) {
  orc::Fmm f;
  f.init(nx, ny, goxd, gozd, dvxd, dvzd);
  f.fim_coarse = 1;
  int st = f.solve_source(pv, scx, scz);
  if (st) return st;
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) ttn[(size_t)(ix - 1) * f.nnz + (iz - 1)] = f.TTN(iz, ix);
  if (sweeps) *sweeps = f.fim_converged ? f.fim_sweeps : -f.fim_sweeps;   // negative: phase 2 hit the pass limit
  return 0;
}
}
