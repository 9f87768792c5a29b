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
#include <cstdlib>

namespace orc {

static inline float sin_rf(float x) { return (float)std::sin((double)x); }
static const float FIM_INF = 1.0e30f;

// fouds2 with the alive tests replaced by the causal test against tcur (the node's current value); returns the minimum
// over the quadrants that have a usable neighbour, or FIM_INF.  Arithmetic is the reference's, operation for operation.
// (defined after fouds2_pred's declaration in the header; the template body follows)
float Fmm::fouds2_values(int iz, int ix, float tcur, bool second_order) {
  return fouds2_pred(iz, ix, second_order, [&](int z, int x) { return TTN(z, x) < tcur; });
}

// The reference's quadrant solver with its "alive" tests answered by a caller-supplied predicate.
template <class Pred>
float Fmm::fouds2_pred(int iz, int ix, bool second_order, Pred usable) {
  int tsw1 = 0;
  float travm = FIM_INF, trav;
  const float slown = 1.0f / VELN(iz, ix);
  const float ri = earth;
  const float risti = ri * sin_rf(gox + (float)(ix - 1) * dnx);
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
  fim_evals = 0;
  // Phase 1: first-order scheme, monotone (minimum) updates: an upper bound with the right causal structure.
  // Phase 2: the reference's mixed first/second-order solver, values OVERWRITTEN (a second-order extrapolation from
  // not-yet-converged neighbours can undershoot; with minimum-only updates such a transient would be frozen in),
  // repeated until a whole pass changes nothing bit for bit.
  if (fim_coarse == 2) {
    // ACTIVE-LIST variant (what a GPU fast-iterative kernel does): only nodes one of whose stencil neighbours changed
    // are re-evaluated.  FIFO work lists, same update rules, same fixed point; counts the evaluations it needs.
    std::vector<int> q;
    std::vector<char> inq((size_t)nnx * nnz + (size_t)ld * nnx, 0);
    auto push = [&](int iz, int ix) {
      if (iz < 1 || iz > nnz || ix < 1 || ix > nnx) return;
      if (fixed[id(iz, ix)] == 1) return;
      const size_t k = (size_t)(ix - 1) * ld + (iz - 1);
      if (!inq[k]) { inq[k] = 1; q.push_back((int)k); }
    };
    for (int phase = 1; phase <= 2; ++phase) {
      q.clear(); std::fill(inq.begin(), inq.end(), 0);
      for (int ix = 1; ix <= nnx; ++ix)
        for (int iz = 1; iz <= nnz; ++iz) {
          if (phase == 1) { if (TTN(iz, ix) < FIM_INF) { push(iz - 1, ix); push(iz + 1, ix); push(iz, ix - 1); push(iz, ix + 1); push(iz, ix); } }
          else push(iz, ix);
        }
      size_t head = 0;
      const size_t cap = (size_t)400 * nnx * nnz;
      while (head < q.size() && q.size() < cap) {
        const int k = q[head++];
        inq[k] = 0;
        const int ix = k / ld + 1, iz = k % ld + 1;
        const char kind = fixed[id(iz, ix)];
        const float cur = TTN(iz, ix);
        float t;
        if (kind == 2 && !band_is_recomputed(iz, ix, injected[id(iz, ix)])) t = injected[id(iz, ix)];
        else if (phase == 1) { ++fim_evals; t = std::min(cur, fouds2_values(iz, ix, cur, false)); }
        else {
          fim_evals += 2;
          t = fouds2_values(iz, ix, FIM_INF, true);
          t = fouds2_values(iz, ix, t, true);
          if (t >= FIM_INF) t = cur;
        }
        if (t != cur) {
          TTN(iz, ix) = t;
          for (int d = 1; d <= (phase == 1 ? 1 : 2); ++d) { push(iz - d, ix); push(iz + d, ix); push(iz, ix - d); push(iz, ix + d); }
        }
      }
      if (phase == 2) fim_converged = head >= q.size();
      if (head > 0x3fffffff) break;
    }
    fim_sweeps = 1;
    for (int ix = 1; ix <= nnx; ++ix)
      for (int iz = 1; iz <= nnz; ++iz) NSTS(iz, ix) = 0;
    n_accept += (long)nnx * nnz;
    return OK;
  }
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
            fim_evals += phase == 1 ? 1 : 2;
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

namespace orc {

// ---------------------------------------------------------------------------------------------------------------
// ORDER EXPERIMENT.  In the reference's march the value a node ends with is a function of the acceptance ORDER only:
// it is fouds2 evaluated when the last of its direct neighbours was accepted before its own pop, with exactly the nodes
// accepted up to that moment alive (CalSurfG.f90:394-417, 557-729).  If that rule is right (checked bit for bit below),
// a solve can be REPLAYED from a rank per node with purely local work -- parallel over the dependency wavefronts
// instead of serial over the heap -- provided the ranks can be predicted.  This measures how well "rank = position in
// the sorted list of (approximate) arrival times" predicts them, and how deep the dependency graph is.
struct OrderStats {
  long popped, rule_mismatch;             // nodes accepted by the coarse march; nodes where the rank rule != reference
  long pairs, pair_ties, pair_inversions; // interacting (stencil) pairs: equal final keys / order contradicts the keys
  long sorted_exact_mismatch;             // replay with rank = sort(final reference values): nodes != reference
  long sorted_fim_mismatch;               // replay with rank = sort(fixed-point values): nodes != reference
  long sorted_fim_rank_errors;            // ... replay steps that found no accepted direct neighbour (impossible order)
  long dag_levels;                        // longest dependency chain of the replay
  long fim_passes;
  long fim_evals;                         // quadrant-solver evaluations of the fixed-point solve (Gauss-Seidel upper bound)
  // local, order-free checks on the replay made from the fixed-point ranks (what a GPU path could run per node):
  long verify_order_flags;                // interacting pairs whose replayed values are tied or contradict the ranks
  long verify_key_increase_flags;         // nodes inside a key-increase window that interact with a node later in that window
  long key_increase_events;               // overwrites of a heap key by a larger value
  long harmless_tie_groups;               // groups of exactly tied interacting nodes whose order provably does not matter
};

namespace {
struct Replay {
  Fmm& f;
  const std::vector<int>& s0;        // status handed to the coarse march
  const std::vector<float>& t0;      // values handed to it
  size_t id(int iz, int ix) const { return (size_t)(ix - 1) * f.ld + (iz - 1); }
  // values from ranks; returns replay steps that had no accepted direct neighbour although the node was far
  // exit_rank >= 0 (refined source box): the node with that rank is the one the march stopped at -- it is alive with
  // the value of its last trial but never updated its neighbours (CalSurfG.f90:362-382); nodes without a rank that
  // touch an accepted node are CLOSE and keep the trial of their last update (their status is returned in close_out).
  long run(const std::vector<int>& rank, std::vector<int>* level_out, int exit_rank = -1,
           std::vector<char>* close_out = nullptr) {
    const int nnx = f.nnx, nnz = f.nnz;
    const int trig_limit = exit_rank >= 0 ? exit_rank : 0x7fffffff;     // ranks below this update their neighbours
    std::vector<std::pair<int, int>> order;     // (rank, linear id)
    for (int ix = 1; ix <= nnx; ++ix)
      for (int iz = 1; iz <= nnz; ++iz) {
        const size_t k = id(iz, ix);
        if (s0[k] >= 0) f.ttn[k] = t0[k]; else f.ttn[k] = FIM_INF;
        if (rank[k] >= 0) order.push_back({rank[k], (int)k});
      }
    std::sort(order.begin(), order.end());
    std::vector<int> level;
    if (level_out) level.assign(f.ttn.size(), 0);
    long impossible = 0;
    for (const auto& e : order) {
      const int k = e.second, r = e.first;
      const int ix = k / f.ld + 1, iz = k % f.ld + 1;
      int tstar = -1;
      bool any = false;
      const int dz[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
      for (int q = 0; q < 4; ++q) {
        const int z = iz + dz[q], x = ix + dx[q];
        if (z < 1 || z > nnz || x < 1 || x > nnx) continue;
        const int rn = rank[id(z, x)];
        if (rn >= 0 && rn < r && rn < trig_limit) { any = true; tstar = std::max(tstar, rn); }
      }
      if (!any) {                         // never recomputed by the march: a band node keeps its injected value
        if (s0[k] < 0) ++impossible;
        continue;
      }
      auto alive = [&](int z, int x) {
        const size_t n = id(z, x);
        return s0[n] == 0 || (rank[n] >= 0 && rank[n] <= tstar);
      };
      f.ttn[k] = f.fouds2_pred(iz, ix, true, alive);
      if (level_out) {
        int lv = 0;
        const int sz[8] = {-1, -2, 1, 2, 0, 0, 0, 0}, sx[8] = {0, 0, 0, 0, -1, -2, 1, 2};
        for (int q = 0; q < 8; ++q) {
          const int z = iz + sz[q], x = ix + sx[q];
          if (z < 1 || z > nnz || x < 1 || x > nnx) continue;
          if (rank[id(z, x)] >= 0 && alive(z, x)) lv = std::max(lv, level[id(z, x)]);
        }
        level[k] = lv + 1;
      }
    }
    if (close_out) {
      close_out->assign(f.ttn.size(), 0);
      const int dz[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
      for (int ix = 1; ix <= nnx; ++ix)
        for (int iz = 1; iz <= nnz; ++iz) {
          const size_t k = id(iz, ix);
          if (rank[k] >= 0 || s0[k] == 0) continue;        // accepted in this march / before it began
          int tstar = -1;
          for (int q = 0; q < 4; ++q) {
            const int z = iz + dz[q], x = ix + dx[q];
            if (z < 1 || z > nnz || x < 1 || x > nnx) continue;
            const int rn = rank[id(z, x)];
            if (rn >= 0 && rn < trig_limit) tstar = std::max(tstar, rn);
          }
          if (tstar < 0) { if (s0[k] > 0) (*close_out)[k] = 1; continue; }     // untouched band node stays close
          (*close_out)[k] = 1;
          f.ttn[k] = f.fouds2_pred(iz, ix, true, [&](int z, int x) {
            const size_t n = id(z, x);
            return s0[n] == 0 || (rank[n] >= 0 && rank[n] <= tstar);
          });
        }
    }
    if (level_out) *level_out = level;
    return impossible;
  }
};

// Checks on a finished replay (f.ttn = replayed values, rank = the ranks it used); every step is local to a node or a
// sort, nothing walks the heap.
//  (a) every interacting pair must be STRICTLY ordered by value the way the ranks say (ties between interacting nodes
//      are broken by heap position in the reference: not predictable);
//  (b) the keys a node had in the heap -- its injected value if it was in the initial band, then one trial per direct
//      neighbour accepted before it -- normally only decrease.  When a trial is LARGER than the key it overwrites
//      (CalSurfG.f90:728 assigns, updtree :864 only sifts up) the node keeps a heap slot justified by the old key: the
//      edge to its children may be invalid, and nodes below it with keys in the window [old, new) are popped late, at
//      the latest when the frontier reaches `new`.  Such a delay changes values only if a node whose key lies in a
//      window interacts with a node whose key lies between its own and the end of the (merged) window.  Windows are
//      merged transitively (a delayed node may itself hide others).
void verify_replay(Fmm& f, const std::vector<int>& s0, const std::vector<float>& t0, const std::vector<int>& rank,
                   long* order_flags, long* increase_flags, long* increase_events, int exit_rank = -1,
                   long* harmless_tie_groups = nullptr) {
  auto id = [&](int iz, int ix) { return (size_t)(ix - 1) * f.ld + (iz - 1); };
  *order_flags = 0; *increase_flags = 0; *increase_events = 0;
  if (harmless_tie_groups) *harmless_tie_groups = 0;
  const int sz[8] = {-1, -2, 1, 2, 0, 0, 0, 0}, sx[8] = {0, 0, 0, 0, -1, -2, 1, 2};
  const int trig_limit = exit_rank >= 0 ? exit_rank : 0x7fffffff;
  std::vector<int> rk = rank;                     // mutable copy for the tie permutations
  // the rule for one node under the ranks rk (what Replay::run computes, without writing)
  auto rule_value = [&](size_t k) -> float {
    const int ix = (int)(k / f.ld) + 1, iz = (int)(k % f.ld) + 1;
    const int r = rk[k] >= 0 ? rk[k] : 0x7fffffff;
    int tstar = -1;
    const int dz[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
    for (int q = 0; q < 4; ++q) {
      const int z = iz + dz[q], x = ix + dx[q];
      if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
      const int rn = rk[id(z, x)];
      if (rn >= 0 && rn < r && rn < trig_limit) tstar = std::max(tstar, rn);
    }
    if (tstar < 0) return s0[k] >= 0 ? t0[k] : FIM_INF;
    return f.fouds2_pred(iz, ix, true, [&](int z, int x) {
      const size_t n = id(z, x);
      return s0[n] == 0 || (rk[n] >= 0 && rk[n] <= tstar);
    });
  };
  std::vector<std::pair<float, int>> tied;         // (value, node) of every node in an exactly tied interacting pair
  std::vector<std::pair<float, float>> win;
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) {
      const size_t k = id(iz, ix);
      const int r = rank[k];
      if (r < 0) continue;
      for (int q = 0; q < 8; ++q) {
        const int z = iz + sz[q], x = ix + sx[q];
        if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
        const int rn = rank[id(z, x)];
        if (rn < 0 || rn > r) continue;
        if (f.ttn[id(z, x)] == f.ttn[k]) { tied.push_back({f.ttn[k], (int)k}); tied.push_back({f.ttn[k], (int)id(z, x)}); }
        else if (!(f.ttn[id(z, x)] < f.ttn[k])) ++*order_flags;
      }
      int trig[4] = {0, 0, 0, 0}, nt = 0;
      const int dz[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
      for (int q = 0; q < 4; ++q) {
        const int z = iz + dz[q], x = ix + dx[q];
        if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
        const int rn = rank[id(z, x)];
        if (rn >= 0 && rn < r) trig[nt++] = rn;
      }
      for (int a = 1; a < nt; ++a)                      // insertion sort of <= 4 entries
        for (int b = a; b > 0 && trig[b - 1] > trig[b]; --b) std::swap(trig[b - 1], trig[b]);
      float prev = s0[k] > 0 ? t0[k] : FIM_INF;          // key already in the heap at the hand-off, if any
      for (int t = 0; t < nt; ++t) {
        const int ts = trig[t];
        const float trial = f.fouds2_pred(iz, ix, true, [&](int z, int x) {
          const size_t n = id(z, x);
          return s0[n] == 0 || (rank[n] >= 0 && rank[n] <= ts);
        });
        if (trial > prev) { win.push_back({prev, trial}); ++*increase_events; }
        prev = trial;
      }
    }
  // The stopping rule of the refined box is a GLOBAL event (the first edge node to reach the heap root): a key equal to
  // the stopping node's key anywhere in the heap makes the accepted set depend on heap position.
  if (exit_rank >= 0) {
    float ve = FIM_INF;
    size_t ke = 0;
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] == exit_rank) { ve = f.ttn[k]; ke = k; }
    if (ve < FIM_INF)
      for (int ix = 1; ix <= f.nnx; ++ix)
        for (int iz = 1; iz <= f.nnz; ++iz)
          if (id(iz, ix) != ke && f.ttn[id(iz, ix)] == ve) ++*order_flags;
  }
  // Exact ties between interacting nodes: the reference breaks them by heap position.  A tie is harmless if every order
  // of the tied nodes gives the same values for them and for every node that can see them (their stencil neighbours
  // accepted later, and close nodes): checked by brute force over the permutations of small groups, locally.
  std::sort(tied.begin(), tied.end());
  tied.erase(std::unique(tied.begin(), tied.end()), tied.end());
  for (size_t a0 = 0; a0 < tied.size();) {
    size_t a1 = a0;
    while (a1 < tied.size() && tied[a1].first == tied[a0].first) ++a1;
    const size_t g = a1 - a0;
    bool hazard = g > 4;
    if (!hazard) {
      std::vector<int> mem, ranks0, watch;
      for (size_t i = a0; i < a1; ++i) { mem.push_back(tied[i].second); ranks0.push_back(rk[tied[i].second]); }
      int rmin = 0x7fffffff;
      for (int r0 : ranks0) rmin = std::min(rmin, r0);
      watch = mem;
      for (int m : mem) {
        const int ix = m / f.ld + 1, iz = m % f.ld + 1;
        for (int q = 0; q < 8; ++q) {
          const int z = iz + sz[q], x = ix + sx[q];
          if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
          const size_t n = id(z, x);
          if (s0[n] == 0) continue;
          if (rk[n] >= 0 && rk[n] < rmin) continue;            // accepted before any of them: cannot see them
          watch.push_back((int)n);
        }
      }
      std::sort(watch.begin(), watch.end());
      watch.erase(std::unique(watch.begin(), watch.end()), watch.end());
      std::vector<int> perm(g);
      for (size_t i = 0; i < g; ++i) perm[i] = (int)i;
      std::sort(ranks0.begin(), ranks0.end());
      // order the members by their current rank so that the identity permutation is the replay's own order
      std::sort(mem.begin(), mem.end(), [&](int x, int y) { return rank[x] < rank[y]; });
      do {
        for (size_t i = 0; i < g; ++i) rk[mem[i]] = ranks0[perm[i]];
        for (int n : watch) {
          if (rk[n] < 0 && s0[n] < 0) {                         // unranked: only close nodes carry a value
            bool touches = false;
            const int ix = n / f.ld + 1, iz = n % f.ld + 1;
            const int dz[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
            for (int q = 0; q < 4; ++q) {
              const int z = iz + dz[q], x = ix + dx[q];
              if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
              if (rk[id(z, x)] >= 0 && rk[id(z, x)] < trig_limit) touches = true;
            }
            if (!touches) continue;
          }
          if (rule_value((size_t)n) != f.ttn[n]) { hazard = true; break; }
        }
        if (hazard) break;
      } while (std::next_permutation(perm.begin(), perm.end()));
      for (size_t i = 0; i < g; ++i) rk[mem[i]] = rank[mem[i]];
    }
    if (hazard) ++*order_flags;
    else if (harmless_tie_groups) ++*harmless_tie_groups;
    a0 = a1;
  }
  if (win.empty()) return;
  std::sort(win.begin(), win.end());
  std::vector<std::pair<float, float>> merged;
  for (const auto& w : win) {
    if (!merged.empty() && w.first <= merged.back().second) merged.back().second = std::max(merged.back().second, w.second);
    else merged.push_back(w);
  }
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) {
      const size_t k = id(iz, ix);
      if (rank[k] < 0) continue;
      const float v = f.ttn[k];
      auto it = std::upper_bound(merged.begin(), merged.end(), std::make_pair(v, FIM_INF));
      if (it == merged.begin()) continue;
      --it;
      if (!(v >= it->first && v < it->second)) continue;      // this node's key lies in no hazard window
      const float end = it->second;
      for (int q = 0; q < 8; ++q) {
        const int z = iz + sz[q], x = ix + sx[q];
        if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
        if (rank[id(z, x)] < 0) continue;
        const float vn = f.ttn[id(z, x)];
        if (vn >= v && vn < end) { ++*increase_flags; break; }
      }
    }
}

std::vector<int> ranks_by_value(const Fmm& f, const std::vector<float>& val, const std::vector<int>& true_rank) {
  std::vector<std::pair<float, int>> v;
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) {
      const size_t k = (size_t)(ix - 1) * f.ld + (iz - 1);
      if (true_rank[k] >= 0) v.push_back({val[k], (int)k});       // the set of nodes the march accepts is known a priori
    }
  std::stable_sort(v.begin(), v.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
  std::vector<int> r(true_rank.size(), -1);
  for (size_t i = 0; i < v.size(); ++i) r[v[i].second] = (int)i;
  return r;
}
}  // namespace
}  // namespace orc

// prefix > 0: the first `prefix` accepts of the coarse march are taken from the reference (= marched serially) and only
// the rest is predicted and replayed.
extern "C" int orc_fmm_order_stats(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                                   float scx, float scz, int prefix, orc::OrderStats* out) {
  using namespace orc;
  Fmm f;
  f.init(nx, ny, goxd, gozd, dvxd, dvzd);
  std::vector<int> rank(f.ttn.size(), -1), s0;
  std::vector<float> t0;
  f.rec_rank = &rank; f.rec_init_nsts = &s0; f.rec_init_ttn = &t0;
  int st = f.solve_source(pv, scx, scz);
  if (st) return st;
  f.rec_rank = nullptr; f.rec_init_nsts = nullptr; f.rec_init_ttn = nullptr;
  const std::vector<float> ref = f.ttn;
  OrderStats S{};
  auto idx = [&](int iz, int ix) { return (size_t)(ix - 1) * f.ld + (iz - 1); };
  auto count_mismatch = [&]() {
    long m = 0;
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz)
        if (f.ttn[idx(iz, ix)] != ref[idx(iz, ix)]) ++m;
    return m;
  };
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz)
      if (rank[idx(iz, ix)] >= 0) ++S.popped;
  Replay R{f, s0, t0};
  // (1) the rank rule itself, with the reference's own order
  std::vector<int> level;
  R.run(rank, &level);
  S.rule_mismatch = count_mismatch();
  for (int v : level) S.dag_levels = std::max<long>(S.dag_levels, v);
  // (2) interacting pairs: do the final keys tell the order?
  const int sz[8] = {-1, -2, 1, 2, 0, 0, 0, 0}, sx[8] = {0, 0, 0, 0, -1, -2, 1, 2};
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) {
      const int r = rank[idx(iz, ix)];
      if (r < 0) continue;
      for (int q = 0; q < 8; ++q) {
        const int z = iz + sz[q], x = ix + sx[q];
        if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
        const int rn = rank[idx(z, x)];
        if (rn < 0 || rn > r) continue;                 // each unordered pair once: N accepted before X
        ++S.pairs;
        if (ref[idx(z, x)] == ref[idx(iz, ix)]) ++S.pair_ties;
        else if (ref[idx(z, x)] > ref[idx(iz, ix)]) ++S.pair_inversions;
      }
    }
  // (3) ranks predicted from the exact final values
  R.run(ranks_by_value(f, ref, rank), nullptr);
  S.sorted_exact_mismatch = count_mismatch();
  // optional serial prefix: state of the march after `prefix` accepts, by the (exact) rule
  if (prefix > 0 && prefix < S.popped) {
    std::vector<int> rcut(rank.size(), -1);
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] >= 0 && rank[k] < prefix) rcut[k] = rank[k];
    std::vector<char> close;
    R.run(rcut, nullptr, prefix, &close);
    std::vector<int> s1(rank.size(), -1);
    std::vector<float> t1(rank.size(), 0.0f);
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz) {
        const size_t k = idx(iz, ix);
        if (s0[k] == 0 || rcut[k] >= 0) { s1[k] = 0; t1[k] = f.ttn[k]; }
        else if (close[k]) { s1[k] = 1; t1[k] = f.ttn[k]; }
      }
    s0 = s1; t0 = t1;
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] >= 0 && rank[k] < prefix) rank[k] = -1;     // accepted before the replayed part began
  }
  // (4) ranks predicted from the fixed-point (fast-iterative) values
  f.nsts = s0; f.ttn = t0;
  if (const char* e = std::getenv("ORC_FIM_ACTIVE_LIST")) f.fim_coarse = std::atoi(e) ? 2 : 0;
  f.travel_fim();
  S.fim_passes = f.fim_sweeps;
  S.fim_evals = f.fim_evals;
  const std::vector<float> fimv = f.ttn;
  const std::vector<int> prank = ranks_by_value(f, fimv, rank);
  S.sorted_fim_rank_errors = R.run(prank, nullptr);
  S.sorted_fim_mismatch = count_mismatch();
  verify_replay(f, s0, t0, prank, &S.verify_order_flags, &S.verify_key_increase_flags, &S.key_increase_events, -1,
                &S.harmless_tie_groups);
  *out = S;
  return 0;
}

// RANK ITERATION.  Instead of paying a whole fixed-point solve for the ranks, start from any cheap guess of the arrival
// order (guess = a field whose sorted order is the first prediction: distance from the source, the same source's field
// at the neighbouring period, ...) and iterate  ranks -> replay -> values -> sort -> ranks  until the ranks stop
// changing; each round costs one solver evaluation per node plus a sort.  Reports the rounds needed and whether the
// fixed point is the reference's field.
extern "C" int orc_fmm_rank_iteration(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                                      float scx, float scz, int prefix, const float* guess /* (nnz,nnx) column-major */,
                                      int max_rounds, long* rounds, long* mismatch, long* flags, long* popped) {
  using namespace orc;
  Fmm f;
  f.init(nx, ny, goxd, gozd, dvxd, dvzd);
  std::vector<int> rank(f.ttn.size(), -1), s0;
  std::vector<float> t0;
  f.rec_rank = &rank; f.rec_init_nsts = &s0; f.rec_init_ttn = &t0;
  int st = f.solve_source(pv, scx, scz);
  if (st) return st;
  f.rec_rank = nullptr; f.rec_init_nsts = nullptr; f.rec_init_ttn = nullptr;
  const std::vector<float> ref = f.ttn;
  auto idx = [&](int iz, int ix) { return (size_t)(ix - 1) * f.ld + (iz - 1); };
  long np = 0;
  for (int v : rank) if (v >= 0) ++np;
  *popped = np;
  Replay R{f, s0, t0};
  if (prefix > 0 && prefix < np) {
    std::vector<int> rcut(rank.size(), -1);
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] >= 0 && rank[k] < prefix) rcut[k] = rank[k];
    std::vector<char> close;
    R.run(rcut, nullptr, prefix, &close);
    std::vector<int> s1(rank.size(), -1);
    std::vector<float> t1(rank.size(), 0.0f);
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz) {
        const size_t k = idx(iz, ix);
        if (s0[k] == 0 || rcut[k] >= 0) { s1[k] = 0; t1[k] = f.ttn[k]; }
        else if (close[k]) { s1[k] = 1; t1[k] = f.ttn[k]; }
      }
    s0 = s1; t0 = t1;
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] >= 0 && rank[k] < prefix) rank[k] = -1;
  }
  std::vector<float> g(f.ttn.size(), 0.0f);
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) g[idx(iz, ix)] = guess[(size_t)(ix - 1) * f.nnz + (iz - 1)];
  std::vector<int> rk = ranks_by_value(f, g, rank);
  *rounds = 0;
  for (int it = 0; it < max_rounds; ++it) {
    R.run(rk, nullptr);
    ++*rounds;
    const std::vector<float> v = f.ttn;
    std::vector<int> nr = ranks_by_value(f, v, rank);
    if (nr == rk) break;
    rk = nr;
  }
  R.run(rk, nullptr);
  long m = 0;
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz)
      if (f.ttn[idx(iz, ix)] != ref[idx(iz, ix)]) ++m;
  *mismatch = m;
  long of = 0, kf = 0, ke = 0;
  verify_replay(f, s0, t0, rk, &of, &kf, &ke);
  *flags = of + kf;
  return 0;
}

namespace orc {
namespace {
// the refined source box exactly as solve_source sets it up (FwdTraveltimeCPS.f90:493-560), without the march
int setup_refined(Fmm& f, const double* pv, float x, float z) {
  f.nnx = f.nnx_c; f.nnz = f.nnz_c; f.dnx = f.dnx_c; f.dnz = f.dnz_c; f.gox = f.gox_c; f.goz = f.goz_c;
  f.gridder(pv);
  int isx = (int)((x - f.gox) / f.dnx) + 1;
  int isz = (int)((z - f.goz) / f.dnz) + 1;
  if (isx < 1 || isx > f.nnx || isz < 1 || isz > f.nnz) return ERR_SOURCE_OUTSIDE;
  if (isx == f.nnx) isx = isx - 1;
  if (isz == f.nnz) isz = isz - 1;
  f.vnl = isx - f.sgs; if (f.vnl < 1) f.vnl = 1;
  f.vnr = isx + f.sgs; if (f.vnr > f.nnx) f.vnr = f.nnx;
  f.vnt = isz - f.sgs; if (f.vnt < 1) f.vnt = 1;
  f.vnb = isz + f.sgs; if (f.vnb > f.nnz) f.vnb = f.nnz;
  f.nrnx = (f.vnr - f.vnl) * f.sgdl + 1;
  f.nrnz = (f.vnb - f.vnt) * f.sgdl + 1;
  f.drnx = f.dvx / (float)(f.gdx * f.sgdl);
  f.drnz = f.dvz / (float)(f.gdz * f.sgdl);
  f.gorx = f.gox + f.dnx * (float)(f.vnl - 1);
  f.gorz = f.goz + f.dnz * (float)(f.vnt - 1);
  f.nnx = f.nrnx; f.nnz = f.nrnz; f.dnx = f.drnx; f.dnz = f.drnz; f.gox = f.gorx; f.goz = f.gorz;
  f.bsplrefine();
  return OK;
}
// the march stops when a node on an edge of the box that is not an edge of the model reaches the heap root
bool is_exit_node(const Fmm& f, int iz, int ix) {
  if (ix == 1 && f.vnl != 1) return true;
  if (ix == f.nnx && f.vnr != f.nnx) return true;      // literal (CalSurfG.f90:369-371)
  if (iz == 1 && f.vnt != 1) return true;
  if (iz == f.nnz && f.vnb != f.nnz) return true;
  return false;
}
}  // namespace
}  // namespace orc

// ORDER EXPERIMENT on the refined source box (129 x 129 around the source, 40x finer than the model grid), including
// its stopping rule and the trial values of the close nodes it hands to the coarse grid.
// prefix > 0: the first `prefix` accepts are taken from the reference (= marched serially with the heap, as today) and
// only the rest of the box is predicted and replayed -- the hazards of this stage sit around the source cell.
extern "C" int orc_fmm_order_stats_refined(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                                           float scx, float scz, int prefix, orc::OrderStats* out) {
  using namespace orc;
  const bool guess_mode = prefix >= 1000;      // prefix = 1000 + p: rank iteration from the distance guess, serial prefix p
  if (guess_mode) prefix -= 1000;
  Fmm f;
  f.init(nx, ny, goxd, gozd, dvxd, dvzd);
  int st = setup_refined(f, pv, scx, scz);
  if (st) return st;
  std::vector<int> rank(f.ttn.size(), -1);
  f.rec_rank = &rank; f.rec_count = 0;
  st = f.travel(scx, scz, 1);
  if (st) return st;
  f.rec_rank = nullptr;
  const int npop = f.rec_count;
  const std::vector<float> ref = f.ttn;
  const std::vector<int> refs = f.nsts;
  auto idx = [&](int iz, int ix) { return (size_t)(ix - 1) * f.ld + (iz - 1); };
  // initial state of this march: everything far except the four corners of the source cell (close, analytic times)
  std::vector<int> s0(f.ttn.size(), -1);
  std::vector<float> t0(f.ttn.size(), 0.0f);
  {
    int isx = (int)((scx - f.gox) / f.dnx) + 1, isz = (int)((scz - f.goz) / f.dnz) + 1;
    if (isx == f.nnx) isx = isx - 1;
    if (isz == f.nnz) isz = isz - 1;
    float vss[2][2];
    for (int i = 1; i <= 2; ++i)
      for (int j = 1; j <= 2; ++j) vss[i - 1][j - 1] = f.VELN(isz - 1 + j, isx - 1 + i);
    const float dsx = (scx - f.gox) - (float)(isx - 1) * f.dnx, dsz = (scz - f.goz) - (float)(isz - 1) * f.dnz;
    const float vsrc = f.bilinear(vss, dsx, dsz);
    for (int i = 1; i <= 2; ++i)
      for (int j = 1; j <= 2; ++j) {
        const float ax = dsx - (float)(i - 1) * f.dnx, az = dsz - (float)(j - 1) * f.dnz;
        const float ds = std::sqrt(ax * ax + az * az);
        s0[idx(isz - 1 + j, isx - 1 + i)] = 1;
        t0[idx(isz - 1 + j, isx - 1 + i)] = 2.0f * ds / (vss[i - 1][j - 1] + vsrc);
      }
  }
  // the node the march stopped at is alive but was never counted by the hook: give it the last rank
  int exit_rank = -1;
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz)
      if (refs[idx(iz, ix)] == 0 && rank[idx(iz, ix)] < 0) { rank[idx(iz, ix)] = npop; exit_rank = npop; }
  OrderStats S{};
  auto compare = [&](const std::vector<int>& rk, const std::vector<char>& close) {
    long m = 0;
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz) {
        const size_t k = idx(iz, ix);
        const int want = refs[k] == 0 ? 0 : (refs[k] > 0 ? 1 : -1);
        const int got = (rk[k] >= 0 || s0[k] == 0) ? 0 : (close[k] ? 1 : -1);
        if (want != got) { ++m; continue; }
        if (want >= 0 && f.ttn[k] != ref[k]) ++m;          // alive values and the trial values of close nodes
      }
    return m;
  };
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz)
      if (rank[idx(iz, ix)] >= 0) ++S.popped;
  Replay R{f, s0, t0};
  std::vector<int> level;
  std::vector<char> close;
  R.run(rank, &level, exit_rank, &close);
  S.rule_mismatch = compare(rank, close);
  for (int v : level) S.dag_levels = std::max<long>(S.dag_levels, v);
  const int sz[8] = {-1, -2, 1, 2, 0, 0, 0, 0}, sx[8] = {0, 0, 0, 0, -1, -2, 1, 2};
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) {
      const int r = rank[idx(iz, ix)];
      if (r < 0) continue;
      for (int q = 0; q < 8; ++q) {
        const int z = iz + sz[q], x = ix + sx[q];
        if (z < 1 || z > f.nnz || x < 1 || x > f.nnx) continue;
        const int rn = rank[idx(z, x)];
        if (rn < 0 || rn > r) continue;
        ++S.pairs;
        if (ref[idx(z, x)] == ref[idx(iz, ix)]) ++S.pair_ties;
        else if (ref[idx(z, x)] > ref[idx(iz, ix)]) ++S.pair_inversions;
      }
    }
  // optional serial prefix: state of the march after `prefix` accepts, by the (exact) rule
  if (prefix > 0 && prefix < npop) {
    std::vector<int> rcut(rank.size(), -1);
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] >= 0 && rank[k] < prefix) rcut[k] = rank[k];
    R.run(rcut, nullptr, prefix, &close);
    std::vector<int> s1(rank.size(), -1);
    std::vector<float> t1(rank.size(), 0.0f);
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz) {
        const size_t k = idx(iz, ix);
        if (rcut[k] >= 0) { s1[k] = 0; t1[k] = f.ttn[k]; }
        else if (close[k]) { s1[k] = 1; t1[k] = f.ttn[k]; }
      }
    s0 = s1; t0 = t1;          // Replay R refers to s0 / t0: from here on the march starts at this state
    for (size_t k = 0; k < rank.size(); ++k)
      if (rank[k] >= 0 && rank[k] < prefix) rank[k] = -2;      // accepted before the replayed part began
  }
  // ranks: sorted order of a field, cut at the first stopping-edge node
  std::vector<int> prank;
  int pexit = -1;
  auto ranks_from = [&](const std::vector<float>& val) {
    std::vector<std::pair<float, int>> v;
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz)
        if (val[idx(iz, ix)] < FIM_INF && s0[idx(iz, ix)] != 0) v.push_back({val[idx(iz, ix)], (int)idx(iz, ix)});
    std::stable_sort(v.begin(), v.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
    std::vector<int> r(f.ttn.size(), -1);
    pexit = -1;
    for (size_t i = 0; i < v.size(); ++i) {
      const int k = v[i].second;
      r[k] = (int)i;
      if (is_exit_node(f, k % f.ld + 1, k / f.ld + 1)) { pexit = (int)i; break; }
    }
    return r;
  };
  if (guess_mode) {
    // RANK ITERATION from a trivial guess (distance from the source in grid metric): ranks -> replay -> sort -> ranks.
    // Close nodes enter the next sort with their trial keys (the keys the heap would hold), far nodes with +infinity.
    int isx = (int)((scx - f.gox) / f.dnx), isz = (int)((scz - f.goz) / f.dnz);
    std::vector<float> cur(f.ttn.size(), FIM_INF);
    for (int ix = 1; ix <= f.nnx; ++ix)
      for (int iz = 1; iz <= f.nnz; ++iz) {
        const float ax = ((float)(ix - 1 - isx) - 0.5f) * f.dnx * f.earth;
        const float az = ((float)(iz - 1 - isz) - 0.5f) * f.dnz * f.earth * sin_rf(f.gox + (float)(ix - 1) * f.dnx);
        cur[idx(iz, ix)] = std::sqrt(ax * ax + az * az);
      }
    prank = ranks_from(cur);
    for (int it = 0; it < 60; ++it) {
      R.run(prank, nullptr, pexit, &close);
      ++S.fim_passes;                                      // rounds of the rank iteration
      for (int ix = 1; ix <= f.nnx; ++ix)
        for (int iz = 1; iz <= f.nnz; ++iz) {
          const size_t k = idx(iz, ix);
          cur[k] = (prank[k] >= 0 || close[k]) ? f.ttn[k] : FIM_INF;
        }
      const int old_exit = pexit;
      std::vector<int> nr = ranks_from(cur);
      if (nr == prank && pexit == old_exit) break;
      prank = nr;
    }
  } else {
    // ranks predicted from the order-free fixed point on the whole box
    f.nsts = s0; f.ttn = t0;
    if (const char* e = std::getenv("ORC_FIM_ACTIVE_LIST")) f.fim_coarse = std::atoi(e) ? 2 : 0;
    f.travel_fim();
    S.fim_passes = f.fim_sweeps;
    S.fim_evals = f.fim_evals;
    const std::vector<float> fimv = f.ttn;
    prank = ranks_from(fimv);
  }
  f.nsts = s0;
  S.sorted_fim_rank_errors = R.run(prank, nullptr, pexit, &close);
  S.sorted_fim_mismatch = compare(prank, close);
  S.sorted_exact_mismatch = -1;                          // not evaluated for this stage
  verify_replay(f, s0, t0, prank, &S.verify_order_flags, &S.verify_key_increase_flags, &S.key_increase_events, pexit,
                &S.harmless_tie_groups);
  *out = S;
  return 0;
}

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
