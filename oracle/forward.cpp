// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the orchestrators and depth-kernel drivers:
//   depthkernel          src/src_inv_iso_joint/CalSurfG.f90:1-139
//   depthkernelTI        src/src_forward/depthkernelTI.f90:2-112
//   refineLayerMdl       src/src_forward/FwdTraveltimeCPS.f90:62-112
//   FwdObsTraveltimeCPS  src/src_forward/FwdTraveltimeCPS.f90:208-784
//   CalSurfG             src/src_inv_iso_joint/CalSurfG.f90:909-1421
//   CalSurfGAnisoJoint   src/src_inv_iso_joint/CalSurfGAniso_Joint.f90:209-827
// The (period, source) loop may be split over host threads (each source is
// independent in the reference too; it just never exploits that).  Results do
// not depend on the thread count.
#include "oracle.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <chrono>
#include <algorithm>

namespace orc {

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// x**n as GCC -O expands __builtin_powif (tree-ssa-math-opts powi table)
static inline float p2(float x) { return x * x; }
static inline float p3(float x) { return x * (x * x); }
static inline float p4(float x) { return (x * x) * (x * x); }
static inline float p5(float x) { return (x * x) * (x * (x * x)); }

// depthkernelTI.f90:53-59 == CalSurfG.f90:48-54
void brocher(float vs, float* vp_out, float* rho_out) {
  float vp = 0.9409f + 2.0947f * vs - 0.8206f * p2(vs) + 0.2683f * p3(vs) - 0.0251f * p4(vs);
  float rho = 1.6612f * vp - 0.4721f * p2(vp) + 0.0671f * p3(vp) - 0.0043f * p4(vp) + 0.000106f * p5(vp);
  *vp_out = vp;
  *rho_out = rho;
}

// FwdTraveltimeCPS.f90:62-112 (refineLayerMdl) == CalSurfG.f90:2317 (refineGrid2LayerMdl)
void refine_layer_mdl(float minthk0, int mmax, const float* dep, const float* vp,
                      const float* vs, const float* rho, int* rmax, float* rdep, float* rvp,
                      float* rvs, float* rrho, float* rthk, int* nsublay) {
  int k = 0;
  float initdep = 0.0f;
  for (int i = 1; i <= mmax - 1; ++i) {
    float thk = dep[i] - dep[i - 1];
    float minthk = thk / minthk0;
    int ns = (int)((thk + 1.0e-4f) / minthk) + 1;
    if (nsublay) nsublay[i - 1] = ns;
    float newthk = thk / (float)ns;
    for (int j = 1; j <= ns; ++j) {
      k = k + 1;
      rthk[k - 1] = newthk;
      rdep[k - 1] = initdep + rthk[k - 1];
      initdep = rdep[k - 1];
      rvp[k - 1] = vp[i - 1] + (float)(2 * j - 1) * (vp[i] - vp[i - 1]) / (float)(2 * ns);
      rvs[k - 1] = vs[i - 1] + (float)(2 * j - 1) * (vs[i] - vs[i - 1]) / (float)(2 * ns);
      rrho[k - 1] = rho[i - 1] + (float)(2 * j - 1) * (rho[i] - rho[i - 1]) / (float)(2 * ns);
    }
  }
  k = k + 1;
  rthk[k - 1] = 0.0f;
  rvp[k - 1] = vp[mmax - 1];
  rvs[k - 1] = vs[mmax - 1];
  rrho[k - 1] = rho[mmax - 1];
  rdep[k - 1] = dep[mmax - 1];
  *rmax = k;
}

template <class F>
static void parallel_for(int n, int nthreads, F f) {
  if (nthreads <= 1 || n <= 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> th;
  nthreads = std::min(nthreads, n);
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=]() {
      for (int i = t; i < n; i += nthreads) f(i);
    });
  for (auto& x : th) x.join();
}

// CalSurfG.f90:1-139.  nthreads mirrors the reference's one OpenMP loop
// (serial semantics of the SAVEd variables are used in every thread).
int depthkernel(int nx, int ny, int nz, const float* vel, double* pvRc, double* sen_vs,
                double* sen_vp, double* sen_rho, int kmaxRc, const double* tRc,
                const float* depz, float minthk, int nthreads, long* neval_out) {
  const int NLc = 200;
  if (nz > NLc) return ERR_LAYERS;
  const size_t nxy = (size_t)nx * ny;
  const int mmax = nz;
  const float dlnVs = 0.01f, dlnVp = 0.01f, dlnrho = 0.01f;
  std::vector<long> nev((size_t)nx * ny, 0);
  std::vector<int> stv((size_t)nx * ny, 0);
  parallel_for(nx * ny, nthreads, [&](int node) {
    const int ii = node % nx + 1, jj = node / nx + 1;
    std::vector<float> vsz(nz), vpz(nz), rhoz(nz), depm(nz), vsm(nz), vpm(nz), rhom(nz);
    float rdep[NLc], rvp[NLc], rvs[NLc], rrho[NLc], rthk[NLc];
    double cg1[60], cg2[60], cgRc[60];
    int rmax;
    long ne = 0, n1;
    for (int k = 0; k < nz; ++k) vsz[k] = vel[(size_t)(ii - 1) + (size_t)(jj - 1) * nx + (size_t)k * nxy];
    for (int k = 0; k < nz; ++k) brocher(vsz[k], &vpz[k], &rhoz[k]);
    refine_layer_mdl(minthk, mmax, depz, vpz.data(), vsz.data(), rhoz.data(), &rmax, rdep, rvp, rvs, rrho, rthk, nullptr);
    int st = surfdisp96(rthk, rvp, rvs, rrho, rmax, 1, 2, 1, 0, kmaxRc, tRc, cgRc, &n1);
    if (st) { stv[node] = st; return; }
    ne += n1;
    const size_t p = (size_t)(jj - 1) * nx + (ii - 1);
    for (int k = 0; k < kmaxRc; ++k) pvRc[p + (size_t)k * nxy] = cgRc[k];
    for (int k = 0; k < mmax; ++k) {
      depm[k] = depz[k]; vsm[k] = vsz[k]; vpm[k] = vpz[k]; rhom[k] = rhoz[k];
    }
    auto run = [&](double* cg) {
      refine_layer_mdl(minthk, mmax, depm.data(), vpm.data(), vsm.data(), rhom.data(), &rmax, rdep, rvp, rvs, rrho, rthk, nullptr);
      surfdisp96(rthk, rvp, rvs, rrho, rmax, 1, 2, 1, 0, kmaxRc, tRc, cg, &n1);
      ne += n1;
    };
    for (int i = 0; i < mmax; ++i) {
      vsm[i] = vsz[i] - 0.5f * dlnVs * vsz[i];
      run(cg1);
      vsm[i] = vsz[i] + 0.5f * dlnVs * vsz[i];
      run(cg2);
      vsm[i] = vsz[i];
      for (int nn = 0; nn < kmaxRc; ++nn)
        sen_vs[p + (size_t)nn * nxy + (size_t)i * nxy * kmaxRc] = (cg2[nn] - cg1[nn]) / (double)(dlnVs * vsz[i]);
      vpm[i] = vpz[i] - 0.5f * dlnVp * vpz[i];
      run(cg1);
      vpm[i] = vpz[i] + 0.5f * dlnVp * vpz[i];
      run(cg2);
      vpm[i] = vpz[i];
      for (int nn = 0; nn < kmaxRc; ++nn)
        sen_vp[p + (size_t)nn * nxy + (size_t)i * nxy * kmaxRc] = (cg2[nn] - cg1[nn]) / (double)(dlnVp * vpz[i]);
      rhom[i] = rhoz[i] - 0.5f * dlnrho * rhoz[i];
      run(cg1);
      rhom[i] = rhoz[i] + 0.5f * dlnrho * rhoz[i];
      run(cg2);
      rhom[i] = rhoz[i];
      for (int nn = 0; nn < kmaxRc; ++nn)
        sen_rho[p + (size_t)nn * nxy + (size_t)i * nxy * kmaxRc] = (cg2[nn] - cg1[nn]) / (double)(dlnrho * rhoz[i]);
    }
    nev[node] = ne;
  });
  long tot = 0;
  for (size_t i = 0; i < nev.size(); ++i) { tot += nev[i]; if (stv[i]) return stv[i]; }
  if (neval_out) *neval_out = tot;
  return OK;
}

// depthkernelTI.f90:2-112 (serial in the reference: tregn96 keeps COMMON state)
int depthkernel_ti(int nx, int ny, int nz, const float* vel, double* pvRc, int kmaxRc,
                   const double* tRc, const float* depz, float minthk, float* Lsen_Gsc,
                   int nthreads) {
  const int NLc = 200, NPc = 60;
  if (nz > NLc || kmaxRc > NPc) return ERR_LAYERS;
  const size_t nxy = (size_t)nx * ny;
  const int mmax = nz;
  std::memset(pvRc, 0, sizeof(double) * nxy * kmaxRc);
  std::memset(Lsen_Gsc, 0, sizeof(float) * nxy * kmaxRc * (nz - 1));
  std::vector<int> stv(nxy, 0);
  parallel_for(nx * ny, nthreads, [&](int node) {
    const int ii = node % nx + 1, jj = node / nx + 1;
    const size_t post = (size_t)ii + (size_t)(jj - 1) * nx;  // 1-based
    std::vector<float> vsz(nz), vpz(nz), rhoz(nz);
    float rdep[NLc], rvp[NLc], rvs[NLc], rrho[NLc], rthk[NLc];
    float TA[NLc], TC[NLc], TF[NLc], TL[NLc], TN[NLc], TRho[NLc];
    float qp[NLc], qs[NLc], etap[NLc], etas[NLc], frefp[NLc], frefs[NLc];
    int nsublay[NLc];
    double cgRc[NPc];
    float t_in[NPc], cp_in[NPc];
    std::vector<float> dcdah((size_t)NPc * NLc, 0.f), dcdbv((size_t)NPc * NLc, 0.f), dcdn((size_t)NPc * NLc, 0.f);
    int rmax;
    for (int k = 0; k < nz; ++k) vsz[k] = vel[(size_t)(ii - 1) + (size_t)(jj - 1) * nx + (size_t)k * nxy];
    for (int k = 0; k < nz; ++k) brocher(vsz[k], &vpz[k], &rhoz[k]);
    refine_layer_mdl(minthk, mmax, depz, vpz.data(), vsz.data(), rhoz.data(), &rmax, rdep, rvp, rvs, rrho, rthk, nsublay);
    int st = surfdisp96(rthk, rvp, rvs, rrho, rmax, 1, 2, 1, 0, kmaxRc, tRc, cgRc, nullptr);
    if (st) { stv[node] = st; return; }
    for (int k = 0; k < kmaxRc; ++k) pvRc[(post - 1) + (size_t)k * nxy] = cgRc[k];
    for (int i = 0; i < rmax; ++i) {
      TA[i] = rrho[i] * (rvp[i] * rvp[i]);
      TC[i] = TA[i];
      TL[i] = rrho[i] * (rvs[i] * rvs[i]);
      TN[i] = TL[i];
      TF[i] = 1.0f * (TA[i] - 2 * TL[i]);
      TRho[i] = rrho[i];
      qp[i] = 150.0f; qs[i] = 50.0f; etap[i] = 0.0f; etas[i] = 0.0f; frefp[i] = 1.0f; frefs[i] = 1.0f;
    }
    for (int k = 0; k < kmaxRc; ++k) { cp_in[k] = (float)cgRc[k]; t_in[k] = (float)tRc[k]; }
    st = tregn96(rmax, rthk, TA, TC, TF, TL, TN, TRho, qp, qs, etap, etas, frefp, frefs,
                 kmaxRc, t_in, cp_in, dcdah.data(), dcdbv.data(), dcdn.data());
    if (st) { stv[node] = st; return; }
    for (int i = 1; i <= kmaxRc; ++i) {
      int k = 0;
      for (int j = 1; j <= nz - 1; ++j) {
        float& L = Lsen_Gsc[(post - 1) + (size_t)(i - 1) * nxy + (size_t)(j - 1) * nxy * kmaxRc];
        for (int jjj = 1; jjj <= nsublay[j - 1]; ++jjj) {
          k = k + 1;
          const size_t q = (size_t)(i - 1) + (size_t)(k - 1) * NPc;
          float den = (TA[k - 1] - 2.0f * TL[k - 1]);
          den = den * den;
          float dcR_dA = 0.5f / (rrho[k - 1] * rvp[k - 1]) * dcdah[q] - TF[k - 1] / den * dcdn[q];
          float dcR_dL = 0.5f / (rrho[k - 1] * rvs[k - 1]) * dcdbv[q] + 2.0f * TF[k - 1] / den * dcdn[q];
          L = L + dcR_dA * TA[k - 1] + dcR_dL * TL[k - 1];
        }
      }
    }
  });
  for (size_t i = 0; i < nxy; ++i) if (stv[i]) return stv[i];
  return OK;
}

namespace {
struct Coo { std::vector<float> rw; std::vector<int> row, col; };

struct SrcItem { int knumi, srcnum; long row0; };
}  // namespace

int gbuild(GBuild& g) {
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  const size_t nxy = (size_t)nx * ny;
  const int kmax = g.sv.kmax, nsrc = g.sv.nsrc, nrcf = g.sv.nrcf;
  const int nvx = nx - 2, nvz = ny - 2;
  const long nparpi = (long)nvx * nvz * (nz - 1);
  const float ftol = 1e-4f;
  g.times = StageTimes{0, 0, 0, 0, 0, 0};
  g.nar = 0;
  g.rbint = 0;
  double t0 = now_s();
  if (!g.precomputed) {
    int st = OK;
    if (g.mode == 0 || g.mode == 2)
      st = depthkernel_ti(nx, ny, nz, g.vels, g.pvRc, g.kmaxRc, g.tRc, g.depz, g.minthk, g.Lsen_Gsc, g.nthreads);
    if (st) return st;
    if (g.mode == 1 || g.mode == 2)
      st = depthkernel(nx, ny, nz, g.vels, g.pvRc, g.sen_vs, g.sen_vp, g.sen_rho, g.kmaxRc, g.tRc, g.depz, g.minthk, g.nthreads, nullptr);
    if (st) return st;
  }
  g.times.kernels_s = now_s() - t0;

  // (period, source) work list in the reference's loop order with the first
  // global row id (count1) of every source
  std::vector<SrcItem> items;
  long count1 = 0;
  for (int knumi = 1; knumi <= kmax; ++knumi)
    for (int srcnum = 1; srcnum <= g.sv.nsrcsurf1[knumi - 1]; ++srcnum) {
      items.push_back({knumi, srcnum, count1});
      count1 += g.sv.nrc1[(size_t)(srcnum - 1) + (size_t)(knumi - 1) * nsrc];
    }
  const long dall = count1;
  const int nthreads = std::max(1, std::min<int>(g.nthreads, (int)items.size()));
  // contiguous chunks balanced by ray count
  std::vector<size_t> bounds(nthreads + 1, items.size());
  bounds[0] = 0;
  {
    size_t it = 0;
    for (int t = 1; t < nthreads; ++t) {
      long target = dall * t / nthreads;
      while (it < items.size() && items[it].row0 < target) ++it;
      bounds[t] = it;
    }
  }
  std::vector<Coo> coo(nthreads);
  std::vector<int> stv(nthreads, 0), rbv(nthreads, 0);
  std::vector<StageTimes> tv(nthreads, StageTimes{0, 0, 0, 0, 0, 0});

  // GcCol/GsCol (FwdTraveltimeCPS.f90:753-758)
  std::vector<float> GcCol, GsCol;
  if (g.mode == 0) {
    GcCol.assign(nparpi, 0.f); GsCol.assign(nparpi, 0.f);
    for (int jj = 1; jj <= ny - 2; ++jj)
      for (int kk = 1; kk <= nx - 2; ++kk)
        for (int k = 1; k <= nz - 1; ++k) {
          size_t src = (size_t)(kk - 1) + (size_t)(jj - 1) * (nx - 2) + (size_t)(k - 1) * (nx - 2) * (ny - 2);
          size_t nn = (size_t)(k - 1) * nvx * nvz + (size_t)(jj - 1) * nvx + kk;
          GcCol[nn - 1] = g.Gctrue[src];
          GsCol[nn - 1] = g.Gstrue[src];
        }
  }

  auto worker = [&](int t) {
    Fmm f;
    f.init(nx, ny, g.goxd, g.gozd, g.dvxd, g.dvzd);
#ifdef ORC_WITH_EXPERIMENTS
    if (const char* e = std::getenv("ORC_FIM_EXPERIMENT")) f.fim_coarse = std::atoi(e);   // fim_experiment.cpp
#endif
    const size_t nf = (size_t)(nvz + 2) * (nvx + 2);
    const int ldf = nvz + 2;
    std::vector<float> fdm(nf), fdmc(nf), fdms(nf);
    std::vector<double> velf(nxy);
    std::vector<float> coe_a(nz), coe_rho(nz);
    std::vector<int> cells;   // footprint cells with |fdm| >= ftol, in (jj,kk) order
    Coo& out = coo[t];
    StageTimes& tm = tv[t];
    for (size_t it = bounds[t]; it < bounds[t + 1]; ++it) {
      const int knumi = items[it].knumi, srcnum = items[it].srcnum;
      const size_t sk = (size_t)(srcnum - 1) + (size_t)(knumi - 1) * nsrc;
      const int per = g.sv.periods[sk];
      for (size_t p = 0; p < nxy; ++p) velf[p] = g.pvRc[p + (size_t)(per - 1) * nxy];
      const float x = g.sv.scxf[sk], z = g.sv.sczf[sk];
      double ta = now_s();
      int st = f.solve_source(velf.data(), x, z);
      double tb = now_s();
      tm.dice_fmm_s += tb - ta;
      if (st) { stv[t] = st; return; }
      long row = items[it].row0;
      const int nr = g.sv.nrc1[sk];
      for (int istep = 1; istep <= nr; ++istep) {
        const size_t rk = (size_t)(istep - 1) + (size_t)(srcnum - 1) * nrcf + (size_t)(knumi - 1) * nrcf * nsrc;
        const float rx = g.sv.rcxf[rk], rz = g.sv.rczf[rk];
        double tc = now_s();
        float cbst1;
        st = f.srtimes(x, z, rx, rz, &cbst1);
        if (st) { stv[t] = st; return; }
        row = row + 1;  // count1
        g.dsurf[row - 1] = cbst1;
        st = f.rpaths(x, z, fdm.data(), fdmc.data(), fdms.data(), rx, rz, g.mode != 1);
        if (st) { stv[t] = st; return; }
        double td = now_s();
        tm.trace_s += td - tc;
        cells.clear();
        for (int jj = 1; jj <= nvz; ++jj)
          for (int kk = 1; kk <= nvx; ++kk)
            if (std::fabs(fdm[(size_t)kk * ldf + jj]) >= ftol) cells.push_back(jj * (nvx + 2) + kk);
        if (g.mode == 0) {
          // obsTgc/obsTgs = MATMUL(GGc,GcCol)/MATMUL(GGs,GsCol): ascending-column
          // float32 accumulation; zero entries of the dense row add +0.0
          float sgc = 0.0f, sgs = 0.0f;
          for (int k = 1; k <= nz - 1; ++k)
            for (int c : cells) {
              const int jj = c / (nvx + 2), kk = c % (nvx + 2);
              const size_t node = (size_t)jj * (nvx + 2) + kk + 1;
              const float L = g.Lsen_Gsc[(node - 1) + (size_t)(knumi - 1) * nxy + (size_t)(k - 1) * nxy * g.kmaxRc];
              const size_t nn = (size_t)(k - 1) * nvx * nvz + (size_t)(jj - 1) * nvx + kk;
              sgc = sgc + (L * fdmc[(size_t)kk * ldf + jj]) * GcCol[nn - 1];
              sgs = sgs + (L * fdms[(size_t)kk * ldf + jj]) * GsCol[nn - 1];
            }
          g.obsTaa[row - 1] = sgc + sgs;
        } else {
          const int nblk = (g.mode == 2) ? 3 : 1;
          for (int blk = 0; blk < nblk; ++blk)
            for (int k = 1; k <= nz - 1; ++k)
              for (int c : cells) {
                const int jj = c / (nvx + 2), kk = c % (nvx + 2);
                const size_t node = (size_t)jj * (nvx + 2) + kk + 1;
                float val;
                if (blk == 0) {
                  const float v = g.vels[(size_t)kk + (size_t)jj * nx + (size_t)(k - 1) * nxy];  // vels(kk+1,jj+1,k)
                  const float ca = (2.0947f - (0.8206f * 2) * v + (0.2683f * 3) * p2(v) - (0.0251f * 4) * p3(v));
                  const float vpft = 0.9409f + 2.0947f * v - 0.8206f * p2(v) + 0.2683f * p3(v) - 0.0251f * p4(v);
                  const float cr = ca * (1.6612f - (0.4721f * 2) * vpft + (0.0671f * 3) * p2(vpft) -
                                         (0.0043f * 4) * p3(vpft) + (0.000106f * 5) * p4(vpft));
                  const size_t q = (node - 1) + (size_t)(knumi - 1) * nxy + (size_t)(k - 1) * nxy * g.kmaxRc;
                  const double r = (g.sen_vp[q] * (double)ca + g.sen_rho[q] * (double)cr + g.sen_vs[q]) *
                                   (double)fdm[(size_t)kk * ldf + jj];
                  val = (float)r;
                } else {
                  const float L = g.Lsen_Gsc[(node - 1) + (size_t)(knumi - 1) * nxy + (size_t)(k - 1) * nxy * g.kmaxRc];
                  val = L * (blk == 1 ? fdmc[(size_t)kk * ldf + jj] : fdms[(size_t)kk * ldf + jj]);
                }
                if (std::fabs(val) > ftol) {
                  const long nn = (long)blk * nparpi + (long)(k - 1) * nvx * nvz + (long)(jj - 1) * nvx + kk;
                  out.rw.push_back(val);
                  out.row.push_back((int)row);
                  out.col.push_back((int)nn);
                }
              }
        }
        tm.assemble_s += now_s() - td;
      }
    }
    tm.n_accept = f.n_accept;
    tm.n_steps = f.n_steps;
    rbv[t] = f.rbint;
  };
  if (nthreads == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
  }
  for (int t = 0; t < nthreads; ++t) {
    if (stv[t]) return stv[t];
    g.rbint |= rbv[t];
    g.times.dice_fmm_s = std::max(g.times.dice_fmm_s, tv[t].dice_fmm_s);
    g.times.trace_s = std::max(g.times.trace_s, tv[t].trace_s);
    g.times.assemble_s = std::max(g.times.assemble_s, tv[t].assemble_s);
    g.times.n_accept += tv[t].n_accept;
    g.times.n_steps += tv[t].n_steps;
  }
  if (g.mode != 0) {
    long nar = 0;
    for (int t = 0; t < nthreads; ++t) nar += (long)coo[t].rw.size();
    g.nar = nar;
    if (nar > g.maxnar) return ERR_NNZ_OVERFLOW;
    long o = 0;
    for (int t = 0; t < nthreads; ++t) {
      const size_t n = coo[t].rw.size();
      if (n) {
        std::memcpy(g.rw + o, coo[t].rw.data(), n * sizeof(float));
        std::memcpy(g.iw_row + o, coo[t].row.data(), n * sizeof(int));
        std::memcpy(g.col + o, coo[t].col.data(), n * sizeof(int));
      }
      o += (long)n;
    }
  }
  // tRcV (FwdTraveltimeCPS.f90:771-778)
  if (g.tRcV) {
    for (int tt = 1; tt <= g.kmaxRc; ++tt)
      for (int jj = 1; jj <= ny - 2; ++jj)
        for (int ii = 1; ii <= nx - 2; ++ii)
          g.tRcV[(size_t)(jj - 1) * (nx - 2) + (ii - 1) + (size_t)(tt - 1) * (nx - 2) * (ny - 2)] =
              g.pvRc[(size_t)jj * nx + ii + (size_t)(tt - 1) * nxy];
  }
  return OK;
}

}  // namespace orc
