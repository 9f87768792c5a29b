// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the reference's 2-D eikonal / ray-tracing stage:
//   module globalp                src/src_inv_iso_joint/CalSurfG.f90:151-217
//   module traveltime (FMM)       CalSurfG.f90:234-893
//   gridder / bsplrefine          CalSurfG.f90:1423 / :1525
//   srtimes / rpaths / bilinear   CalSurfG.f90:1599 / :1735 / :2293
//   rpathsAzim / azdist           rpathsAzim.f90:16 / :687
//   per-source refined->coarse    FwdTraveltimeCPS.f90:467-645
// All module-global state lives in one struct so that independent sources can
// be solved on independent host threads (the reference is single threaded).
// Every REAL(KIND=i10) is float; arithmetic order follows the Fortran source.
#pragma once
#include <vector>
#include <cstdint>

namespace orc {

enum Status {
  OK = 0,
  ERR_SOURCE_OUTSIDE = 1,    // CalSurfG.f90:287-293 / FwdTraveltimeCPS.f90:498-504
  ERR_RECEIVER_OUTSIDE = 2,  // CalSurfG.f90:1649-1655, rpathsAzim.f90:193-216
  ERR_NNZ_OVERFLOW = 3,      // Main_Jt.f90:523
  ERR_BAD_ARG = 4,
  ERR_LAYERS = 5
};

struct Fmm {
  // ---- globalp scalars ----
  int nvx = 0, nvz = 0, nnx = 0, nnz = 0, fom = 1, gdx = 5, gdz = 5;
  int vnl = 0, vnr = 0, vnt = 0, vnb = 0, nrnx = 0, nrnz = 0, sgdl = 8, rbint = 0;
  int nnxr = 0, nnzr = 0, asgr = 1, sgs = 8;
  float gox = 0, goz = 0, dnx = 0, dnz = 0, dvx = 0, dvz = 0, snb = 0.5f, earth = 6371.0f;
  float goxd = 0, gozd = 0, dvxd = 0, dvzd = 0, dnxd = 0, dnzd = 0;
  float drnx = 0, drnz = 0, gorx = 0, gorz = 0;
  float dnxr = 0, dnzr = 0, goxr = 0, gozr = 0;
  // coarse-grid constants kept across sources (the reference backs them up in
  // nnxb/nnzb/dnxb/dnzb/goxb/gozb around each refined solve)
  int nnx_c = 0, nnz_c = 0;
  float dnx_c = 0, dnz_c = 0, gox_c = 0, goz_c = 0;
  // ---- arrays (column-major, first index = z) ----
  int ld = 0;        // leading dimension of veln/ttn/nsts (>= any nnz used)
  int ldr = 0;       // leading dimension of ttnr/nstsr
  int ldv = 0;       // leading dimension of velv = nvz+2
  std::vector<float> velv, veln, velnb, ttn, ttnr;
  std::vector<int> nsts, nstsr;
  // ---- binary heap ----
  int ntr = 0;
  std::vector<int> btg_px, btg_pz;
  // ---- EXPERIMENT switch (fim_experiment.cpp; never set by tests of the reference path) ----
  // 1: the coarse-grid continuation is solved as a fast-iterative fixed point instead of the heap march
  std::vector<int>* rec_rank = nullptr;   // experiment: if set, rec_rank[(ix-1)*ld+(iz-1)] = pop counter of travel()
  int rec_count = 0;
  std::vector<int>* rec_init_nsts = nullptr;   // experiment: status / values handed to the coarse march
  std::vector<float>* rec_init_ttn = nullptr;
  int fim_coarse = 0;
  long fim_sweeps = 0;   // Gauss-Seidel passes the last fixed-point solve took
  int fim_converged = 0;
  long fim_evals = 0;    // quadrant-solver evaluations the last fixed-point solve made (cost model)
  // ---- statistics for the bench (not in the reference) ----
  long n_accept = 0;   // nodes set alive
  long n_steps = 0;    // ray steps taken

  inline float& VELV(int i, int j) { return velv[(size_t)j * ldv + i]; }          // velv(0:nvz+1,0:nvx+1)
  inline float& VELN(int iz, int ix) { return veln[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float& VELNB(int iz, int ix) { return velnb[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float& TTN(int iz, int ix) { return ttn[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline int& NSTS(int iz, int ix) { return nsts[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float& TTNR(int iz, int ix) { return ttnr[(size_t)(ix - 1) * ldr + (iz - 1)]; }
  inline int& NSTSR(int iz, int ix) { return nstsr[(size_t)(ix - 1) * ldr + (iz - 1)]; }

  // Part 1 of the orchestrators (FwdTraveltimeCPS.f90:346-409)
  void init(int nx, int ny, float goxdf, float gozdf, float dvxdf, float dvzdf);
  void gridder(const double* pv);         // CalSurfG.f90:1423
  void bsplrefine();                      // CalSurfG.f90:1525
  int travel(float scx, float scz, int urg);  // CalSurfG.f90:258
  void fouds2(int iz, int ix);            // CalSurfG.f90:557
  void addtree(int iz, int ix);           // CalSurfG.f90:738
  void downtree();                        // CalSurfG.f90:786
  void updtree(int iz, int ix);           // CalSurfG.f90:864
  float bilinear(const float nv[2][2], float dsx, float dsz);  // CalSurfG.f90:2293
  // the per-source block of the orchestrators (FwdTraveltimeCPS.f90:467-645):
  // gridder + refined solve + hand-off + coarse solve
  int solve_source(const double* pv, float x, float z);
  int travel_fim();                       // fim_experiment.cpp (experiment, not the reference's algorithm)
  float fouds2_values(int iz, int ix, float tcur, bool second_order);   // fim_experiment.cpp
  template <class Pred> float fouds2_pred(int iz, int ix, bool second_order, Pred usable);   // fim_experiment.cpp
  int srtimes(float scx, float scz, float rcx1, float rcz1, float* cbst1);  // CalSurfG.f90:1599
  // rpathsAzim.f90:16 (azim=true) / rpaths CalSurfG.f90:1735 (azim=false)
  int rpaths(float scx, float scz, float* fdm, float* fdmc, float* fdms,
             float surfrcx, float surfrcz, bool azim);
};

// rpathsAzim.f90:687 (implicit typing: stalat..baz, piby2, predel are REAL*4)
void azdist(float stalat, float stalon, float evtlat, float evtlon,
            float* delta, float* az, float* baz);
// delsph.f90:1
float delsph(float flat1, float flon1, float flat2, float flon2);

}  // namespace orc
