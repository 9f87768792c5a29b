// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the forward-modelling hot path of DAzimSurfTomo
// (reference: /root/reference, Fortran).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Parity pinning: the restatement is checked against the golden files the
// reference ships in example/test1_syn_foward/output/ (see tests/golden/ and
// tests/test_oracle_golden.py).  What those files do not pin -- the
// isotropic finite-difference kernels sen_vs/vp/rho and the COO triplets -- is
// pinned end to end by the inversion results the reference ships for test2 /
// test3 (inversion.cpp header, scripts/pin_inversion.py, tests/test_inversion.py).
#pragma once
#include "fmm2d.hpp"

namespace orc {

// surfdisp96.f:52 (Rayleigh phase branch). cg holds float32-rounded values.
int surfdisp96(const float* thkm, const float* vpm, const float* vsm, const float* rhom,
               int nlayer, int iflsph, int iwave, int mode, int igr, int kmax,
               const double* t, double* cg, long* neval);

// tregn96_subroutine.f:49.  dcdah/dcdbv/dcdn are (NP=60, NL=200) column-major.
int tregn96(int mmax, const float* thk, const float* TA, const float* TC, const float* TF,
            const float* TL, const float* TN, const float* TRho, const float* qp,
            const float* qs, const float* etap, const float* etas, const float* frefp,
            const float* frefs, int kmax, const float* t_in, const float* cp_in,
            float* dcdah, float* dcdbv, float* dcdn);

// CalSurfG.f90:2317 (refineGrid2LayerMdl) == FwdTraveltimeCPS.f90:62 (refineLayerMdl)
void refine_layer_mdl(float minthk0, int mmax, const float* dep, const float* vp,
                      const float* vs, const float* rho, int* rmax, float* rdep, float* rvp,
                      float* rvs, float* rrho, float* rthk, int* nsublay);
// depthkernelTI.f90:53-59 / CalSurfG.f90:48-54
void brocher(float vs, float* vp, float* rho);

// CalSurfG.f90:1 ; tables are Fortran column-major (nx*ny, kmax, nz)
int depthkernel(int nx, int ny, int nz, const float* vel, double* pvRc, double* sen_vs,
                double* sen_vp, double* sen_rho, int kmaxRc, const double* tRc,
                const float* depz, float minthk, int nthreads, long* neval);
// depthkernelTI.f90:2 ; Lsen_Gsc (nx*ny, kmax, nz-1) column-major float
int depthkernel_ti(int nx, int ny, int nz, const float* vel, double* pvRc, int kmaxRc,
                   const double* tRc, const float* depz, float minthk, float* Lsen_Gsc,
                   int nthreads);

struct Survey {  // the station tables the Fortran drivers pass down
  int kmax, nsrc, nrcf;
  const int* periods;    // (nsrc,kmax)
  const int* nrc1;       // (nsrc,kmax)
  const int* nsrcsurf1;  // (kmax)
  const float* scxf;     // (nsrc,kmax) colatitude rad
  const float* sczf;     // (nsrc,kmax) longitude rad
  const float* rcxf;     // (nrcf,nsrc,kmax)
  const float* rczf;
};

struct StageTimes { double kernels_s, dice_fmm_s, trace_s, assemble_s; long n_accept, n_steps; };

// mode 0: FwdObsTraveltimeCPS (FwdTraveltimeCPS.f90:208) -> dsurf, obsTaa
// mode 1: CalSurfG (CalSurfG.f90:909)                      -> dsurf, COO (nparpi cols)
// mode 2: CalSurfGAnisoJoint (CalSurfGAniso_Joint.f90:209) -> dsurf, COO (3*nparpi cols)
// COO: rw/col 0-based arrays of length nar, iw_row = 1-based row id per entry
// (the reference stores it at iw(2:nar+1)).  Tables may be passed in
// (precomputed != 0) to time/verify the FMM+ray stage alone.
struct GBuild {
  int mode;
  int nx, ny, nz;
  const float* vels;              // (nx,ny,nz)
  float goxd, gozd, dvxd, dvzd;
  int kmaxRc;
  const double* tRc;
  const float* depz;
  float minthk;
  Survey sv;
  const float* Gctrue; const float* Gstrue;   // mode 0, (nx-2,ny-2,nz-1)
  int precomputed;                // tables below already filled
  double* pvRc;                   // (nx*ny,kmax)
  double* sen_vs; double* sen_vp; double* sen_rho;   // (nx*ny,kmax,nz) modes 1,2
  float* Lsen_Gsc;                // (nx*ny,kmax,nz-1) modes 0,2
  // outputs
  float* dsurf; float* obsTaa;    // (dall)
  double* tRcV;                   // ((nx-2)(ny-2),kmax) or null
  float* rw; int* iw_row; int* col; long maxnar; long nar;
  int nthreads;                   // sources are independent; 1 = reference order
  int rbint;
  StageTimes times;
};
int gbuild(GBuild& g);

}  // namespace orc
