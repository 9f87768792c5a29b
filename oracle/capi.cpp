// ORACLE (test infrastructure, NOT product code).  Plain C entry points for
// ctypes (tests/, bench.py cpu_baseline leg).
#include "oracle.h"
#include <cstring>

extern "C" {

int orc_surfdisp96(const float* thk, const float* vp, const float* vs, const float* rho, int nlayer,
                   int iflsph, int iwave, int mode, int igr, int kmax, const double* t, double* cg,
                   long* neval) {
  return orc::surfdisp96(thk, vp, vs, rho, nlayer, iflsph, iwave, mode, igr, kmax, t, cg, neval);
}

int orc_tregn96(int mmax, const float* thk, const float* TA, const float* TC, const float* TF,
                const float* TL, const float* TN, const float* TRho, const float* qp, const float* qs,
                const float* etap, const float* etas, const float* frefp, const float* frefs, int kmax,
                const float* t_in, const float* cp_in, float* dcdah, float* dcdbv, float* dcdn) {
  return orc::tregn96(mmax, thk, TA, TC, TF, TL, TN, TRho, qp, qs, etap, etas, frefp, frefs, kmax,
                      t_in, cp_in, dcdah, dcdbv, dcdn);
}

void orc_refine_layer_mdl(float minthk0, int mmax, const float* dep, const float* vp, const float* vs,
                          const float* rho, int* rmax, float* rdep, float* rvp, float* rvs, float* rrho,
                          float* rthk, int* nsublay) {
  orc::refine_layer_mdl(minthk0, mmax, dep, vp, vs, rho, rmax, rdep, rvp, rvs, rrho, rthk, nsublay);
}

void orc_brocher(float vs, float* vp, float* rho) { orc::brocher(vs, vp, rho); }

int orc_depthkernel(int nx, int ny, int nz, const float* vel, double* pvRc, double* sen_vs, double* sen_vp,
                    double* sen_rho, int kmaxRc, const double* tRc, const float* depz, float minthk,
                    int nthreads, long* neval) {
  return orc::depthkernel(nx, ny, nz, vel, pvRc, sen_vs, sen_vp, sen_rho, kmaxRc, tRc, depz, minthk, nthreads, neval);
}

int orc_depthkernel_ti(int nx, int ny, int nz, const float* vel, double* pvRc, int kmaxRc, const double* tRc,
                       const float* depz, float minthk, float* Lsen_Gsc, int nthreads) {
  return orc::depthkernel_ti(nx, ny, nz, vel, pvRc, kmaxRc, tRc, depz, minthk, Lsen_Gsc, nthreads);
}

float orc_delsph(float a, float b, float c, float d) { return orc::delsph(a, b, c, d); }

void orc_azdist(float stalat, float stalon, float evtlat, float evtlon, float* delta, float* az, float* baz) {
  orc::azdist(stalat, stalon, evtlat, evtlon, delta, az, baz);
}

// One (period, source) solve exposing the fields the ray tracer consumes.
// ttn/nsts: coarse (nnz,nnx) column-major; ttnr/nstsr: refined (129,129) with
// the used extent (nnzr,nnxr) returned in geom[0..1]; geom = nnzr,nnxr,vnl,vnr,vnt,vnb
int orc_fmm_source(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                   float scx, float scz, float* veln, float* ttn, int* nsts, float* ttnr, int* nstsr,
                   int* geom, float* fgeom) {
  orc::Fmm f;
  f.init(nx, ny, goxd, gozd, dvxd, dvzd);
  int st = f.solve_source(pv, scx, scz);
  if (st) return st;
  for (int ix = 1; ix <= f.nnx; ++ix)
    for (int iz = 1; iz <= f.nnz; ++iz) {
      size_t o = (size_t)(ix - 1) * f.nnz + (iz - 1);
      veln[o] = f.VELN(iz, ix);
      ttn[o] = f.TTN(iz, ix);
      nsts[o] = f.NSTS(iz, ix);
    }
  for (int ix = 1; ix <= f.nnxr; ++ix)
    for (int iz = 1; iz <= f.nnzr; ++iz) {
      size_t o = (size_t)(ix - 1) * f.ldr + (iz - 1);
      ttnr[o] = f.TTNR(iz, ix);
      nstsr[o] = f.NSTSR(iz, ix);
    }
  geom[0] = f.nnzr; geom[1] = f.nnxr; geom[2] = f.vnl; geom[3] = f.vnr; geom[4] = f.vnt; geom[5] = f.vnb;
  geom[6] = f.nnz; geom[7] = f.nnx;
  fgeom[0] = f.goxr; fgeom[1] = f.gozr; fgeom[2] = f.dnxr; fgeom[3] = f.dnzr;
  fgeom[4] = f.gox; fgeom[5] = f.goz; fgeom[6] = f.dnx; fgeom[7] = f.dnz;
  return 0;
}

// One ray: solve the source, then srtimes + rpaths(Azim); fdm* are (nvz+2,nvx+2) column-major
int orc_ray(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv, float scx,
            float scz, float rcx, float rcz, int azim, float* tt, float* fdm, float* fdmc, float* fdms,
            long* nsteps) {
  orc::Fmm f;
  f.init(nx, ny, goxd, gozd, dvxd, dvzd);
  int st = f.solve_source(pv, scx, scz);
  if (st) return st;
  st = f.srtimes(scx, scz, rcx, rcz, tt);
  if (st) return st;
  st = f.rpaths(scx, scz, fdm, fdmc, fdms, rcx, rcz, azim != 0);
  if (nsteps) *nsteps = f.n_steps;
  return st;
}

struct orc_gbuild_args {
  int mode, nx, ny, nz;
  const float* vels;
  float goxd, gozd, dvxd, dvzd;
  int kmaxRc;
  const double* tRc;
  const float* depz;
  float minthk;
  int kmax, nsrc, nrcf;
  const int* periods; const int* nrc1; const int* nsrcsurf1;
  const float* scxf; const float* sczf; const float* rcxf; const float* rczf;
  const float* Gctrue; const float* Gstrue;
  int precomputed;
  double* pvRc; double* sen_vs; double* sen_vp; double* sen_rho; float* Lsen_Gsc;
  float* dsurf; float* obsTaa; double* tRcV;
  float* rw; int* iw_row; int* col; long maxnar; long nar;
  int nthreads; int rbint;
  double t_kernels, t_dice_fmm, t_trace, t_assemble;
  long n_accept, n_steps;
};

int orc_gbuild(orc_gbuild_args* a) {
  orc::GBuild g;
  g.mode = a->mode; g.nx = a->nx; g.ny = a->ny; g.nz = a->nz; g.vels = a->vels;
  g.goxd = a->goxd; g.gozd = a->gozd; g.dvxd = a->dvxd; g.dvzd = a->dvzd;
  g.kmaxRc = a->kmaxRc; g.tRc = a->tRc; g.depz = a->depz; g.minthk = a->minthk;
  g.sv.kmax = a->kmax; g.sv.nsrc = a->nsrc; g.sv.nrcf = a->nrcf;
  g.sv.periods = a->periods; g.sv.nrc1 = a->nrc1; g.sv.nsrcsurf1 = a->nsrcsurf1;
  g.sv.scxf = a->scxf; g.sv.sczf = a->sczf; g.sv.rcxf = a->rcxf; g.sv.rczf = a->rczf;
  g.Gctrue = a->Gctrue; g.Gstrue = a->Gstrue; g.precomputed = a->precomputed;
  g.pvRc = a->pvRc; g.sen_vs = a->sen_vs; g.sen_vp = a->sen_vp; g.sen_rho = a->sen_rho;
  g.Lsen_Gsc = a->Lsen_Gsc; g.dsurf = a->dsurf; g.obsTaa = a->obsTaa; g.tRcV = a->tRcV;
  g.rw = a->rw; g.iw_row = a->iw_row; g.col = a->col; g.maxnar = a->maxnar; g.nar = 0;
  g.nthreads = a->nthreads; g.rbint = 0;
  int st = orc::gbuild(g);
  a->nar = g.nar; a->rbint = g.rbint;
  a->t_kernels = g.times.kernels_s; a->t_dice_fmm = g.times.dice_fmm_s;
  a->t_trace = g.times.trace_s; a->t_assemble = g.times.assemble_s;
  a->n_accept = g.times.n_accept; a->n_steps = g.times.n_steps;
  return st;
}

}  // extern "C"
