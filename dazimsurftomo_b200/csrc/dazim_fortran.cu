// gfortran-ABI drop-in symbols (lower case + trailing underscore, every
// argument by reference, LOGICAL = 4-byte int) over the C ABI.  They replace,
// at link time, the reference's
//   FwdObsTraveltimeCPS  (src/src_forward/FwdTraveltimeCPS.f90:208-211)
//   CalSurfG             (src/src_inv_iso_joint/CalSurfG.f90:909-912)
//   CalSurfGAnisoJoint   (src/src_inv_iso_joint/CalSurfGAniso_Joint.f90:209-212)
//   depthkernel          (src/src_inv_iso_joint/CalSurfG.f90:1-2)
//   depthkernelTI        (src/src_forward/depthkernelTI.f90:2)
// Error convention of the reference: print the message and STOP.
#include "../../include/dazim_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

static dazim_handle* g_handle = nullptr;

static dazim_handle* handle() {
  if (!g_handle) {
    int dev = 0;
    if (const char* e = getenv("DAZIM_DEVICE")) dev = atoi(e);
    int st = dazim_create(&g_handle, dev);
    if (st) {
      fprintf(stderr, " dazim_b200: cannot initialise CUDA device %d: %s\n TERMINATING PROGRAM!!!\n", dev,
              dazim_strerror(st));
      exit(1);
    }
  }
  return g_handle;
}

// DAZIM_DEVICES="0,1,2,3" or "all": the three orchestrators spread ONE call over several GPUs (dazim_gbuild_multi);
// unset or a single device: dazim_gbuild on DAZIM_DEVICE (default 0).
#include <cuda_runtime.h>
#include <vector>
static std::vector<int> device_list() {
  std::vector<int> out;
  const char* e = getenv("DAZIM_DEVICES");
  if (!e || !*e) return out;
  if (!strcmp(e, "all")) {
    int n = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess)
      for (int i = 0; i < n; ++i) out.push_back(i);
    return out;
  }
  for (const char* q = e; *q;) {
    char* end = nullptr;
    const long v = strtol(q, &end, 10);
    if (end == q) break;
    out.push_back((int)v);
    q = (*end == ',') ? end + 1 : end;
  }
  return out;
}
static int gbuild_any(int mode, const dazim_problem* p, dazim_tables* tb, const float* Gc, const float* Gs, float* dsurf,
                      float* obsTaa, double* tRcV, dazim_coo* coo) {
  const std::vector<int> dev = device_list();
  if (dev.size() > 1)
    return dazim_gbuild_multi((int)dev.size(), dev.data(), mode, p, tb, 0, Gc, Gs, dsurf, obsTaa, tRcV, coo, nullptr);
  return dazim_gbuild(handle(), mode, p, tb, 0, Gc, Gs, dsurf, obsTaa, tRcV, coo);
}

// writepath (FwdTraveltimeCPS.f90:673-686, rpathsAzim.f90:617-625) dumps every ray's nodes to
// raypath_refmdl_<T>s.dat; all shipped examples run with it off.  It is a diagnostic, not an input of the
// inversion: say so loudly instead of silently ignoring the request.
static void warn_writepath(int flag) {
  if (flag) {
    printf(" dazim_b200: writepath requested -- ray-path files (raypath_refmdl_*s.dat) are NOT written by the GPU path\n");
    fflush(stdout);
  }
}

static void stop_on(int st, const char* where) {
  if (st == DAZIM_OK) return;
  fprintf(stdout, " %s\n TERMINATING PROGRAM!!! (%s)\n", dazim_strerror(st), where);
  fflush(stdout);
  exit(1);
}

static void fill_problem(dazim_problem& p, int* nx, int* ny, int* nz, float* vels, float* goxdf, float* gozdf,
                         float* dvxdf, float* dvzdf, int* kmaxRc, double* tRc, int* periods, float* depz,
                         float* minthk, float* scxf, float* sczf, float* rcxf, float* rczf, int* nrc1,
                         int* nsrcsurf1, int* kmax, int* nsrcsurf, int* nrcf) {
  std::memset(&p, 0, sizeof(p));
  p.nx = *nx; p.ny = *ny; p.nz = *nz; p.vels = vels; p.goxd = *goxdf; p.gozd = *gozdf; p.dvxd = *dvxdf;
  p.dvzd = *dvzdf; p.kmaxRc = *kmaxRc; p.tRc = tRc; p.depz = depz; p.minthk = *minthk; p.kmax = *kmax;
  p.nsrc = *nsrcsurf; p.nrcf = *nrcf; p.periods = periods; p.nrc1 = nrc1; p.nsrcsurf1 = nsrcsurf1;
  p.scxf = scxf; p.sczf = sczf; p.rcxf = rcxf; p.rczf = rczf;
}

// Dense copies GVs/GGc/GGs(dall,nparpi) are only read by the reference's residual
// statistics (CalSigamNorm.f90); they are rebuilt from the sparse triplets when the
// caller passes them (without the stale-coefficient bug of SURVEY Q6).
static void densify(const dazim_coo& c, long long dall, long long nparpi, float* GVs, float* GGc, float* GGs) {
  float* blk[3] = {GVs, GGc, GGs};
  for (int b = 0; b < 3; ++b)
    if (blk[b]) std::memset(blk[b], 0, sizeof(float) * (size_t)dall * (size_t)nparpi);
  for (long long k = 0; k < c.nar; ++k) {
    const long long col = c.col[k] - 1, row = c.iw_row[k] - 1;
    const int b = (int)(col / nparpi);
    if (b < 3 && blk[b]) blk[b][(size_t)row + (size_t)(col % nparpi) * (size_t)dall] = c.rw[k];
  }
}

extern "C" void depthkernel_(int* nx, int* ny, int* nz, float* vel, double* pvRc, double* sen_vsRc,
                             double* sen_vpRc, double* sen_rhoRc, int* iwave, int* igr, int* kmaxRc, double* tRc,
                             float* depz, float* minthk) {
  if (*iwave != 2 || *igr != 0) stop_on(DAZIM_EBADARG, "depthkernel: only Rayleigh phase velocity");
  stop_on(dazim_depthkernel(handle(), *nx, *ny, *nz, vel, pvRc, sen_vsRc, sen_vpRc, sen_rhoRc, *kmaxRc, tRc, depz,
                            *minthk), "depthkernel");
}

extern "C" void depthkernelti_(int* nx, int* ny, int* nz, float* vel, double* pvRc, int* iwave, int* igr,
                               int* kmaxRc, double* tRc, float* depz, float* minthk, float* Lsen_Gsc) {
  if (*iwave != 2 || *igr != 0) stop_on(DAZIM_EBADARG, "depthkernelTI: only Rayleigh phase velocity");
  stop_on(dazim_depthkernel_ti(handle(), *nx, *ny, *nz, vel, pvRc, *kmaxRc, tRc, depz, *minthk, Lsen_Gsc),
          "depthkernelTI");
}

extern "C" void fwdobstraveltimecps_(int* nx, int* ny, int* nz, int* nparpi, float* vels, float* Gctrue,
                                     float* Gstrue, float* dsurf, float* obsTaa, int* dall, int* rmax, double* tRcV,
                                     float* Lsen_Gsc, float* goxdf, float* gozdf, float* dvxdf, float* dvzdf,
                                     int* kmaxRc, double* tRc, int* periods, float* depz, float* minthk, float* scxf,
                                     float* sczf, float* rcxf, float* rczf, int* nrc1, int* nsrcsurf1, int* kmax,
                                     int* nsrcsurf, int* nrcf, int* writepath) {
  (void)nparpi; (void)dall; (void)rmax;
  warn_writepath(*writepath);
  dazim_problem p;
  fill_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, scxf, sczf, rcxf,
               rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf);
  dazim_tables tb;
  std::memset(&tb, 0, sizeof(tb));
  tb.Lsen_Gsc = Lsen_Gsc;
  printf("  DepthkernelTI begin!\n");
  stop_on(gbuild_any(0, &p, &tb, Gctrue, Gstrue, dsurf, obsTaa, tRcV, nullptr), "FwdObsTraveltimeCPS");
  printf("  DepthkernelTI time cost= %13.1f s\n", dazim_last_times(handle())->kernels_ms * 1e-3);
}

extern "C" void calsurfg_(int* nx, int* ny, int* nz, int* nparpi, float* vels, int* iw, float* rw, int* col,
                          float* dsurf, float* GVs, int* dall, float* goxdf, float* gozdf, float* dvxdf, float* dvzdf,
                          int* kmaxRc, double* tRc, int* periods, float* depz, float* minthk, float* scxf,
                          float* sczf, float* rcxf, float* rczf, int* nrc1, int* nsrcsurf1, int* kmax, int* nsrcsurf,
                          int* nrcf, int* nar) {
  dazim_problem p;
  fill_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, scxf, sczf, rcxf,
               rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf);
  dazim_tables tb;
  std::memset(&tb, 0, sizeof(tb));
  dazim_coo c;
  c.rw = rw; c.iw_row = iw + 1; c.col = col; c.nar = 0;
  // the caller sized rw/col by spfra (Main_Jt.f90:325); the bound is not passed down, so trust it like the reference
  c.maxnar = (long long)1 << 62;
  stop_on(gbuild_any(1, &p, &tb, nullptr, nullptr, dsurf, nullptr, nullptr, &c), "CalSurfG");
  if (c.nar > 2147483647ll) stop_on(DAZIM_ENNZ_OVERFLOW, "CalSurfG: nar exceeds the reference's default INTEGER");
  *nar = (int)c.nar;
  if (GVs) densify(c, *dall, *nparpi, GVs, nullptr, nullptr);
}

extern "C" void calsurfganisojoint_(int* nx, int* ny, int* nz, int* nparpi, float* vels, int* iw, float* rw, int* col,
                                    float* dsurf, float* GVs, float* GGc, float* GGs, float* Lsen_Gsc, int* dall,
                                    int* rmax, double* tRcV, float* goxdf, float* gozdf, float* dvxdf, float* dvzdf,
                                    int* kmaxRc, double* tRc, int* periods, float* depz, float* minthk, float* scxf,
                                    float* sczf, float* rcxf, float* rczf, int* nrc1, int* nsrcsurf1, int* kmax,
                                    int* nsrcsurf, int* nrcf, int* nar, int* writepath) {
  (void)rmax;
  warn_writepath(*writepath);
  dazim_problem p;
  fill_problem(p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, tRc, periods, depz, minthk, scxf, sczf, rcxf,
               rczf, nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf);
  dazim_tables tb;
  std::memset(&tb, 0, sizeof(tb));
  tb.Lsen_Gsc = Lsen_Gsc;
  dazim_coo c;
  c.rw = rw; c.iw_row = iw + 1; c.col = col; c.nar = 0;
  c.maxnar = (long long)1 << 62;
  stop_on(gbuild_any(2, &p, &tb, nullptr, nullptr, dsurf, nullptr, tRcV, &c), "CalSurfGAnisoJoint");
  if (c.nar > 2147483647ll) stop_on(DAZIM_ENNZ_OVERFLOW, "CalSurfGAnisoJoint: nar exceeds the reference's default INTEGER");
  *nar = (int)c.nar;
  if (GVs || GGc || GGs) densify(c, *dall, *nparpi, GVs, GGc, GGs);
  if (dazim_last_times(handle())->rbint) printf(" ray path along the boundary, dangerous!!\n");
}

// LSMRmodule::LSMR (lsmrModule.f90:36): iw(1) = nnz, iw(2:nnz+1) = rows, iw(nnz+2:2nnz+1) = columns (aprod.f90:22-27)
extern "C" void __lsmrmodule_MOD_lsmr(int* m, int* n, int* leniw, int* lenrw, int* iw, float* rw, float* b, float* damp,
                                      float* atol, float* btol, float* conlim, int* itnlim, int* localSize, int* nout,
                                      float* x, int* istop, int* itn, float* normA, float* condA, float* normr,
                                      float* normAr, float* normx) {
  (void)leniw; (void)lenrw; (void)nout;
  const long long nnz = iw[0];
  dazim_lsmr_info info;
  std::memset(&info, 0, sizeof(info));
  stop_on(dazim_lsmr(handle(), *m, *n, nnz, iw + 1, iw + 1 + nnz, rw, b, *damp, *atol, *btol, *conlim, *itnlim,
                     *localSize, x, &info), "LSMR");
  *istop = info.istop; *itn = info.itn; *normA = info.normA; *condA = info.condA; *normr = info.normr;
  *normAr = info.normAr; *normx = info.normx;
}

// CalDdatSigma (CalSigamNorm.f90:2; called at Main_Jt.f90:460)
extern "C" void calddatsigma_(int* dall, float* obst, float* cbst, float* sigmaT, float* meandeltaT) {
  stop_on(dazim_cal_ddat_sigma(handle(), *dall, obst, cbst, sigmaT, meandeltaT), "CalDdatSigma");
}

// TikhonovRegularization (TikhRegul.f90:2; Main_Jt.f90:513): rows live at iw(2:), iso_inv is a LOGICAL (4-byte)
extern "C" void tikhonovregularization_(int* nx, int* ny, int* nz, int* maxvp, int* dall, int* nar, float* rw, int* iw,
                                        int* col, int* count3, int* iso_inv, float* weightGcs, float* weightVs) {
  long long n = *nar;
  stop_on(dazim_tikhonov(handle(), 0, *nx, *ny, *nz, *maxvp, *dall, &n, rw, iw + 1, col, nullptr, count3, *iso_inv != 0,
                         *weightGcs, *weightVs), "TikhonovRegularization");
  if (n > 2147483647ll) stop_on(DAZIM_ENNZ_OVERFLOW, "TikhonovRegularization: nar exceeds the reference's default INTEGER");
  *nar = (int)n;
}

// TikhRegul_joint (TikhRegul.f90:108; Main_Jt.f90:515)
extern "C" void tikhregul_joint_(int* nx, int* ny, int* nz, int* maxvp, int* dall, int* nar, float* rw, int* iw, int* col,
                                 int* narVs, int* count3, float* weightGcs, float* weightVs) {
  long long n = *nar, nvs = -1;
  stop_on(dazim_tikhonov(handle(), 1, *nx, *ny, *nz, *maxvp, *dall, &n, rw, iw + 1, col, &nvs, count3, 0, *weightGcs,
                         *weightVs), "TikhRegul_joint");
  if (n > 2147483647ll) stop_on(DAZIM_ENNZ_OVERFLOW, "TikhRegul_joint: nar exceeds the reference's default INTEGER");
  *nar = (int)n;
  *narVs = (int)nvs;
}
