// Host side of the C ABI declared in include/dazim_b200.h: scalar prologues of
// the reference orchestrators (float32, same operation order), work-list
// construction, device memory, kernel sequencing and timing.  No compute
// happens on the host beyond what the reference itself does once per call in
// scalar code (grid constants, per-source box bounds, sin() of grid rows).
#include "../../include/dazim_b200.h"
#include "dazim_dev.h"
#include "dazim_tps.h"
#include "dazim_inv.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace dz {
// kernels (dazim_fmm.cu / dazim_trace.cu / dazim_th.cu)
cudaError_t upload_basis(const float* ub, const float* cb);
cudaError_t launch_dice_coarse(const GridC& g, int nper, const float* velv, float* veln, float* slow, cudaStream_t st);
cudaError_t fmm_max_ctas(int hcap, int spc, int nsm, int* nctas);
cudaError_t launch_fmm(const FmmArgs& A, int nctas, cudaStream_t st);
cudaError_t fmm_duo_max_ctas(int hcap, int nsm, int minb, int* nctas);
cudaError_t launch_fmm_duo(const FmmArgs& A, int nctas, int minb, cudaStream_t st);
cudaError_t fmm_tps_max_ctas(int hcap, int nsm, int coh, int qs, int* nctas);
cudaError_t launch_fmm_tps(const TpsArgs& A, int nctas, int coh, int qs, cudaStream_t st);
cudaError_t launch_decode_status(const unsigned* E, const int* hpos, size_t n, int nnz_tiled, float* ttn, int* nsts,
                                 cudaStream_t st);
cudaError_t launch_trace(const TraceArgs& A, bool azim, int nblocks, cudaStream_t st);
cudaError_t launch_coef(int nx, int ny, int nz, const float* vels, float* ca, float* cr, cudaStream_t st);
cudaError_t launch_assemble(const AsmArgs& A, bool fill, cudaStream_t st);
cudaError_t launch_taa(const TaaArgs& A, cudaStream_t st);
cudaError_t scan_rowptr(const long long* counts, long long* rowptr, int n, void* tmp, size_t* tmp_bytes, cudaStream_t st);
// Thomson-Haskell stage (dazim_th.cu)
int th_depthkernel(cudaStream_t st, int nx, int ny, int nz, const float* vel, double* pvRc, double* sen_vs,
                   double* sen_vp, double* sen_rho, int kmaxRc, const double* tRc, const float* depz, float minthk,
                   float* ms, long long* nlaunch);
int th_depthkernel_ti(cudaStream_t st, int nx, int ny, int nz, const float* vel, double* pvRc, int kmaxRc,
                      const double* tRc, const float* depz, float minthk, float* Lsen_Gsc, float* ms,
                      long long* nlaunch);
int th_surfdisp96(cudaStream_t st, int nprof, int nlayer, const float* thk, const float* vp, const float* vs,
                  const float* rho, int kmax, const double* t, double* cg);
}  // namespace dz
#include "dazim_coll.h"
namespace dzl {
int lsmr_solve(cudaStream_t st, int m, int n, long long nnz, const int* row, const int* col, const float* rw,
               const float* b, float damp, float atol, float btol, float conlim, int itnlim, int localSize, float* x,
               dazim_lsmr_info* info, bool coo_on_device, const Coll* coll = nullptr);
}

using namespace dz;

#define CK(x)                                                     \
  do {                                                            \
    cudaError_t e_ = (x);                                         \
    if (e_ != cudaSuccess) return DAZIM_ECUDA + (int)e_;          \
  } while (0)

static const float PI_F = 3.1415926535898f;   // CalSurfG.f90:166
static inline float sin_r(float x) { return (float)std::sin((double)x); }
static inline float cube(float x) { return x * (x * x); }
static void bspl_basis(float u, float* b) {
  b[0] = cube(1.0f - u) / 6.0f;
  b[1] = (4.0f - 6.0f * (u * u) + 3.0f * cube(u)) / 6.0f;
  b[2] = (1.0f + 3.0f * u + 3.0f * (u * u) - 3.0f * cube(u)) / 6.0f;
  b[3] = cube(u) / 6.0f;
}

struct dazim_handle {
  int dev;
  cudaStream_t st;
  dazim_times times;
  int nsm;
};

// Device buffers come from the stream-ordered allocator (cudaMallocAsync on the library stream;
// the default pool keeps freed blocks, dazim_create raises its release threshold), so the
// multi-GB workspaces of consecutive calls are recycled instead of going back to the driver.
static thread_local cudaStream_t g_alloc_stream = nullptr;
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t st = nullptr;
  cudaError_t alloc(size_t cnt) {
    release();
    n = cnt;
    st = g_alloc_stream;
    if (cnt == 0) return cudaSuccess;
    return cudaMallocAsync((void**)&p, cnt * sizeof(T), st);
  }
  void release() {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    n = 0;
  }
  ~DBuf() { release(); }
};

// memory the allocator could hand out right now: free device memory + what the pool holds but does not use
static cudaError_t available_bytes(int dev, size_t* out) {
  size_t free_b = 0, total_b = 0;
  cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
  if (e != cudaSuccess) return e;
  cudaMemPool_t pool;
  unsigned long long reserved = 0, used = 0;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
    free_b += (size_t)(reserved - used);
  *out = free_b;
  return cudaSuccess;
}

// FwdTraveltimeCPS.f90:346-380 (identical prologue in CalSurfG / CalSurfGAnisoJoint)
static GridC make_grid(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd) {
  GridC g;
  g.gdx = 5; g.gdz = 5; g.sgdl = 8; g.sgs = 8; g.earth = 6371.0f;
  g.nvx = nx - 2; g.nvz = ny - 2;
  g.dvx = dvxd * PI_F / 180.0f;
  g.dvz = dvzd * PI_F / 180.0f;
  g.gox = (90.0f - goxd) * PI_F / 180.0f;
  g.goz = gozd * PI_F / 180.0f;
  g.nnx = (g.nvx - 1) * g.gdx + 1;
  g.nnz = (g.nvz - 1) * g.gdz + 1;
  g.dnx = g.dvx / (float)g.gdx;
  g.dnz = g.dvz / (float)g.gdz;
  // dpl of srtimes / rpathsAzim (CalSurfG.f90:1650-1654, rpathsAzim.f90:156-161)
  float dpl = g.dnx * g.earth;
  float rd1 = g.dnz * g.earth * sin_r(g.gox);
  if (rd1 < dpl) dpl = rd1;
  rd1 = g.dnz * g.earth * sin_r(g.gox + (float)(g.nnx - 1) * g.dnx);
  if (rd1 < dpl) dpl = rd1;
  g.dpl_full = dpl;
  g.dpl_half = 0.5f * dpl;
  return g;
}

// per-source scalar prologue, FwdTraveltimeCPS.f90:493-530 + CalSurfG.f90:282-295
static int make_src(const GridC& g, float x, float z, SrcRec& s) {
  s.scx = x; s.scz = z;
  int isx = (int)((x - g.gox) / g.dnx) + 1;
  int isz = (int)((z - g.goz) / g.dnz) + 1;
  if (isx < 1 || isx > g.nnx || isz < 1 || isz > g.nnz) return DAZIM_ESOURCE_OUTSIDE;
  s.isx_c = isx; s.isz_c = isz;
  if (isx == g.nnx) isx = isx - 1;
  if (isz == g.nnz) isz = isz - 1;
  s.isx_cc = isx; s.isz_cc = isz;
  s.drx_c = (x - g.gox) - (float)(s.isx_c - 1) * g.dnx;
  s.drz_c = (z - g.goz) - (float)(s.isz_c - 1) * g.dnz;
  s.vnl = isx - g.sgs; if (s.vnl < 1) s.vnl = 1;
  s.vnr = isx + g.sgs; if (s.vnr > g.nnx) s.vnr = g.nnx;
  s.vnt = isz - g.sgs; if (s.vnt < 1) s.vnt = 1;
  s.vnb = isz + g.sgs; if (s.vnb > g.nnz) s.vnb = g.nnz;
  s.nnxr = (s.vnr - s.vnl) * g.sgdl + 1;
  s.nnzr = (s.vnb - s.vnt) * g.sgdl + 1;
  s.dnxr = g.dvx / (float)(g.gdx * g.sgdl);
  s.dnzr = g.dvz / (float)(g.gdz * g.sgdl);
  s.goxr = g.gox + g.dnx * (float)(s.vnl - 1);
  s.gozr = g.goz + g.dnz * (float)(s.vnt - 1);
  int rx = (int)((x - s.goxr) / s.dnxr) + 1;
  int rz = (int)((z - s.gozr) / s.dnzr) + 1;
  s.isx_t = rx; s.isz_t = rz;
  if (rx < 1 || rx > s.nnxr || rz < 1 || rz > s.nnzr) return DAZIM_ESOURCE_OUTSIDE;
  if (rx == s.nnxr) rx = rx - 1;
  if (rz == s.nnzr) rz = rz - 1;
  s.isx_r = rx; s.isz_r = rz;
  s.dsx_r = (x - s.goxr) - (float)(rx - 1) * s.dnxr;
  s.dsz_r = (z - s.gozr) - (float)(rz - 1) * s.dnzr;
  return DAZIM_OK;
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct dazim_plan {
  dazim_handle* h = nullptr;
  int mode = 0, nx = 0, ny = 0, nz = 0, kmaxRc = 0;
  GridC g;
  std::vector<SrcRec> src;            // owned units, loop order; ray0 = global row
  std::vector<RayRec> ray;            // batch-ordered rays (src index batch-local)
  std::vector<float> h_velv;          // staging of real(pvRc) (kept so that a model update can reuse it)
  std::vector<long long> batch_src0;  // batch boundaries in src (size nb+1)
  std::vector<long long> batch_ray0;  // batch boundaries in ray
  long long row0 = 0, nrow = 0;
  long long nnz = 0;
  bool azim = true;
  int emit_all = 0;
  // device inputs
  DBuf<SrcRec> d_src; DBuf<RayRec> d_ray; DBuf<int> d_row_knumi;
  DBuf<float> d_velv, d_veln_c, d_slow_c, d_risti_c, d_risti_r;
  DBuf<double> d_sen_vs, d_sen_vp, d_sen_rho; DBuf<float> d_lsen, d_vels, d_coe_a, d_coe_rho, d_gc, d_gs;
  // fmm / trace workspaces
  DBuf<unsigned> d_E_c, d_E_r;            // per solve: encoded travel-time/status fields
  DBuf<int> d_hpos_c, d_hpos_r, d_slot_of; DBuf<float> d_slow_r; DBuf<int2> d_hspill;   // per slot
  int nctas = 0;
  DBuf<unsigned short> d_map; DBuf<int> d_skey; DBuf<float> d_sval;
  int duo = 0;   // latency mode: one two-warp CTA per solve (k_fmm_duo)
  int duo_minb = 10;   // its register budget: 10 CTAs per SM (96 registers) or 16 (64 registers, 1.6 x the solves in flight)
  int tps = 0;   // one heap lane per solve (dazim_tps.h): k_fmm_coh (cohort kernel, the default) or k_fmm_tps
  int coh = 8;   // solves per heap warp of the cohort kernel (8 / 16 / 32; 8 measured best); 0: the one-thread-per-solve kernel
  int coh_qs = 1;   // stencil threads per neighbour (1 / 2 / 4; measured: the stencil side is DRAM-latency bound, more threads do not help)
  DBuf<int> d_hpos_r_out;                       // per solve, test seam only (tps)
  int hcap = 512, spc = 2, hspill = 0, cap = 0, trace_blocks = 0, maxB = 0;
  // footprint pool + outputs
  DBuf<int> d_fp_off, d_fp_cnt, d_fp_cell; DBuf<float> d_fp_fdm, d_fp_fdmc, d_fp_fdms;
  unsigned long long pool_cap = 0;
  DBuf<unsigned long long> d_counters;   // [0] pool_used [1] n_accept [2] n_steps
  DBuf<int> d_icnt;                       // [0] work counter [1] flags
  DBuf<float> d_dsurf, d_taa, d_val; DBuf<int> d_col, d_rowid; DBuf<long long> d_nnz_row, d_rowptr;
  DBuf<unsigned char> d_scan_tmp; size_t scan_tmp_bytes = 0;
  long long val_cap = 0;
  cudaEvent_t ev[8];
  bool ev_ok = false;
};

static void plan_free(dazim_plan* p) {
  if (!p) return;
  if (p->ev_ok) for (int i = 0; i < 8; ++i) cudaEventDestroy(p->ev[i]);
  delete p;
}

extern "C" int dazim_create(dazim_handle** out, int device) {
  if (!out) return DAZIM_EBADARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return DAZIM_ECUDA + (int)e;
  if (n == 0 || device >= n || device < 0) return DAZIM_ECUDA + (int)cudaErrorNoDevice;
  CK(cudaSetDevice(device));
  dazim_handle* h = new dazim_handle();
  h->dev = device;
  std::memset(&h->times, 0, sizeof(h->times));
  e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return DAZIM_ECUDA + (int)e; }
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, device);
  h->nsm = pr.multiProcessorCount;
  // B-spline basis tables (CalSurfG.f90:1469-1488, :1540-1555), float32 like the reference
  float ub[41 * 4], cb[6 * 4];
  for (int j = 1; j <= 41; ++j) { float u = 40.0f; u = (float)(j - 1) / u; bspl_basis(u, &ub[(j - 1) * 4]); }
  for (int i = 1; i <= 6; ++i) { float u = 5.0f; u = (float)(i - 1) / u; bspl_basis(u, &cb[(i - 1) * 4]); }
  e = upload_basis(ub, cb);
  if (e != cudaSuccess) { cudaStreamDestroy(h->st); delete h; return DAZIM_ECUDA + (int)e; }
  *out = h;
  return DAZIM_OK;
}

extern "C" int dazim_host_alloc(void** p, unsigned long long bytes) {
  if (!p) return DAZIM_EBADARG;
  cudaError_t e = cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault);
  return e == cudaSuccess ? DAZIM_OK : DAZIM_ECUDA + (int)e;
}
extern "C" void dazim_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" void dazim_destroy(dazim_handle* h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  cudaStreamDestroy(h->st);
  delete h;
}

extern "C" const dazim_times* dazim_last_times(const dazim_handle* h) { return h ? &h->times : nullptr; }

extern "C" const char* dazim_strerror(int code) {
  switch (code) {
    case DAZIM_OK: return "ok";
    case DAZIM_ESOURCE_OUTSIDE: return "Source lies outside bounds of model";
    case DAZIM_ERECEIVER_OUTSIDE: return "Receiver lies outside model";
    case DAZIM_ENNZ_OVERFLOW: return "nar > maxnar: please increase sparsity fraction (spfra)";
    case DAZIM_EBADARG: return "bad argument";
    case DAZIM_ELAYERS: return "too many layers (NL=200) or periods (NP=60)";
    case DAZIM_EHEAP: return "narrow band exceeded heap workspace";
    case DAZIM_EFOOTPRINT: return "ray footprint exceeded workspace";
    case DAZIM_ENOROOT: return "improper initial value in disper - no zero found";
    case DAZIM_ENCCL: return "NCCL not available or a collective failed";
    default: break;
  }
  if (code >= DAZIM_ECUDA) return cudaGetErrorString((cudaError_t)(code - DAZIM_ECUDA));
  return "unknown";
}

// kernel configuration of the round-1 eikonal kernels (two-warp latency kernel when every solve is resident, else the
// half-warp throughput kernel with the shared heap halved while that buys more solves in flight)
static int legacy_fmm_config(dazim_plan* P, long long nsrc, int hneed, int hmin, int* nctas_out) {
  dazim_handle* h = P->h;
  const GridC& g = P->g;
  int nctas = 1;
  // Measured on S200 (profiles/r2_k3_kernel_choice.md): every solve resident in the two-warp kernel wins up to its
  // capacity -- 1 480 solves with 96 registers, 2 368 with 64 (1 000 solves: 1.91 s / 2.01 s against 2.68 s for the
  // half-warp kernel; 2 000 solves: 4.13 s in two waves of the 96-register build, 3.16 s in one wave of the 64-register
  // build, 3.67 s half-warp) -- beyond that the half-warp throughput kernel.
  P->hcap = hneed;
  P->spc = 1;
  P->duo = 1;
  P->duo_minb = 10;
  if (const char* e = getenv("DAZIM_DUO_MINB")) P->duo_minb = atoi(e) >= 16 ? 16 : 10;
  CK(fmm_duo_max_ctas(P->hcap, h->nsm, P->duo_minb, &nctas));
  while (nctas < nsrc && P->hcap > 2048) {     // every solve resident with a (rarely spilling) smaller heap?
    P->hcap /= 2;
    CK(fmm_duo_max_ctas(P->hcap, h->nsm, P->duo_minb, &nctas));
  }
  if (nctas < nsrc && P->duo_minb == 10 && !getenv("DAZIM_DUO_MINB")) {
    // 64-register build with a shared heap small enough for 16 CTAs per SM (2 368 solves per chip): a second wave
    // costs a whole solve chain however few solves it holds, so the tier only applies when every solve is resident
    const int h16 = std::min(P->hcap, 1536);
    int n16 = 0;
    CK(fmm_duo_max_ctas(h16, h->nsm, 16, &n16));
    if (n16 >= nsrc) { P->duo_minb = 16; P->hcap = h16; nctas = n16; }
  }
  if (nctas < nsrc) {
    P->duo = 0;
    P->hcap = hneed;
    P->spc = 2;
    CK(fmm_max_ctas(P->hcap, 2, h->nsm, &nctas));
    while (P->hcap > hmin && (long long)nctas * 2 < nsrc) {
      P->hcap /= 2;
      CK(fmm_max_ctas(P->hcap, 2, h->nsm, &nctas));
    }
  }
  if (getenv("DAZIM_HCAP") || getenv("DAZIM_SPC") || getenv("DAZIM_DUO") || getenv("DAZIM_NCTAS")) {
    if (const char* e = getenv("DAZIM_DUO")) P->duo = atoi(e) ? 1 : 0;
    if (const char* e = getenv("DAZIM_SPC")) P->spc = atoi(e) == 1 ? 1 : 2;
    if (P->duo) P->spc = 1;
    if (const char* e = getenv("DAZIM_HCAP")) P->hcap = std::max(64, (atoi(e) + 1) & ~1);
    if (P->duo) CK(fmm_duo_max_ctas(P->hcap, h->nsm, P->duo_minb, &nctas));
    else CK(fmm_max_ctas(P->hcap, P->spc, h->nsm, &nctas));
    if (const char* e = getenv("DAZIM_NCTAS")) nctas = std::max(1, std::min(nctas, atoi(e)));
  }
  if (nctas < 1) return DAZIM_EBADARG;
  P->hspill = std::max(0, 8 * (g.nnx + g.nnz) + 1024 - P->hcap) + 16;
  *nctas_out = nctas;
  return DAZIM_OK;
}

// ---------------------------------------------------------------------------
static int plan_build(dazim_handle* h, int mode, const dazim_problem* p, const dazim_tables* tb,
                      const float* Gc, const float* Gs, long long sb, long long se, int emit_all,
                      dazim_plan** out) {
  if (!h || !p || !tb || !out) return DAZIM_EBADARG;
  if (mode < 0 || mode > 2) return DAZIM_EBADARG;
  if (p->nx < 5 || p->ny < 5 || p->nz < 2) return DAZIM_EBADARG;
  if (!tb->pvRc) return DAZIM_EBADARG;
  if ((mode == 0 || mode == 2) && !tb->Lsen_Gsc) return DAZIM_EBADARG;
  if ((mode == 1 || mode == 2) && (!tb->sen_vs || !tb->sen_vp || !tb->sen_rho)) return DAZIM_EBADARG;
  if (mode == 0 && (!Gc || !Gs) && !emit_all) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  g_alloc_stream = h->st;
  dazim_plan* P = new dazim_plan();
  P->h = h; P->mode = mode; P->nx = p->nx; P->ny = p->ny; P->nz = p->nz; P->kmaxRc = p->kmaxRc;
  P->azim = (mode != 1);
  P->emit_all = emit_all;
  P->g = make_grid(p->nx, p->ny, p->goxd, p->gozd, p->dvxd, p->dvzd);
  const GridC& g = P->g;
  const size_t nxy = (size_t)p->nx * p->ny;
  const size_t ncoarse = (size_t)g.nnx * g.nnz;
  h->times.h2d_bytes = 0;
  // ---- work list in the reference's loop order (FwdTraveltimeCPS.f90:464-465) ----
  long long unit = 0, count1 = 0;
  std::vector<int> row_knumi;
  bool first = true;
  for (int knumi = 1; knumi <= p->kmax; ++knumi)
    for (int srcnum = 1; srcnum <= p->nsrcsurf1[knumi - 1]; ++srcnum, ++unit) {
      const size_t sk = (size_t)(srcnum - 1) + (size_t)(knumi - 1) * p->nsrc;
      const int nr = p->nrc1[sk];
      if (unit >= sb && (se < 0 || unit < se)) {
        SrcRec s;
        std::memset(&s, 0, sizeof(s));
        int st = make_src(g, p->scxf[sk], p->sczf[sk], s);
        if (st) { plan_free(P); return st; }
        s.period = p->periods[sk] - 1;
        s.knumi = knumi - 1;
        if (s.period < 0 || s.period >= p->kmaxRc) { plan_free(P); return DAZIM_EBADARG; }
        if (first) { P->row0 = count1; first = false; }
        s.ray0 = (int)(count1 - P->row0);
        s.nray = nr;
        P->src.push_back(s);
        for (int i = 0; i < nr; ++i) row_knumi.push_back(knumi - 1);
      }
      count1 += nr;
    }
  P->nrow = (long long)row_knumi.size();
  const long long nsrc = (long long)P->src.size();

  // ---- workspace sizing ----
  // heap capacity in shared memory: bounded by the largest narrow band (~3 x grid edge; measured max 2.7 x on
  // S200).  Few solves (latency-bound): one solve per CTA with the whole heap in shared memory.  Many solves
  // (issue-bound): two solves per CTA and the capacity halved while that buys more solves in flight (measured
  // optimum 512 on S200: profiles/).
  int hneed = std::min(4096, std::max(256, pow2ceil(3 * std::max(std::max(g.nnx, g.nnz), REF_LD))));
  int hmin = std::min(hneed, 512);
  if (const char* e = getenv("DAZIM_HCAP_MIN")) hmin = std::max(64, std::min(hneed, atoi(e)));
  const size_t ncf = coarse_field_size(g.nnx, g.nnz);     // E_c / hpos_c in the interleaved layout
  P->hspill = 0;
  int nctas = 1;
  // ---- kernel choice (measured on the B200, profiles/r2_k3_kernel_choice.md).  The round-1 kernels keep the whole
  //      narrow band in shared memory when few solves compete (two-warp latency kernel: 1.6-2.1 us per accept) and
  //      pack 5 920 solves per chip otherwise (half-warp throughput kernel).  The cohort kernel (dazim_tps.h: one heap
  //      LANE per solve, neighbour records computed one node ahead) holds 8 288 solves per chip; on a grid whose band
  //      cannot live in shared memory anyway it beats the throughput kernel at every size measured (S200 grid: 3 000
  //      solves 3.54 s against 3.99 s, 4 000: 3.89 / 4.72, 8 000: 5.18 / 8.97) and loses only to the two-warp kernel
  //      while that one holds every solve resident (<= 2 368 solves) and on small grids (test1 shape: 91 ms against
  //      50 ms).  DAZIM_TPS=1/0 forces the choice; the legacy knobs (DAZIM_DUO / DAZIM_SPC) imply DAZIM_TPS=0. ----
  int auto_tps = 0;
  {
    int nc_l = 1;
    int st_l = legacy_fmm_config(P, nsrc, hneed, hmin, &nc_l);
    if (st_l) { plan_free(P); return st_l; }
    auto_tps = (!P->duo && hneed >= 2048) ? 1 : 0;
  }
  P->tps = auto_tps;
  if (getenv("DAZIM_DUO") || getenv("DAZIM_SPC")) P->tps = 0;
  if (const char* e = getenv("DAZIM_TPS")) P->tps = atoi(e) ? 1 : 0;
  const int hspill_full = 8 * (g.nnx + g.nnz) + 1024 + 16;     // generous bound on the narrow band (measured 2.7 x edge)
  if (P->tps) {
    if (const char* e = getenv("DAZIM_COH")) P->coh = atoi(e) ? P->coh : 0;
    if (const char* e = getenv("DAZIM_COH_LANES")) { const int v = atoi(e); if (P->coh) P->coh = (v <= 8) ? 8 : (v <= 16 ? 16 : 32); }
    if (const char* e = getenv("DAZIM_COH_QS")) P->coh_qs = std::max(1, atoi(e));
    const int L = P->coh ? P->coh : 32;                  // solves per CTA
    const long long nres = std::max<long long>(1, nsrc);
    const int ctas_needed = (int)((nres + L - 1) / L);
    // shared heap part: as large as possible while every solve is resident (else as many resident as the SM holds)
    int sm_budget = 227 * 1024;
    if (const char* e = getenv("DAZIM_SM_SLACK")) sm_budget -= std::max(0, std::min(65536, atoi(e)));   // shared memory left to co-resident kernels
    int per_sm = std::max(1, (ctas_needed + h->nsm - 1) / h->nsm);
    per_sm = std::min(per_sm, 64 / L > 0 ? 64 / L : 1);       // at most 64 solves per SM
    if (const char* e = getenv("DAZIM_TPS_PER_SM")) per_sm = std::max(1, std::min(16, atoi(e)));
    const int xch_bytes = P->coh ? (4 * L + 32 + 4 * L + 32 * L) * 4 + 128 : 0;
    P->hcap = std::min(hneed, (int)((sm_budget / per_sm - 1024 - xch_bytes) / (8 * L)));
    if (const char* e = getenv("DAZIM_HCAP")) P->hcap = std::max(8, std::min(880, atoi(e)));
    P->hcap = std::max(8, P->hcap & ~1);            // even: a sibling pair never straddles shared / spilled
    CK(fmm_tps_max_ctas(P->hcap, h->nsm, P->coh, P->coh_qs, &nctas));
    if (nctas < 1) { plan_free(P); return DAZIM_EBADARG; }
    nctas = std::min(nctas, ctas_needed);
    P->hspill = (std::max(16, hspill_full) + 1) & ~1;
    if (const char* e = getenv("DAZIM_TPS_HSPILL")) P->hspill = (std::max(4, atoi(e)) + 1) & ~1;   // test hook: force the overflow fall-back
  }
  if (!P->tps) {
    int st_l = legacy_fmm_config(P, nsrc, hneed, hmin, &nctas);
    if (st_l) { plan_free(P); return st_l; }
  }
  const int tpsL = P->coh ? P->coh : 32;
  const long long npairs_all = P->tps ? (nsrc + tpsL - 1) / tpsL : (nsrc + P->spc - 1) / P->spc;
  size_t free_b = 0;
  CK(available_bytes(h->dev, &free_b));
  double budget = 0.60 * (double)free_b;
  if (const char* e = getenv("DAZIM_WS_GB")) budget = std::min(budget, atof(e) * 1e9);
  double per_src = (double)ncf * 4 + (double)REF_N * 4 + REF_LD * 4 + 64;
  double per_slot = (double)ncf * 4 + (double)REF_N * 8 + (double)P->hspill * 8;
  if (P->tps) {   // everything is per solve: no slot workspaces, no hpos fields
    per_src += (double)REF_N * 4 + (double)P->hspill * 8 + (emit_all ? (double)REF_N * 4 : 0.0);
    per_slot = 0;
  }
  nctas = (int)std::min<long long>(nctas, std::max<long long>(npairs_all, 1));
  long long maxB = (long long)((budget - 2.0 * nctas * per_slot) / per_src);   // 2 slots per CTA are always laid out
  if (maxB < 2) maxB = 2;
  if (const char* e = getenv("DAZIM_BATCH")) maxB = std::max(1, atoi(e));
  maxB = std::min<long long>(maxB, std::max<long long>(nsrc, 1));
  if (P->tps) nctas = (int)std::min<long long>(nctas, (maxB + tpsL - 1) / tpsL);
  else nctas = (int)std::min<long long>(nctas, (maxB + P->spc - 1) / P->spc);
  P->nctas = nctas;
  P->maxB = (int)maxB;
  // batches + per-batch ray lists (long rays first so that a warp holds rays of similar length)
  P->batch_src0.push_back(0);
  P->batch_ray0.push_back(0);
  double pool_est = 0;
  const int ncell = (g.nvz + 2) * (g.nvx + 2);
  for (long long b0 = 0; b0 < nsrc; b0 += maxB) P->batch_src0.push_back(std::min(nsrc, b0 + maxB));
  if (nsrc == 0) P->batch_src0.push_back(0);
  // receivers: second pass over the loop to fill RayRec (keeps the code above simple)
  {
    std::vector<std::vector<std::pair<float, RayRec>>> per_batch(P->batch_src0.size() - 1);
    long long unit2 = 0, si = 0;
    for (int knumi = 1; knumi <= p->kmax; ++knumi)
      for (int srcnum = 1; srcnum <= p->nsrcsurf1[knumi - 1]; ++srcnum, ++unit2) {
        if (!(unit2 >= sb && (se < 0 || unit2 < se))) continue;
        const SrcRec& sr = P->src[si];
        const int b = (int)(si / maxB);
        for (int i = 0; i < sr.nray; ++i) {
          const size_t rk = (size_t)i + (size_t)(srcnum - 1) * p->nrcf + (size_t)(knumi - 1) * p->nrcf * p->nsrc;
          RayRec r;
          r.rcx = p->rcxf[rk]; r.rcz = p->rczf[rk];
          r.src = (int)(si - (long long)b * maxB);
          r.row = sr.ray0 + i;
          const float ddx = std::fabs(r.rcx - sr.scx) / g.dvx, ddz = std::fabs(r.rcz - sr.scz) / g.dvz;
          pool_est += std::min<double>(ncell, 48.0 + 10.0 * (ddx + ddz));
          per_batch[b].push_back({ddx + ddz, r});
        }
        ++si;
      }
    for (auto& v : per_batch) {
      std::stable_sort(v.begin(), v.end(), [](const std::pair<float, RayRec>& a, const std::pair<float, RayRec>& b) {
        return a.first > b.first;
      });
      for (auto& e : v) P->ray.push_back(e.second);
      P->batch_ray0.push_back((long long)P->ray.size());
    }
  }
  if (emit_all) pool_est = (double)P->nrow * ncell;
  P->pool_cap = (unsigned long long)(pool_est * 1.5) + 1024;
  if (const char* e = getenv("DAZIM_POOL_SCALE")) P->pool_cap = (unsigned long long)(P->pool_cap * atof(e));
  if (P->pool_cap > 2000000000ull) P->pool_cap = 2000000000ull;
  P->cap = std::min(ncell, 64 + 8 * (g.nvx + g.nvz));
  if (P->cap > 65535) P->cap = 65535;
  P->trace_blocks = h->nsm * 4;
  {
    long long maxrays = 0;
    for (size_t b = 0; b + 1 < P->batch_ray0.size(); ++b) maxrays = std::max(maxrays, P->batch_ray0[b + 1] - P->batch_ray0[b]);
    const long long need = (maxrays + 127) / 128;
    if (need < P->trace_blocks) P->trace_blocks = (int)std::max<long long>(1, need);
  }
  const size_t nthr = (size_t)P->trace_blocks * 128;

  // ---- device allocations + uploads ----
  cudaStream_t st = h->st;
#define UP(buf, hostptr, cnt)                                                                   \
  do {                                                                                          \
    CK((buf).alloc(cnt));                                                                       \
    if ((cnt) > 0) {                                                                            \
      CK(cudaMemcpyAsync((buf).p, hostptr, (cnt) * sizeof(*(buf).p), cudaMemcpyHostToDevice, st)); \
      h->times.h2d_bytes += (long long)((cnt) * sizeof(*(buf).p));                              \
    }                                                                                           \
  } while (0)
  int rc = DAZIM_OK;
  auto body = [&]() -> int {
    UP(P->d_src, P->src.data(), P->src.size());
    UP(P->d_ray, P->ray.data(), P->ray.size());
    UP(P->d_row_knumi, row_knumi.data(), row_knumi.size());
    // velv(i,j) = real(pv(i*(nvx+2)+j+1)) (CalSurfG.f90:1455) for every period
    P->h_velv.resize(nxy * p->kmaxRc);
    for (size_t i = 0; i < P->h_velv.size(); ++i) P->h_velv[i] = (float)tb->pvRc[i];
    UP(P->d_velv, P->h_velv.data(), P->h_velv.size());
    std::vector<float> ric(g.nnx);
    for (int ix = 1; ix <= g.nnx; ++ix) ric[ix - 1] = g.earth * sin_r(g.gox + (float)(ix - 1) * g.dnx);
    UP(P->d_risti_c, ric.data(), ric.size());
    std::vector<float> rir((size_t)nsrc * REF_LD, 0.0f);
    for (long long s = 0; s < nsrc; ++s) {
      const SrcRec& sr = P->src[s];
      for (int ix = 1; ix <= sr.nnxr; ++ix)
        rir[(size_t)s * REF_LD + ix - 1] = g.earth * sin_r(sr.goxr + (float)(ix - 1) * sr.dnxr);
    }
    UP(P->d_risti_r, rir.data(), rir.size());
    const size_t nlay = (size_t)p->nz - 1;
    if (mode == 1 || mode == 2) {
      UP(P->d_sen_vs, tb->sen_vs, nxy * p->kmaxRc * p->nz);
      UP(P->d_sen_vp, tb->sen_vp, nxy * p->kmaxRc * p->nz);
      UP(P->d_sen_rho, tb->sen_rho, nxy * p->kmaxRc * p->nz);
      UP(P->d_vels, p->vels, nxy * p->nz);
      CK(P->d_coe_a.alloc(nxy * nlay));
      CK(P->d_coe_rho.alloc(nxy * nlay));
    }
    if (mode == 0 || mode == 2) UP(P->d_lsen, tb->Lsen_Gsc, nxy * p->kmaxRc * nlay);
    if (mode == 0 && Gc && Gs) {
      UP(P->d_gc, Gc, (size_t)(p->nx - 2) * (p->ny - 2) * nlay);
      UP(P->d_gs, Gs, (size_t)(p->nx - 2) * (p->ny - 2) * nlay);
    }
    CK(P->d_veln_c.alloc(ncoarse * p->kmaxRc));
    CK(P->d_slow_c.alloc(ncoarse * p->kmaxRc));
    const size_t B = (size_t)P->maxB, nslot = 2 * (size_t)P->nctas;
    CK(P->d_E_r.alloc(B * REF_N));
    CK(P->d_E_c.alloc(B * coarse_field_size(g.nnx, g.nnz)));
    if (P->tps) {
      CK(P->d_slow_r.alloc(B * REF_N));
      CK(P->d_hspill.alloc(B * P->hspill));
      if (emit_all) CK(P->d_hpos_r_out.alloc(B * REF_N));
    } else {
      CK(P->d_hpos_c.alloc(nslot * coarse_field_size(g.nnx, g.nnz)));
      CK(P->d_hpos_r.alloc(nslot * REF_N));
      CK(P->d_slow_r.alloc(nslot * REF_N));
      CK(P->d_hspill.alloc(nslot * P->hspill));
    }
    CK(P->d_slot_of.alloc(B));
    CK(P->d_map.alloc(nthr * ncell));
    CK(cudaMemsetAsync(P->d_map.p, 0, nthr * ncell * sizeof(unsigned short), st));
    CK(P->d_skey.alloc(nthr * P->cap));
    CK(P->d_sval.alloc(nthr * 3 * P->cap));
    const size_t nrow = (size_t)std::max<long long>(P->nrow, 1);
    CK(P->d_fp_off.alloc(nrow));
    CK(P->d_fp_cnt.alloc(nrow));
    CK(P->d_fp_cell.alloc(P->pool_cap));
    CK(P->d_fp_fdm.alloc(P->pool_cap));
    if (P->azim) { CK(P->d_fp_fdmc.alloc(P->pool_cap)); CK(P->d_fp_fdms.alloc(P->pool_cap)); }
    CK(P->d_counters.alloc(4));
    CK(P->d_icnt.alloc(4));
    CK(P->d_dsurf.alloc(nrow));
    if (mode == 0) CK(P->d_taa.alloc(nrow));
    if (mode != 0) {
      CK(P->d_nnz_row.alloc(nrow));
      CK(P->d_rowptr.alloc(nrow + 1));
      P->scan_tmp_bytes = 0;
      CK(scan_rowptr(P->d_nnz_row.p, P->d_rowptr.p, (int)nrow, nullptr, &P->scan_tmp_bytes, st));
      CK(P->d_scan_tmp.alloc(P->scan_tmp_bytes + 16));
    }
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&P->ev[i]));
    P->ev_ok = true;
    CK(cudaStreamSynchronize(st));
    return DAZIM_OK;
  };
  rc = body();
#undef UP
  if (rc) { plan_free(P); return rc; }
  *out = P;
  return DAZIM_OK;
}

// events that are destroyed on every exit path
struct EventList {
  std::vector<cudaEvent_t> v;
  cudaError_t make(cudaEvent_t* e) {
    cudaError_t r = cudaEventCreate(e);
    if (r == cudaSuccess) v.push_back(*e);
    return r;
  }
  cudaEvent_t operator[](size_t i) const { return v[i]; }
  ~EventList() { for (auto e : v) cudaEventDestroy(e); }
};

static const int ST_POOL_OVERFLOW = -1000;   // internal: ray-footprint pool too small, plan_run grows it and re-runs

static int plan_run_once(dazim_plan* P) {
  dazim_handle* h = P->h;
  CK(cudaSetDevice(h->dev));
  g_alloc_stream = h->st;
  cudaStream_t st = h->st;
  const GridC& g = P->g;
  dazim_times& T = h->times;
  T.dice_ms = T.fmm_ms = T.trace_ms = T.assemble_ms = T.total_ms = 0;
  T.n_fmm_launch = T.n_trace_launch = T.n_launch = 0;
  CK(cudaMemsetAsync(P->d_counters.p, 0, 4 * sizeof(unsigned long long), st));
  CK(cudaMemsetAsync(P->d_icnt.p, 0, 4 * sizeof(int), st));
  CK(cudaEventRecord(P->ev[0], st));
  CK(launch_dice_coarse(g, P->kmaxRc, P->d_velv.p, P->d_veln_c.p, P->d_slow_c.p, st));
  T.n_launch++;
  if (P->mode != 0) {
    CK(launch_coef(P->nx, P->ny, P->nz, P->d_vels.p, P->d_coe_a.p, P->d_coe_rho.p, st));
    T.n_launch++;
  }
  CK(cudaEventRecord(P->ev[1], st));
  const size_t nb = P->batch_src0.size() - 1;
  EventList bev;   // per-batch fmm/trace boundaries
  for (size_t b = 0; b < nb; ++b) {
    const long long s0 = P->batch_src0[b], s1 = P->batch_src0[b + 1];
    const long long r0 = P->batch_ray0[b], r1 = P->batch_ray0[b + 1];
    cudaEvent_t e0, e1, e2;
    CK(bev.make(&e0)); CK(bev.make(&e1)); CK(bev.make(&e2));
    CK(cudaEventRecord(e0, st));
    FmmArgs F;
    F.g = g; F.src = P->d_src.p + s0; F.nsrc = (int)(s1 - s0); F.velv = P->d_velv.p; F.slow_c = P->d_slow_c.p;
    F.risti_c = P->d_risti_c.p; F.risti_r = P->d_risti_r.p + (size_t)s0 * REF_LD;
    F.E_c = P->d_E_c.p; F.E_r = P->d_E_r.p; F.hpos_c = P->d_hpos_c.p; F.hpos_r = P->d_hpos_r.p;
    F.slow_r = P->d_slow_r.p; F.hspill = P->d_hspill.p; F.hspill_n = P->hspill; F.hcap = P->hcap; F.spc = P->spc;
    F.queue = P->d_icnt.p + 2; F.slot_of = P->d_slot_of.p;
    F.flags = P->d_icnt.p + 1; F.n_accept = P->d_counters.p + 1;
    if (F.nsrc > 0) {
      // far = 0xFFFFFFFF everywhere on the coarse grids of this batch; the refined boxes reset themselves
      CK(cudaMemsetAsync(P->d_E_c.p, 0xFF, (size_t)F.nsrc * coarse_field_size(g.nnx, g.nnz) * sizeof(unsigned), st));
      CK(cudaMemsetAsync(P->d_icnt.p + 2, 0, sizeof(int), st));
      if (P->tps) {
        TpsArgs A;
        A.g = g; A.src = F.src; A.nsrc = F.nsrc; A.velv = F.velv; A.slow_c = F.slow_c; A.risti_c = F.risti_c;
        A.risti_r = F.risti_r; A.E_c = F.E_c; A.E_r = F.E_r; A.slow_r = P->d_slow_r.p; A.hspill = P->d_hspill.p;
        A.hspill_n = P->hspill; A.hcap = P->hcap; A.hpos_r_out = P->d_hpos_r_out.p;
        A.flags = F.flags; A.n_accept = F.n_accept;
        A.prof = getenv("DAZIM_COH_PROF") ? atoi(getenv("DAZIM_COH_PROF")) : 0;
        A.pf2 = getenv("DAZIM_COH_PF2") ? atoi(getenv("DAZIM_COH_PF2")) : 1;
        const int Lc = P->coh ? P->coh : 32;
        CK(launch_fmm_tps(A, std::min(P->nctas, (F.nsrc + Lc - 1) / Lc), P->coh, P->coh_qs, st));
        T.n_launch++;      // + k_tps_init
      }
      else if (P->duo) CK(launch_fmm_duo(F, std::min(P->nctas, F.nsrc), P->duo_minb, st));
      else CK(launch_fmm(F, std::min(P->nctas, (F.nsrc + P->spc - 1) / P->spc), st));
      T.n_launch++; T.n_fmm_launch++;
    }
    CK(cudaEventRecord(e1, st));
    if (r1 > r0) {
      CK(cudaMemsetAsync(P->d_icnt.p, 0, sizeof(int), st));
      TraceArgs A;
      A.g = g; A.src = P->d_src.p + s0; A.ray = P->d_ray.p + r0; A.nray = (int)(r1 - r0);
      A.veln_c = P->d_veln_c.p; A.ttn_c = reinterpret_cast<const float*>(P->d_E_c.p);
      A.ttn_r = reinterpret_cast<const float*>(P->d_E_r.p);
      A.map = P->d_map.p; A.skey = P->d_skey.p; A.sval = P->d_sval.p; A.cap = P->cap; A.dsurf = P->d_dsurf.p;
      A.fp_off = P->d_fp_off.p; A.fp_cnt = P->d_fp_cnt.p; A.fp_cell = P->d_fp_cell.p; A.fp_fdm = P->d_fp_fdm.p;
      A.fp_fdmc = P->d_fp_fdmc.p; A.fp_fdms = P->d_fp_fdms.p; A.pool_cap = P->pool_cap;
      A.pool_used = P->d_counters.p; A.counter = P->d_icnt.p; A.flags = P->d_icnt.p + 1;
      A.n_steps = P->d_counters.p + 2; A.emit_all = P->emit_all;
      CK(launch_trace(A, P->azim, P->trace_blocks, st));
      T.n_launch++; T.n_trace_launch++;
    }
    CK(cudaEventRecord(e2, st));
  }
  CK(cudaEventRecord(P->ev[2], st));
  // status checks need the counters: one small D2H
  unsigned long long cnt[4];
  int icnt[4];
  CK(cudaMemcpyAsync(cnt, P->d_counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(icnt, P->d_icnt.p, sizeof(icnt), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (size_t b = 0; b < nb; ++b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, bev[3 * b], bev[3 * b + 1]); T.fmm_ms += ms;
    cudaEventElapsedTime(&ms, bev[3 * b + 1], bev[3 * b + 2]); T.trace_ms += ms;
  }
  T.n_accept = (long long)cnt[1];
  T.n_steps = (long long)cnt[2];
  T.rbint = icnt[1] & 1;
  if (icnt[1] & 16) return DAZIM_EHEAP;
  if (icnt[1] & 2) return DAZIM_ERECEIVER_OUTSIDE;
  if (icnt[1] & 4) return DAZIM_EFOOTPRINT;
  if (icnt[1] & 8) return ST_POOL_OVERFLOW;       // pool overflow: plan_run retries with a bigger pool
  CK(cudaEventRecord(P->ev[3], st));
  P->nnz = 0;
  if (P->nrow > 0 && !P->emit_all) {
    if (P->mode == 0) {
      TaaArgs A;
      A.nx = P->nx; A.ny = P->ny; A.nz = P->nz; A.nvx = g.nvx; A.nvz = g.nvz; A.kmax = P->kmaxRc;
      A.nrow = (int)P->nrow; A.row0 = 0; A.row_knumi = P->d_row_knumi.p; A.fp_off = P->d_fp_off.p;
      A.fp_cnt = P->d_fp_cnt.p; A.fp_cell = P->d_fp_cell.p; A.fp_fdmc = P->d_fp_fdmc.p; A.fp_fdms = P->d_fp_fdms.p;
      A.lsen = P->d_lsen.p; A.gc = P->d_gc.p; A.gs = P->d_gs.p; A.taa = P->d_taa.p;
      CK(launch_taa(A, st));
      T.n_launch++;
    } else {
      AsmArgs A;
      A.mode = P->mode; A.nx = P->nx; A.ny = P->ny; A.nz = P->nz; A.nvx = g.nvx; A.nvz = g.nvz; A.kmax = P->kmaxRc;
      A.row0 = 0; A.nrow = (int)P->nrow; A.row_knumi = P->d_row_knumi.p; A.fp_off = P->d_fp_off.p;
      A.fp_cnt = P->d_fp_cnt.p; A.fp_cell = P->d_fp_cell.p; A.fp_fdm = P->d_fp_fdm.p; A.fp_fdmc = P->d_fp_fdmc.p;
      A.fp_fdms = P->d_fp_fdms.p; A.sen_vs = P->d_sen_vs.p; A.sen_vp = P->d_sen_vp.p; A.sen_rho = P->d_sen_rho.p;
      A.lsen = P->d_lsen.p; A.coe_a = P->d_coe_a.p; A.coe_rho = P->d_coe_rho.p; A.nnz_row = P->d_nnz_row.p;
      A.rowptr = nullptr; A.val = nullptr; A.col = nullptr; A.rowid = nullptr;
      CK(launch_assemble(A, false, st));
      size_t tb = P->scan_tmp_bytes;
      CK(scan_rowptr(P->d_nnz_row.p, P->d_rowptr.p, (int)P->nrow, P->d_scan_tmp.p, &tb, st));
      long long last_off = 0;
      long long last_cnt = 0;
      CK(cudaMemcpyAsync(&last_off, P->d_rowptr.p + (P->nrow - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(&last_cnt, P->d_nnz_row.p + (P->nrow - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      P->nnz = last_off + last_cnt;
      CK(cudaMemcpyAsync(P->d_rowptr.p + P->nrow, &P->nnz, sizeof(long long), cudaMemcpyHostToDevice, st));
      // headroom: the Tikhonov rows of dazim_plan_iterate are appended behind G (at most 7 entries per model cell)
      const long long reg_room = 21ll * g.nvx * g.nvz * (P->nz - 1);
      if (P->nnz + reg_room > P->val_cap) {
        P->val_cap = P->nnz + P->nnz / 8 + 1024 + reg_room;
        CK(P->d_val.alloc(P->val_cap));
        CK(P->d_col.alloc(P->val_cap));
        CK(P->d_rowid.alloc(P->val_cap));
      }
      A.rowptr = P->d_rowptr.p; A.val = P->d_val.p; A.col = P->d_col.p; A.rowid = P->d_rowid.p;
      CK(launch_assemble(A, true, st));
      T.n_launch += 4;
    }
  }
  CK(cudaEventRecord(P->ev[4], st));
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, P->ev[0], P->ev[1]); T.dice_ms = ms;
  cudaEventElapsedTime(&ms, P->ev[3], P->ev[4]); T.assemble_ms = ms;
  cudaEventElapsedTime(&ms, P->ev[0], P->ev[4]); T.total_ms = ms;
  return DAZIM_OK;
}

// The footprint pool is sized from a straight-line estimate at plan creation; rays of a later outer iteration can
// bend more.  Grow it (x4, up to three times) and re-run instead of failing a plan that is reused across iterations.
static int plan_run(dazim_plan* P) {
  int st = DAZIM_OK;
  for (int attempt = 0; attempt < 4; ++attempt) {
    st = plan_run_once(P);
    if (st == DAZIM_EHEAP && P->tps) {
      // the narrow band outgrew the spill workspace of the cohort kernel: re-run on the round-1 kernels
      CK(cudaSetDevice(P->h->dev));
      g_alloc_stream = P->h->st;
      const GridC& g = P->g;
      const int hneed = std::min(4096, std::max(256, pow2ceil(3 * std::max(std::max(g.nnx, g.nnz), REF_LD))));
      int nctas = 1;
      P->tps = 0;
      int st_l = legacy_fmm_config(P, (long long)P->maxB, hneed, std::min(hneed, 512), &nctas);
      if (st_l) return st_l;
      nctas = (int)std::min<long long>(nctas, ((long long)P->maxB + P->spc - 1) / P->spc);
      P->nctas = nctas;
      const size_t nslot = 2 * (size_t)nctas;
      CK(P->d_hpos_c.alloc(nslot * coarse_field_size(g.nnx, g.nnz)));
      CK(P->d_hpos_r.alloc(nslot * REF_N));
      CK(P->d_slow_r.alloc(nslot * REF_N));
      CK(P->d_hspill.alloc(nslot * P->hspill));
      continue;
    }
    if (st != ST_POOL_OVERFLOW) return st;
    if (attempt == 3) break;
    CK(cudaSetDevice(P->h->dev));
    g_alloc_stream = P->h->st;
    if (P->pool_cap >= 2000000000ull) break;                       // fp_off is a 32-bit offset
    P->pool_cap = std::min(P->pool_cap * 4ull, 2000000000ull);
    CK(P->d_fp_cell.alloc(P->pool_cap));
    CK(P->d_fp_fdm.alloc(P->pool_cap));
    if (P->azim) { CK(P->d_fp_fdmc.alloc(P->pool_cap)); CK(P->d_fp_fdms.alloc(P->pool_cap)); }
  }
  return DAZIM_EFOOTPRINT;
}

extern "C" int dazim_plan_create(dazim_handle* h, int mode, const dazim_problem* p, const dazim_tables* tables,
                                 const float* Gc, const float* Gs, long long sb, long long se, dazim_plan** plan) {
  return plan_build(h, mode, p, tables, Gc, Gs, sb, se, 0, plan);
}
extern "C" int dazim_plan_run(dazim_plan* plan) { return plan ? plan_run(plan) : DAZIM_EBADARG; }
extern "C" long long dazim_plan_rows(const dazim_plan* plan, long long* row0) {
  if (!plan) return 0;
  if (row0) *row0 = plan->row0;
  return plan->nrow;
}
extern "C" long long dazim_plan_nnz(const dazim_plan* plan) { return plan ? plan->nnz : 0; }
extern "C" void dazim_plan_destroy(dazim_plan* plan) {
  if (plan) { cudaSetDevice(plan->h->dev); plan_free(plan); }
}

extern "C" int dazim_plan_fetch(dazim_plan* P, float* dsurf, float* taa, long long* rowptr, int* col, float* val) {
  if (!P) return DAZIM_EBADARG;
  dazim_handle* h = P->h;
  CK(cudaSetDevice(h->dev));
  cudaStream_t st = h->st;
  h->times.d2h_bytes = 0;
  const size_t n = (size_t)P->nrow;
  if (dsurf && n) { CK(cudaMemcpyAsync(dsurf, P->d_dsurf.p, n * 4, cudaMemcpyDeviceToHost, st)); h->times.d2h_bytes += n * 4; }
  if (taa && n && P->mode == 0) { CK(cudaMemcpyAsync(taa, P->d_taa.p, n * 4, cudaMemcpyDeviceToHost, st)); h->times.d2h_bytes += n * 4; }
  if (P->mode != 0 && n) {
    if (rowptr) { CK(cudaMemcpyAsync(rowptr, P->d_rowptr.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st)); h->times.d2h_bytes += (n + 1) * 8; }
    if (col && P->nnz) { CK(cudaMemcpyAsync(col, P->d_col.p, (size_t)P->nnz * 4, cudaMemcpyDeviceToHost, st)); h->times.d2h_bytes += P->nnz * 4; }
    if (val && P->nnz) { CK(cudaMemcpyAsync(val, P->d_val.p, (size_t)P->nnz * 4, cudaMemcpyDeviceToHost, st)); h->times.d2h_bytes += P->nnz * 4; }
  }
  CK(cudaStreamSynchronize(st));
  return DAZIM_OK;
}

// name of the eikonal kernel this plan launches (bench.py / profiles name the dominant kernel by it)
extern "C" const char* dazim_plan_eikonal_kernel(const dazim_plan* P) {
  if (!P) return "";
  if (P->tps) return P->coh == 8 ? "k_fmm_coh<8>" : (P->coh == 16 ? "k_fmm_coh<16>" : (P->coh == 32 ? "k_fmm_coh<32>" : "k_fmm_tps"));
  return P->duo ? (P->duo_minb >= 16 ? "k_fmm_duo<16>" : "k_fmm_duo<10>") : (P->spc == 2 ? "k_fmm<2>" : "k_fmm<1>");
}

extern "C" int dazim_plan_device_ptrs(dazim_plan* P, void** dsurf, void** taa, void** rowptr, void** col, void** val) {
  if (!P) return DAZIM_EBADARG;
  if (dsurf) *dsurf = P->d_dsurf.p;
  if (taa) *taa = P->d_taa.p;
  if (rowptr) *rowptr = P->d_rowptr.p;
  if (col) *col = P->d_col.p;
  if (val) *val = P->d_val.p;
  return DAZIM_OK;
}

// ---------------------------------------------------------------------------
extern "C" int dazim_depthkernel(dazim_handle* h, int nx, int ny, int nz, const float* vel, double* pvRc,
                                 double* sen_vs, double* sen_vp, double* sen_rho, int kmaxRc, const double* tRc,
                                 const float* depz, float minthk) {
  if (!h) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  long long nl = 0;
  return th_depthkernel(h->st, nx, ny, nz, vel, pvRc, sen_vs, sen_vp, sen_rho, kmaxRc, tRc, depz, minthk,
                        &h->times.kernels_ms, &nl);
}

extern "C" int dazim_depthkernel_ti(dazim_handle* h, int nx, int ny, int nz, const float* vel, double* pvRc,
                                    int kmaxRc, const double* tRc, const float* depz, float minthk, float* Lsen_Gsc) {
  if (!h) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  long long nl = 0;
  return th_depthkernel_ti(h->st, nx, ny, nz, vel, pvRc, kmaxRc, tRc, depz, minthk, Lsen_Gsc, &h->times.kernels_ms, &nl);
}

extern "C" int dazim_surfdisp96(dazim_handle* h, int nprof, int nlayer, const float* thk, const float* vp,
                                const float* vs, const float* rho, int kmax, const double* t, double* cg) {
  if (!h) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  return th_surfdisp96(h->st, nprof, nlayer, thk, vp, vs, rho, kmax, t, cg);
}

extern "C" int dazim_gbuild(dazim_handle* h, int mode, const dazim_problem* p, dazim_tables* tables,
                            int tables_precomputed, const float* Gc, const float* Gs, float* dsurf, float* obsTaa,
                            double* tRcV, dazim_coo* coo) {
  if (!h || !p || !tables) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  const size_t nxy = (size_t)p->nx * p->ny;
  float kms = 0.0f;
  long long klaunch = 0;
  // scratch tables when the caller does not want them back
  std::vector<double> own_pv, own_s[3];
  std::vector<float> own_l;
  dazim_tables tb = *tables;
  if (!tb.pvRc) { own_pv.resize(nxy * p->kmaxRc); tb.pvRc = own_pv.data(); }
  if (!tables_precomputed) {
    if (mode == 0 || mode == 2) {
      if (!tb.Lsen_Gsc) { own_l.resize(nxy * p->kmaxRc * (p->nz - 1)); tb.Lsen_Gsc = own_l.data(); }
      float ms = 0;
      int st = th_depthkernel_ti(h->st, p->nx, p->ny, p->nz, p->vels, tb.pvRc, p->kmaxRc, p->tRc, p->depz,
                                 p->minthk, tb.Lsen_Gsc, &ms, &klaunch);
      if (st) return st;
      kms += ms;
    }
    if (mode == 1 || mode == 2) {
      double** ps[3] = {&tb.sen_vs, &tb.sen_vp, &tb.sen_rho};
      for (int i = 0; i < 3; ++i)
        if (!*ps[i]) { own_s[i].resize(nxy * p->kmaxRc * p->nz); *ps[i] = own_s[i].data(); }
      float ms = 0;
      int st = th_depthkernel(h->st, p->nx, p->ny, p->nz, p->vels, tb.pvRc, tb.sen_vs, tb.sen_vp, tb.sen_rho,
                              p->kmaxRc, p->tRc, p->depz, p->minthk, &ms, &klaunch);
      if (st) return st;
      kms += ms;
    }
  }
  dazim_plan* P = nullptr;
  int st = DAZIM_OK;
  st = plan_build(h, mode, p, &tb, Gc, Gs, 0, -1, 0, &P);
  if (st) return st;
  st = plan_run(P);       // grows the ray-footprint pool and re-runs if the estimate was too small
  if (st) { if (P) plan_free(P); return st; }
  h->times.kernels_ms = kms;
  h->times.n_launch += klaunch;
  if (coo && mode != 0) {
    coo->nar = P->nnz;
    if (P->nnz > coo->maxnar) { plan_free(P); return DAZIM_ENNZ_OVERFLOW; }
  }
  st = dazim_plan_fetch(P, dsurf, obsTaa, nullptr, coo ? coo->col : nullptr, coo ? coo->rw : nullptr);
  if (!st && coo && mode != 0 && coo->iw_row && P->nnz) {
    cudaError_t e = cudaMemcpy(coo->iw_row, P->d_rowid.p, (size_t)P->nnz * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) st = DAZIM_ECUDA + (int)e;
    h->times.d2h_bytes += P->nnz * 4;
  }
  plan_free(P);
  if (st) return st;
  // tRcV (FwdTraveltimeCPS.f90:771-778): interior nodes of the phase-velocity table
  if (tRcV) {
    const int nx = p->nx, ny = p->ny;
    for (int tt = 1; tt <= p->kmaxRc; ++tt)
      for (int jj = 1; jj <= ny - 2; ++jj)
        for (int ii = 1; ii <= nx - 2; ++ii)
          tRcV[(size_t)(jj - 1) * (nx - 2) + (ii - 1) + (size_t)(tt - 1) * (nx - 2) * (ny - 2)] =
              tb.pvRc[(size_t)jj * nx + ii + (size_t)(tt - 1) * nxy];
  }
  return DAZIM_OK;
}

// ---------------------------------------------------------------------------
// Multi-GPU behind the C ABI (what a Fortran / C host that links libdazim_b200.so gets: Main_Jt.f90:403-406 calls ONE
// subroutine).  Single process, one host thread + one stream per device:
//   stage A  depth kernels on strips of grid rows (nodes are independent), tables merged on the host;
//   stage B  contiguous (period, source) ranges balanced by ray count, one plan per device, no collective;
//   exchange the caller owns dsurf / rw / iw / col on the host (the reference's driver does), so every device copies
//            its row block straight to its offset of the caller's arrays: the "all-gather" is the D2H copy that the
//            single-GPU call does anyway, now over ndev PCIe links at once.
// Results are bit-identical to the single-device call for any ndev (row blocks are disjoint, no floating-point
// reduction crosses devices).
#include <map>
#include <mutex>
#include <thread>
static std::mutex g_multi_mu;
static std::map<int, dazim_handle*> g_multi_handles;

static int multi_handle(int dev, dazim_handle** out) {
  std::lock_guard<std::mutex> lk(g_multi_mu);
  auto it = g_multi_handles.find(dev);
  if (it != g_multi_handles.end()) { *out = it->second; return DAZIM_OK; }
  dazim_handle* h = nullptr;
  int st = dazim_create(&h, dev);
  if (st) return st;
  g_multi_handles[dev] = h;
  *out = h;
  return DAZIM_OK;
}

__global__ void k_add_row_offset(int* rowid, long long n, int off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rowid[i] += off;
}

extern "C" int dazim_gbuild_multi(int ndev, const int* devices, int mode, const dazim_problem* p, dazim_tables* tables,
                                  int tables_precomputed, const float* Gc, const float* Gs, float* dsurf, float* obsTaa,
                                  double* tRcV, dazim_coo* coo, dazim_times* times_max) {
  if (ndev < 1 || !devices || !p || !tables) return DAZIM_EBADARG;
  if (mode < 0 || mode > 2) return DAZIM_EBADARG;
  std::vector<dazim_handle*> H(ndev, nullptr);
  for (int r = 0; r < ndev; ++r) {
    for (int q = 0; q < r; ++q) if (devices[q] == devices[r]) return DAZIM_EBADARG;
    int st = multi_handle(devices[r], &H[r]);
    if (st) return st;
  }
  const int nx = p->nx, ny = p->ny, nz = p->nz, kmax = p->kmaxRc;
  const size_t nxy = (size_t)nx * ny;
  std::vector<double> own_pv, own_s[3];
  std::vector<float> own_l;
  dazim_tables tb = *tables;
  if (!tb.pvRc) { own_pv.resize(nxy * kmax); tb.pvRc = own_pv.data(); }
  std::vector<int> status(ndev, DAZIM_OK);
  std::vector<float> kms(ndev, 0.0f);
  // ---- stage A: strips of grid rows ----
  if (!tables_precomputed) {
    if ((mode == 0 || mode == 2) && !tb.Lsen_Gsc) { own_l.resize(nxy * kmax * (nz - 1)); tb.Lsen_Gsc = own_l.data(); }
    if (mode == 1 || mode == 2) {
      double** ps[3] = {&tb.sen_vs, &tb.sen_vp, &tb.sen_rho};
      for (int i = 0; i < 3; ++i)
        if (!*ps[i]) { own_s[i].resize(nxy * kmax * nz); *ps[i] = own_s[i].data(); }
    }
    auto stageA = [&](int r) {
      const int j0 = (int)((long long)ny * r / ndev), j1 = (int)((long long)ny * (r + 1) / ndev), nj = j1 - j0;
      if (nj <= 0) return;
      if (cudaSetDevice(H[r]->dev) != cudaSuccess) { status[r] = DAZIM_ECUDA; return; }
      const size_t ns = (size_t)nx * nj;
      std::vector<float> vel(ns * nz);
      for (int k = 0; k < nz; ++k)
        std::memcpy(&vel[(size_t)k * ns], p->vels + (size_t)k * nxy + (size_t)j0 * nx, ns * sizeof(float));
      std::vector<double> pv(ns * kmax);
      long long nl = 0;
      auto scatter2 = [&](const double* src, double* dst, int nlast) {     // (ns, kmax[, nlast]) -> (nxy, kmax[, nlast])
        for (int c = 0; c < kmax * nlast; ++c)
          std::memcpy(dst + (size_t)c * nxy + (size_t)j0 * nx, src + (size_t)c * ns, ns * sizeof(double));
      };
      if (mode == 0 || mode == 2) {
        std::vector<float> L(ns * kmax * (nz - 1));
        float ms = 0;
        int st = th_depthkernel_ti(H[r]->st, nx, nj, nz, vel.data(), pv.data(), kmax, p->tRc, p->depz, p->minthk, L.data(), &ms, &nl);
        if (st) { status[r] = st; return; }
        kms[r] += ms;
        for (int c = 0; c < kmax * (nz - 1); ++c)
          std::memcpy(tb.Lsen_Gsc + (size_t)c * nxy + (size_t)j0 * nx, &L[(size_t)c * ns], ns * sizeof(float));
      }
      if (mode == 1 || mode == 2) {
        std::vector<double> a(ns * kmax * nz), b(ns * kmax * nz), c3(ns * kmax * nz);
        float ms = 0;
        int st = th_depthkernel(H[r]->st, nx, nj, nz, vel.data(), pv.data(), a.data(), b.data(), c3.data(), kmax, p->tRc,
                                p->depz, p->minthk, &ms, &nl);
        if (st) { status[r] = st; return; }
        kms[r] += ms;
        scatter2(a.data(), tb.sen_vs, nz); scatter2(b.data(), tb.sen_vp, nz); scatter2(c3.data(), tb.sen_rho, nz);
      }
      scatter2(pv.data(), tb.pvRc, 1);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r) th.emplace_back(stageA, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; ++r) if (status[r]) return status[r];
  }
  // ---- stage B: contiguous unit ranges balanced by rays ----
  std::vector<long long> rays;      // per unit in loop order
  for (int knumi = 1; knumi <= p->kmax; ++knumi)
    for (int srcnum = 1; srcnum <= p->nsrcsurf1[knumi - 1]; ++srcnum)
      rays.push_back(p->nrc1[(size_t)(srcnum - 1) + (size_t)(knumi - 1) * p->nsrc]);
  const long long nunit = (long long)rays.size();
  long long total = 0;
  for (auto v : rays) total += v;
  std::vector<long long> bound(ndev + 1, nunit);
  bound[0] = 0;
  {
    long long acc = 0, u = 0;
    for (int r = 1; r < ndev; ++r) {
      const long long target = total * r / ndev;
      while (u < nunit && acc + rays[u] / 2 < target) { acc += rays[u]; ++u; }
      bound[r] = u;
    }
  }
  std::vector<dazim_plan*> P(ndev, nullptr);
  auto stageB = [&](int r) {
    if (bound[r + 1] <= bound[r]) return;
    int st = plan_build(H[r], mode, p, &tb, Gc, Gs, bound[r], bound[r + 1], 0, &P[r]);
    if (!st) st = plan_run(P[r]);
    status[r] = st;
  };
  {
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r) th.emplace_back(stageB, r);
    for (auto& t : th) t.join();
  }
  int st = DAZIM_OK;
  for (int r = 0; r < ndev; ++r) if (status[r] && !st) st = status[r];
  std::vector<long long> off(ndev + 1, 0);
  for (int r = 0; r < ndev; ++r) off[r + 1] = off[r] + (P[r] ? P[r]->nnz : 0);
  if (!st && coo && mode != 0) {
    coo->nar = off[ndev];
    if (off[ndev] > coo->maxnar) st = DAZIM_ENNZ_OVERFLOW;
  }
  if (!st) {
    auto fetch = [&](int r) {
      if (!P[r]) return;
      dazim_plan* Q = P[r];
      int s2 = dazim_plan_fetch(Q, dsurf ? dsurf + Q->row0 : nullptr, (obsTaa && mode == 0) ? obsTaa + Q->row0 : nullptr, nullptr,
                                (coo && mode != 0) ? coo->col + off[r] : nullptr, (coo && mode != 0) ? coo->rw + off[r] : nullptr);
      if (!s2 && coo && mode != 0 && coo->iw_row && Q->nnz) {
        cudaStream_t cs = Q->h->st;
        if (Q->row0) k_add_row_offset<<<(unsigned)((Q->nnz + 255) / 256), 256, 0, cs>>>(Q->d_rowid.p, Q->nnz, (int)Q->row0);
        cudaError_t e = cudaMemcpyAsync(coo->iw_row + off[r], Q->d_rowid.p, (size_t)Q->nnz * 4, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
        if (e != cudaSuccess) s2 = DAZIM_ECUDA + (int)e;
        Q->h->times.d2h_bytes += Q->nnz * 4;
      }
      status[r] = s2;
    };
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r) th.emplace_back(fetch, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; ++r) if (status[r] && !st) st = status[r];
  }
  if (times_max) {
    std::memset(times_max, 0, sizeof(*times_max));
    for (int r = 0; r < ndev; ++r) {
      const dazim_times& T = H[r]->times;
      times_max->kernels_ms = std::max(times_max->kernels_ms, kms[r]);
      times_max->dice_ms = std::max(times_max->dice_ms, T.dice_ms);
      times_max->fmm_ms = std::max(times_max->fmm_ms, T.fmm_ms);
      times_max->trace_ms = std::max(times_max->trace_ms, T.trace_ms);
      times_max->assemble_ms = std::max(times_max->assemble_ms, T.assemble_ms);
      times_max->total_ms = std::max(times_max->total_ms, T.total_ms);
      if (P[r]) { times_max->n_accept += T.n_accept; times_max->n_steps += T.n_steps; times_max->n_launch += T.n_launch;
                  times_max->n_fmm_launch += T.n_fmm_launch; times_max->n_trace_launch += T.n_trace_launch;
                  times_max->h2d_bytes += T.h2d_bytes; times_max->d2h_bytes += T.d2h_bytes; times_max->rbint |= T.rbint; }
    }
  }
  for (int r = 0; r < ndev; ++r) if (P[r]) plan_free(P[r]);
  if (st) return st;
  if (tRcV)
    for (int tt = 1; tt <= kmax; ++tt)
      for (int jj = 1; jj <= ny - 2; ++jj)
        for (int ii = 1; ii <= nx - 2; ++ii)
          tRcV[(size_t)(jj - 1) * (nx - 2) + (ii - 1) + (size_t)(tt - 1) * (nx - 2) * (ny - 2)] =
              tb.pvRc[(size_t)jj * nx + ii + (size_t)(tt - 1) * nxy];
  return DAZIM_OK;
}

extern "C" int dazim_lsmr(dazim_handle* h, int m, int n, long long nnz, const int* iw_row, const int* col,
                          const float* rw, const float* b, float damp, float atol, float btol, float conlim, int itnlim,
                          int localSize, float* x, dazim_lsmr_info* info) {
  if (!h) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  return dzl::lsmr_solve(h->st, m, n, nnz, iw_row, col, rw, b, damp, atol, btol, conlim, itnlim, localSize, x, info, false);
}

extern "C" int dazim_plan_lsmr(dazim_plan* P, const float* b, float damp, float atol, float btol, float conlim,
                               int itnlim, int localSize, float* x, dazim_lsmr_info* info) {
  if (!P || P->mode == 0 || P->nnz <= 0) return DAZIM_EBADARG;
  dazim_handle* h = P->h;
  CK(cudaSetDevice(h->dev));
  const int nblk = (P->mode == 2) ? 3 : 1;
  const long long ncol = (long long)nblk * P->g.nvx * P->g.nvz * (P->nz - 1);
  return dzl::lsmr_solve(h->st, (int)P->nrow, (int)ncol, P->nnz, P->d_rowid.p, P->d_col.p, P->d_val.p, b, damp, atol, btol,
                         conlim, itnlim, localSize, x, info, true);
}

extern "C" int dazim_lsmr_rows(dazim_handle* h, dazim_comm* comm, int m_local, long long m_total, int n, long long nnz,
                               const int* iw_row, const int* col, const float* rw, const float* b, float damp, float atol,
                               float btol, float conlim, int itnlim, int localSize, float* x, dazim_lsmr_info* info) {
  if (!h || !comm || m_total < m_local) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  const dzl::Coll coll{comm, dzc::sum_f32, dzc::sum_f64, m_total};
  return dzl::lsmr_solve(h->st, m_local, n, nnz, iw_row, col, rw, b, damp, atol, btol, conlim, itnlim, localSize, x, info,
                         false, &coll);
}

extern "C" int dazim_plan_lsmr_rows(dazim_plan* P, dazim_comm* comm, long long m_total, const float* b, float damp,
                                    float atol, float btol, float conlim, int itnlim, int localSize, float* x,
                                    dazim_lsmr_info* info) {
  if (!P || !comm || P->mode == 0 || P->nnz <= 0 || m_total < P->nrow) return DAZIM_EBADARG;
  dazim_handle* h = P->h;
  CK(cudaSetDevice(h->dev));
  const int nblk = (P->mode == 2) ? 3 : 1;
  const long long ncol = (long long)nblk * P->g.nvx * P->g.nvz * (P->nz - 1);
  const dzl::Coll coll{comm, dzc::sum_f32, dzc::sum_f64, m_total};
  return dzl::lsmr_solve(h->st, (int)P->nrow, (int)ncol, P->nnz, P->d_rowid.p, P->d_col.p, P->d_val.p, b, damp, atol, btol,
                         conlim, itnlim, localSize, x, info, true, &coll);
}

// ---------------------------------------------------------------------------
// Outer-iteration tail: Main_Jt.f90:416-727 on the device-resident G of a plan.

extern "C" int dazim_plan_update_model(dazim_plan* P, const float* vels, const dazim_tables* tb) {
  if (!P || !tb || !tb->pvRc) return DAZIM_EBADARG;
  const int mode = P->mode;
  if ((mode == 0 || mode == 2) && !tb->Lsen_Gsc) return DAZIM_EBADARG;
  if ((mode == 1 || mode == 2) && (!tb->sen_vs || !tb->sen_vp || !tb->sen_rho || !vels)) return DAZIM_EBADARG;
  dazim_handle* h = P->h;
  CK(cudaSetDevice(h->dev));
  cudaStream_t st = h->st;
  const size_t nxy = (size_t)P->nx * P->ny, nlay = (size_t)P->nz - 1, k = (size_t)P->kmaxRc;
  h->times.h2d_bytes = 0;
  P->h_velv.resize(nxy * k);
  for (size_t i = 0; i < P->h_velv.size(); ++i) P->h_velv[i] = (float)tb->pvRc[i];
  CK(cudaMemcpyAsync(P->d_velv.p, P->h_velv.data(), nxy * k * sizeof(float), cudaMemcpyHostToDevice, st));
  h->times.h2d_bytes += (long long)(nxy * k * sizeof(float));
  if (mode == 1 || mode == 2) {
    const size_t nb = nxy * k * P->nz * sizeof(double);
    CK(cudaMemcpyAsync(P->d_sen_vs.p, tb->sen_vs, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(P->d_sen_vp.p, tb->sen_vp, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(P->d_sen_rho.p, tb->sen_rho, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(P->d_vels.p, vels, nxy * P->nz * sizeof(float), cudaMemcpyHostToDevice, st));
    h->times.h2d_bytes += (long long)(3 * nb + nxy * P->nz * sizeof(float));
  }
  if (mode == 0 || mode == 2) {
    CK(cudaMemcpyAsync(P->d_lsen.p, tb->Lsen_Gsc, nxy * k * nlay * sizeof(float), cudaMemcpyHostToDevice, st));
    h->times.h2d_bytes += (long long)(nxy * k * nlay * sizeof(float));
  }
  CK(cudaStreamSynchronize(st));
  return DAZIM_OK;
}

extern "C" long long dazim_tikh_offset(int i, int j, int k, int nvx, int nvz, int nzm1) {
  return dzi::tikh_offset(i, j, k, nvx, nvz, nzm1);
}
extern "C" long long dazim_tikh_block_entries(int nvx, int nvz, int nzm1) {
  return dzi::tikh_block_entries(nvx, nvz, nzm1);
}

// Tikhonov rows behind `base` entries of (val, col, rowid); returns entries appended, rows appended and the entry
// count after the dVs block (joint).  TikhRegul.f90:2-105 (joint = 0) / :108-209 (joint = 1).
static int tikh_append(cudaStream_t st, int joint, int iso_inv, int nx, int ny, int nz, int maxvp, int dall, long long base,
                       float weightGcs, float weightVs, float* val, int* col, int* rowid, long long* appended,
                       int* count3, long long* after_vs) {
  const int nvx = nx - 2, nvz = ny - 2, nzm1 = nz - 1;
  const int ncell = nvx * nvz * nzm1;
  const long long per = dzi::tikh_block_entries(nvx, nvz, nzm1);
  long long o = base;
  int rows = 0;
  *after_vs = -1;
  if (joint) {
    CK(dzi::launch_tikh(nvx, nvz, nzm1, o, dall + rows + 1, 0, weightVs, val, col, rowid, st));
    o += per; rows += ncell;
    *after_vs = o;
    for (int sc = 1; sc <= 2; ++sc) {
      CK(dzi::launch_tikh(nvx, nvz, nzm1, o, dall + rows + 1, sc * maxvp, weightGcs, val, col, rowid, st));
      o += per; rows += ncell;
    }
  } else if (iso_inv) {
    CK(dzi::launch_tikh(nvx, nvz, nzm1, o, dall + rows + 1, 0, weightVs, val, col, rowid, st));
    o += per; rows += ncell;
  } else {
    for (int sc = 1; sc <= 2; ++sc) {
      CK(dzi::launch_tikh(nvx, nvz, nzm1, o, dall + rows + 1, (sc - 1) * maxvp, weightGcs, val, col, rowid, st));
      o += per; rows += ncell;
    }
  }
  *appended = o - base;
  *count3 = rows;
  return DAZIM_OK;
}

// The system the iteration tail works on: CSR rows of G in HBM (all rows, global row ids 1..nrow in rowid), with
// room behind the nnz entries of val/col/rowid for the regularisation rows, and the reference travel times.
struct IterSystem {
  int nx, ny, nz;
  long long nrow, nnz, cap;
  const long long* rowptr;   // [nrow+1]
  int* col; float* val; int* rowid;   // [cap]
  const float* dsurf;        // [nrow]
  float* vels;               // [nx*ny*nz] device workspace for the model update
};

// Row-distributed form (rd != nullptr): Y holds THIS rank's block of rows (global rows rd->row0 + 1 .. rd->row0 + Y.nrow of
// rd->dall); obst and every per-row output have the length of the WHOLE system.  What is O(rows) -- synthetic times,
// residuals, weights, their statistics -- is made complete on every rank by a sum over zero-padded arrays (x + 0 is
// exact) and then computed redundantly by the same kernels in the same order as on one GPU, so everything but the LSMR
// solution is bit-identical to the single-GPU tail; what is O(non-zeros) -- row scaling, DWS, products with G, the
// solve -- stays with the rank that owns the rows.  The regularisation rows live on the last rank.
struct RowsDist {
  const dzl::Coll* coll;
  long long row0, dall;
  int rank, nranks;
};
#define CKC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

static int iterate_core(dazim_handle* h, const IterSystem& Y, const float* obst, const dazim_iter_params* prm, float* vsf,
                        float* dv, float* gcf, float* gsf, float* dws, float* sigmaT, float* resbst, float* fwdTvs,
                        float* fwdTaa, dazim_iter_stats* S, const RowsDist* rd = nullptr) {
  CK(cudaSetDevice(h->dev));
  g_alloc_stream = h->st;
  cudaStream_t st = h->st;
  const int iso = prm->iso_inv ? 1 : 0;
  const int nx = Y.nx, ny = Y.ny, nz = Y.nz;
  const int maxvp = (nx - 2) * (ny - 2) * (nz - 1);
  const int nblk = iso ? 1 : 3;
  const int n = nblk * maxvp;
  const long long dall_ll = rd ? rd->dall : Y.nrow;
  if (dall_ll > 0x7fffffffll - 3ll * maxvp || (rd && (rd->row0 < 0 || rd->row0 + Y.nrow > rd->dall))) return DAZIM_EBADARG;
  const int dall = (int)dall_ll;               // rows of the whole system
  const int dl = (int)Y.nrow;                  // rows held here
  const long long r0 = rd ? rd->row0 : 0;
  const bool own_reg = !rd || rd->rank == rd->nranks - 1;
  std::memset(S, 0, sizeof(*S));
  struct Events {                               // destroyed on every return path
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); }
  } ev;
  for (auto& x : ev.e) CK(cudaEventCreate(&x));
  cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2], e3 = ev.e[3];
  DBuf<float> d_obst, d_cbst, d_tdata, d_dt, d_sig, d_w, d_stats, d_b, d_dv, d_gcf, d_gsf, d_tvs, d_taa, d_res, d_resw,
      d_lm, d_lmw, d_dws;
  DBuf<double> d_partial, d_dwsacc;
  DBuf<float> d_dsyn;
  const long long reg_entries = (long long)nblk * dzi::tikh_block_entries(nx - 2, ny - 2, nz - 1);
  if (own_reg && Y.nnz + reg_entries > Y.cap) return DAZIM_ENNZ_OVERFLOW;   // the caller reserves this room (plan_run does)
  CK(d_obst.alloc(dall)); CK(d_cbst.alloc(dall)); CK(d_tdata.alloc(dall)); CK(d_dt.alloc(dall)); CK(d_sig.alloc(dall));
  CK(d_w.alloc(dall)); CK(d_stats.alloc(64)); CK(d_b.alloc((size_t)dall + (size_t)nblk * maxvp)); CK(d_dv.alloc(n));
  CK(d_gcf.alloc(maxvp)); CK(d_gsf.alloc(maxvp)); CK(d_tvs.alloc(dall)); CK(d_taa.alloc(dall)); CK(d_res.alloc(dall));
  CK(d_resw.alloc(dall)); CK(d_lm.alloc(reg_entries)); CK(d_lmw.alloc(reg_entries)); CK(d_partial.alloc(1024));
  if (iso && dws) { CK(d_dws.alloc(maxvp)); CK(d_dwsacc.alloc(maxvp)); }
  float hs[64];
  std::memset(hs, 0, sizeof(hs));
  CK(cudaMemsetAsync(d_stats.p, 0, sizeof(float) * 64, st));     // slots a mode does not use read back as zero
  CK(cudaMemcpyAsync(d_obst.p, obst, sizeof(float) * dall, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(e0, st));
  const float* dsyn = Y.dsurf;
  if (rd) {                                    // the synthetic times of every rank's rays
    CK(d_dsyn.alloc(dall));
    CK(cudaMemsetAsync(d_dsyn.p, 0, sizeof(float) * dall, st));
    CK(cudaMemcpyAsync(d_dsyn.p + r0, Y.dsurf, sizeof(float) * dl, cudaMemcpyDeviceToDevice, st));
    CKC(rd->coll->sum_f32(rd->coll->ctx, d_dsyn.p, (size_t)dall, st));
    dsyn = d_dsyn.p;
  }
  // ---- residual of the reference model, CalDdatSigma, weights (Main_Jt.f90:425-469) ----
  CK(dzi::launch_resid(dall, d_obst.p, dsyn, d_cbst.p, d_tdata.p, d_dt.p, st));
  {
    const float* arrs[2] = {d_cbst.p, d_dt.p};
    CK(dzi::launch_seq_stats(2, arrs, dall, d_stats.p, st));                       // [0..2] cbst, [3..5] deltaT
  }
  CK(dzi::launch_norm2(d_cbst.p, dall, d_partial.p, d_stats.p + 42, st));          // ||cbst|| before weighting
  CK(dzi::launch_sigma(dall, d_dt.p, d_obst.p, d_stats.p + 3, d_sig.p, d_w.p, d_cbst.p, st));
  {
    const float* arrs[2] = {d_w.p, d_cbst.p};
    CK(dzi::launch_seq_stats(2, arrs, dall, d_stats.p + 6, st));                   // [6] sum w, [10] sum |cbst_w|
  }
  CK(cudaEventRecord(e2, st));
  CK(dzi::launch_scale_rows(Y.nrow, Y.rowptr, d_w.p + r0, Y.val, st));
  CK(cudaEventRecord(e3, st));
  if (iso && dws) {
    CK(dzi::launch_dws(Y.nnz, Y.col, Y.val, maxvp, d_dwsacc.p, d_dws.p, st));
    if (rd) {
      CKC(rd->coll->sum_f64(rd->coll->ctx, d_dwsacc.p, (size_t)maxvp, st));
      CK(dzi::launch_dws_finish(maxvp, d_dwsacc.p, d_dws.p, st));
    }
  }
  // ---- regularisation rows behind G (Main_Jt.f90:507-520) ----
  long long appended = 0, after_vs = -1;
  int count3 = 0;
  if (own_reg) {
    int rc = tikh_append(st, iso ? 0 : 1, iso, nx, ny, nz, maxvp, dl, Y.nnz, prm->weightGcs, prm->weightVs, Y.val,
                         Y.col, Y.rowid, &appended, &count3, &after_vs);
    if (rc) return rc;
  }
  const long long nar1 = Y.nnz, nar = Y.nnz + appended;     // of this rank's block
  const int m = dl + count3;
  long long nar1_all = nar1, nar_all = nar, count3_all = count3;
  if (rd) {                                    // sizes of the whole system (integers below 2^53: exact in double)
    double hc[3] = {(double)nar1, (double)nar, (double)count3};
    CK(cudaMemcpyAsync(d_partial.p, hc, sizeof(hc), cudaMemcpyHostToDevice, st));
    CKC(rd->coll->sum_f64(rd->coll->ctx, d_partial.p, 3, st));
    CK(cudaMemcpyAsync(hc, d_partial.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    nar1_all = (long long)hc[0]; nar_all = (long long)hc[1]; count3_all = (long long)hc[2];
  }
  CK(cudaMemcpyAsync(d_b.p, d_cbst.p + r0, sizeof(float) * dl, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(d_b.p + dl, 0, sizeof(float) * (size_t)count3, st));
  // ---- LSMR on [W G ; L] (Main_Jt.f90:527-564) ----
  float atol = prm->atol, btol = prm->btol, conlim = prm->conlim;
  int itnlim = prm->itnlim, localSize = prm->localSize;
  if (prm->use_ref_controls) {
    if (iso) { atol = 1e-3f; btol = 1e-3f; conlim = 1200.0f; itnlim = 1000; localSize = n / 4; }
    else { atol = 1e-5f; btol = 1e-4f; conlim = 200.0f; itnlim = 500; localSize = 10; }
  }
  {
    dzl::Coll cl;
    if (rd) { cl = *rd->coll; cl.m_total = (long long)dall + count3_all; }
    int rc = dzl::lsmr_solve(st, m, n, nar, Y.rowid, Y.col, Y.val, d_b.p, prm->damp, atol, btol, conlim,
                             itnlim, localSize, d_dv.p, &S->lsmr, true, rd ? &cl : nullptr);
    if (rc) return rc;
  }
  // ---- model update (Main_Jt.f90:582-620) ----
  CK(cudaMemcpyAsync(Y.vels, vsf, sizeof(float) * (size_t)nx * ny * nz, cudaMemcpyHostToDevice, st));
  CK(dzi::launch_model_update(nx, ny, nz, iso, d_dv.p, Y.vels, prm->minvel, prm->maxvel, d_gcf.p, d_gsf.p, st));
  // ---- ||Lm|| and the residual of the solution from the sparse rows (CalSigamNorm.f90) ----
  const long long nre = nar - nar1, nre_vs = iso ? nre : (own_reg ? after_vs - nar1 : 0);
  CK(dzi::launch_lm_terms(nre, nre_vs, Y.val + nar1, Y.col + nar1, d_dv.p, prm->weightVs, prm->weightGcs, d_lm.p,
                          d_lmw.p, st));
  if (iso) {
    CK(dzi::launch_norm2(d_lm.p, nre, d_partial.p, d_stats.p + 36, st));
    CK(dzi::launch_norm2(d_lmw.p, nre, d_partial.p, d_stats.p + 37, st));
  } else {
    CK(dzi::launch_norm2(d_lm.p, nre_vs, d_partial.p, d_stats.p + 32, st));
    CK(dzi::launch_norm2(d_lmw.p, nre_vs, d_partial.p, d_stats.p + 33, st));
    CK(dzi::launch_norm2(d_lm.p + nre_vs, nre - nre_vs, d_partial.p, d_stats.p + 34, st));
    CK(dzi::launch_norm2(d_lmw.p + nre_vs, nre - nre_vs, d_partial.p, d_stats.p + 35, st));
    CK(dzi::launch_norm2(d_lm.p, nre, d_partial.p, d_stats.p + 36, st));
    CK(dzi::launch_norm2(d_lmw.p, nre, d_partial.p, d_stats.p + 37, st));
  }
  if (rd) {
    // the model norms exist on the rank that owns the regularisation rows (zeros elsewhere): the sum hands them round
    CKC(rd->coll->sum_f32(rd->coll->ctx, d_stats.p + 32, 6, st));
    CK(cudaMemsetAsync(d_tvs.p, 0, sizeof(float) * dall, st)); CK(cudaMemsetAsync(d_taa.p, 0, sizeof(float) * dall, st));
    CK(cudaMemsetAsync(d_res.p, 0, sizeof(float) * dall, st)); CK(cudaMemsetAsync(d_resw.p, 0, sizeof(float) * dall, st));
  }
  CK(dzi::launch_resid_rows(Y.nrow, Y.rowptr, Y.col, Y.val, d_dv.p, d_w.p + r0, maxvp, nblk, d_tdata.p + r0, d_tvs.p + r0,
                            d_taa.p + r0, d_res.p + r0, d_resw.p + r0, st));
  if (rd) {
    CKC(rd->coll->sum_f32(rd->coll->ctx, d_tvs.p, (size_t)dall, st)); CKC(rd->coll->sum_f32(rd->coll->ctx, d_taa.p, (size_t)dall, st));
    CKC(rd->coll->sum_f32(rd->coll->ctx, d_res.p, (size_t)dall, st)); CKC(rd->coll->sum_f32(rd->coll->ctx, d_resw.p, (size_t)dall, st));
  }
  {
    const float* arrs[3] = {d_res.p, d_taa.p, d_tvs.p};
    CK(dzi::launch_seq_stats(3, arrs, dall, d_stats.p + 12, st));                  // [12..14] res, [16] |taa|, [19] |tvs|
  }
  CK(dzi::launch_norm2(d_res.p, dall, d_partial.p, d_stats.p + 40, st));
  CK(dzi::launch_norm2(d_resw.p, dall, d_partial.p, d_stats.p + 41, st));
  CK(cudaEventRecord(e1, st));
  // ---- results to the host ----
  CK(cudaMemcpyAsync(hs, d_stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(vsf, Y.vels, sizeof(float) * (size_t)nx * ny * nz, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(dv, d_dv.p, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  if (!iso && gcf) CK(cudaMemcpyAsync(gcf, d_gcf.p, sizeof(float) * maxvp, cudaMemcpyDeviceToHost, st));
  if (!iso && gsf) CK(cudaMemcpyAsync(gsf, d_gsf.p, sizeof(float) * maxvp, cudaMemcpyDeviceToHost, st));
  if (iso && dws) CK(cudaMemcpyAsync(dws, d_dws.p, sizeof(float) * maxvp, cudaMemcpyDeviceToHost, st));
  if (sigmaT) CK(cudaMemcpyAsync(sigmaT, d_sig.p, sizeof(float) * dall, cudaMemcpyDeviceToHost, st));
  if (resbst) CK(cudaMemcpyAsync(resbst, d_res.p, sizeof(float) * dall, cudaMemcpyDeviceToHost, st));
  if (fwdTvs) CK(cudaMemcpyAsync(fwdTvs, d_tvs.p, sizeof(float) * dall, cudaMemcpyDeviceToHost, st));
  if (fwdTaa) CK(cudaMemcpyAsync(fwdTaa, d_taa.p, sizeof(float) * dall, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  const float fd = (float)dall;
  S->before[0] = hs[1] / fd;                     // abs mean
  S->before[1] = std::sqrt(hs[2] / fd);          // std
  S->before[2] = hs[42] / std::sqrt(fd);         // RMS = dnrm2/sqrt(real(dall))
  S->before[3] = hs[0] / fd;                     // mean
  S->meandeltaT = hs[3] / fd;
  S->mean_weight = hs[6] / fd;
  S->meanabs_weighted = hs[10] / fd;
  S->after[0] = hs[13] / fd;
  S->after[1] = std::sqrt(hs[14] / fd);
  S->after[2] = hs[40] / std::sqrt(fd);
  S->after[3] = hs[12] / fd;
  S->meanabs_Taa = hs[16] / fd;
  S->meanabs_Tvs = hs[19] / fd;
  for (int q = 0; q < 6; ++q) S->norms[q] = hs[32 + q];
  S->res2Nm = hs[40];
  S->resW2Nm = hs[41];
  S->nar1 = nar1_all; S->nar = nar_all; S->count3 = (int)count3_all;
  cudaEventElapsedTime(&S->step_ms, e0, e1);
  cudaEventElapsedTime(&S->scale_ms, e2, e3);
  return DAZIM_OK;
}

extern "C" int dazim_plan_iterate(dazim_plan* P, const float* obst, const dazim_iter_params* prm, float* vsf, float* dv,
                                  float* gcf, float* gsf, float* dws, float* sigmaT, float* resbst, float* fwdTvs,
                                  float* fwdTaa, dazim_iter_stats* S) {
  if (!P || !obst || !prm || !vsf || !dv || !S) return DAZIM_EBADARG;
  if (P->mode != 1 && P->mode != 2) return DAZIM_EBADARG;
  if ((P->mode == 1) != (prm->iso_inv != 0)) return DAZIM_EBADARG;
  if (P->row0 != 0 || P->nrow < 1 || P->nnz < 1) return DAZIM_EBADARG;
  IterSystem Y;
  Y.nx = P->nx; Y.ny = P->ny; Y.nz = P->nz; Y.nrow = P->nrow; Y.nnz = P->nnz; Y.cap = P->val_cap;
  Y.rowptr = P->d_rowptr.p; Y.col = P->d_col.p; Y.val = P->d_val.p; Y.rowid = P->d_rowid.p; Y.dsurf = P->d_dsurf.p;
  Y.vels = P->d_vels.p;
  return iterate_core(P->h, Y, obst, prm, vsf, dv, gcf, gsf, dws, sigmaT, resbst, fwdTvs, fwdTaa, S);
}

// The tail with the rows of G left where they were built: no gather of G, row-distributed LSMR.
extern "C" int dazim_plan_iterate_rows(dazim_plan* P, dazim_comm* comm, long long dall_total, const float* obst,
                                       const dazim_iter_params* prm, float* vsf, float* dv, float* gcf, float* gsf,
                                       float* dws, float* sigmaT, float* resbst, float* fwdTvs, float* fwdTaa,
                                       dazim_iter_stats* S) {
  if (!P || !comm || !obst || !prm || !vsf || !dv || !S) return DAZIM_EBADARG;
  if (P->mode != 1 && P->mode != 2) return DAZIM_EBADARG;
  if ((P->mode == 1) != (prm->iso_inv != 0)) return DAZIM_EBADARG;
  if (P->nrow < 1 || P->nnz < 1 || dall_total < P->row0 + P->nrow) return DAZIM_EBADARG;
  IterSystem Y;
  Y.nx = P->nx; Y.ny = P->ny; Y.nz = P->nz; Y.nrow = P->nrow; Y.nnz = P->nnz; Y.cap = P->val_cap;
  Y.rowptr = P->d_rowptr.p; Y.col = P->d_col.p; Y.val = P->d_val.p; Y.rowid = P->d_rowid.p; Y.dsurf = P->d_dsurf.p;
  Y.vels = P->d_vels.p;
  const dzl::Coll coll{comm, dzc::sum_f32, dzc::sum_f64, 0};
  const RowsDist rd{&coll, P->row0, dall_total, dazim_comm_rank(comm), dazim_comm_size(comm)};
  return iterate_core(P->h, Y, obst, prm, vsf, dv, gcf, gsf, dws, sigmaT, resbst, fwdTvs, fwdTaa, S, &rd);
}

// The same tail on a system the caller holds in HBM -- the row blocks of several ranks after the NCCL all-gather
// ("exchange before the solver").  All d_* are DEVICE pointers on the handle's device, complete before the
// call (synchronise the producing stream); d_val / d_col / d_rowid have cap >= nnz + regularisation entries
// (dazim_tikh_block_entries x 1 or 3); they are modified (weighted, rows appended).
extern "C" int dazim_iterate_device(dazim_handle* h, int nx, int ny, int nz, long long nrow, long long nnz, long long cap,
                                    const long long* d_rowptr, int* d_col, float* d_val, int* d_rowid,
                                    const float* d_dsurf, const float* obst, const dazim_iter_params* prm, float* vsf,
                                    float* dv, float* gcf, float* gsf, float* dws, float* sigmaT, float* resbst,
                                    float* fwdTvs, float* fwdTaa, dazim_iter_stats* S) {
  if (!h || !d_rowptr || !d_col || !d_val || !d_rowid || !d_dsurf || !obst || !prm || !vsf || !dv || !S) return DAZIM_EBADARG;
  if (nx < 5 || ny < 5 || nz < 2 || nrow < 1 || nnz < 1 || cap < nnz) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  g_alloc_stream = h->st;
  DBuf<float> d_vels;
  CK(d_vels.alloc((size_t)nx * ny * nz));
  IterSystem Y;
  Y.nx = nx; Y.ny = ny; Y.nz = nz; Y.nrow = nrow; Y.nnz = nnz; Y.cap = cap;
  Y.rowptr = d_rowptr; Y.col = d_col; Y.val = d_val; Y.rowid = d_rowid; Y.dsurf = d_dsurf; Y.vels = d_vels.p;
  return iterate_core(h, Y, obst, prm, vsf, dv, gcf, gsf, dws, sigmaT, resbst, fwdTvs, fwdTaa, S);
}

extern "C" int dazim_cal_ddat_sigma(dazim_handle* h, int dall, const float* obst, const float* cbst, float* sigmaT,
                                    float* meandeltaT) {
  if (!h || dall < 1 || !obst || !cbst || !sigmaT || !meandeltaT) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  g_alloc_stream = h->st;
  cudaStream_t st = h->st;
  DBuf<float> d_obst, d_cbst, d_dt, d_sig, d_w, d_stats;
  CK(d_obst.alloc(dall)); CK(d_cbst.alloc(dall)); CK(d_dt.alloc(dall)); CK(d_sig.alloc(dall)); CK(d_w.alloc(dall));
  CK(d_stats.alloc(8));
  CK(cudaMemcpyAsync(d_obst.p, obst, sizeof(float) * dall, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_cbst.p, cbst, sizeof(float) * dall, cudaMemcpyHostToDevice, st));
  CK(dzi::launch_delta(dall, d_cbst.p, d_obst.p, d_dt.p, st));
  const float* arrs[1] = {d_dt.p};
  CK(dzi::launch_seq_stats(1, arrs, dall, d_stats.p, st));
  CK(dzi::launch_sigma(dall, d_dt.p, d_obst.p, d_stats.p, d_sig.p, d_w.p, d_cbst.p, st));
  float hs[3];
  CK(cudaMemcpyAsync(hs, d_stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(sigmaT, d_sig.p, sizeof(float) * dall, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  *meandeltaT = hs[0] / (float)dall;
  return DAZIM_OK;
}

extern "C" int dazim_tikhonov(dazim_handle* h, int joint, int nx, int ny, int nz, int maxvp, int dall, long long* nar,
                              float* rw, int* iw_row, int* col, long long* narVs, int* count3, int iso_inv,
                              float weightGcs, float weightVs) {
  if (!h || !nar || !rw || !iw_row || !col || !count3 || nx < 3 || ny < 3 || nz < 2 || *nar < 0) return DAZIM_EBADARG;
  if (maxvp != (nx - 2) * (ny - 2) * (nz - 1)) return DAZIM_EBADARG;
  CK(cudaSetDevice(h->dev));
  g_alloc_stream = h->st;
  cudaStream_t st = h->st;
  const int nblk = joint ? 3 : (iso_inv ? 1 : 2);
  const long long tot = (long long)nblk * dzi::tikh_block_entries(nx - 2, ny - 2, nz - 1);
  DBuf<float> d_val; DBuf<int> d_col, d_row;
  CK(d_val.alloc(tot)); CK(d_col.alloc(tot)); CK(d_row.alloc(tot));
  long long appended = 0, after_vs = -1;
  int rc = tikh_append(st, joint, iso_inv, nx, ny, nz, maxvp, dall, 0, weightGcs, weightVs, d_val.p, d_col.p, d_row.p,
                       &appended, count3, &after_vs);
  if (rc) return rc;
  CK(cudaMemcpyAsync(rw + *nar, d_val.p, sizeof(float) * appended, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(col + *nar, d_col.p, sizeof(int) * appended, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(iw_row + *nar, d_row.p, sizeof(int) * appended, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (narVs) *narVs = joint ? *nar + after_vs : -1;
  *nar += appended;
  return DAZIM_OK;
}

// ---------------------------------------------------------------------------
// test seams
static int mini_problem(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, int n, const float* scx,
                        const float* scz, const float* rcx, const float* rcz, dazim_problem& p,
                        std::vector<int>& periods, std::vector<int>& nrc1, std::vector<int>& ns1) {
  std::memset(&p, 0, sizeof(p));
  p.nx = nx; p.ny = ny; p.nz = 2; p.goxd = goxd; p.gozd = gozd; p.dvxd = dvxd; p.dvzd = dvzd;
  p.kmaxRc = 1; p.kmax = 1; p.nsrc = n; p.nrcf = 1;
  periods.assign(n, 1); nrc1.assign(n, rcx ? 1 : 0); ns1.assign(1, n);
  p.periods = periods.data(); p.nrc1 = nrc1.data(); p.nsrcsurf1 = ns1.data();
  p.scxf = scx; p.sczf = scz; p.rcxf = rcx ? rcx : scx; p.rczf = rcz ? rcz : scz;
  return 0;
}

extern "C" int dazim_fmm_solve(dazim_handle* h, int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                               const double* pv, int n, const float* scx, const float* scz, float* veln, float* ttn,
                               int* nsts, float* ttnr, int* nstsr, int* geom) {
  if (!h || !pv || n <= 0) return DAZIM_EBADARG;
  dazim_problem p;
  std::vector<int> pe, nr, ns;
  mini_problem(nx, ny, goxd, gozd, dvxd, dvzd, n, scx, scz, nullptr, nullptr, p, pe, nr, ns);
  dazim_tables tb;
  std::memset(&tb, 0, sizeof(tb));
  tb.pvRc = const_cast<double*>(pv);
  std::vector<float> dummyL((size_t)nx * ny, 0.0f);
  tb.Lsen_Gsc = dummyL.data();
  dazim_plan* P = nullptr;
  int st = plan_build(h, 0, &p, &tb, nullptr, nullptr, 0, -1, 1, &P);
  if (st) return st;
  if (P->batch_src0.size() != 2) { plan_free(P); return DAZIM_EBADARG; }   // test seam: single batch only
  st = plan_run(P);
  if (st) { plan_free(P); return st; }
  const size_t nc = (size_t)P->g.nnx * P->g.nnz;
  cudaError_t e = cudaSuccess;
  if (veln && e == cudaSuccess) e = cudaMemcpy(veln, P->d_veln_c.p, nc * 4, cudaMemcpyDeviceToHost);
  // decode K3's (E, hpos) encoding into the reference's (ttn, nsts) pair; hpos lives in the slot that solved s
  {
    std::vector<int> slot_of(n);
    if (e == cudaSuccess && !P->tps) e = cudaMemcpy(slot_of.data(), P->d_slot_of.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
    DBuf<float> d_t; DBuf<int> d_s;
    if (e == cudaSuccess) e = d_t.alloc(std::max(nc, (size_t)REF_N));
    if (e == cudaSuccess) e = d_s.alloc(std::max(nc, (size_t)REF_N));
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
      // the coarse hpos of a slot is only valid for the LAST solve it ran; the final coarse field is all alive anyway
      e = launch_decode_status(P->d_E_c.p + (size_t)i * coarse_field_size(P->g.nnx, P->g.nnz), nullptr, nc, P->g.nnz,
                               d_t.p, d_s.p, h->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
      if (ttn && e == cudaSuccess) e = cudaMemcpy(ttn + (size_t)i * nc, d_t.p, nc * 4, cudaMemcpyDeviceToHost);
      if (nsts && e == cudaSuccess) e = cudaMemcpy(nsts + (size_t)i * nc, d_s.p, nc * 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) break;
      const bool own_slot = ((size_t)P->spc * (size_t)P->nctas >= (size_t)n);   // one solve per slot: refined heap slots are intact
      const int* hp = nullptr;
      if (P->tps) hp = P->d_hpos_r_out.p ? P->d_hpos_r_out.p + (size_t)i * REF_N : nullptr;
      else if (own_slot) hp = P->d_hpos_r.p + (size_t)slot_of[i] * REF_N;
      e = launch_decode_status(P->d_E_r.p + (size_t)i * REF_N, hp, REF_N, 0, d_t.p, d_s.p, h->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
      if (ttnr && e == cudaSuccess) e = cudaMemcpy(ttnr + (size_t)i * REF_N, d_t.p, (size_t)REF_N * 4, cudaMemcpyDeviceToHost);
      if (nstsr && e == cudaSuccess) e = cudaMemcpy(nstsr + (size_t)i * REF_N, d_s.p, (size_t)REF_N * 4, cudaMemcpyDeviceToHost);
    }
  }
  if (geom)
    for (int i = 0; i < n; ++i) {
      const SrcRec& s = P->src[i];
      int* gg = geom + 8 * i;
      gg[0] = s.nnzr; gg[1] = s.nnxr; gg[2] = s.vnl; gg[3] = s.vnr; gg[4] = s.vnt; gg[5] = s.vnb;
      gg[6] = P->g.nnz; gg[7] = P->g.nnx;
    }
  plan_free(P);
  return e == cudaSuccess ? DAZIM_OK : DAZIM_ECUDA + (int)e;
}

// TEST SEAM, host only: the thread-per-solve eikonal code of dazim_tps.h (the very functions k_fmm_tps runs, they are
// __host__ __device__) executed on the CPU for ONE source, so that the logic can be compared with the oracle on a
// machine without a GPU (tests/test_tps_host_twin.py).  No product entry point calls this; it needs no device.
// The cohort kernel's "records computed ahead" protocol (coh_march_heap / coh_march_stencil in dazim_fmm.cu) replayed
// serially on the host: the four neighbour records of the PREDICTED next node are gathered either before (when = 0)
// or after (when = 1) the current node's updates are applied -- the two ends of the window in which the stencil threads
// read E on the device --, used in the next round if (node, key) were predicted right, and patched exactly as the heap
// lane patches them (a neighbour the previous round inserted is "close, position unknown").  stats: rounds predicted,
// not predicted, records patched.
template <int URG>
static void twin_march_ahead(TpsState& S, const TpsGrid& G, unsigned long long& nacc, const int when, long long* stats) {
  TpsNb spec[4];
  int spec_node = -1, spec_key = 0;
  int ins[4] = {-1, -1, -1, -1};
  for (;;) {
    TpsPre P;
    if (!tps_pre<URG>(S, G, nacc, P)) break;
    TpsNb N[4];
    const bool hit = (P.pn == spec_node && (int)P.tself == spec_key);
    for (int q = 0; q < 4; ++q) N[q] = hit ? spec[q] : tps_neighbour<URG>(G, P.ix, P.iz, P.tself, q);
    stats[hit ? 0 : 1] += 1;
    for (int q = 0; q < 4; ++q)
      if (N[q].qst == -1 && (N[q].co == ins[0] || N[q].co == ins[1] || N[q].co == ins[2] || N[q].co == ins[3])) {
        N[q].qst = 1; N[q].qid = 0; stats[2] += 1;
      }
    for (int q = 0; q < 4; ++q) ins[q] = (N[q].qst == -1) ? N[q].co : -1;
    tps_pop<false>(S, P, true);
    auto gather = [&]() {
      spec_node = P.pred; spec_key = P.predk & 0x7fffffff;
      if (P.pred < 0) return;
      int px, pz;
      ndecode<URG>(P.pred, G.ld, G.inv_ld, px, pz);
      for (int q = 0; q < 4; ++q) spec[q] = tps_neighbour<URG>(G, px, pz, (unsigned)spec_key, q);
    };
    if (when == 0) gather();
    const bool go = tps_apply<URG, false>(S, G, N, true);
    if (when != 0) gather();
    if (!go) break;
  }
}

static int fmm_host_twin_impl(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv, float scx,
                              float scz, int hcap, int hspill_n, float* ttn, int* nsts, float* ttnr, int* nstsr, int* geom,
                              long long* n_accept, int ahead_when, long long* ahead_stats);

extern "C" int dazim_debug_fmm_host_twin(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                                         float scx, float scz, int hcap, int hspill_n, float* ttn, int* nsts, float* ttnr,
                                         int* nstsr, int* geom, long long* n_accept) {
  return fmm_host_twin_impl(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, hcap, hspill_n, ttn, nsts, ttnr, nstsr, geom,
                            n_accept, -1, nullptr);
}
extern "C" int dazim_debug_fmm_host_twin_ahead(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                                               float scx, float scz, int hcap, int hspill_n, int when, float* ttn, int* nsts,
                                               float* ttnr, int* nstsr, int* geom, long long* n_accept, long long* stats) {
  if ((when != 0 && when != 1) || !stats) return DAZIM_EBADARG;
  stats[0] = stats[1] = stats[2] = 0;
  return fmm_host_twin_impl(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, hcap, hspill_n, ttn, nsts, ttnr, nstsr, geom,
                            n_accept, when, stats);
}

static int fmm_host_twin_impl(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv, float scx,
                              float scz, int hcap, int hspill_n, float* ttn, int* nsts, float* ttnr, int* nstsr, int* geom,
                              long long* n_accept, int ahead_when, long long* ahead_stats) {
  if (!pv || nx < 5 || ny < 5 || hcap < 8 || (hcap & 1) || (hspill_n & 1) || hspill_n < 4) return DAZIM_EBADARG;
  const GridC g = make_grid(nx, ny, goxd, gozd, dvxd, dvzd);
  SrcRec sr;
  std::memset(&sr, 0, sizeof(sr));
  int st = make_src(g, scx, scz, sr);
  if (st) return st;
  float ub[41 * 4], cb[6 * 4];
  for (int j = 1; j <= 41; ++j) { float u = 40.0f; u = (float)(j - 1) / u; bspl_basis(u, &ub[(j - 1) * 4]); }
  for (int i = 1; i <= 6; ++i) { float u = 5.0f; u = (float)(i - 1) / u; bspl_basis(u, &cb[(i - 1) * 4]); }
  const size_t nxy = (size_t)nx * ny, nc = (size_t)g.nnx * g.nnz, ncf = coarse_field_size(g.nnx, g.nnz);
  std::vector<float> velv(nxy), slow_c(nc), ric(g.nnx), rir(REF_LD, 0.0f), slow_r(REF_N, 0.0f);
  for (size_t i = 0; i < nxy; ++i) velv[i] = (float)pv[i];
  for (size_t i = 0; i < nc; ++i) slow_c[i] = 1.0f / dice_coarse_node(g, velv.data(), cb, (int)i);
  for (int ix = 1; ix <= g.nnx; ++ix) ric[ix - 1] = g.earth * sin_r(g.gox + (float)(ix - 1) * g.dnx);
  for (int ix = 1; ix <= sr.nnxr; ++ix) rir[ix - 1] = g.earth * sin_r(sr.goxr + (float)(ix - 1) * sr.dnxr);
  std::vector<unsigned> E_c(ncf, E_FAR), E_r(REF_N, E_FAR);
  for (int e = 0; e < sr.nnxr * sr.nnzr; ++e) {
    const int idm1 = e % sr.nnzr + 1, idm2 = e / sr.nnzr + 1;
    slow_r[(size_t)(idm2 - 1) * REF_LD + (idm1 - 1)] = 1.0f / refined_vel_t(g, sr, velv.data(), ub, idm1, idm2);
  }
  std::vector<int2> sm((size_t)hcap), gl((size_t)std::max(1, hspill_n));
  std::vector<int> hpos_r(REF_N, 0);
  TpsState S;
  S.sm = sm.data(); S.stride = 1; S.gl = gl.data(); S.hcap = hcap; S.htot = hcap + hspill_n - 2;
  S.E = nullptr; S.overflow = 0; S.prof = nullptr; S.pt0 = 0;
  unsigned long long nacc = 0;
  tps_source_init(S, g, sr, velv.data(), ub, E_r.data());
  {
    const TpsGrid G = tps_grid_refined(g, sr, slow_r.data(), rir.data(), E_r.data());
    if (ahead_when < 0) { while (tps_step<1>(S, G, nacc)) {} }
    else twin_march_ahead<1>(S, G, nacc, ahead_when, ahead_stats);
  }
  if (!S.overflow) {
    tps_refined_finish(S, E_r.data(), hpos_r.data());
    tps_handoff(S, g, sr, E_r.data(), E_c.data());
    const TpsGrid G = tps_grid_coarse(g, slow_c.data(), ric.data(), E_c.data());
    if (ahead_when < 0) { while (tps_step<2>(S, G, nacc)) {} }
    else twin_march_ahead<2>(S, G, nacc, ahead_when, ahead_stats);
  }
  if (S.overflow) return DAZIM_EHEAP;
  auto decode = [](unsigned e, int hp, float& t, int& stt) {
    if (e == E_FAR) { stt = -1; t = 0.0f; }
    else if ((int)e >= 0) { stt = 0; t = tps_as_float((int)e); }
    else { stt = hp; t = tps_as_float((int)(e & ~E_SIGN)); }
  };
  for (int ix = 0; ix < g.nnx; ++ix)
    for (int iz = 0; iz < g.nnz; ++iz) {
      float t; int stt;
      decode(E_c[cidx(ix, iz, g.nnz)], 1, t, stt);
      if (ttn) ttn[(size_t)ix * g.nnz + iz] = t;
      if (nsts) nsts[(size_t)ix * g.nnz + iz] = stt;
    }
  for (int i = 0; i < REF_N; ++i) {
    float t; int stt;
    decode(E_r[i], hpos_r[i], t, stt);
    if (ttnr) ttnr[i] = t;
    if (nstsr) nstsr[i] = stt;
  }
  if (geom) { geom[0] = sr.nnzr; geom[1] = sr.nnxr; geom[2] = sr.vnl; geom[3] = sr.vnr; geom[4] = sr.vnt; geom[5] = sr.vnb; geom[6] = g.nnz; geom[7] = g.nnx; }
  if (n_accept) *n_accept = (long long)nacc;
  return DAZIM_OK;
}

extern "C" int dazim_raytrace(dazim_handle* h, int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                              const double* pv, int n, const float* scx, const float* scz, const float* rcx,
                              const float* rcz, int azim, float* tt, float* fdm, float* fdmc, float* fdms) {
  if (!h || !pv || n <= 0 || !rcx || !rcz) return DAZIM_EBADARG;
  dazim_problem p;
  std::vector<int> pe, nr, ns;
  mini_problem(nx, ny, goxd, gozd, dvxd, dvzd, n, scx, scz, rcx, rcz, p, pe, nr, ns);
  // receivers: rcxf(nrcf=1, nsrc=n, kmax=1) == rcx[n]
  dazim_tables tb;
  std::memset(&tb, 0, sizeof(tb));
  tb.pvRc = const_cast<double*>(pv);
  std::vector<float> dummyL((size_t)nx * ny, 0.0f);
  std::vector<double> dummyS((size_t)nx * ny * 2, 0.0);
  tb.Lsen_Gsc = dummyL.data();
  tb.sen_vs = tb.sen_vp = tb.sen_rho = dummyS.data();
  std::vector<float> vels((size_t)nx * ny * 2, 3.0f);
  p.vels = vels.data();
  dazim_plan* P = nullptr;
  int st = plan_build(h, azim ? 0 : 1, &p, &tb, nullptr, nullptr, 0, -1, 1, &P);
  if (st) return st;
  st = plan_run(P);
  if (st) { plan_free(P); return st; }
  std::vector<int> off(n), cnt(n);
  cudaError_t e = cudaMemcpy(off.data(), P->d_fp_off.p, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(cnt.data(), P->d_fp_cnt.p, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && tt) e = cudaMemcpy(tt, P->d_dsurf.p, n * 4, cudaMemcpyDeviceToHost);
  unsigned long long used = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&used, P->d_counters.p, 8, cudaMemcpyDeviceToHost);
  std::vector<int> cell(used);
  std::vector<float> v0(used), v1(used), v2(used);
  if (e == cudaSuccess && used) e = cudaMemcpy(cell.data(), P->d_fp_cell.p, used * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && used) e = cudaMemcpy(v0.data(), P->d_fp_fdm.p, used * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && used && azim) e = cudaMemcpy(v1.data(), P->d_fp_fdmc.p, used * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && used && azim) e = cudaMemcpy(v2.data(), P->d_fp_fdms.p, used * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) {
    const int nvx = nx - 2, nvz = ny - 2, ldc = nvx + 2, ldf = nvz + 2;
    const size_t nf = (size_t)(nvz + 2) * (nvx + 2);
    if (fdm) std::memset(fdm, 0, nf * n * 4);
    if (fdmc) std::memset(fdmc, 0, nf * n * 4);
    if (fdms) std::memset(fdms, 0, nf * n * 4);
    for (int r = 0; r < n; ++r)
      for (int i = 0; i < cnt[r]; ++i) {
        const size_t q = (size_t)off[r] + i;
        const int z = cell[q] / ldc, x = cell[q] % ldc;
        const size_t o = (size_t)r * nf + (size_t)x * ldf + z;   // (0:nvz+1,0:nvx+1) column-major
        if (fdm) fdm[o] = v0[q];
        if (fdmc && azim) fdmc[o] = v1[q];
        if (fdms && azim) fdms[o] = v2[q];
      }
  }
  plan_free(P);
  return e == cudaSuccess ? DAZIM_OK : DAZIM_ECUDA + (int)e;
}
