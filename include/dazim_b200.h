/* dazim_b200.h -- C ABI of the B200-native forward-modelling path of DAzimSurfTomo.
 *
 * The reference (Chuanming-Liu/DAzimSurfTomo) is Fortran and has no FFI; the
 * seam this library replaces is the Fortran external-procedure boundary
 * between the drivers and the forward-modelling subroutines:
 *
 *   FwdObsTraveltimeCPS   src/src_forward/FwdTraveltimeCPS.f90:208   (called MainForward.f90:372)
 *   CalSurfG              src/src_inv_iso_joint/CalSurfG.f90:909     (called Main_Jt.f90:398)
 *   CalSurfGAnisoJoint    src/src_inv_iso_joint/CalSurfGAniso_Joint.f90:209 (called Main_Jt.f90:403)
 *   depthkernel           src/src_inv_iso_joint/CalSurfG.f90:1
 *   depthkernelTI         src/src_forward/depthkernelTI.f90:2
 *
 * All arrays are Fortran column-major with the reference's shapes; all
 * pointers are HOST pointers unless a function says otherwise; the caller owns
 * every buffer.  Functions return 0 on success or a DAZIM_E* code that mirrors
 * the reference's STOP sites.  There is no CPU fallback: every entry point
 * fails with DAZIM_ECUDA when no CUDA device is usable.
 */
#ifndef DAZIM_B200_H
#define DAZIM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
  DAZIM_OK = 0,
  DAZIM_ESOURCE_OUTSIDE = 1,   /* "Source lies outside bounds of model", CalSurfG.f90:287-293 */
  DAZIM_ERECEIVER_OUTSIDE = 2, /* "Receiver lies outside model", CalSurfG.f90:1649-1655, rpathsAzim.f90:193-216 */
  DAZIM_ENNZ_OVERFLOW = 3,     /* nar > maxnar, "increase sparsity fraction", Main_Jt.f90:523 */
  DAZIM_EBADARG = 4,
  DAZIM_ELAYERS = 5,           /* more than NL=200 layers / NP=60 periods, surfdisp96.f:57-59 */
  DAZIM_EHEAP = 6,             /* narrow band larger than the heap workspace */
  DAZIM_EFOOTPRINT = 7,        /* a ray touched more control points than the footprint workspace holds */
  DAZIM_ENOROOT = 9,           /* surfdisp96 found no root (reference prints a warning and zero-fills) */
  DAZIM_ENCCL = 10,            /* NCCL not loadable / a collective failed (row-distributed solver) */
  DAZIM_ECUDA = 100            /* + cudaError_t */
};

typedef struct dazim_handle dazim_handle;

/* Station / geometry tables exactly as the Fortran drivers hold them
 * (MainForward.f90:239-281).  Leading dimensions nsrc, nrcf are the callers'. */
typedef struct dazim_problem {
  int nx, ny, nz;            /* model grid incl. edge nodes */
  const float* vels;         /* (nx,ny,nz) Vs km/s */
  float goxd, gozd;          /* upper-left (lat, lon) degrees */
  float dvxd, dvzd;          /* grid spacing degrees */
  int kmaxRc;                /* number of periods */
  const double* tRc;         /* (kmaxRc) periods, s */
  const float* depz;         /* (nz) depths km */
  float minthk;              /* "sublayers" of para.in */
  int kmax, nsrc, nrcf;
  const int* periods;        /* (nsrc,kmax) */
  const int* nrc1;           /* (nsrc,kmax) */
  const int* nsrcsurf1;      /* (kmax) */
  const float* scxf;         /* (nsrc,kmax) colatitude rad */
  const float* sczf;         /* (nsrc,kmax) longitude rad */
  const float* rcxf;         /* (nrcf,nsrc,kmax) */
  const float* rczf;         /* (nrcf,nsrc,kmax) */
} dazim_problem;

/* Depth-kernel tables (outputs of depthkernel / depthkernelTI). */
typedef struct dazim_tables {
  double* pvRc;              /* (nx*ny,kmaxRc) phase velocity */
  double* sen_vs;            /* (nx*ny,kmaxRc,nz) dc/dVs ; may be NULL in forward mode */
  double* sen_vp;
  double* sen_rho;
  float* Lsen_Gsc;           /* (nx*ny,kmaxRc,nz-1) ; may be NULL in iso mode */
} dazim_tables;

/* Sparse G in the reference's COO convention (rw(k), row iw(k+1), col(k)),
 * rows ascending, columns ascending inside a row; all indices 1-based. */
typedef struct dazim_coo {
  float* rw;                 /* (maxnar) */
  int* iw_row;               /* (maxnar) row of entry k; the Fortran shim passes iw+1 */
  int* col;                  /* (maxnar) */
  long long maxnar;
  long long nar;             /* out */
} dazim_coo;

/* Per-stage device times of the last call (CUDA events on the library stream), ms. */
typedef struct dazim_times {
  float kernels_ms;          /* K1/K2 Thomson-Haskell */
  float dice_ms;             /* K0 */
  float fmm_ms;              /* K3 */
  float trace_ms;            /* K4 */
  float assemble_ms;         /* K5 */
  float total_ms;            /* first launch to last kernel end, excluding H2D/D2H */
  long long n_accept;        /* accepted FMM nodes */
  long long n_steps;         /* ray steps */
  long long n_fmm_launch, n_trace_launch, n_launch; /* kernel launches */
  long long h2d_bytes, d2h_bytes;
  int rbint;                 /* a ray hugged the model boundary (reference warning) */
} dazim_times;

int dazim_create(dazim_handle** h, int device);
/* page-locked host buffers for callers that want full-speed PCIe copies of the (multi-GB) COO
 * outputs; plain malloc'ed / Fortran arrays work too, only slower */
int dazim_host_alloc(void** p, unsigned long long bytes);
void dazim_host_free(void* p);
void dazim_destroy(dazim_handle* h);
const char* dazim_strerror(int code);
const dazim_times* dazim_last_times(const dazim_handle* h);

/* --- L2: depth kernels ---------------------------------------------------- */
/* depthkernel, CalSurfG.f90:1-139 */
int dazim_depthkernel(dazim_handle* h, int nx, int ny, int nz, const float* vel, double* pvRc,
                      double* sen_vs, double* sen_vp, double* sen_rho, int kmaxRc, const double* tRc,
                      const float* depz, float minthk);
/* depthkernelTI, depthkernelTI.f90:2-112 */
int dazim_depthkernel_ti(dazim_handle* h, int nx, int ny, int nz, const float* vel, double* pvRc,
                         int kmaxRc, const double* tRc, const float* depz, float minthk,
                         float* Lsen_Gsc);
/* surfdisp96 (surfdisp96.f:52), Rayleigh phase branch, nprof layered profiles at once:
 * thk/vp/vs/rho (nlayer,nprof) column-major, cg (kmax,nprof) */
int dazim_surfdisp96(dazim_handle* h, int nprof, int nlayer, const float* thk, const float* vp,
                     const float* vs, const float* rho, int kmax, const double* t, double* cg);

/* --- L1: orchestrators ---------------------------------------------------- */
/* mode 0: FwdObsTraveltimeCPS  -> dsurf, obsTaa            (needs pvRc, Lsen_Gsc; Gc/Gs)
 * mode 1: CalSurfG             -> dsurf, COO (nparpi cols) (needs pvRc, sen_*)
 * mode 2: CalSurfGAnisoJoint   -> dsurf, COO (3*nparpi)    (needs all tables)
 * tables_precomputed=0: the tables are computed on the GPU first (and returned
 * in *tables where non-NULL); =1: *tables are inputs and only the
 * dice + eikonal + ray + assembly stages run. */
int dazim_gbuild(dazim_handle* h, int mode, const dazim_problem* p, dazim_tables* tables,
                 int tables_precomputed, const float* Gctrue, const float* Gstrue, float* dsurf,
                 float* obsTaa, double* tRcV, dazim_coo* coo);

/* Multi-GPU variant of dazim_gbuild for a host that makes ONE subroutine call per outer iteration (the reference's
 * drivers: Main_Jt.f90:398-406, MainForward.f90:372-375).  Single process; the ndev devices listed in devices[] each
 * get one host thread, one stream, a strip of grid rows for the depth kernels and a contiguous (period, source) range
 * balanced by ray count; every device copies its row block straight into the caller's arrays at its offset (rows,
 * columns and values come out in the reference's order, bit-identical to the single-device call for any ndev).
 * times_max (may be NULL) receives the per-stage maximum over the devices and the summed counters.  The gfortran
 * drop-in symbols use it when the environment variable DAZIM_DEVICES lists more than one device ("0,1,2,3" or "all"). */
int dazim_gbuild_multi(int ndev, const int* devices, int mode, const dazim_problem* p, dazim_tables* tables,
                       int tables_precomputed, const float* Gctrue, const float* Gstrue, float* dsurf, float* obsTaa,
                       double* tRcV, dazim_coo* coo, dazim_times* times_max);

/* --- test seams ------------------------------------------------------------ */
/* Eikonal solves for n sources on one phase-velocity map pv (nx*ny doubles).
 * Outputs per source: ttn/nsts coarse (nnz,nnx), ttnr/nstsr (129,129), geom[8]
 * = nnzr,nnxr,vnl,vnr,vnt,vnb,nnz,nnx.  Any output may be NULL. */
int dazim_fmm_solve(dazim_handle* h, int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                    const double* pv, int n, const float* scx, const float* scz, float* veln,
                    float* ttn, int* nsts, float* ttnr, int* nstsr, int* geom);

/* Host-only test seam: the eikonal code of the thread-per-solve kernel (csrc/dazim_tps.h, __host__ __device__) run on
 * the CPU for ONE source, with a shared heap part of hcap entries and hspill_n spilled entries.  Same outputs as
 * dazim_fmm_solve for n = 1.  Lets the CPU-only test suite compare the kernel's logic with the oracle; never called by
 * a product entry point and needs no device. */
int dazim_debug_fmm_host_twin(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                              float scx, float scz, int hcap, int hspill_n, float* ttn, int* nsts, float* ttnr,
                              int* nstsr, int* geom, long long* n_accept);
/* The same with the cohort kernel's "neighbour records computed ahead" protocol replayed serially: the records of the
 * predicted next node are gathered before (when = 0) or after (when = 1) the current updates are applied, used if the
 * prediction holds and patched as the kernel's heap lane patches them.  stats[3] = rounds predicted, not predicted,
 * records patched.  Host-only test seam. */
int dazim_debug_fmm_host_twin_ahead(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, const double* pv,
                                    float scx, float scz, int hcap, int hspill_n, int when, float* ttn, int* nsts,
                                    float* ttnr, int* nstsr, int* geom, long long* n_accept, long long* stats);
/* Rays: n (source, receiver) pairs on one map; outputs travel time tt(n) and dense
 * Frechet maps fdm/fdmc/fdms (nvz+2, nvx+2, n) column-major (zero where untouched). */
int dazim_raytrace(dazim_handle* h, int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                   const double* pv, int n, const float* scx, const float* scz, const float* rcx,
                   const float* rcz, int azim, float* tt, float* fdm, float* fdmc, float* fdms);

/* --- device-resident plan (inputs uploaded once; used by bench.py for the
 * HBM-resident figure and by the multi-GPU driver) ---------------------------- */
typedef struct dazim_plan dazim_plan;
/* src_begin/src_end: half-open range of (period,source) units in loop order
 * owned by this rank (0, -1 = all). */
int dazim_plan_create(dazim_handle* h, int mode, const dazim_problem* p, const dazim_tables* tables,
                      const float* Gctrue, const float* Gstrue, long long src_begin,
                      long long src_end, dazim_plan** plan);
int dazim_plan_run(dazim_plan* plan);                 /* all kernels, results stay in HBM */
long long dazim_plan_rows(const dazim_plan* plan, long long* row0);   /* rows owned, first row */
long long dazim_plan_nnz(const dazim_plan* plan);
/* copy results of the last run to host (any pointer may be NULL);
 * rowptr has rows+1 entries (0-based offsets into val/col), col is 1-based */
int dazim_plan_fetch(dazim_plan* plan, float* dsurf, float* obsTaa, long long* rowptr, int* col,
                     float* val);
/* device pointers of the last run's outputs (for NCCL gathers without a host hop) */
/* Name of the eikonal kernel the plan launches ("k_fmm_coh<8>", "k_fmm_duo", "k_fmm<2>", ...): chosen from the
 * number of solves and the grid size. */
const char* dazim_plan_eikonal_kernel(const dazim_plan* plan);
int dazim_plan_device_ptrs(dazim_plan* plan, void** dsurf, void** obsTaa, void** rowptr, void** col,
                           void** val);
void dazim_plan_destroy(dazim_plan* plan);

/* --- next stage: the sparse least-squares solve that consumes G ------------------
 * LSMR (src/src_inv_iso_joint/lsmrModule.f90:36) with the reference's COO products (aprod.f90:7), single
 * precision like the reference (lsmrDataModule.f90:21).  A is m x n, given as nnz triplets
 * (iw_row[k], col[k], rw[k]) with 1-based indices in any order -- exactly iw(2:nar+1), iw(nar+2:2nar+1), rw of the
 * reference (Main_Jt.f90:529-532).  x (n) is the output.  Stopping rules, istop codes and the norm / condition
 * estimates are the reference's. */
typedef struct dazim_lsmr_info {
  int istop, itn;
  float normA, condA, normr, normAr, normx;
  float setup_ms;            /* upload + CSR/CSC construction (device time) */
  float solve_ms;            /* the iterations (device time) */
} dazim_lsmr_info;
int dazim_lsmr(dazim_handle* h, int m, int n, long long nnz, const int* iw_row, const int* col, const float* rw,
               const float* b, float damp, float atol, float btol, float conlim, int itnlim, int localSize,
               float* x, dazim_lsmr_info* info);

/* the same solve on the G row block of a plan's last run: the matrix never leaves the GPU (b has plan rows entries,
 * x has nparpi (iso) or 3*nparpi (joint) entries) */
int dazim_plan_lsmr(dazim_plan* plan, const float* b, float damp, float atol, float btol, float conlim, int itnlim,
                    int localSize, float* x, dazim_lsmr_info* info);

/* --- row-distributed LSMR (SURVEY 8f-1 "keeping G resident and row-distributed", 8e: "if the solver is later
 * distributed by rows the gather of G disappears and only an n-vector all-reduce per LSMR iteration remains") --------
 * One process per GPU.  Every rank holds a block of ROWS of A (its rays' rows of G, any share of the regularisation
 * rows) and the matching entries of b; x and every n-vector of lsmrModule.f90 are replicated.  Per iteration the ranks
 * exchange one n-vector (A^T u: ncclAllReduce, float32 sum) and one scalar (||u||^2, float64); both sit in stream order
 * between the solver's kernels and are captured into the per-iteration CUDA graph.  All ranks return the same x / info
 * bit for bit; against the single-GPU solve the sums associate differently (tolerance, not bit, parity).
 * The communicator: rank 0 calls dazim_comm_unique_id and hands the DAZIM_COMM_ID_BYTES to every rank by any means
 * (torch.distributed broadcast, MPI_Bcast, a file); every rank then calls dazim_comm_create with its device. */
#define DAZIM_COMM_ID_BYTES 128
typedef struct dazim_comm dazim_comm;
int dazim_comm_unique_id(unsigned char* id /* [DAZIM_COMM_ID_BYTES] */);
int dazim_comm_create(int device, const unsigned char* id, int rank, int nranks, dazim_comm** out);
void dazim_comm_destroy(dazim_comm* comm);
int dazim_comm_rank(const dazim_comm* comm);
int dazim_comm_size(const dazim_comm* comm);
/* dazim_lsmr on a row block: m_local rows with LOCAL 1-based row ids in iw_row, b_local (m_local); m_total = rows of
 * the whole system (sizes the reorthogonalisation window like the reference: min(localSize, m, n)).  Collective: every
 * rank of the communicator must call it with the same n and controls, and every rank needs at least one row
 * (m_local >= 1).  A rank that fails its argument checks returns early while the others wait in the first collective:
 * validate on the host before the call, as with any collective. */
int dazim_lsmr_rows(dazim_handle* h, dazim_comm* comm, int m_local, long long m_total, int n, long long nnz_local,
                    const int* iw_row, const int* col, const float* rw, const float* b_local, float damp, float atol,
                    float btol, float conlim, int itnlim, int localSize, float* x, dazim_lsmr_info* info);
/* the same on the G row block a plan built for its share of the sources (resident in HBM, never gathered) */
int dazim_plan_lsmr_rows(dazim_plan* plan, dazim_comm* comm, long long m_total, const float* b_local, float damp,
                         float atol, float btol, float conlim, int itnlim, int localSize, float* x,
                         dazim_lsmr_info* info);

/* --- next stage: the rest of one outer iteration of Main_Jt.f90 ----------------------------
 * Everything the reference does between two G builds (Main_Jt.f90:416-727), on the device-resident G of a plan:
 * residual and its statistics, CalDdatSigma, data weighting of b and of the rows of G, DWS sums, Tikhonov rows
 * appended behind G, LSMR, model update with clamps, model norms and the residual of the solution computed from the
 * sparse rows (no dense GVs/GGc/GGs).  G never leaves the GPU. */
typedef struct dazim_iter_params {
  int iso_inv;               /* 1: CalSurfG system (plan mode 1), TikhonovRegularization; 0: joint (plan mode 2), TikhRegul_joint */
  float weightVs, weightGcs; /* para.in smoothing weights (Main_Jt.f90:174-175, :508-510) */
  float damp;                /* LSMR damping (Main_Jt.f90:176) */
  float minvel, maxvel;      /* Vs clamps (Main_Jt.f90:168, :593-594) */
  int use_ref_controls;      /* 1: atol/btol/conlim/itnlim/localSize of Main_Jt.f90:542-554; 0: the five fields below */
  float atol, btol, conlim;
  int itnlim, localSize;
} dazim_iter_params;

typedef struct dazim_iter_stats {
  float before[4];           /* abs mean, std, RMS, mean of obst - dsyn        (Main_Jt.f90:432-437) */
  float after[4];            /* the same of the residual after the solve       (Main_Jt.f90:720-725) */
  float meandeltaT;          /* CalDdatSigma's mean |dT/T| */
  float mean_weight;         /* sum(datweight)/dall */
  float meanabs_weighted;    /* sum(|cbst|)/dall after weighting */
  float norms[6];            /* VsNorm2, VswNorm2, GcsNorm2, GcswNorm2, Mnorm2, MwNorm2 (Calmodel2Norm / ...Joint) */
  float res2Nm, resW2Nm;     /* ||Gm-d||, ||W(Gm-d)||                         (CalVsReslNorm / CalReslNormJoint) */
  float meanabs_Taa, meanabs_Tvs;
  long long nar1, nar;       /* non-zeros before / after the regularisation rows */
  int count3;                /* regularisation rows */
  dazim_lsmr_info lsmr;
  float step_ms;             /* device time of the whole step */
  float scale_ms;            /* of which: the row scaling of G by the data weights (8 B per non-zero) */
} dazim_iter_stats;

/* New model for an existing plan (same geometry): re-uploads vels and the depth-kernel tables; the work list,
 * workspaces and ray order are kept.  Call between outer iterations instead of destroying the plan. */
int dazim_plan_update_model(dazim_plan* plan, const float* vels, const dazim_tables* tables);

/* One outer-iteration tail on the G of the plan's last run (the plan must own every row: src range 0,-1).
 * obst (rows) observed travel times in row order.  vsf (nx,ny,nz) is updated in place; dv (nparpi or 3*nparpi)
 * receives the clipped LSMR solution; gcf/gsf ((nx-2),(ny-2),(nz-1); joint only), dws (nparpi; iso only) and the
 * per-row arrays sigmaT, resbst, fwdTvs, fwdTaa may be NULL.  Afterwards the plan's val[] holds the WEIGHTED
 * rows until the next dazim_plan_run. */
int dazim_plan_iterate(dazim_plan* plan, const float* obst, const dazim_iter_params* prm, float* vsf, float* dv,
                       float* gcf, float* gsf, float* dws, float* sigmaT, float* resbst, float* fwdTvs,
                       float* fwdTaa, dazim_iter_stats* stats);

/* The same with the rows of G left on the ranks that built them (each rank's plan covers its share of the sources;
 * no gather of G): obst and the per-row outputs have dall_total entries (the whole system, loop order), the
 * regularisation rows live on the last rank, LSMR runs row-distributed (dazim_lsmr_rows).  Everything that is O(rows)
 * is completed on every rank and computed in the single-GPU order, so statistics and weights are bit-identical to
 * dazim_plan_iterate; the solution differs by the association of the LSMR sums (1e-7 relative).  Collective. */
int dazim_plan_iterate_rows(dazim_plan* plan, dazim_comm* comm, long long dall_total, const float* obst,
                            const dazim_iter_params* prm, float* vsf, float* dv, float* gcf, float* gsf, float* dws,
                            float* sigmaT, float* resbst, float* fwdTvs, float* fwdTaa, dazim_iter_stats* stats);

/* The same on a system the caller holds in HBM: the CSR row blocks of several ranks after the NCCL all-gather
 * (all rows; d_rowid = 1-based global row id of every entry).  d_* are DEVICE pointers on the handle's device and
 * must be complete before the call; d_val / d_col / d_rowid need cap >= nnz + (1 or 3) x dazim_tikh_block_entries and
 * are modified (rows weighted, regularisation rows appended).  Host arrays as in dazim_plan_iterate. */
int dazim_iterate_device(dazim_handle* h, int nx, int ny, int nz, long long nrow, long long nnz, long long cap,
                         const long long* d_rowptr, int* d_col, float* d_val, int* d_rowid, const float* d_dsurf,
                         const float* obst, const dazim_iter_params* prm, float* vsf, float* dv, float* gcf, float* gsf,
                         float* dws, float* sigmaT, float* resbst, float* fwdTvs, float* fwdTaa, dazim_iter_stats* stats);

/* The two small subroutines on their own (host arrays), for callers that keep the stock Fortran loop:
 * CalDdatSigma (CalSigamNorm.f90:2) and TikhonovRegularization / TikhRegul_joint (TikhRegul.f90:2 / :108).
 * dazim_tikhonov appends to rw / iw_row (= iw+1) / col behind *nar entries and advances *nar; joint != 0 selects
 * TikhRegul_joint (narVs is then set, iso_inv ignored). */
int dazim_cal_ddat_sigma(dazim_handle* h, int dall, const float* obst, const float* cbst, float* sigmaT,
                         float* meandeltaT);
int dazim_tikhonov(dazim_handle* h, int joint, int nx, int ny, int nz, int maxvp, int dall, long long* nar, float* rw,
                   int* iw_row, int* col, long long* narVs, int* count3, int iso_inv, float weightGcs, float weightVs);
/* test seam (no GPU needed): offset of cell (i,j,k)'s row inside one Tikhonov block and the block's entry count */
long long dazim_tikh_offset(int i, int j, int k, int nvx, int nvz, int nzm1);
long long dazim_tikh_block_entries(int nvx, int nvz, int nzm1);

/* --- gfortran-ABI drop-in symbols (lower case + underscore, all by reference) --- */
void fwdobstraveltimecps_(int* nx, int* ny, int* nz, int* nparpi, float* vels, float* Gctrue,
                          float* Gstrue, float* dsurf, float* obsTaa, int* dall, int* rmax,
                          double* tRcV, float* Lsen_Gsc, float* goxdf, float* gozdf, float* dvxdf,
                          float* dvzdf, int* kmaxRc, double* tRc, int* periods, float* depz,
                          float* minthk, float* scxf, float* sczf, float* rcxf, float* rczf,
                          int* nrc1, int* nsrcsurf1, int* kmax, int* nsrcsurf, int* nrcf,
                          int* writepath);
void calsurfg_(int* nx, int* ny, int* nz, int* nparpi, float* vels, int* iw, float* rw, int* col,
               float* dsurf, float* GVs, int* dall, float* goxdf, float* gozdf, float* dvxdf,
               float* dvzdf, int* kmaxRc, double* tRc, int* periods, float* depz, float* minthk,
               float* scxf, float* sczf, float* rcxf, float* rczf, int* nrc1, int* nsrcsurf1,
               int* kmax, int* nsrcsurf, int* nrcf, int* nar);
void calsurfganisojoint_(int* nx, int* ny, int* nz, int* nparpi, float* vels, int* iw, float* rw,
                         int* col, float* dsurf, float* GVs, float* GGc, float* GGs,
                         float* Lsen_Gsc, int* dall, int* rmax, double* tRcV, float* goxdf,
                         float* gozdf, float* dvxdf, float* dvzdf, int* kmaxRc, double* tRc,
                         int* periods, float* depz, float* minthk, float* scxf, float* sczf,
                         float* rcxf, float* rczf, int* nrc1, int* nsrcsurf1, int* kmax,
                         int* nsrcsurf, int* nrcf, int* nar, int* writepath);
void depthkernel_(int* nx, int* ny, int* nz, float* vel, double* pvRc, double* sen_vsRc,
                  double* sen_vpRc, double* sen_rhoRc, int* iwave, int* igr, int* kmaxRc,
                  double* tRc, float* depz, float* minthk);
void depthkernelti_(int* nx, int* ny, int* nz, float* vel, double* pvRc, int* iwave, int* igr,
                    int* kmaxRc, double* tRc, float* depz, float* minthk, float* Lsen_Gsc);
/* module procedure LSMRmodule::LSMR as gfortran mangles it (lsmrModule.f90:36; called at Main_Jt.f90:562) */
void __lsmrmodule_MOD_lsmr(int* m, int* n, int* leniw, int* lenrw, int* iw, float* rw, float* b, float* damp,
                           float* atol, float* btol, float* conlim, int* itnlim, int* localSize, int* nout,
                           float* x, int* istop, int* itn, float* normA, float* condA, float* normr,
                           float* normAr, float* normx);

/* CalDdatSigma (CalSigamNorm.f90:2; called at Main_Jt.f90:460), TikhonovRegularization (TikhRegul.f90:2; Main_Jt.f90:513)
 * and TikhRegul_joint (TikhRegul.f90:108; Main_Jt.f90:515) -- external (non-module) subroutines */
void calddatsigma_(int* dall, float* obst, float* cbst, float* sigmaT, float* meandeltaT);
void tikhonovregularization_(int* nx, int* ny, int* nz, int* maxvp, int* dall, int* nar, float* rw, int* iw, int* col,
                             int* count3, int* iso_inv, float* weightGcs, float* weightVs);
void tikhregul_joint_(int* nx, int* ny, int* nz, int* maxvp, int* dall, int* nar, float* rw, int* iw, int* col,
                      int* narVs, int* count3, float* weightGcs, float* weightVs);

#ifdef __cplusplus
}
#endif
#endif
