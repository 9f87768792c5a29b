// Device-side shared declarations for the B200 forward-modelling path.
// Parity-critical float32 code: the whole library is compiled with
// --fmad=false (the reference is x86-64/SSE2 without FMA) and default
// IEEE-correct division / sqrt.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dz {

// Coarse propagation grid + B-spline control grid constants (module globalp,
// reference CalSurfG.f90:151-217; values set at FwdTraveltimeCPS.f90:346-380).
struct GridC {
  int nvx, nvz;        // control-grid interior vertices (nx-2, ny-2)
  int nnx, nnz;        // coarse propagation nodes in x (colatitude) and z (longitude)
  int gdx, gdz;        // dicing (5)
  int sgdl, sgs;       // source-grid dicing level (8) and extent (8)
  float gox, goz, dnx, dnz, dvx, dvz, earth;
  float dpl_half;      // 0.5*min cell width, rpathsAzim.f90:156-161 (host computed)
  float dpl_full;      // min cell width, srtimes CalSurfG.f90:1650-1654
};

// Per (period, source) record, filled by the host in float32 exactly like the
// scalar prologue of the reference (FwdTraveltimeCPS.f90:493-530).
struct SrcRec {
  float scx, scz;            // source colatitude / longitude (rad)
  int period;                // 0-based index into the velocity-map tables (periods(srcnum,knumi)-1)
  int knumi;                 // 0-based period-loop index (selects kernel-table slice)
  int vnl, vnr, vnt, vnb;    // coarse bounds of the refined box
  int nnxr, nnzr;            // refined node counts
  float goxr, gozr, dnxr, dnzr;
  int isx_r, isz_r;          // source cell in the refined grid (travel, CalSurfG.f90:282-295)
  float dsx_r, dsz_r;        // source offset inside that cell
  int isx_t, isz_t;          // source cell used by the tracer (rpathsAzim.f90:146-147, no edge clamp)
  int isx_c, isz_c;          // source cell in the coarse grid (srtimes, no clamp)
  int isx_cc, isz_cc;        // clamped coarse source cell (for the near-source velocity)
  float drx_c, drz_c;        // source offset in the clamped coarse cell
  int ray0, nray;            // first global row and number of receivers
};

struct RayRec {
  float rcx, rcz;   // receiver colatitude / longitude (rad)
  int src;          // index into SrcRec (batch-local)
  int row;          // global 0-based row id (count1-1)
};

// Layout of the per-solve COARSE fields (E_c: status + travel time, hpos_c: heap back pointers).
// 8 x-columns are interleaved per z row: offset(ix, iz) = ((ix >> 3) * nnz + iz) * 8 + (ix & 7), 0-based.
// One 32-byte sector then holds the 8 x-neighbours of a node and the z-neighbours sit in the adjacent
// sectors, so the 7 x 7 diamond an accept step gathers comes from 1-2 contiguous 224-byte runs (one or
// two DRAM rows) instead of 7 columns 4 KB apart.  The eikonal stage at full occupancy is bound by the
// rate of random DRAM accesses, which is what this layout attacks.
__host__ __device__ inline size_t coarse_field_size(int nnx, int nnz) { return (size_t)((nnx + 7) >> 3) * (size_t)nnz * 8; }
__host__ __device__ inline int cidx(int ix, int iz, int nnz) { return (((ix >> 3) * nnz + iz) << 3) + (ix & 7); }

static const int REF_LD = 129;           // leading dimension of refined fields: 2*sgs*sgdl+1
static const int REF_N = REF_LD * REF_LD;

struct FmmArgs {
  GridC g;
  const SrcRec* src;          // [nsrc]
  int nsrc;
  const float* velv;          // [nper][(nvz+2)(nvx+2)]
  const float* slow_c;        // [nper][nnx*nnz]  1/veln (fouds2's slown)
  const float* risti_c;       // [nnx]   earth*sin(gox+(ix-1)*dnx), host computed
  const float* risti_r;       // [nsrc][REF_LD]
  // per-solve fields (persist for the ray tracer).  One 32-bit word per node:
  // alive = +t, close = t | sign bit, far = 0xFFFFFFFF (preset by the host)
  unsigned* E_c;              // [nsrc][nnx*nnz]
  unsigned* E_r;              // [nsrc][REF_N]   (ld REF_LD)
  // per-slot workspaces (slot = resident half-warp = 2*blockIdx.x + half)
  int* hpos_c;                // [nslot][nnx*nnz]  heap position of close nodes
  int* hpos_r;                // [nslot][REF_N]
  float* slow_r;              // [nslot][REF_N]    refined slowness
  int2* hspill;               // [nslot][hspill_n] heap entries beyond the shared capacity
  int hspill_n;
  int hcap;                   // heap entries held in shared memory per solve (positions 1..hcap-1)
  int spc;                    // solves per CTA: 2 (one per half-warp; throughput mode) or 1 (half 1 idle: no
                              // cross-solve divergence; used when every solve can be resident anyway)
  int* queue;                 // work counter (solve pairs)
  int* slot_of;               // optional [nsrc]: slot that solved source s (test seam)
  int* flags;                 // bit4 (16): heap overflow
  unsigned long long* n_accept;  // total accepted nodes (statistics)
};

struct TraceArgs {
  GridC g;
  const SrcRec* src;
  const RayRec* ray;        // batch rays in processing order (long rays first)
  int nray;
  const float* veln_c;      // [nper][nnx*nnz]
  const float* ttn_c;       // [nsrc][nnx*nnz]  final coarse field (all alive => plain travel times)
  const float* ttn_r;       // [nsrc][REF_N]    refined field in K3's encoding: alive iff sign bit clear
  // per-thread scratch
  unsigned short* map;      // [nthreads][ncell]
  int* skey;                // [nthreads][cap]
  float* sval;              // [nthreads][3][cap]
  int cap;
  // outputs
  float* dsurf;             // [dall] by global row
  int* fp_off;              // [dall]
  int* fp_cnt;              // [dall]
  int* fp_cell;             // pool: jj*(nvx+2)+kk
  float* fp_fdm; float* fp_fdmc; float* fp_fdms;
  unsigned long long pool_cap;
  unsigned long long* pool_used;
  int* counter;             // work counter
  int* flags;               // bit0 rbint, bit1 receiver outside, bit2 slot overflow, bit3 pool overflow
  unsigned long long* n_steps;
  int emit_all;             // test seam: emit every touched control point (no interior/threshold filter)
};

struct AsmArgs {
  int mode;                 // 1 iso (1 block), 2 joint (3 blocks)
  int nx, ny, nz, nvx, nvz, kmax;
  int row0, nrow;           // rows [row0, row0+nrow) handled by this launch
  const int* row_knumi;     // [dall] period-loop index of each row (0-based)
  const int* fp_off; const int* fp_cnt;
  const int* fp_cell; const float* fp_fdm; const float* fp_fdmc; const float* fp_fdms;
  const double* sen_vs; const double* sen_vp; const double* sen_rho;   // (nx*ny,kmax,nz)
  const float* lsen;        // (nx*ny,kmax,nz-1)
  const float* coe_a; const float* coe_rho;   // (nx*ny, nz-1) at node index jj*nx+kk (vels(kk+1,jj+1,k))
  long long* nnz_row;       // [nrow] counts (pass 1); 64-bit so that the row-pointer scan cannot wrap at 2^31 non-zeros
  const long long* rowptr;  // [nrow+1] (pass 2)
  float* val; int* col;     // CSR outputs (pass 2)
  int* rowid;               // optional COO row ids, 1-based (pass 2)
};

// forward mode: obsTaa(row) = sum_n GGc(row,n)*Gc(n) + sum_n GGs(row,n)*Gs(n), ascending n, float32
struct TaaArgs {
  int nx, ny, nz, nvx, nvz, kmax, nrow, row0;
  const int* row_knumi; const int* fp_off; const int* fp_cnt; const int* fp_cell;
  const float* fp_fdmc; const float* fp_fdms; const float* lsen;
  const float* gc; const float* gs;   // (nx-2,ny-2,nz-1)
  float* taa;
};

}  // namespace dz
