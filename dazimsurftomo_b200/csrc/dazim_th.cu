// placeholder until K1/K2 land
#include "dazim_dev.h"
namespace dz {
int th_depthkernel(cudaStream_t, int, int, int, const float*, double*, double*, double*, double*, int, const double*, const float*, float, float*, long long*) { return 4; }
int th_depthkernel_ti(cudaStream_t, int, int, int, const float*, double*, int, const double*, const float*, float, float*, float*, long long*) { return 4; }
int th_surfdisp96(cudaStream_t, int, int, const float*, const float*, const float*, const float*, int, const double*, double*) { return 4; }
}
