#include "common.cuh"
#include "log_v2.cuh"
using namespace cd;
__global__ void k_log_v1(const double* x, double* out) { out[threadIdx.x] = log_pos(x[threadIdx.x]); }
__global__ void k_log_v2(const double* x, double* out)
{
    __shared__ double tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = kLogTab[i];
    __syncthreads();
    out[threadIdx.x] = log_pos_v2(x[threadIdx.x], tab);
}
