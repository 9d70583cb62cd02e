#include "common.cuh"
namespace cd { __constant__ CdDesign c_des; }
#include "posterior.cuh"
#include "posterior_v2.cuh"
#include "posterior_v3.cuh"
using namespace cd;
template <int P, int V>
__global__ void k(const double* a, const double* ys, const double* mus, int S, double* out)
{
    extern __shared__ double sh[];
    double* y = sh + threadIdx.x; double* m = sh + 128 * 32 + threadIdx.x;
    for (int j = 0; j < S; j++) { y[j * 128] = ys[j * 128 + threadIdx.x]; m[j * 128] = mus[j * 128 + threadIdx.x]; }
    double lp, dlp;
    if (V == 1) eval_post<P, true>(a[threadIdx.x], y, m, 128, S, 0.0, 1.0, false, lp, dlp);
    else if (V == 2) eval_post_v2<P, true>(a[threadIdx.x], y, m, 128, S, 0.0, 1.0, false, lp, dlp);
    else {
        double* tab = sh + 2 * 128 * 32;                      // 128 x {rc, -log rc}
        for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = kLogTab[i];
        __syncthreads();
        eval_post_v3<P, true>(a[threadIdx.x], y, m, 128, S, 0.0, 1.0, false, tab, lp, dlp);
    }
    out[threadIdx.x] = lp + dlp;
}
template __global__ void k<1, 1>(const double*, const double*, const double*, int, double*);
template __global__ void k<1, 2>(const double*, const double*, const double*, int, double*);
template __global__ void k<2, 1>(const double*, const double*, const double*, int, double*);
template __global__ void k<2, 2>(const double*, const double*, const double*, int, double*);
template __global__ void k<1, 3>(const double*, const double*, const double*, int, double*);
template __global__ void k<2, 3>(const double*, const double*, const double*, int, double*);
