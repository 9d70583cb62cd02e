/* Stand-in for <cuda_runtime.h> so that chicdiff_b200/csrc/common.cuh compiles with g++: the accuracy tests in
 * tests/test_device_math.py run the device math (log_pos, rcp_pos, the shift-10 gamma rationals, dnbinom_mu_log)
 * on the CPU over argument ranges no synthetic data set reaches.  Test tooling; the product never includes it. */
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __constant__ const
#define __restrict__
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
