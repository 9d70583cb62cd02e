// universe.cu -- region universe on the device (getRegionUniverse, chicdiff.R:369-426).
//
// Every filtered peak (baitID, oeID) becomes the window .expandAvoidBait(bait, oe, RUexpand)
// (chicdiff.R:353-367): oe-s .. oe+s, stopping two fragments short of the bait when the bait is within
// s+1; rows beyond the last fragment (:402) or on another chromosome than the bait (:417; also IDs below the
// first fragment, which the rmap join turns into NA) are dropped.  regionID = row index of the peak.
// Output is the CSR structure the rest of the path consumes: row_off[m+1], row_bait[R], row_oe[R].
#include "kernels.h"
#include <cub/device/device_scan.cuh>

namespace cd {

__device__ __forceinline__ bool ru_window(int32_t bait, int32_t oe, int s, int64_t& lo, int64_t& hi)
{
    const int64_t d = (int64_t)bait - (int64_t)oe;
    const int64_t ad = d < 0 ? -d : d;
    if (ad > (int64_t)s + 1) { lo = (int64_t)oe - s; hi = (int64_t)oe + s; return true; }
    if (oe > bait) { lo = (int64_t)bait + 2; hi = (int64_t)oe + s; return true; }
    if (oe < bait) { lo = (int64_t)oe - s; hi = (int64_t)bait - 2; return true; }
    return false;                                   // stop("Invalid parameters ...")
}

__device__ __forceinline__ bool ru_keep(int64_t f, int chrb, int64_t F, int32_t id0, const int32_t* __restrict__ chr)
{
    const int64_t k = f - id0;
    return k >= 0 && k < F && chr[k] == chrb;
}

__global__ void __launch_bounds__(256)
ru_count_kernel(int64_t m, const int32_t* __restrict__ peak_bait, const int32_t* __restrict__ peak_oe, int s,
                int64_t F, int32_t id0, const int32_t* __restrict__ chr, int64_t* __restrict__ counts,
                int32_t* __restrict__ status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t bait = peak_bait[i], oe = peak_oe[i];
    int64_t lo, hi, c = 0;
    const int64_t kb = (int64_t)bait - id0;
    if (!ru_window(bait, oe, s, lo, hi)) { atomicOr(status, 1); counts[i] = 0; return; }
    if (kb < 0 || kb >= F) { atomicOr(status, 2); counts[i] = 0; return; }
    const int chrb = chr[kb];
    for (int64_t f = lo; f <= hi; f++) c += ru_keep(f, chrb, F, id0, chr) ? 1 : 0;
    counts[i] = c;
}

__global__ void __launch_bounds__(256)
ru_fill_kernel(int64_t m, const int32_t* __restrict__ peak_bait, const int32_t* __restrict__ peak_oe, int s,
               int64_t F, int32_t id0, const int32_t* __restrict__ chr, const int64_t* __restrict__ row_off,
               int32_t* __restrict__ row_bait, int32_t* __restrict__ row_oe)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t bait = peak_bait[i], oe = peak_oe[i];
    int64_t lo, hi;
    const int64_t kb = (int64_t)bait - id0;
    if (!ru_window(bait, oe, s, lo, hi) || kb < 0 || kb >= F) return;
    const int chrb = chr[kb];
    int64_t w = row_off[i];
    for (int64_t f = lo; f <= hi; f++)
        if (ru_keep(f, chrb, F, id0, chr)) { row_bait[w] = bait; row_oe[w] = (int32_t)f; w++; }
}

cudaError_t ru_launch_count(int64_t m, const int32_t* peak_bait, const int32_t* peak_oe, int s, int64_t F, int32_t id0,
                            const int32_t* chr, int64_t* counts, int32_t* status, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    ru_count_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, peak_bait, peak_oe, s, F, id0, chr, counts, status);
    return cudaGetLastError();
}

// row_off[0..m] = exclusive prefix sum of counts[0..m-1] with the total at [m] (counts has m+1 entries, last = 0)
cudaError_t ru_launch_scan(int64_t m, const int64_t* counts, int64_t* row_off, void* tmp, size_t& tmp_bytes, cudaStream_t st)
{
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, row_off, (int)(m + 1), st);
}

cudaError_t ru_launch_fill(int64_t m, const int32_t* peak_bait, const int32_t* peak_oe, int s, int64_t F, int32_t id0,
                           const int32_t* chr, const int64_t* row_off, int32_t* row_bait, int32_t* row_oe, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    ru_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, peak_bait, peak_oe, s, F, id0, chr, row_off, row_bait, row_oe);
    return cudaGetLastError();
}

}  // namespace cd
