// aggregate.cu -- stage 1: aggregation of per-fragment counts and expected-background offsets
// into other-end regions, for every sample (chicdiff.R:1540-1547: the by=(baitID, regionID,
// sample) group-by with N = sum(N), FullMean = sum(FullMean)).
//
// Input rows are region-contiguous (sorted by regionID, then otherEndID), regions are CSR
// segments row_off[n+1]; per-sample columns N_rows[s][R] (int32) and FM_rows[s][R] (fp64, NaN =
// NA) are sample-major.  A CTA owns kRegions consecutive regions, i.e. one contiguous run of
// rows.  For each sample it streams that run into shared memory with 16-byte cp.async copies
// (coalesced, double-buffered across samples so the copy of sample s+1 overlaps the sums of
// sample s), then every thread sums its own region's segment from shared memory in row order:
// integer counts are exact (int64 accumulator, overflow -> NA_integer_ like R), FullMean sums
// are sequential in otherEndID order exactly like the reference and propagate NA.
//
// Algorithmic bytes per region: W*(S*12) read + 8 (row_off) + S*12 written (W = rows/region).
#include "kernels.h"
#include <cuda_pipeline.h>

namespace cd {

constexpr int kAggThreads = 256;
constexpr int kAggRegions = 256;       // regions per CTA (one per thread)
constexpr int kAggCap = 3072;          // rows per shared-memory stage (>= 256 * 11)
constexpr int kAggStages = 2;

struct AggStage {
    int32_t nbuf[kAggCap + 4];
    double fbuf[kAggCap + 2];
};

// copy cnt elements starting at src into dst[pad ...] where pad makes the 16-byte groups of
// global memory land on 16-byte groups of shared memory
template <typename T>
__device__ __forceinline__ void stage_copy(T* dst, const T* __restrict__ src, int cnt, int tid)
{
    constexpr int G = 16 / sizeof(T);
    const int pad = (int)(((uintptr_t)src / sizeof(T)) & (G - 1));
    int head = (G - pad) & (G - 1);
    if (head > cnt) head = cnt;
    const int groups = (cnt - head) / G;
    const int tail = cnt - head - groups * G;
    T* d = dst + pad;
    if (tid < head) __pipeline_memcpy_async(d + tid, src + tid, sizeof(T));
    for (int g = tid; g < groups; g += kAggThreads)
        __pipeline_memcpy_async(d + head + g * G, src + head + g * G, 16);
    if (tid < tail) __pipeline_memcpy_async(d + head + groups * G + tid, src + head + groups * G + tid, sizeof(T));
}

template <typename T> __device__ __forceinline__ int stage_pad(const T* src)
{
    constexpr int G = 16 / sizeof(T);
    return (int)(((uintptr_t)src / sizeof(T)) & (G - 1));
}

__global__ void __launch_bounds__(kAggThreads)
aggregate_kernel(int64_t n, int S, const int64_t* __restrict__ row_off, int64_t R,
                 const int32_t* __restrict__ N_rows, const double* __restrict__ FM_rows,
                 int32_t* __restrict__ K, double* __restrict__ FM)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AggStage* stages = reinterpret_cast<AggStage*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * kAggRegions;
    const int64_t i1 = (i0 + kAggRegions < n) ? i0 + kAggRegions : n;
    const int64_t i = i0 + tid;
    const int64_t r0 = row_off[i0], r1 = row_off[i1];
    const int nrows = (int)(r1 - r0);
    int seg_lo = 0, seg_hi = 0;
    if (i < i1) { seg_lo = (int)(row_off[i] - r0); seg_hi = (int)(row_off[i + 1] - r0); }
    const int nchunk = (nrows + kAggCap - 1) / kAggCap;          // 1 unless regions are unusually wide
    const int ntask = S * (nchunk > 0 ? nchunk : 1);

    auto issue = [&](int task) {
        if (task < ntask && nchunk > 0) {
            const int s = task / nchunk, c = task - s * nchunk;
            const int c0 = c * kAggCap;
            const int cnt = (nrows - c0 < kAggCap) ? nrows - c0 : kAggCap;
            AggStage& st = stages[task % kAggStages];
            stage_copy<int32_t>(st.nbuf, N_rows + (int64_t)s * R + r0 + c0, cnt, tid);
            stage_copy<double>(st.fbuf, FM_rows + (int64_t)s * R + r0 + c0, cnt, tid);
        }
        __pipeline_commit();
    };

    issue(0);
    int64_t acc = 0;
    double facc = 0.0;
    for (int task = 0; task < ntask; task++) {
        issue(task + 1);
        __pipeline_wait_prior(1);
        __syncthreads();
        const int s = (nchunk > 0) ? task / nchunk : task;
        const int c = (nchunk > 0) ? task - s * nchunk : 0;
        if (c == 0) { acc = 0; facc = 0.0; }
        if (nchunk > 0) {
            const int c0 = c * kAggCap;
            const int c1 = (c0 + kAggCap < nrows) ? c0 + kAggCap : nrows;
            const AggStage& st = stages[task % kAggStages];
            const int32_t* nb = st.nbuf + stage_pad(N_rows + (int64_t)s * R + r0 + c0) - c0;
            const double* fb = st.fbuf + stage_pad(FM_rows + (int64_t)s * R + r0 + c0) - c0;
            const int lo = seg_lo > c0 ? seg_lo : c0;
            const int hi = seg_hi < c1 ? seg_hi : c1;
            for (int r = lo; r < hi; r++) {
                acc += nb[r];
                facc += fb[r];
            }
        }
        if (c == nchunk - 1 || nchunk == 0) {
            if (i < i1) {
                K[(int64_t)s * n + i] = (acc > 2147483647LL || acc < -2147483647LL) ? INT32_MIN : (int32_t)acc;
                FM[(int64_t)s * n + i] = facc;
            }
        }
        __syncthreads();      // stage is free for the copy issued two tasks ahead
    }
}

cudaError_t launch_aggregate(int64_t n, int S, const int64_t* row_off, int64_t R, const int32_t* N_rows,
                             const double* FM_rows, int32_t* K, double* FM, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(AggStage) * kAggStages;
    // (the attribute is per device: set it on every launch, like the other launchers, so that contexts on several
    //  devices of one process all get it)
    cudaError_t e = cudaFuncSetAttribute(aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (n + kAggRegions - 1) / kAggRegions;
    aggregate_kernel<<<(unsigned)blocks, kAggThreads, smem, st>>>(n, S, row_off, R, N_rows, FM_rows, K, FM);
    return cudaGetLastError();
}

}  // namespace cd
