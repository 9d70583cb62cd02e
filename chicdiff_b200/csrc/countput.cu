// countput.cu -- the per-condition (baitID, otherEndID) summary that getFullRegionData writes to
// <outprefix>_countput.Rds and plotDiffBaits reads (chicdiff.R:708-735, 755-770; :1139-1160, 1176-1188):
// over all CHiCAGO rows (with a distance) of the replicates of one condition,
//     Nav = mean(N), Bav = mean(Bmean), score = max(score), oeID_mid = (start + end) / 2
// grouped by (baitID, otherEndID), the means running over the replicates in which the pair occurs, groups in
// order of first appearance in the row-bound table (data.table's `by` order).
// The replicate tables are concatenated, radix-sorted by the 64-bit (bait, oe) key with the global row index
// as payload (stable, so equal keys stay in replicate order), reduced by key, and the groups are put back
// into first-appearance order by sorting on the smallest row index of each group.
#include "kernels.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace cd {

__global__ void __launch_bounds__(256)
cp_keys_kernel(int64_t rows, int64_t base, const int32_t* __restrict__ bait, const int32_t* __restrict__ oe,
               unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    keys[base + i] = ((unsigned long long)(unsigned int)bait[i] << 32) | (unsigned long long)(unsigned int)oe[i];
    idx[base + i] = (unsigned int)(base + i);
}

__global__ void __launch_bounds__(256)
cp_heads_kernel(int64_t T, const unsigned long long* __restrict__ keys, int64_t* __restrict__ head)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per sorted row that starts a group
__global__ void __launch_bounds__(256)
cp_reduce_kernel(int64_t T, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idx,
                 const int64_t* __restrict__ head, const int64_t* __restrict__ slot,
                 const int32_t* __restrict__ N, const double* __restrict__ Bmean, const double* __restrict__ score,
                 unsigned long long* __restrict__ g_key, double* __restrict__ g_nav, double* __restrict__ g_bav,
                 double* __restrict__ g_score, unsigned int* __restrict__ g_first)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T || !head[i]) return;
    const unsigned long long k = keys[i];
    double sn = 0.0, sb = 0.0, mx = -INFINITY;
    bool mx_na = false;
    unsigned int first = 0xffffffffu;
    int cnt = 0;
    for (int64_t j = i; j < T && keys[j] == k; j++) {
        const unsigned int r = idx[j];
        sn += (double)N[r];
        sb += Bmean[r];
        const double sc = score[r];
        if (isnan(sc)) mx_na = true; else if (sc > mx) mx = sc;
        if (r < first) first = r;
        cnt++;
    }
    const int64_t g = slot[i];
    g_key[g] = k;
    g_nav[g] = sn / cnt;
    g_bav[g] = sb / cnt;
    g_score[g] = mx_na ? NAN : mx;
    g_first[g] = first;
}

__global__ void __launch_bounds__(256)
cp_gather_kernel(int64_t G, const unsigned int* __restrict__ order, const unsigned long long* __restrict__ g_key,
                 const double* __restrict__ g_nav, const double* __restrict__ g_bav, const double* __restrict__ g_score,
                 int64_t F, int32_t id0, const int32_t* __restrict__ frag_start, const int32_t* __restrict__ frag_end,
                 int32_t* __restrict__ o_bait, int32_t* __restrict__ o_oe, double* __restrict__ o_nav,
                 double* __restrict__ o_bav, double* __restrict__ o_score, double* __restrict__ o_mid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    const unsigned int g = order[i];
    const unsigned long long k = g_key[g];
    const int32_t oe = (int32_t)(unsigned int)(k & 0xffffffffull);
    o_bait[i] = (int32_t)(unsigned int)(k >> 32);
    o_oe[i] = oe;
    o_nav[i] = g_nav[g]; o_bav[i] = g_bav[g]; o_score[i] = g_score[g];
    const int64_t f = (int64_t)oe - id0;
    o_mid[i] = (f >= 0 && f < F) ? ((double)frag_start[f] + (double)frag_end[f]) / 2.0 : NAN;
}

__global__ void cp_iota_kernel(int64_t G, unsigned int* __restrict__ v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G) v[i] = (unsigned int)i;
}

static inline unsigned nb(int64_t n) { return (unsigned)((n + 255) / 256); }

cudaError_t cp_launch_keys(int64_t rows, int64_t base, const int32_t* bait, const int32_t* oe, unsigned long long* keys,
                           unsigned int* idx, cudaStream_t st)
{
    if (rows > 0) cp_keys_kernel<<<nb(rows), 256, 0, st>>>(rows, base, bait, oe, keys, idx);
    return cudaGetLastError();
}

cudaError_t cp_sort_pairs_u64(void* tmp, size_t& bytes, const unsigned long long* kin, unsigned long long* kout,
                              const unsigned int* vin, unsigned int* vout, int64_t n, cudaStream_t st)
{
    return cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (int)n, 0, 64, st);
}

cudaError_t cp_sort_pairs_u32(void* tmp, size_t& bytes, const unsigned int* kin, unsigned int* kout,
                              const unsigned int* vin, unsigned int* vout, int64_t n, cudaStream_t st)
{
    return cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (int)n, 0, 32, st);
}

cudaError_t cp_scan_i64(void* tmp, size_t& bytes, const int64_t* in, int64_t* out, int64_t n, cudaStream_t st)
{
    return cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, st);
}

cudaError_t cp_launch_heads(int64_t T, const unsigned long long* keys, int64_t* head, cudaStream_t st)
{
    if (T > 0) cp_heads_kernel<<<nb(T), 256, 0, st>>>(T, keys, head);
    return cudaGetLastError();
}

cudaError_t cp_launch_reduce(int64_t T, const unsigned long long* keys, const unsigned int* idx, const int64_t* head,
                             const int64_t* slot, const int32_t* N, const double* Bmean, const double* score,
                             unsigned long long* g_key, double* g_nav, double* g_bav, double* g_score,
                             unsigned int* g_first, cudaStream_t st)
{
    if (T > 0) cp_reduce_kernel<<<nb(T), 256, 0, st>>>(T, keys, idx, head, slot, N, Bmean, score, g_key, g_nav, g_bav, g_score, g_first);
    return cudaGetLastError();
}

cudaError_t cp_launch_iota(int64_t G, unsigned int* v, cudaStream_t st)
{
    if (G > 0) cp_iota_kernel<<<nb(G), 256, 0, st>>>(G, v);
    return cudaGetLastError();
}

cudaError_t cp_launch_gather(int64_t G, const unsigned int* order, const unsigned long long* g_key, const double* g_nav,
                             const double* g_bav, const double* g_score, int64_t F, int32_t id0, const int32_t* frag_start,
                             const int32_t* frag_end, int32_t* o_bait, int32_t* o_oe, double* o_nav, double* o_bav,
                             double* o_score, double* o_mid, cudaStream_t st)
{
    if (G > 0) cp_gather_kernel<<<nb(G), 256, 0, st>>>(G, order, g_key, g_nav, g_bav, g_score, F, id0, frag_start, frag_end,
                                                    o_bait, o_oe, o_nav, o_bav, o_score, o_mid);
    return cudaGetLastError();
}

}  // namespace cd
