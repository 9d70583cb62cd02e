// tables.cu -- the per-replicate look-up tables of getFullRegionData1 built on the device from one replicate's raw
// CHiCAGO columns (chicdiff.R:632-634, 659-692, 828-853).
//
// The reference keys the CHiCAGO table x by (baitID, otherEndID) (:632) and then takes
//   per baitID        the first (s_j, tblb)           x[, list(s_j = s_j[1], tblb = tblb[1]), by = "baitID"]        (:659)
//   per otherEndID    the first (s_i, tlb)            x[, list(s_i = s_i[1], tlb = tlb[1]), by = "otherEndID"]     (:668)
//   per (tblb, tlb)   the first Tmean, after the stable re-key setkey(x, tlb, tblb)                                 (:678-680)
// and left-joins the per-pair counts of the .chinput file (:828-853).  data.table does this with radix sorts of the
// whole table (1e7-1e8 rows per replicate genome-wide).  "First in key order" needs no sort: it is the row that
// minimises (baitID, otherEndID, input position) within its group, found here with one 64-bit atomicMin per row and
// group -- the key's group part is constant inside a group, so the packed remainder of the order fits 64 bits:
//   group baitID      -> min (otherEndID << 32 | row)
//   group otherEndID  -> min (baitID << 32 | row)
//   group (tblb, tlb) -> min (baitID << 32 | otherEndID), then min row among the rows that attain it (second pass)
// A second small kernel gathers the winners' values into the per-fragment tables of cd_sample_tables.  The sparse count
// rows do need their (baitID, otherEndID) order for the merge-walk of assemble.cu: one CUB radix sort of packed keys,
// then a lower bound per fragment gives the CSR offsets.
#include "kernels.h"

namespace cd {

static inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }
constexpr unsigned long long kNone = ~0ull;

__global__ void __launch_bounds__(256)
tb_first_kernel(int64_t m, const int32_t* __restrict__ bait, const int32_t* __restrict__ oe, const int32_t* __restrict__ tblb,
                const int32_t* __restrict__ tlb, int64_t F, int32_t id0, int n_tblb, int n_tlb,
                unsigned long long* __restrict__ best_bait, unsigned long long* __restrict__ best_oe,
                unsigned long long* __restrict__ best_tt, int32_t* __restrict__ status)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int64_t b = (int64_t)bait[r] - id0, o = (int64_t)oe[r] - id0;
    if (b < 0 || b >= F || o < 0 || o >= F) { atomicOr(status, 1); return; }      // a fragment that is not in the rmap
    atomicMin(best_bait + b, ((unsigned long long)(unsigned)oe[r] << 32) | (unsigned long long)r);
    atomicMin(best_oe + o, ((unsigned long long)(unsigned)bait[r] << 32) | (unsigned long long)r);
    const int tb = tblb[r], tl = tlb[r];
    if (tb >= 0 && tl >= 0) {
        if (tb >= n_tblb || tl >= n_tlb) { atomicOr(status, 2); return; }
        atomicMin(best_tt + (size_t)tb * n_tlb + tl, ((unsigned long long)(unsigned)bait[r] << 32) | (unsigned long long)(unsigned)oe[r]);
    }
}

// second pass for the (tblb, tlb) groups: the smallest row among those whose (baitID, otherEndID) is the group's minimum
__global__ void __launch_bounds__(256)
tb_first_tt_row_kernel(int64_t m, const int32_t* __restrict__ bait, const int32_t* __restrict__ oe, const int32_t* __restrict__ tblb,
                       const int32_t* __restrict__ tlb, int n_tlb, const unsigned long long* __restrict__ best_tt,
                       unsigned long long* __restrict__ best_tt_row)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int tb = tblb[r], tl = tlb[r];
    if (tb < 0 || tl < 0) return;
    const unsigned long long key = ((unsigned long long)(unsigned)bait[r] << 32) | (unsigned long long)(unsigned)oe[r];
    if (best_tt[(size_t)tb * n_tlb + tl] == key) atomicMin(best_tt_row + (size_t)tb * n_tlb + tl, (unsigned long long)r);
}

__global__ void __launch_bounds__(256)
tb_fill_kernel(int64_t F, int n_tt, const unsigned long long* __restrict__ best_bait, const unsigned long long* __restrict__ best_oe,
               const unsigned long long* __restrict__ best_tt_row, const double* __restrict__ s_j_rows, const int32_t* __restrict__ tblb_rows,
               const double* __restrict__ s_i_rows, const int32_t* __restrict__ tlb_rows, const double* __restrict__ tmean_rows,
               double* __restrict__ s_j, int32_t* __restrict__ tblb, double* __restrict__ s_i, int32_t* __restrict__ tlb,
               double* __restrict__ tmean)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) {
        const unsigned long long kb = best_bait[f], ko = best_oe[f];
        if (kb == kNone) { s_j[f] = NAN; tblb[f] = -1; }
        else { const unsigned long long r = kb & 0xffffffffull; s_j[f] = s_j_rows[r]; tblb[f] = tblb_rows[r]; }
        if (ko == kNone) { s_i[f] = NAN; tlb[f] = -1; }
        else { const unsigned long long r = ko & 0xffffffffull; s_i[f] = s_i_rows[r]; tlb[f] = tlb_rows[r]; }
    }
    if (f < n_tt) {
        const unsigned long long r = best_tt_row[f];
        tmean[f] = (r == kNone) ? NAN : tmean_rows[r];
    }
}

cudaError_t tb_launch_first(int64_t m, const int32_t* bait, const int32_t* oe, const int32_t* tblb, const int32_t* tlb, int64_t F,
                            int32_t id0, int n_tblb, int n_tlb, unsigned long long* best /* 2 F + 2 n_tblb n_tlb words */,
                            int32_t* status, cudaStream_t st)
{
    const size_t nt = (size_t)n_tblb * n_tlb;
    cudaError_t e = cudaMemsetAsync(best, 0xff, sizeof(unsigned long long) * (2 * (size_t)F + 2 * nt), st);
    if (e != cudaSuccess || m == 0) return e;
    unsigned long long* bb = best; unsigned long long* bo = best + F; unsigned long long* bt = best + 2 * F; unsigned long long* btr = bt + nt;
    tb_first_kernel<<<blocks_for(m, 256), 256, 0, st>>>(m, bait, oe, tblb, tlb, F, id0, n_tblb, n_tlb, bb, bo, bt, status);
    tb_first_tt_row_kernel<<<blocks_for(m, 256), 256, 0, st>>>(m, bait, oe, tblb, tlb, n_tlb, bt, btr);
    return cudaGetLastError();
}

cudaError_t tb_launch_fill(int64_t F, int n_tblb, int n_tlb, const unsigned long long* best, const double* s_j_rows,
                           const int32_t* tblb_rows, const double* s_i_rows, const int32_t* tlb_rows, const double* tmean_rows,
                           double* s_j, int32_t* tblb, double* s_i, int32_t* tlb, double* tmean, cudaStream_t st)
{
    const size_t nt = (size_t)n_tblb * n_tlb;
    const int64_t work = F > (int64_t)nt ? F : (int64_t)nt;
    tb_fill_kernel<<<blocks_for(work, 256), 256, 0, st>>>(F, (int)nt, best, best + F, best + 2 * F + nt, s_j_rows, tblb_rows, s_i_rows,
                                                        tlb_rows, tmean_rows, s_j, tblb, s_i, tlb, tmean);
    return cudaGetLastError();
}

// ---- sparse counts: rows with N of pairs whose bait is a fragment of the rmap, in (baitID, otherEndID) order, CSR by bait ----
__global__ void __launch_bounds__(256)
tb_count_keys_kernel(int64_t m, const int32_t* __restrict__ bait, const int32_t* __restrict__ oe, int64_t F, int32_t id0,
                     unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int64_t b = (int64_t)bait[r] - id0;
    // rows of baits outside the rmap sort to the end and are cut off (chicdiff.R:831: the .chinput is restricted to the RU baits)
    keys[r] = (b < 0 || b >= F) ? kNone : (((unsigned long long)b << 32) | (unsigned long long)(unsigned)oe[r]);
    idx[r] = (unsigned int)r;
}

__global__ void __launch_bounds__(256)
tb_count_offsets_kernel(int64_t F, int64_t m, const unsigned long long* __restrict__ sorted_keys, int64_t* __restrict__ cnt_off)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f > F) return;
    const unsigned long long target = (unsigned long long)f << 32;          // first key of bait f (f == F: end of the valid rows)
    int64_t lo = 0, hi = m;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    cnt_off[f] = lo;
}

__global__ void __launch_bounds__(256)
tb_count_gather_kernel(int64_t m_valid, const unsigned long long* __restrict__ sorted_keys, const unsigned int* __restrict__ sorted_idx,
                       const int32_t* __restrict__ N_rows, int32_t* __restrict__ cnt_oe, int32_t* __restrict__ cnt_N)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m_valid) return;
    cnt_oe[k] = (int32_t)(sorted_keys[k] & 0xffffffffull);
    cnt_N[k] = N_rows[sorted_idx[k]];
}

cudaError_t tb_launch_count_keys(int64_t m, const int32_t* bait, const int32_t* oe, int64_t F, int32_t id0, unsigned long long* keys,
                                 unsigned int* idx, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    tb_count_keys_kernel<<<blocks_for(m, 256), 256, 0, st>>>(m, bait, oe, F, id0, keys, idx);
    return cudaGetLastError();
}

cudaError_t tb_launch_count_offsets(int64_t F, int64_t m, const unsigned long long* sorted_keys, int64_t* cnt_off, cudaStream_t st)
{
    tb_count_offsets_kernel<<<blocks_for(F + 1, 256), 256, 0, st>>>(F, m, sorted_keys, cnt_off);
    return cudaGetLastError();
}

cudaError_t tb_launch_count_gather(int64_t m_valid, const unsigned long long* sorted_keys, const unsigned int* sorted_idx,
                                   const int32_t* N_rows, int32_t* cnt_oe, int32_t* cnt_N, cudaStream_t st)
{
    if (m_valid == 0) return cudaSuccess;
    tb_count_gather_kernel<<<blocks_for(m_valid, 256), 256, 0, st>>>(m_valid, sorted_keys, sorted_idx, N_rows, cnt_oe, cnt_N);
    return cudaGetLastError();
}

}  // namespace cd
