// assemble.cu -- per-replicate assembly of the region-universe rows fused with stage 1.
//
// What getFullRegionData1() does per replicate (chicdiff.R:609-702, 820-910) -- the keyed join of the
// region-universe pairs against the CHiCAGO table, the per-bait (s_j, tblb) and per-other-end (s_i, tlb)
// look-ups, the Tmean table with its min-imputation, the distance recomputation, Chicago's Bmean =
// s_j s_i f(d), FullMean = Bmean + Tmean and the count merge with zero fill -- followed directly by the
// region sums of DESeq2Wrap (chicdiff.R:1540-1547), without ever materialising the long table:
// one lane = one (region, replicate); it binary-searches the replicate's sparse count rows of its bait
// once, then merge-walks them along the region's fragments (both are sorted by otherEndID) while
// evaluating the expected background of every fragment, and writes K[s][i], FullMean[s][i].
// Optionally the per-row columns are written too (FullRegionData for saveAuxData / plots).
#include "kernels.h"

namespace cd {

__device__ __forceinline__ double dist_fun_dev(double d, const double* __restrict__ p)
{
    const double l = log(d);
    double out;
    if (l > p[5]) out = p[8] + l * p[9];
    else if (l < p[4]) out = p[6] + l * p[7];
    else out = p[0] + p[1] * l + p[2] * (l * l) + p[3] * (l * l * l);
    return exp(out);
}

__global__ void __launch_bounds__(128)
assemble_kernel(int64_t n, int S, const int64_t* __restrict__ row_off, int64_t R,
                const int32_t* __restrict__ row_bait, const int32_t* __restrict__ row_oe,
                int64_t F, int32_t frag_id0, const int32_t* __restrict__ frag_chr,
                const int32_t* __restrict__ frag_start, const int32_t* __restrict__ frag_end,
                const AssembleTables* __restrict__ tabs,
                int32_t* __restrict__ K, double* __restrict__ FM, double* __restrict__ avDist,
                int32_t* __restrict__ N_rows, double* __restrict__ FM_rows, double* __restrict__ BM_rows,
                int32_t* __restrict__ status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= n) return;
    const AssembleTables t = tabs[s];
    const int64_t r0 = row_off[i], r1 = row_off[i + 1];
    int64_t acc = 0;
    double facc = 0.0, dsum = 0.0;
    if (r1 > r0) {
        const int32_t bait = row_bait[r0];
        const int64_t b = (int64_t)bait - frag_id0;
        if (b < 0 || b >= F) { atomicOr(status, 1); return; }
        const double sj = t.s_j[b];
        const int tb = t.tblb[b];
        const int chrb = frag_chr[b];
        const double midb2 = (double)frag_start[b] + (double)frag_end[b];
        // R round(0.5 * (start + end)) of the bait, half to even (chicdiff.R:871)
        const double midb = rint(0.5 * midb2);
        const int64_t hi = t.cnt_off[b + 1];
        int64_t ptr = t.cnt_off[b];
        {   // first sparse row of this bait with otherEndID >= the region's first fragment
            int64_t lo = ptr, h2 = hi;
            const int32_t key = row_oe[r0];
            while (lo < h2) {
                const int64_t mid = (lo + h2) >> 1;
                if (t.cnt_oe[mid] < key) lo = mid + 1; else h2 = mid;
            }
            ptr = lo;
        }
        for (int64_t r = r0; r < r1; r++) {
            const int32_t oe = row_oe[r];
            const int64_t o = (int64_t)oe - frag_id0;
            if (o < 0 || o >= F || row_bait[r] != bait) { atomicOr(status, (o < 0 || o >= F) ? 1 : 2); return; }
            while (ptr < hi && t.cnt_oe[ptr] < oe) ptr++;
            const int32_t cnt = (ptr < hi && t.cnt_oe[ptr] == oe) ? t.cnt_N[ptr] : 0;
            const bool cis = (frag_chr[o] == chrb);
            const double mido2 = (double)frag_start[o] + (double)frag_end[o];
            const double d = rint((mido2 - midb2) / 2.0);           // chicdiff.R:648
            double si = t.s_i[o];
            if (isnan(si)) si = 1.0;                                    // chicdiff.R:672
            const int tl = t.tlb[o];
            double tm = NAN;
            if (tb >= 0) tm = (tl >= 0) ? t.tmean[tb * t.n_tlb + tl] : t.tmin[tb];
            double bm = cis ? sj * si * dist_fun_dev(fabs(d), t.distfun) : 0.0;
            if (isnan(sj)) bm = NAN;                                    // chicdiff.R:702
            const double fm = bm + tm;
            acc += cnt;
            facc += fm;
            dsum += cis ? rint(0.5 * mido2) - midb : NAN;              // chicdiff.R:878-881
            if (N_rows) { N_rows[(int64_t)s * R + r] = cnt; FM_rows[(int64_t)s * R + r] = fm; BM_rows[(int64_t)s * R + r] = bm; }
        }
    }
    K[(int64_t)s * n + i] = (acc > 2147483647LL) ? INT32_MIN : (int32_t)acc;
    FM[(int64_t)s * n + i] = facc;
    if (s == 0 && avDist) avDist[i] = (r1 > r0) ? dsum / (double)(r1 - r0) : NAN;
}

// min over the non-NA entries of every Tmean row (chicdiff.R:689-691)
__global__ void tmin_kernel(int n_tblb, int n_tlb, const double* __restrict__ tmean, double* __restrict__ tmin)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_tblb) return;
    double m = NAN;
    for (int b = 0; b < n_tlb; b++) {
        const double v = tmean[a * n_tlb + b];
        if (!isnan(v) && (isnan(m) || v < m)) m = v;
    }
    tmin[a] = m;
}

cudaError_t launch_tmin(int n_tblb, int n_tlb, const double* tmean, double* tmin, cudaStream_t st)
{
    if (n_tblb <= 0) return cudaSuccess;
    tmin_kernel<<<(n_tblb + 63) / 64, 64, 0, st>>>(n_tblb, n_tlb, tmean, tmin);
    return cudaGetLastError();
}

cudaError_t launch_assemble(int64_t n, int S, const int64_t* row_off, int64_t R, const int32_t* row_bait,
                            const int32_t* row_oe, int64_t F, int32_t frag_id0, const int32_t* frag_chr,
                            const int32_t* frag_start, const int32_t* frag_end, const AssembleTables* tabs_dev,
                            int32_t* K, double* FM, double* avDist, int32_t* N_rows, double* FM_rows,
                            double* BM_rows, int32_t* status, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)S);
    assemble_kernel<<<grid, 128, 0, st>>>(n, S, row_off, R, row_bait, row_oe, F, frag_id0, frag_chr, frag_start, frag_end,
                                          tabs_dev, K, FM, avDist, N_rows, FM_rows, BM_rows, status);
    return cudaGetLastError();
}

}  // namespace cd
