// offsets.cu -- size factors, stage 2 (per-region normalisation offsets), the parametric
// dispersion trend passes and the deterministic reductions they need.
//
//   chicdiff.R:1561-1562  estimateSizeFactors: log-ratio matrix (medians are taken by the host
//                         orchestration after a device sort)
//   chicdiff.R:1583-1589  FullMean scaling factors, rows with NA replaced by the size factors
//   chicdiff.R:1614-1615, theta mix with the size factors and row geometric-mean rescale
//              1635-1638
//   DESeq2 parametricDispersionFit / dispersionFunction<-: Gamma(identity) IRLS sums, fitted
//                         trend and log residuals
//
// All reductions are two-stage with a fixed block count and fixed summation order, so results
// are bit-reproducible from run to run (no floating-point atomics).
#include "kernels.h"

namespace cd {

static inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* dst)
{
    __shared__ double sh[NV][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) sh[k][wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0.0;
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; w++) x += sh[threadIdx.x][w];
        dst[threadIdx.x] = x;
    }
    __syncthreads();
}

// final stage: out[k] = sum_b partial[b*nv + k]; one warp per value, lanes stride over the blocks
// and a fixed shuffle tree joins them, so the summation order never changes
__global__ void __launch_bounds__(32) final_reduce_kernel(int nblocks, int nv, const double* __restrict__ partial,
                                                          double* __restrict__ out)
{
    const int k = blockIdx.x;
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) x += partial[(size_t)b * nv + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if (threadIdx.x == 0) out[k] = x;
}

// ---------------------------------------------------------------------------------------
// masked column sums: out[s] = sum_i M[s][i] over rows with mask[i] == 0 ; out[S] = #rows
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
masked_colsums_kernel(int64_t n, int S, const double* __restrict__ M, const uint8_t* __restrict__ mask,
                      double* __restrict__ partial)
{
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = (lo + chunk < n) ? lo + chunk : n;
    for (int s = 0; s <= S; s++) {
        double v[1] = {0.0};
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const bool use = (mask == nullptr) || ((mask[i] & CD_FLAG_ALLZERO) == 0);
            if (use) v[0] += (s < S) ? M[(int64_t)s * n + i] : 1.0;
        }
        block_reduce_store<1>(v, partial + (size_t)blockIdx.x * (S + 1) + s);
    }
}

cudaError_t launch_masked_colsums(int64_t n, int S, const double* M, const uint8_t* mask, double* partial,
                                  double* out, cudaStream_t st)
{
    masked_colsums_kernel<<<kReduceBlocks, 256, 0, st>>>(n, S, M, mask, partial);
    final_reduce_kernel<<<S + 1, 32, 0, st>>>(kReduceBlocks, S + 1, partial, out);
    return cudaGetLastError();
}

cudaError_t launch_sum_nan(int64_t n, const double* v, double* partial, double* out, cudaStream_t st)
{
    return launch_masked_colsums(n, 1, v, nullptr, partial, out, st);
}

// ---------------------------------------------------------------------------------------
// size factors: LR[s][i] = log K[s][i] - mean_s' log K[s'][i] for rows with every K > 0,
// +inf otherwise (sorted to the end; the host reads the median of the finite prefix)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
log_ratios_kernel(int64_t n, int S, const int32_t* __restrict__ K, double* __restrict__ LR)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    bool ok = true;
    for (int s = 0; s < S; s++) {
        const int32_t k = K[(int64_t)s * n + i];
        ok = ok && (k > 0);
        acc += log((double)k);
    }
    const double lgm = acc / S;
    for (int s = 0; s < S; s++)
        LR[(int64_t)s * n + i] = ok ? log((double)K[(int64_t)s * n + i]) - lgm : INFINITY;
}

cudaError_t launch_log_ratios(int64_t n, int S, const int32_t* K, double* LR, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    log_ratios_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, S, K, LR);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// stage 2: normalisation factors.  mode 0 standard, 1 fullmean, 2 combined(theta)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
norm_factors_kernel(int64_t n, int S, const double* __restrict__ FMagg, const double* __restrict__ sf,
                    int mode, double theta, double* __restrict__ nf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) {
        for (int s = 0; s < S; s++) nf[(int64_t)s * n + i] = sf[s];
        return;
    }
    double acc = 0.0;
    for (int s = 0; s < S; s++) acc += log(FMagg[(int64_t)s * n + i]);
    const double g = exp(acc / S);
    bool anyna = false;
    for (int s = 0; s < S; s++) {
        const double m3 = FMagg[(int64_t)s * n + i] / g;
        anyna = anyna || isnan(m3);
    }
    if (mode == 1) {
        for (int s = 0; s < S; s++) nf[(int64_t)s * n + i] = anyna ? sf[s] : FMagg[(int64_t)s * n + i] / g;
        return;
    }
    double acc2 = 0.0;
    for (int s = 0; s < S; s++) {
        const double m3 = anyna ? sf[s] : FMagg[(int64_t)s * n + i] / g;
        acc2 += log(m3 * (1.0 - theta) + sf[s] * theta);
    }
    const double g2 = exp(acc2 / S);
    for (int s = 0; s < S; s++) {
        const double m3 = anyna ? sf[s] : FMagg[(int64_t)s * n + i] / g;
        nf[(int64_t)s * n + i] = (m3 * (1.0 - theta) + sf[s] * theta) / g2;
    }
}

cudaError_t launch_norm_factors(int64_t n, int S, const double* FMagg, const double* sf, int mode, double theta,
                                double* nf, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    norm_factors_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, S, FMagg, sf, mode, theta, nf);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// parametric trend: one pass of glm.fit(family = Gamma(link = "identity")) at coefficients b
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
trend_pass_kernel(int64_t n, const double* __restrict__ baseMean, const double* __restrict__ dispGeneEst,
                  const uint8_t* __restrict__ flags, double c0, double c1, double b0, double b1,
                  double* __restrict__ partial)
{
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = (lo + chunk < n) ? lo + chunk : n;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        if (flags[i] & CD_FLAG_ALLZERO) continue;
        const double d = dispGeneEst[i];
        if (!(d > 100.0 * kMinDisp)) continue;
        const double x = 1.0 / baseMean[i];
        const double r = d / (c0 + c1 * x);
        if (!((r > 1e-4) && (r < 15.0))) continue;
        const double mu = b0 + b1 * x;
        v[7] += 1.0;
        if (!(mu > 0.0) || !isfinite(mu)) { v[6] += 1.0; continue; }
        const double w = 1.0 / (mu * mu);
        v[0] += w; v[1] += w * x; v[2] += w * x * x;
        v[3] += w * d; v[4] += w * x * d;
        v[5] += -2.0 * (log(d / mu) - (d - mu) / mu);
    }
    block_reduce_store<8>(v, partial + (size_t)blockIdx.x * 8);
}

cudaError_t launch_trend_pass(int64_t n, const double* baseMean, const double* dispGeneEst, const uint8_t* flags,
                              double c0, double c1, double b0, double b1, double* partial, double* out,
                              cudaStream_t st)
{
    trend_pass_kernel<<<kReduceBlocks, 256, 0, st>>>(n, baseMean, dispGeneEst, flags, c0, c1, b0, b1, partial);
    final_reduce_kernel<<<8, 32, 0, st>>>(kReduceBlocks, 8, partial, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// The whole parametricDispersionFit on the device: one cooperative kernel, one CTA per SM.
// Every pass (sums + deviance at coefficients b over the rows kept by the outer coefficients c)
// is a chunked block reduction, a grid barrier, and a fixed-order sum of the per-CTA partials
// that every CTA repeats for itself, so all CTAs take the same branch of glm.fit's control flow
// (start validity, IRLS <= 25, step halving, outer loop <= 11) without a host round trip.
// out[0..1] = coefficients, out[2] = status (0 ok; >0 = reason the reference would fall back to
// a local fit), out[3] = outer iterations, out[4] = passes.
// ---------------------------------------------------------------------------------------
constexpr int kTrendThreads = 512;

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks, unsigned int& phase)
{
    __syncthreads();
    phase++;                                   // every thread keeps the same phase count
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (atomicAdd(bar, 0u) < phase * nblocks) { }
        __threadfence();
    }
    __syncthreads();
}

// Cross-GPU part of a pass (sharded runs): CTA 0 stores this rank's 8 sums straight into every peer's
// mailbox over NVLink (peer pointers from cudaIpcOpenMemHandle), then a sequence word.  EVERY CTA then waits
// until all ranks' slots of the local mailbox carry this pass's sequence and adds them in rank order, so all
// CTAs of all ranks obtain bit-identical totals and follow the same control flow, without a second grid
// barrier.  Slots are double-buffered by pass parity: a peer can only be one pass ahead, because finishing a
// pass needs everybody's contribution to it, and CTA 0 contributes to pass k+1 only after the grid barrier of
// pass k+1, i.e. after all local CTAs have read pass k.  (Between two launches the stream carries other
// collectives -- the medians -- so a peer cannot start the next launch while this one still reads.)
__device__ __forceinline__ void p2p_allreduce8(const TrendP2P& pp, unsigned long long seq, unsigned int parity, double* sh_tot)
{
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (blockIdx.x == 0 && (int)threadIdx.x < pp.nranks) {
        double* dst = pp.peers[threadIdx.x] + ((size_t)parity * pp.nranks + pp.rank) * 16;
#pragma unroll
        for (int k = 0; k < 8; k++) dst[k] = sh_tot[k];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(dst + 8) = seq;
    }
    if ((int)threadIdx.x < pp.nranks) {
        volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
            pp.mymail + ((size_t)parity * pp.nranks + threadIdx.x) * 16 + 8);
        const long long t0 = clock64();
        while (*f != seq) {
            if (clock64() - t0 > 20000000000LL) { timed_out = 1; break; }      // ~10 s: a peer died; give up
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double x = 0.0;
        for (int r = 0; r < pp.nranks; r++)
            x += *reinterpret_cast<volatile double*>(pp.mymail + ((size_t)parity * pp.nranks + r) * 16 + threadIdx.x);
        if (timed_out) x = (threadIdx.x == 6) ? 1.0 : NAN;          // reads as "invalid" in the control flow below
        sh_tot[threadIdx.x] = x;
    }
    __syncthreads();
}

// One pass: the sums glm.fit needs at coefficients b over the rows kept by the outer coefficients c.
// xs is n doubles of scratch that carries 1/baseMean between passes: NaN = row never used (all-zero region or
// dispersion at the floor), negative = excluded by the current outer coefficients.  `refresh` = 2 on the first
// pass of the launch (fill xs), 1 on the first pass of an outer iteration (re-decide the sign), 0 otherwise.
// Every row is always handled by the same thread, so xs needs no synchronisation.
__device__ __forceinline__ void trend_pass_device(int64_t n, const double* __restrict__ baseMean,
                                                  const double* __restrict__ dispGeneEst, const uint8_t* __restrict__ flags,
                                                  double* __restrict__ xs, int refresh,
                                                  double c0, double c1, double b0, double b1, double* partial_base,
                                                  unsigned int* bar, unsigned int& phase, double* sh_tot /*8, shared*/,
                                                  const TrendP2P& pp, unsigned long long& pass_no)
{
    // partials are double-buffered by pass parity: a CTA can be at most one pass ahead of the slowest
    // one (there is a barrier in every pass), so one barrier per pass is enough
    double* partial = partial_base + (size_t)(phase & 1u) * 8 * gridDim.x;
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = (lo + chunk < n) ? lo + chunk : n;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double xv;
        const double d = dispGeneEst[i];
        if (refresh == 2) {
            xv = NAN;
            if (!(flags[i] & CD_FLAG_ALLZERO) && (d > 100.0 * kMinDisp)) xv = 1.0 / baseMean[i];
        } else {
            xv = xs[i];
        }
        if (refresh) {
            if (xv != xv) { if (refresh == 2) xs[i] = xv; continue; }
            const double x = fabs(xv);
            const double r = d / (c0 + c1 * x);
            xv = ((r > 1e-4) && (r < 15.0)) ? x : -x;
            xs[i] = xv;
        }
        if (!(xv > 0.0)) continue;
        const double x = xv;
        const double mu = b0 + b1 * x;
        v[7] += 1.0;
        if (!(mu > 0.0) || !isfinite(mu)) { v[6] += 1.0; continue; }
        double w, t;
        if (mu > 1e-100 && mu < 1e100) { const double inv = rcp_pos(mu); w = inv * inv; t = d * inv; }
        else { w = 1.0 / (mu * mu); t = d / mu; }
        const double wx = w * x;
        v[0] += w; v[1] += wx; v[2] += wx * x;
        v[3] += w * d; v[4] += wx * d;
        // unit deviance of the Gamma family: -2 (log(d/mu) - (d - mu)/mu)
        const double lt = (t > 1e-300 && t < 1e300) ? log_pos(t) : log(t);
        v[5] += -2.0 * (lt - (t - 1.0));
    }
    // block reduction (fixed order)
    __shared__ double sh[8][kTrendThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double x = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) sh[k][wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double x = 0.0;
        for (int w = 0; w < kTrendThreads / 32; w++) x += sh[threadIdx.x][w];
        __stcg(partial + (size_t)blockIdx.x * 8 + threadIdx.x, x);
    }
    grid_barrier(bar, gridDim.x, phase);
    // fixed-order sum of the per-CTA partials, one warp per value; every CTA repeats it for itself
    if (wid < 8) {
        double x = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) x += __ldcg(partial + (size_t)b * 8 + wid);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) sh_tot[wid] = x;
    }
    __syncthreads();
    if (pp.nranks > 1) {
        pass_no++;
        p2p_allreduce8(pp, (pp.epoch << 32) | pass_no, (unsigned int)(pass_no & 1ull), sh_tot);
    }
}

__global__ void __launch_bounds__(kTrendThreads)
trend_fit_kernel(int64_t n, const double* __restrict__ baseMean, const double* __restrict__ dispGeneEst,
                 const uint8_t* __restrict__ flags, double* __restrict__ xs, double* partial, unsigned int* bar, double* out,
                 TrendP2P pp)
{
    __shared__ double tot[8];
    unsigned int phase = 0;
    unsigned long long pass_no = 0;
    double c0 = 0.1, c1 = 1.0;
    int iter = 0, status = 0, passes = 0;
    while (true) {
        double b0 = c0, b1 = c1, ob0 = c0, ob1 = c1;
        trend_pass_device(n, baseMean, dispGeneEst, flags, xs, passes == 0 ? 2 : 1, c0, c1, b0, b1, partial, bar, phase, tot, pp,
                          pass_no);
        passes++;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = tot[k];
        if (v[7] < 2.0) { status = 1; break; }
        if (v[6] > 0.0) { status = 2; break; }
        double devold = v[5];
        bool conv = false;
        for (int it = 0; it < 25 && status == 0; it++) {
            const double det = v[0] * v[2] - v[1] * v[1];
            double nb0 = (v[2] * v[3] - v[1] * v[4]) / det;
            double nb1 = (v[0] * v[4] - v[1] * v[3]) / det;
            double w[8];
            int halv = 0;
            while (true) {
                trend_pass_device(n, baseMean, dispGeneEst, flags, xs, 0, c0, c1, nb0, nb1, partial, bar, phase, tot, pp, pass_no);
                passes++;
#pragma unroll
                for (int k = 0; k < 8; k++) w[k] = tot[k];
                if (w[6] == 0.0 && isfinite(w[5])) break;
                if (++halv > 25) { status = 3; break; }
                nb0 = 0.5 * (nb0 + ob0); nb1 = 0.5 * (nb1 + ob1);
            }
            if (status) break;
            b0 = nb0; b1 = nb1;
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = w[k];
            const double dev = w[5];
            if (fabs(dev - devold) / (fabs(dev) + 0.1) < 1e-8) { conv = true; break; }
            devold = dev; ob0 = b0; ob1 = b1;
        }
        if (status) break;
        const double oc0 = c0, oc1 = c1;
        c0 = b0; c1 = b1;
        if (!(c0 > 0.0 && c1 > 0.0)) { status = 4; break; }
        const double l0 = log(c0 / oc0), l1 = log(c1 / oc1);
        if ((l0 * l0 + l1 * l1 < 1e-6) && conv) break;
        iter++;
        if (iter > 10) { status = 5; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = c0; out[1] = c1; out[2] = (double)status; out[3] = (double)(iter + 1); out[4] = (double)passes;
    }
}

cudaError_t launch_trend_fit(int64_t n, const double* baseMean, const double* dispGeneEst, const uint8_t* flags,
                             double* xs, double* partial, unsigned int* bar, double* out, const TrendP2P& pp_in, cudaStream_t st)
{
    TrendP2P pp = pp_in;
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return cudaErrorNotSupported;
    cudaError_t e = cudaMemsetAsync(bar, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&n, (void*)&baseMean, (void*)&dispGeneEst, (void*)&flags, (void*)&xs, (void*)&partial, (void*)&bar, (void*)&out,
                    (void*)&pp};
    return cudaLaunchCooperativeKernel((const void*)trend_fit_kernel, dim3((unsigned)sms), dim3(kTrendThreads), args, 0, st);
}

// dispFit = a0 + a1/baseMean ; resid = log(dispGeneEst) - log(dispFit) or +inf when excluded ; coefficients on device
__global__ void __launch_bounds__(256)
trend_apply_kernel(int64_t n, const double* __restrict__ baseMean, const double* __restrict__ dispGeneEst,
                   const uint8_t* __restrict__ flags, const double* __restrict__ coefs, double* __restrict__ dispFit,
                   double* __restrict__ resid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i] & CD_FLAG_ALLZERO) { dispFit[i] = NAN; resid[i] = INFINITY; return; }
    const double f = coefs[0] + coefs[1] / baseMean[i];
    const double d = dispGeneEst[i];
    dispFit[i] = f;
    resid[i] = (d >= 100.0 * kMinDisp) ? log(d) - log(f) : INFINITY;
}

cudaError_t launch_trend_apply(int64_t n, const double* baseMean, const double* dispGeneEst, const uint8_t* flags,
                               const double* coefs_dev, double* dispFit, double* resid, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    trend_apply_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, baseMean, dispGeneEst, flags, coefs_dev, dispFit, resid);
    return cudaGetLastError();
}

}  // namespace cd
