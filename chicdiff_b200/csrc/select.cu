// select.cu -- exact medians without sorting, shardable across ranks.
//
// R's median() of the finite entries of B columns (size factors: one column per replicate, chicdiff.R:1561;
// the MAD of the log dispersion residuals inside estimateDispersionsFit) by most-significant-digit radix
// selection on the order-preserving 64-bit image of the doubles: six passes of (11, 11, 11, 11, 11, 9) bits,
// each a shared-memory histogram over the keys that still match the running prefix, then a 2048-bin scan
// that extends the prefix.  The histograms are plain integer counts, so in a sharded run the only exchange
// is an all-reduce of B x 2048 counters per pass; no rank ever needs another rank's values and nothing is
// gathered.  Everything stays on the stream: prefix and remaining rank live in device memory.
//
// The all-reduce is part of the consuming kernel (init / scan / finish): CTA c stores this rank's counters of
// column c into every peer's mailbox over NVLink (pointers from cudaIpcOpenMemHandle), then a sequence word per
// (rank, column); it waits for all ranks' sequence words in its own mailbox and combines their slots.  Integer
// sums and minima do not depend on the order, so every rank continues with identical state.  Mailboxes are
// double-buffered by the parity of the sequence number: a peer can start exchange k+2 only after finishing k+1,
// which needs this rank's contribution to k+1, which stream order places after this rank has read exchange k.
// When peer memory is unavailable the host enqueues NCCL all-reduces between the kernels instead.
//
// Two adjacent order statistics are needed for an even count: after the k1-th value v1 is known, one more
// pass counts the keys <= v1 and finds the smallest key > v1.
#include "kernels.h"

namespace cd {

constexpr int kSelBins = 2048;

__device__ __forceinline__ unsigned long long key_of(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double value_of(unsigned long long k)
{
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// value of column c, element i, after the optional |x - center[c]| transform; +inf / NaN = excluded
__device__ __forceinline__ bool sel_load(const double* __restrict__ base, int64_t stride, int c, int64_t i,
                                         const double* __restrict__ center, double& v)
{
    double x = base[(int64_t)c * stride + i];
    if (!isfinite(x)) return false;
    if (center) x = fabs(x - center[c]);
    v = x;
    return true;
}

// state per column: [0] prefix (finally the key of the k1-th value), [1] remaining rank k, [2] total finite count m
constexpr int kSelState = 8;

__global__ void __launch_bounds__(256)
sel_count_kernel(int64_t n, int B, const double* __restrict__ base, int64_t stride, const double* __restrict__ center,
                 unsigned long long* __restrict__ counts /*B*/)
{
    const int c = blockIdx.y;
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v;
        local += sel_load(base, stride, c, i, center, v) ? 1ull : 0ull;
    }
    for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(counts + c, local);
}


// All threads of the CTA take part.  vals: len <= 2048 words (any address space); entries with index >= min_from are
// combined with min, the others with +.  Returns false (and raises *pp.err) if a peer never answered.
__device__ __forceinline__ bool sel_exchange(const SelP2P& pp, int c, unsigned long long* vals, int len, int min_from)
{
    __shared__ int timed_out;
    const int nr = pp.nranks;
    const size_t nslots = (size_t)2 * nr * kSelP2PMaxCols;
    const size_t par_base = (size_t)(pp.seq & 1ull) * nr;
    const size_t my_slot = (par_base + pp.rank) * kSelP2PMaxCols + c;
    if (threadIdx.x == 0) timed_out = 0;
    for (int r = 0; r < nr; r++) {
        unsigned long long* dst = pp.peers[r] + my_slot * kSelBins;
        for (int b = threadIdx.x; b < len; b += blockDim.x) dst[b] = vals[b];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < nr) {
        *reinterpret_cast<volatile unsigned long long*>(pp.peers[threadIdx.x] + nslots * kSelBins + my_slot) = pp.seq;
        volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
            pp.mymail + nslots * kSelBins + (par_base + threadIdx.x) * kSelP2PMaxCols + c);
        const long long t0 = clock64();
        while (*f != pp.seq) {
            if (clock64() - t0 > 20000000000LL) { timed_out = 1; break; }      // ~10 s: a peer died; give up
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int b = threadIdx.x; b < len; b += blockDim.x) {
        const bool is_min = b >= min_from;
        unsigned long long acc = is_min ? ~0ull : 0ull;
        for (int r = 0; r < nr; r++) {
            const unsigned long long x = __ldcv(pp.mymail + ((par_base + r) * kSelP2PMaxCols + c) * kSelBins + b);
            acc = is_min ? (x < acc ? x : acc) : acc + x;
        }
        vals[b] = acc;
    }
    __syncthreads();
    const bool ok = timed_out == 0;
    if (!ok && threadIdx.x == 0) atomicExch(pp.err, 1ull);
    return ok;
}

// m = total count over all ranks; k1 = (m - 1) / 2 ; state reset.  One CTA per column.
__global__ void __launch_bounds__(64)
sel_init_kernel(int B, unsigned long long* __restrict__ counts, unsigned long long* __restrict__ state,
                unsigned long long* __restrict__ le, unsigned long long* __restrict__ mg, SelP2P pp)
{
    const int c = blockIdx.x;
    bool ok = true;
    if (pp.nranks > 1) ok = sel_exchange(pp, c, counts + c, 1, 1);
    if (threadIdx.x != 0) return;
    const unsigned long long m = ok ? counts[c] : 0ull;
    unsigned long long* s = state + (size_t)c * kSelState;
    s[0] = 0ull; s[1] = (m > 0) ? (m - 1) / 2 : 0ull; s[2] = m;
    le[c] = 0ull; mg[c] = ~0ull;
}

__global__ void __launch_bounds__(256)
sel_hist_kernel(int64_t n, int B, const double* __restrict__ base, int64_t stride, const double* __restrict__ center,
                const unsigned long long* __restrict__ state, int shift, int bits, unsigned long long* __restrict__ hist /*B x 2048*/)
{
    __shared__ unsigned int sh[kSelBins];
    const int c = blockIdx.y;
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = state[(size_t)c * kSelState];
    const int hi_shift = shift + bits;                       // bits above the current digit must equal the prefix
    const unsigned long long mask = (1ull << bits) - 1ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v;
        if (!sel_load(base, stride, c, i, center, v)) continue;
        const unsigned long long k = key_of(v);
        const bool match = (hi_shift >= 64) ? true : ((k >> hi_shift) == prefix);
        if (match) atomicAdd(&sh[(unsigned)((k >> shift) & mask)], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
        if (sh[b]) atomicAdd(hist + (size_t)c * kSelBins + b, (unsigned long long)sh[b]);
}

// extend the prefix by the digit whose cumulative count first exceeds the remaining rank
__global__ void __launch_bounds__(256)
sel_scan_kernel(int B, int bits, unsigned long long* __restrict__ state, unsigned long long* __restrict__ hist, SelP2P pp)
{
    const int c = blockIdx.x;
    __shared__ unsigned long long part[256];
    unsigned long long* h = hist + (size_t)c * kSelBins;
    if (pp.nranks > 1 && !sel_exchange(pp, c, h, kSelBins, kSelBins)) {
        if (threadIdx.x == 0) state[(size_t)c * kSelState + 2] = 0ull;        // median comes out NaN; host sees *pp.err
    }
    const int per = kSelBins / 256;                          // 8 consecutive bins per thread
    unsigned long long mine = 0;
    for (int j = 0; j < per; j++) mine += h[threadIdx.x * per + j];
    part[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* s = state + (size_t)c * kSelState;
        unsigned long long k = s[1], cum = 0;
        int t = 0;
        while (t < 255 && cum + part[t] <= k) { cum += part[t]; t++; }
        int b = t * per;
        while (b < kSelBins - 1 && cum + h[b] <= k) { cum += h[b]; b++; }
        s[0] = (s[0] << bits) | (unsigned long long)b;
        s[1] = k - cum;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) h[b] = 0ull;      // ready for the next pass
}

// after the last digit: state[0] is the key of the k1-th value.  Count keys <= v1 and find min key > v1.
__global__ void __launch_bounds__(256)
sel_next_kernel(int64_t n, int B, const double* __restrict__ base, int64_t stride, const double* __restrict__ center,
                const unsigned long long* __restrict__ state, unsigned long long* __restrict__ le_out,
                unsigned long long* __restrict__ mg_out)
{
    const int c = blockIdx.y;
    const unsigned long long v1 = state[(size_t)c * kSelState];
    unsigned long long le = 0, mg = ~0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v;
        if (!sel_load(base, stride, c, i, center, v)) continue;
        const unsigned long long k = key_of(v);
        if (k <= v1) le++; else if (k < mg) mg = k;
    }
    for (int off = 16; off > 0; off >>= 1) {
        le += __shfl_down_sync(0xffffffffu, le, off);
        const unsigned long long o = __shfl_down_sync(0xffffffffu, mg, off);
        mg = o < mg ? o : mg;
    }
    if ((threadIdx.x & 31) == 0) {
        if (le) atomicAdd(le_out + c, le);
        atomicMin(mg_out + c, mg);
    }
}

// median = v1 (odd count) or (v1 + v2) / 2 ; out[c] = exp(scale * median) or scale * median ; NaN if empty
__global__ void __launch_bounds__(64)
sel_finish_kernel(int B, const unsigned long long* __restrict__ state, const unsigned long long* __restrict__ le,
                  const unsigned long long* __restrict__ mg, double* __restrict__ out, int do_exp, double scale, SelP2P pp)
{
    const int c = blockIdx.x;
    __shared__ unsigned long long lm[2];
    if (threadIdx.x == 0) { lm[0] = le[c]; lm[1] = mg[c]; }
    __syncthreads();
    bool ok = true;
    if (pp.nranks > 1) ok = sel_exchange(pp, c, lm, 2, 1);
    if (threadIdx.x != 0) return;
    const unsigned long long* s = state + (size_t)c * kSelState;
    const unsigned long long m = ok ? s[2] : 0ull;
    double med = NAN;
    if (m > 0) {
        const double v1 = value_of(s[0]);
        if (m & 1ull) med = v1;
        else {
            const unsigned long long k2 = m / 2;
            const double v2 = (lm[0] > k2) ? v1 : value_of(lm[1]);
            med = 0.5 * (v1 + v2);
        }
    }
    med *= scale;
    out[c] = do_exp ? exp(med) : med;
}

static inline dim3 sel_grid(int64_t n, int B)
{
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return dim3((unsigned)blocks, (unsigned)B);
}

cudaError_t sel_launch_count(int64_t n, int B, const double* base, int64_t stride, const double* center,
                             unsigned long long* counts, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)B, st);
    if (e != cudaSuccess) return e;
    if (n > 0) sel_count_kernel<<<sel_grid(n, B), 256, 0, st>>>(n, B, base, stride, center, counts);
    return cudaGetLastError();
}

cudaError_t sel_launch_init(int B, unsigned long long* counts, unsigned long long* state, unsigned long long* hist,
                            unsigned long long* le, unsigned long long* mg, const SelP2P& pp, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (size_t)B * kSelBins, st);
    if (e != cudaSuccess) return e;
    sel_init_kernel<<<B, 64, 0, st>>>(B, counts, state, le, mg, pp);
    return cudaGetLastError();
}

cudaError_t sel_launch_hist(int64_t n, int B, const double* base, int64_t stride, const double* center,
                            const unsigned long long* state, int pass, unsigned long long* hist, cudaStream_t st)
{
    static const int shifts[6] = {53, 42, 31, 20, 9, 0};
    static const int bitsv[6] = {11, 11, 11, 11, 11, 9};
    if (n > 0) sel_hist_kernel<<<sel_grid(n, B), 256, 0, st>>>(n, B, base, stride, center, state, shifts[pass], bitsv[pass], hist);
    return cudaGetLastError();
}

cudaError_t sel_launch_scan(int B, int pass, unsigned long long* state, unsigned long long* hist, const SelP2P& pp,
                            cudaStream_t st)
{
    static const int bitsv[6] = {11, 11, 11, 11, 11, 9};
    sel_scan_kernel<<<B, 256, 0, st>>>(B, bitsv[pass], state, hist, pp);
    return cudaGetLastError();
}

cudaError_t sel_launch_next(int64_t n, int B, const double* base, int64_t stride, const double* center,
                            const unsigned long long* state, unsigned long long* le, unsigned long long* mg, cudaStream_t st)
{
    if (n > 0) sel_next_kernel<<<sel_grid(n, B), 256, 0, st>>>(n, B, base, stride, center, state, le, mg);
    return cudaGetLastError();
}

cudaError_t sel_launch_finish(int B, const unsigned long long* state, const unsigned long long* le, const unsigned long long* mg,
                              double* out, int do_exp, double scale, const SelP2P& pp, cudaStream_t st)
{
    sel_finish_kernel<<<B, 64, 0, st>>>(B, state, le, mg, out, do_exp, scale, pp);
    return cudaGetLastError();
}

}  // namespace cd
