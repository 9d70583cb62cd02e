// select.cu -- exact medians without sorting, shardable across ranks, one kernel per batch of medians.
//
// R's median() of the finite entries of B columns (size factors: one column per replicate, chicdiff.R:1561;
// the MAD of the log dispersion residuals inside estimateDispersionsFit: one column per fit of a batch) by
// most-significant-digit radix selection on the order-preserving 64-bit image of the doubles: six passes of
// (11, 11, 11, 11, 11, 9) bits, each a shared-memory histogram over the keys that still match the running prefix,
// then a 2048-bin scan that extends the prefix.  The histograms are plain integer counts, so in a sharded run the
// only exchange is an all-reduce of B x 2048 counters per pass; no rank ever needs another rank's values and nothing
// is gathered.
//
// Everything -- the count of finite values, the six histogram / scan rounds, the look for the second middle value of
// an even count, the final average -- runs in ONE cooperative kernel with a grid barrier between the phases (round 1
// used 16 launches per median).  In the "column" phases CTA c owns column c: it exchanges that column's counters
// with the other ranks (it stores them into every peer's mailbox over NVLink -- pointers from cudaIpcOpenMemHandle
// or, in a single process, cudaDeviceEnablePeerAccess -- then a sequence word per (rank, column), waits for all
// ranks' sequence words in its own mailbox and combines their slots; integer sums and minima do not depend on the
// order, so every rank continues with identical state), while the other CTAs wait at the barrier.  Mailboxes are
// double-buffered by the parity of the sequence number: a peer can start exchange k+2 only after finishing k+1, which
// needs this rank's contribution to k+1, which this rank makes only after it has read exchange k.
//
// Two adjacent order statistics are needed for an even count: after the k1-th value v1 is known, one more
// pass counts the keys <= v1 and finds the smallest key > v1.
#include "kernels.h"

namespace cd {

constexpr int kSelBins = 2048;
constexpr int kSelThreads = 256;
constexpr int kSelMaxRanks = 16;                    // peer-memory mailboxes exist for up to 16 ranks (cd_comm_init)
constexpr long long kSpinLimit = 20000000000LL;      // ~10 s of SM clocks: a peer died; give up

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

// value after the optional |x - center[c]| transform; +inf / NaN = excluded
__device__ __forceinline__ bool sel_xform(double x, const double* __restrict__ center, int c, double& v)
{
    if (!isfinite(x)) return false;
    if (center) x = fabs(x - center[c]);
    v = x;
    return true;
}

// Every pass streams 8 bytes per value; four values per thread are loaded before any is processed so that enough loads are
// in flight to cover the HBM / L2 latency.  f(value) is called for every included value of column c.
template <typename F>
__device__ __forceinline__ void sel_for_each(const double* __restrict__ base, int64_t stride, int c, int64_t n, int64_t tid, int64_t nthr,
                                             const double* __restrict__ center, F f)
{
    const double* col = base + (int64_t)c * stride;
    int64_t i = tid;
    for (; i + 3 * nthr < n; i += 4 * nthr) {
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) x[u] = col[i + u * nthr];
#pragma unroll
        for (int u = 0; u < 4; u++) { double v; if (sel_xform(x[u], center, c, v)) f(v); }
    }
    for (; i < n; i += nthr) { double v; if (sel_xform(col[i], center, c, v)) f(v); }
}

// state per column: [0] prefix (finally the key of the k1-th value), [1] remaining rank k, [2] total finite count m
constexpr int kSelState = 8;

// Grid barrier of the cooperative kernels.  Returns false when it gave up: a CTA of this grid left early or the error
// word was raised by somebody (a peer-memory exchange that timed out); every spin loop of the global steps watches the
// same word, so one failure ends all of them instead of leaving CTAs spinning.
__device__ __forceinline__ bool grid_barrier_err(unsigned int* bar, unsigned int nblocks, unsigned int& phase,
                                                 unsigned long long* err)
{
    __shared__ int ok_sh;
    __syncthreads();
    phase++;                                   // every thread keeps the same phase count
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        int ok = 1;
        const long long t0 = clock64();
        while (atomicAdd(bar, 0u) < phase * nblocks) {
            if (*reinterpret_cast<volatile unsigned long long*>(err) != 0ull) { ok = 0; break; }
            if (clock64() - t0 > kSpinLimit) { ok = 0; atomicExch(err, 1ull); break; }
        }
        __threadfence();
        ok_sh = ok;
    }
    __syncthreads();
    return ok_sh != 0;
}

// All threads of the CTA take part.  vals: len <= 2048 words (any address space); entries with index >= min_from are
// combined with min, the others with +.  Returns false (and raises *pp.err) if a peer never answered.
__device__ __forceinline__ bool sel_exchange(const SelP2P& pp, unsigned long long seq, int c, unsigned long long* vals,
                                             int len, int min_from)
{
    __shared__ int timed_out;
    const int nr = pp.nranks;
    const size_t nslots = (size_t)2 * nr * kSelP2PMaxCols;
    const size_t par_base = (size_t)(seq & 1ull) * nr;
    const size_t my_slot = (par_base + pp.rank) * kSelP2PMaxCols + c;
    if (threadIdx.x == 0) timed_out = (*reinterpret_cast<volatile unsigned long long*>(pp.err) != 0ull) ? 1 : 0;
    __syncthreads();
    if (timed_out) return false;                  // an earlier exchange already failed: do not wait again
    // The counters were accumulated by other SMs' atomics in L2: read them there (ld.cg), never through this SM's L1,
    // once, all loads of a thread in flight together -- and then store them to every peer.  (A volatile load per peer and
    // word, as in the first version, is one serialised L2 round trip each: ~40 us per exchange at 8 ranks.)
    {
        const bool in_global = __isGlobal(vals);      // (the two-word exchange of the last phase hands over shared memory)
        unsigned long long mine[kSelBins / kSelThreads];
#pragma unroll
        for (int k = 0; k < kSelBins / kSelThreads; k++) {
            const int b = threadIdx.x + k * kSelThreads;
            mine[k] = (b >= len) ? 0ull : (in_global ? __ldcg(vals + b) : *reinterpret_cast<volatile unsigned long long*>(vals + b));
        }
        for (int r = 0; r < nr; r++) {
            unsigned long long* dst = pp.peers[r] + my_slot * kSelBins;
#pragma unroll
            for (int k = 0; k < kSelBins / kSelThreads; k++) {
                const int b = threadIdx.x + k * kSelThreads;
                if (b < len) dst[b] = mine[k];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < nr) {
        *reinterpret_cast<volatile unsigned long long*>(pp.peers[threadIdx.x] + nslots * kSelBins + my_slot) = seq;
        volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
            pp.mymail + nslots * kSelBins + (par_base + threadIdx.x) * kSelP2PMaxCols + c);
        const long long t0 = clock64();
        while (*f != seq) {
            if (*reinterpret_cast<volatile unsigned long long*>(pp.err) != 0ull) { timed_out = 1; break; }
            if (clock64() - t0 > kSpinLimit) { timed_out = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    // the peers' slots arrived through NVLink in this GPU's L2 and are final once their sequence words are seen: plain
    // L2 loads, all ranks' words of a bin in flight together
    for (int b = threadIdx.x; b < len; b += blockDim.x) {
        const bool is_min = b >= min_from;
        const unsigned long long neutral = is_min ? ~0ull : 0ull;
        unsigned long long x[kSelMaxRanks];
#pragma unroll
        for (int r = 0; r < kSelMaxRanks; r++)
            x[r] = (r < nr) ? __ldcg(pp.mymail + ((par_base + r) * kSelP2PMaxCols + c) * kSelBins + b) : neutral;
        unsigned long long acc = neutral;
#pragma unroll
        for (int r = 0; r < kSelMaxRanks; r++) acc = is_min ? (x[r] < acc ? x[r] : acc) : acc + x[r];
        vals[b] = acc;
    }
    __syncthreads();
    const bool ok = timed_out == 0;
    if (!ok && threadIdx.x == 0) atomicExch(pp.err, 1ull);
    return ok;
}

// aux: counts[B], le[B], mg[B].  counts, le and hist arrive zeroed; mg is set here.
__global__ void __launch_bounds__(kSelThreads)
sel_fused_kernel(int64_t n, int B, const double* __restrict__ base, int64_t stride, const double* __restrict__ center,
                 double* __restrict__ out, int do_exp, double scale, unsigned long long* __restrict__ state,
                 unsigned long long* __restrict__ hist, unsigned long long* __restrict__ aux, unsigned int* bar, SelP2P pp)
{
    __shared__ unsigned int sh[kSelBins];
    __shared__ unsigned long long part[kSelThreads];
    __shared__ unsigned long long lm[2];
    unsigned long long* counts = aux;
    unsigned long long* le = aux + B;
    unsigned long long* mg = aux + 2 * B;
    unsigned int phase = 0;
    unsigned long long seq = pp.seq;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31u;

    // ---- finite values per column ----
    for (int c = 0; c < B; c++) {
        unsigned long long local = 0;
        sel_for_each(base, stride, c, n, tid, nthr, center, [&](double) { local++; });
        for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
        if (lane == 0 && local) atomicAdd(counts + c, local);
    }
    if (!grid_barrier_err(bar, gridDim.x, phase, pp.err)) return;
    for (int c = blockIdx.x; c < B; c += gridDim.x) {
        bool ok = true;
        if (pp.nranks > 1) ok = sel_exchange(pp, seq, c, counts + c, 1, 1);
        if (threadIdx.x == 0) {
            const unsigned long long m = ok ? *reinterpret_cast<volatile unsigned long long*>(counts + c) : 0ull;
            unsigned long long* s = state + (size_t)c * kSelState;
            s[0] = 0ull; s[1] = (m > 0) ? (m - 1) / 2 : 0ull; s[2] = m;
            mg[c] = ~0ull;
        }
    }
    seq++;
    if (!grid_barrier_err(bar, gridDim.x, phase, pp.err)) return;

    // ---- six digits ----
    for (int pass = 0; pass < 6; pass++) {
        const int bits = (pass == 5) ? 9 : 11;
        const int shift = (pass == 5) ? 0 : 53 - 11 * pass;
        const int hi_shift = shift + bits;                       // bits above the current digit must equal the prefix
        const unsigned long long mask = (1ull << bits) - 1ull;
        for (int c = 0; c < B; c++) {
            for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) sh[b] = 0u;
            __syncthreads();
            const unsigned long long prefix = __ldcg(state + (size_t)c * kSelState);
            sel_for_each(base, stride, c, n, tid, nthr, center, [&](double v) {
                const unsigned long long k = key_of(v);
                const bool match = (hi_shift >= 64) ? true : ((k >> hi_shift) == prefix);
                if (match) atomicAdd(&sh[(unsigned)((k >> shift) & mask)], 1u);
            });
            __syncthreads();
            for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
                if (sh[b]) atomicAdd(hist + (size_t)c * kSelBins + b, (unsigned long long)sh[b]);
            __syncthreads();
        }
        if (!grid_barrier_err(bar, gridDim.x, phase, pp.err)) return;
        // extend the prefix by the digit whose cumulative count first exceeds the remaining rank
        for (int c = blockIdx.x; c < B; c += gridDim.x) {
            unsigned long long* h = hist + (size_t)c * kSelBins;
            if (pp.nranks > 1 && !sel_exchange(pp, seq, c, h, kSelBins, kSelBins)) {
                if (threadIdx.x == 0) state[(size_t)c * kSelState + 2] = 0ull;    // median comes out NaN; host sees *pp.err
            }
            const int per = kSelBins / kSelThreads;                  // 8 consecutive bins per thread
            unsigned long long mine = 0;
            for (int j = 0; j < per; j++) mine += __ldcg(h + threadIdx.x * per + j);
            part[threadIdx.x] = mine;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long* s = state + (size_t)c * kSelState;
                unsigned long long k = s[1], cum = 0;
                int t = 0;
                while (t < kSelThreads - 1 && cum + part[t] <= k) { cum += part[t]; t++; }
                int b = t * per;
                while (b < kSelBins - 1 && cum + __ldcg(h + b) <= k) { cum += __ldcg(h + b); b++; }
                s[0] = (s[0] << bits) | (unsigned long long)b;
                s[1] = k - cum;
            }
            __syncthreads();
            for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) h[b] = 0ull;      // ready for the next pass
        }
        seq++;
        if (!grid_barrier_err(bar, gridDim.x, phase, pp.err)) return;
    }

    // ---- state[0] is the key of the k1-th value: count keys <= v1 and find the smallest key > v1 ----
    for (int c = 0; c < B; c++) {
        const unsigned long long v1 = __ldcg(state + (size_t)c * kSelState);
        unsigned long long l = 0, g = ~0ull;
        sel_for_each(base, stride, c, n, tid, nthr, center, [&](double v) {
            const unsigned long long k = key_of(v);
            if (k <= v1) l++; else if (k < g) g = k;
        });
        for (int off = 16; off > 0; off >>= 1) {
            l += __shfl_down_sync(0xffffffffu, l, off);
            const unsigned long long o = __shfl_down_sync(0xffffffffu, g, off);
            g = o < g ? o : g;
        }
        if (lane == 0) {
            if (l) atomicAdd(le + c, l);
            if (g != ~0ull) atomicMin(mg + c, g);
        }
    }
    if (!grid_barrier_err(bar, gridDim.x, phase, pp.err)) return;
    // median = v1 (odd count) or (v1 + v2) / 2 ; out[c] = exp(scale * median) or scale * median ; NaN if empty
    for (int c = blockIdx.x; c < B; c += gridDim.x) {
        if (threadIdx.x == 0) { lm[0] = __ldcg(le + c); lm[1] = __ldcg(mg + c); }
        __syncthreads();
        bool ok = true;
        if (pp.nranks > 1) ok = sel_exchange(pp, seq, c, lm, 2, 1);
        if (threadIdx.x == 0) {
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
        __syncthreads();
    }
}

cudaError_t sel_launch_fused(int64_t n, int B, const double* base, int64_t stride, const double* center, double* out,
                             int do_exp, double scale, unsigned long long* state, unsigned long long* hist,
                             unsigned long long* aux, unsigned int* bar, const SelP2P& pp_in, cudaStream_t st)
{
    if (B < 1 || B > kSelP2PMaxCols) return cudaErrorInvalidValue;
    SelP2P pp = pp_in;
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (size_t)B * kSelBins, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(aux, 0, sizeof(unsigned long long) * (size_t)B * 3, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(bar, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sel_fused_kernel, kSelThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (n + kSelThreads * 8 - 1) / (kSelThreads * 8);
    if (blocks < B) blocks = B;
    if (blocks > (int64_t)sms * per_sm) blocks = (int64_t)sms * per_sm;
    if (blocks < 1) blocks = 1;
    void* args[] = {(void*)&n, (void*)&B, (void*)&base, (void*)&stride, (void*)&center, (void*)&out, (void*)&do_exp, (void*)&scale,
                    (void*)&state, (void*)&hist, (void*)&aux, (void*)&bar, (void*)&pp};
    return cudaLaunchCooperativeKernel((const void*)sel_fused_kernel, dim3((unsigned)blocks), dim3(kSelThreads), args, 0, st);
}

}  // namespace cd
