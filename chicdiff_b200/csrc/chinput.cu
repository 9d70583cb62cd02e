// chinput.cu -- .chinput text codec on the device (the reference reads these files with data.table::fread,
// chicdiff.R:828 and :1272): tab/space separated rows `baitID otherEndID N otherEndLen distSign` after an optional
// '#' comment line and a header line; distSign may be NA (trans pairs).
// The file bytes are copied to the device once; one kernel flags line starts, CUB compacts their offsets, one
// lane per line parses its five fields.  Lines that do not start with a digit (comment, header, blank) are
// marked invalid and dropped by a second compaction.  Pure byte work: bound by HBM and the H2D copy.
#include "kernels.h"
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

namespace cd {

__global__ void __launch_bounds__(256)
ch_line_flags_kernel(int64_t nbytes, const char* __restrict__ text, uint8_t* __restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbytes) return;
    flag[i] = (i == 0 || text[i - 1] == '\n') ? 1 : 0;
}

__device__ __forceinline__ bool ch_is_sep(char c) { return c == '\t' || c == ' ' || c == ',' || c == '\r'; }

// parses one optionally signed integer or "NA" starting at p; returns false at end of line
__device__ __forceinline__ bool ch_field(const char* __restrict__ text, int64_t& p, int64_t end, long long& val, bool& na)
{
    while (p < end && ch_is_sep(text[p])) p++;
    if (p >= end || text[p] == '\n') return false;
    na = false;
    if (text[p] == 'N' || text[p] == 'n') {                       // NA / NaN
        na = true;
        while (p < end && !ch_is_sep(text[p]) && text[p] != '\n') p++;
        val = 0;
        return true;
    }
    bool neg = false;
    if (text[p] == '-') { neg = true; p++; } else if (text[p] == '+') p++;
    long long v = 0;
    bool any = false;
    while (p < end && text[p] >= '0' && text[p] <= '9') { v = v * 10 + (text[p] - '0'); p++; any = true; }
    // tolerate a fractional part / exponent written by other tools ("12.0"): skip to the separator
    while (p < end && !ch_is_sep(text[p]) && text[p] != '\n') p++;
    if (!any) { na = true; v = 0; }
    val = neg ? -v : v;
    return true;
}

__global__ void __launch_bounds__(256)
ch_parse_kernel(int64_t nlines, int64_t nbytes, const char* __restrict__ text, const int64_t* __restrict__ line_start,
                int32_t* __restrict__ bait, int32_t* __restrict__ oe, int32_t* __restrict__ N, int32_t* __restrict__ oelen,
                double* __restrict__ dist, uint8_t* __restrict__ valid)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    int64_t p = line_start[l];
    const int64_t end = (l + 1 < nlines) ? line_start[l + 1] : nbytes;
    const char c0 = (p < end) ? text[p] : '\n';
    bool ok = (c0 >= '0' && c0 <= '9');
    long long v[5] = {0, 0, 0, 0, 0};
    bool na[5] = {true, true, true, true, true};
    if (ok) {
        for (int f = 0; f < 5; f++) {
            if (!ch_field(text, p, end, v[f], na[f])) { if (f < 3) ok = false; break; }
        }
        ok = ok && !na[0] && !na[1] && !na[2];
    }
    valid[l] = ok ? 1 : 0;
    bait[l] = (int32_t)v[0]; oe[l] = (int32_t)v[1]; N[l] = (int32_t)v[2];
    oelen[l] = na[3] ? INT32_MIN : (int32_t)v[3];
    dist[l] = na[4] ? NAN : (double)v[4];
}

template <typename T>
static cudaError_t ch_compact(void* tmp, size_t& bytes, const T* in, const uint8_t* flags, T* out, int64_t* n_out, int64_t n,
                              cudaStream_t st)
{
    return cub::DeviceSelect::Flagged(tmp, bytes, in, flags, out, n_out, (int)n, st);
}

cudaError_t ch_launch_line_flags(int64_t nbytes, const char* text, uint8_t* flag, cudaStream_t st)
{
    if (nbytes > 0) ch_line_flags_kernel<<<(unsigned)((nbytes + 255) / 256), 256, 0, st>>>(nbytes, text, flag);
    return cudaGetLastError();
}

cudaError_t ch_line_starts(void* tmp, size_t& bytes, const uint8_t* flag, int64_t* starts, int64_t* n_out, int64_t nbytes,
                           cudaStream_t st)
{
    cub::CountingInputIterator<int64_t> it(0);
    return cub::DeviceSelect::Flagged(tmp, bytes, it, flag, starts, n_out, (int)nbytes, st);
}

cudaError_t ch_launch_parse(int64_t nlines, int64_t nbytes, const char* text, const int64_t* line_start, int32_t* bait,
                            int32_t* oe, int32_t* N, int32_t* oelen, double* dist, uint8_t* valid, cudaStream_t st)
{
    if (nlines > 0)
        ch_parse_kernel<<<(unsigned)((nlines + 255) / 256), 256, 0, st>>>(nlines, nbytes, text, line_start, bait, oe, N, oelen, dist, valid);
    return cudaGetLastError();
}

cudaError_t ch_compact_i32(void* tmp, size_t& bytes, const int32_t* in, const uint8_t* flags, int32_t* out, int64_t* n_out,
                           int64_t n, cudaStream_t st) { return ch_compact<int32_t>(tmp, bytes, in, flags, out, n_out, n, st); }
cudaError_t ch_compact_f64(void* tmp, size_t& bytes, const double* in, const uint8_t* flags, double* out, int64_t* n_out,
                           int64_t n, cudaStream_t st) { return ch_compact<double>(tmp, bytes, in, flags, out, n_out, n, st); }

}  // namespace cd
