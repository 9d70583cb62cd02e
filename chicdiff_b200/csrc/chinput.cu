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

// parses one number ("12", "-4500", "12.0", "1e5", "1.25E+3") or "NA" starting at p; returns false at end of line.
// The value comes back as a double (exact for every integer a .chinput file holds); malformed text reads as NA.
__device__ __forceinline__ bool ch_field(const char* __restrict__ text, int64_t& p, int64_t end, double& val, bool& na)
{
    while (p < end && ch_is_sep(text[p])) p++;
    if (p >= end || text[p] == '\n') return false;
    na = false;
    val = 0.0;
    if (text[p] == 'N' || text[p] == 'n') {                       // NA / NaN
        na = true;
        while (p < end && !ch_is_sep(text[p]) && text[p] != '\n') p++;
        return true;
    }
    bool neg = false;
    if (text[p] == '-') { neg = true; p++; } else if (text[p] == '+') p++;
    double v = 0.0;
    int digits = 0, frac = 0;
    while (p < end && text[p] >= '0' && text[p] <= '9') { v = v * 10.0 + (double)(text[p] - '0'); p++; digits++; }
    if (p < end && text[p] == '.') {
        p++;
        while (p < end && text[p] >= '0' && text[p] <= '9') { v = v * 10.0 + (double)(text[p] - '0'); p++; digits++; frac++; }
    }
    int ex = 0;
    bool bad = digits == 0;
    if (p < end && (text[p] == 'e' || text[p] == 'E')) {
        p++;
        bool eneg = false;
        if (p < end && (text[p] == '-' || text[p] == '+')) { eneg = text[p] == '-'; p++; }
        int ed = 0;
        while (p < end && text[p] >= '0' && text[p] <= '9') { if (ex < 1000) ex = ex * 10 + (text[p] - '0'); p++; ed++; }
        if (ed == 0) bad = true;
        if (eneg) ex = -ex;
    }
    // anything left before the separator is not a number
    if (p < end && !ch_is_sep(text[p]) && text[p] != '\n') {
        bad = true;
        while (p < end && !ch_is_sep(text[p]) && text[p] != '\n') p++;
    }
    if (bad) { na = true; return true; }
    int e10 = ex - frac;
    double scale = 1.0, base = 10.0;
    for (int k = e10 < 0 ? -e10 : e10; k > 0; k >>= 1) { if (k & 1) scale *= base; base *= base; }
    v = (e10 < 0) ? v / scale : v * scale;
    val = neg ? -v : v;
    return true;
}

// an integer column: the value must be a whole number that fits int32
__device__ __forceinline__ bool ch_as_int(double v, bool na, int32_t& out)
{
    if (na || !(v >= -2147483647.0 && v <= 2147483647.0) || v != floor(v)) return false;
    out = (int32_t)v;
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
    double v[5] = {0, 0, 0, 0, 0};
    bool na[5] = {true, true, true, true, true};
    if (ok) {
        for (int f = 0; f < 5; f++) {
            if (!ch_field(text, p, end, v[f], na[f])) { if (f < 3) ok = false; break; }
        }
    }
    // a row whose baitID / otherEndID / N is not a whole number is not a .chinput row: dropped, never truncated
    int32_t ib = 0, io = 0, iN = 0, il = INT32_MIN;
    ok = ok && ch_as_int(v[0], na[0], ib) && ch_as_int(v[1], na[1], io) && ch_as_int(v[2], na[2], iN);
    if (!ch_as_int(v[3], na[3], il)) il = INT32_MIN;
    valid[l] = ok ? 1 : 0;
    bait[l] = ib; oe[l] = io; N[l] = iN;
    oelen[l] = il;
    dist[l] = na[4] ? NAN : v[4];
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
