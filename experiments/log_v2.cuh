// log_v2.cuh -- EXPERIMENT (see experiments/README.md): a table-assisted natural logarithm for the line-search kernel.
//
// log_pos (common.cuh, fdlibm scheme) costs ~10 integer + ~25 FP64 instructions, 7 of them the Newton reciprocal for
// s = f / (2 + f); it runs 2.5-3 times per sample and is ~60 % of the kernel's executed instructions.  Here the
// mantissa is reduced multiplicatively with a 128-entry table instead: x = 2^k m, m rc_i = 1 + r with |r| <= 2^-7,
// log x = k ln2 - log rc_i + log1p(r), log1p by a degree-8 Taylor polynomial.  No reciprocal, 13 FP64 instructions.
// rc_i has 20 significant bits, so fma(m, rc_i, -1) is exact up to its single rounding.  The first interval uses
// rc = 1 (r = m - 1) and the last one is moved to the next binade (r = m/2 - 1), so the result keeps full relative
// accuracy around x = 1, which log(1 + mu alpha) at small alpha needs.
// The table (2 KB) is meant to live in shared memory (one copy per CTA): the lanes of a warp index it with unrelated
// mantissas, which constant memory would serialise.
#pragma once
#include "common.cuh"

namespace cd {

static __constant__ double kLogTab[256] = {
#include "log_table.inc"
};

// x positive and normal; tab = 128 x {rc, -log rc}
__device__ __forceinline__ double log_pos_v2(double x, const double* tab)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int idx = hi >> 13;                         // top 7 bits of the mantissa
    const int up = (hi + 0x2000) & 0x100000;          // set iff idx == 127: treat [2 - 1/64, 2) as [1 - 1/128, 1) of the next binade
    hi |= (up ^ 0x3ff00000);
    k += (up >> 20);
    const double m = __hiloint2double(hi, lo);
    const double rc = tab[2 * idx], lc = tab[2 * idx + 1];
    const double r = fma(m, rc, -1.0);
    double q = fma(r, -1.0 / 8.0, 1.0 / 7.0);
    q = fma(r, q, -1.0 / 6.0);
    q = fma(r, q, 1.0 / 5.0);
    q = fma(r, q, -1.0 / 4.0);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    const double dk = (double)k;
    const double t = fma(dk, kLogC[0], lc);           // k ln2_hi is exact (ln2_hi has 32 trailing zero bits)
    const double u = fma(r * r, q, dk * kLogC[1]);
    return t + (r + u);
}

}  // namespace cd
