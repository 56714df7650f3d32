// Definitions shared by the registration kernels of reg.cu (radix FFT engine, float32 / float64) and reg_tc.cu
// (tensor-core short-axis transforms + warp-level column FFT, float32).
#pragma once

#include "sb_common.cuh"

namespace {

constexpr double kInScale = 1.0 / 65536.0;                       // exact power-of-two input scaling
constexpr double kClamp = 100.0 * 2.220446049250313e-16 * kInScale * kInScale;   // 100*eps64, same scaling as P

struct PairDesc {            // device-side description of one pair of a batch
    const uint16_t* a;       // first pixel of the reference strip
    const uint16_t* b;       // first pixel of the moving strip
    int32_t a_tile, b_tile;  // indices into the min/max table
};

struct CtaBest {             // block-local first maximum and the second-largest value; double keeps the float64 path's resolution
    double val;
    double second;           // largest |cc| of the block at any OTHER pixel (may equal val)
    int32_t idx;
    int32_t pad;
};

struct PeakOut {             // per pair, written by the device, read back by the host
    int32_t coarse_y, coarse_x;
    int32_t fine_y, fine_x;
    float peak, second, runner_up;     // see sb_pair_result
    float fine_peak, fine_second;
    float skew;                        // max / min of the two strips' sums when they shared one packed transform, else 1
};

// normalize_image (:844-855): ((v - min) / (max - min)) * 65535 in float64, truncating cast.
// `maxval` = iinfo(dtype).max of the caller's pixels: 65535, or 255 for (widened) uint8 tiles (:854).
__device__ __forceinline__ int stretch_px_f64(unsigned v, int mn, int mx, int maxval) {
    if (mx <= mn) return 0;                       // 0/0 -> NaN -> undefined cast in the reference; defined as 0
    const double q = __ddiv_rn((double)((int)v - mn), (double)(mx - mn));
    return (int)(q * (double)maxval);
}

// Branch-free core of stretch_px for the general case (0 < max - min < maxval): returns k = trunc(a * maxval / b) from the
// float estimate corrected by the exact integer remainder, and raises `exact` when b divides a * maxval (k != 0) -- the one
// case where the float64 expression of the reference may land on either side of the integer and must be evaluated as
// such (stretch_px_f64).  No conversion-unit instructions: uint -> float and float -> int go through the 2^23 magic number.
__device__ __forceinline__ int stretch_core(unsigned a, unsigned b, float inv, unsigned maxval, bool& exact) {
    const float af = __uint_as_float(0x4B000000u | a) - 8388608.0f;                           // a < 2^16: exact
    unsigned k = __float_as_uint(__fadd_rz(af * inv, 8388608.0f)) - 0x4B000000u;                // trunc(af * inv), < 2^17
    int rem = (int)(a * maxval - k * b);                                                        // k is off by at most one
    const bool lt = rem < 0, ge = rem >= (int)b;
    k = lt ? k - 1u : (ge ? k + 1u : k);
    rem = lt ? rem + (int)b : (ge ? rem - (int)b : rem);
    exact = exact || (rem == 0 && k != 0u);
    return (int)k;
}

// The form the tensor-core converters run (reg_tc.cu, 8 warps x 14 chunks per tile: instruction count is what bounds them).
// The float estimate is biased DOWNWARD -- inv_lo = maxval / b * (1 - 1.5e-6), all rounding errors together stay below
// 2e-7 relative -- so that trunc(a * inv_lo) is k or k - 1 (never above: q * 1.7e-6 < 0.12 for q <= 65535) and ONE compare
// of the exact integer remainder repairs it.  Returns kStretchMagic + k (the float 2^23 bit pattern carrying k in its
// mantissa: sums and differences of two of them cancel the constant) and raises `exact` as stretch_core does.
// magic_b = kStretchMagic * b (mod 2^32), inv_lo = stretch_inv_lo(b, maxval): once per tile.
constexpr unsigned kStretchMagic = 0x4B000000u;
__device__ __forceinline__ float stretch_inv_lo(unsigned b, unsigned maxval) {
    return __fmul_rn(__fdiv_rn((float)maxval, (float)b), 0.9999985f);
}
__device__ __forceinline__ unsigned stretch_bits(unsigned a, unsigned b, float inv_lo, unsigned maxval, unsigned magic_b, bool& exact) {
    const float af = __uint_as_float(kStretchMagic | a) - 8388608.0f;                          // a < 2^16: exact
    unsigned bits = __float_as_uint(__fadd_rz(__fmul_rn(af, inv_lo), 8388608.0f));             // kStretchMagic + k - {0, 1}
    unsigned rem = a * maxval + magic_b - bits * b;                                            // in [0, 2 b)
    const bool ge = rem >= b;
    bits += ge ? 1u : 0u;
    rem -= ge ? b : 0u;
    exact = exact || (rem == 0u && a != 0u);
    return bits;
}

// The integer form the tensor-core converters run since r2 call 23: k = floor(a * maxval / b) = (a * M) >> 32 with the
// 48-bit magic M = ceil(2^32 * maxval / b).  Exact: a * M / 2^32 exceeds the true quotient by less than a * 2^-32 <= 2^-16,
// and a non-integer a * maxval / b lies at least 1 / b > 2^-16 below the next integer (b <= 65535).  Two instructions
// (IMAD.HI on the low word + IMAD on the high word); `exact` as above from the remainder.
struct StretchMagic { unsigned lo, hi; };
__device__ __forceinline__ StretchMagic stretch_magic(unsigned b, unsigned maxval) {          // once per tile (b >= 1)
    const unsigned long long num = (unsigned long long)maxval << 32;
    const unsigned long long m = (num + b - 1) / b;
    return {(unsigned)m, (unsigned)(m >> 32)};
}
__device__ __forceinline__ unsigned stretch_mulhi(unsigned a, unsigned b, StretchMagic m, unsigned maxval, bool& exact) {
    const unsigned k = a * m.hi + __umulhi(a, m.lo);
    exact = exact || (a * maxval - k * b == 0u && a != 0u);
    return k;
}

// The same value without the float64 divide.  With a = v - min, b = max - min the exact quotient a * 65535 / b is
// rational with denominator b <= 65535, so unless it is an integer it lies at least 1 / 65535 away from the next
// one, while the float64 evaluation is off by at most 65535 * 2^-52: trunc() of both agree.  When b divides
// a * 65535 the float64 result can land on either side of the integer -- only then the float64 sequence is run.
// `inv` = maxval / b as float (computed once per tile).
__device__ __forceinline__ int stretch_px(unsigned v, int mn, int mx, float inv, int maxval) {
    if (mx <= mn) return 0;
    const unsigned b = (unsigned)(mx - mn);
    // full-range tile (saturated pixels next to a zero): every quotient is exact, and float64 gives (a / maxval) * maxval
    // == a for all a <= maxval (checked exhaustively for 255 and 65535) -- skip the float64 sequence for the whole tile
    if (b == (unsigned)maxval) return (int)v - mn;
    const unsigned num = (unsigned)((int)v - mn) * (unsigned)maxval;       // < 2^32
    unsigned k = (unsigned)__float2int_rz(__uint2float_rn((unsigned)((int)v - mn)) * inv);
    unsigned rem = num - k * b;                                            // k is off by at most one either way
    if ((int)rem < 0) { --k; rem += b; }
    else if (rem >= b) { ++k; rem -= b; }
    if (rem == 0 && k != 0) return stretch_px_f64(v, mn, mx, maxval);      // exact quotient (rare): reproduce float64 rounding
    return (int)k;
}

template <typename V>
__device__ __forceinline__ void best_update(V& bv, int& bi, V v, int i) {
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}
// first maximum (ties -> lowest index) plus the largest value at any other position
template <typename V>
__device__ __forceinline__ void top2_update(V& bv, int& bi, V& b2, V v, int i) {
    if (v > bv || (v == bv && i < bi)) { b2 = bv; bv = v; bi = i; }
    else if (v > b2) b2 = v;
}
template <typename V>
__device__ __forceinline__ void top2_merge(V& bv, int& bi, V& b2, V ov, int oi, V o2) {
    if (ov > bv || (ov == bv && oi < bi)) { b2 = bv > o2 ? bv : o2; bv = ov; bi = oi; }
    else if (ov > b2) b2 = ov;
}


}  // namespace
