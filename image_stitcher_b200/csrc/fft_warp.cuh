// Warp-level 1024-point complex FFT held in registers (sm_100a).
//
// The column pass of the registration chain transforms lines of 1024 complex values (the strip's long axis for 2048^2
// tiles).  A warp owns a line: lane l holds x[l + 32 j], j = 0..31, so global loads and stores are whole 256-byte
// warp accesses.  1024 = 32 x 32 (four-step FFT):
//
//     1. each lane runs a 32-point FFT over its registers            Y[l][k2] = sum_j x[l + 32 j] W32^(j k2)
//     2. twiddle                                                     Y[l][k2] *= W1024^(l k2)      (table [k2][l])
//     3. 32 x 32 transpose through shared memory (pitch 33 words: conflict-free both ways)
//     4. a second 32-point FFT over the lane index                   X[k2 + 32 k1] = sum_l Y[l][k2] W32^(l k1)
//
// One shared-memory round trip and one __syncwarp per line instead of five block-wide passes with barriers.  The 32-point
// FFT is a fully unrolled radix-2 decimation-in-frequency network; its output is left in bit-reversed REGISTER order
// (slot s holds frequency brev5(s)), which costs nothing because every register index is a compile-time constant.
#pragma once

#include <cuda_runtime.h>

namespace wfft {

__host__ __device__ constexpr int brev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// cos(2 pi m / 32), sin(2 pi m / 32), m = 0..15
__device__ __forceinline__ float w32c(int m) {
    constexpr float t[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                             0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                             -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    return t[m];
}
__device__ __forceinline__ float w32s(int m) {
    constexpr float t[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                             1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
    return t[m];
}

// In-register 32-point FFT, radix-2 decimation in frequency.  Input x[j] in natural order; output frequency k in x[brev5(k)].
// INV = conjugated twiddles (unnormalised inverse).
template <bool INV>
__device__ __forceinline__ void fft32(float2 (&x)[32]) {
#pragma unroll
    for (int span = 16; span >= 1; span >>= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i & span) continue;
            const int m = (i & (span - 1)) * (16 / span);        // twiddle W32^m, m in [0, 16)
            const float2 a = x[i], b = x[i + span];
            x[i] = make_float2(a.x + b.x, a.y + b.y);
            const float dx = a.x - b.x, dy = a.y - b.y;
            // forward: d * (c - i s); inverse: d * (c + i s)
            if (m == 0) {
                x[i + span] = make_float2(dx, dy);
            } else if (m == 8) {                                 // -i (forward) / +i (inverse)
                x[i + span] = INV ? make_float2(-dy, dx) : make_float2(dy, -dx);
            } else {
                const float c = w32c(m), s = INV ? -w32s(m) : w32s(m);
                x[i + span] = make_float2(fmaf(dx, c, dy * s), fmaf(dy, c, -dx * s));
            }
        }
    }
}

// Shared memory of one warp: the transpose buffers (re / im planes, pitch 33)
struct WarpBuf {
    float re[32 * 33];
    float im[32 * 33];
};

// 1024-point FFT of the line a warp holds as x[j] = line[lane + 32 j].  `tw` = table [k2][l] of W1024^(l k2) (forward
// sign; conjugated here for INV).  Result: frequency (lane + 32 k1) in x[brev5(k1)].
template <bool INV>
__device__ __forceinline__ void fft1024(float2 (&x)[32], WarpBuf& buf, const float2* __restrict__ tw, int lane) {
    fft32<INV>(x);
    __syncwarp();                                               // the buffer may still be read by the previous transform
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        const float2 v = x[brev5(k2)];
        float2 w = tw[k2 * 32 + lane];
        if (INV) w.y = -w.y;
        buf.re[k2 * 33 + lane] = v.x * w.x - v.y * w.y;
        buf.im[k2 * 33 + lane] = v.x * w.y + v.y * w.x;
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 32; ++l) x[l] = make_float2(buf.re[lane * 33 + l], buf.im[lane * 33 + l]);
    fft32<INV>(x);
}

}  // namespace wfft
