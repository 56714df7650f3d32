// Shared-memory mixed-radix Stockham FFT for arbitrary line lengths (sm_100a).
//
// The strip extents of the registration path are whatever the acquisition geometry makes them
// (1024 x 214 for 2048^2 tiles at 10 % overlap, 1500 x 314 for 3000^2: 214 = 2*107, 314 = 2*157),
// and parity with the reference needs the circular correlation at EXACTLY those sizes, so the
// engine takes any factorisation: every pass is a radix-R Stockham step (autosort, no bit
// reversal) where R may be a large prime.  A pass computes, for j in [0, N/R), q in [0, R):
//
//     out[(j / Ns) * Ns * R + (j % Ns) + q * Ns] = sum_r in[j + r * N/R] * W_N^{ r * ((j % Ns) * N/(Ns R) + q * N/R) }
//
// (Ns = product of the radices already applied).  All twiddles are N-th roots of unity and come
// from one table W_N[m] = exp(-2 pi i m / N) that the host computes in double precision.
// Work item = (output index, group of LB lines): the twiddle is loaded once and reused for LB
// lines; neighbouring lanes share j, so their input loads are shared-memory broadcasts.
#pragma once

#include <cuda_runtime.h>

struct FftPlan {
    int n;
    int nfac;
    int fac[16];
};

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T2, typename T>
__device__ __forceinline__ T2 mk2(T x, T y) {
    T2 r;
    r.x = x;
    r.y = y;
    return r;
}

// acc += a * w
template <typename T2>
__device__ __forceinline__ void cfma(T2& acc, const T2 a, const T2 w) {
    acc.x = fma(a.x, w.x, acc.x);
    acc.x = fma(-a.y, w.y, acc.x);
    acc.y = fma(a.x, w.y, acc.y);
    acc.y = fma(a.y, w.x, acc.y);
}

// In-place (ping-pong) FFT of `nlines` lines of length plan.n held in shared memory, line stride
// plan.n.  nlines must be a multiple of LB.  Returns the buffer that holds the result.  All
// threads of the block must call; ends with a __syncthreads().
template <typename T2, int LB>
__device__ T2* fft_lines(T2* buf0, T2* buf1, const T2* __restrict__ tw, const FftPlan& plan, int nlines, bool inverse) {
    const int N = plan.n;
    T2* a = buf0;
    T2* b = buf1;
    int Ns = 1;
    const int groups = nlines / LB;
    for (int f = 0; f < plan.nfac; ++f) {
        const int R = plan.fac[f];
        const int M = N / R;
        const int tstep = N / (Ns * R);
        const int items = N * groups;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int g = it / N;
            const int o = it - g * N;
            const int j = o / R, q = o - j * R;          // q fastest: lanes share j -> broadcast input loads
            const int k = j % Ns;
            const int step = (int)(((long long)k * tstep + (long long)q * M) % N);
            T2 acc[LB];
#pragma unroll
            for (int l = 0; l < LB; ++l) acc[l].x = acc[l].y = 0;
            const T2* in = a + (size_t)g * LB * N + j;
            int idx = 0;
            for (int r = 0; r < R; ++r) {
                T2 w = tw[idx];
                if (inverse) w.y = -w.y;
                idx += step;
                if (idx >= N) idx -= N;
#pragma unroll
                for (int l = 0; l < LB; ++l) cfma(acc[l], in[(size_t)l * N + (size_t)r * M], w);
            }
            const int dst = (j / Ns) * Ns * R + k + q * Ns;
#pragma unroll
            for (int l = 0; l < LB; ++l) b[(size_t)(g * LB + l) * N + dst] = acc[l];
        }
        __syncthreads();
        T2* t = a;
        a = b;
        b = t;
        Ns *= R;
    }
    return a;
}
