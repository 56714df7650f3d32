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
//
// The passes are shared-memory-bandwidth bound (ncu, profiles/r1_registration.md), so they are
// organised to reuse every shared-memory load:
//   * radix 4 / radix 2: one thread per butterfly and group of LB lines -- 3 (1) twiddle loads are
//     shared by the LB lines, each input is read once, each output written once;
//   * any other radix (large primes included): a thread owns a QT x LB tile (QT outputs q of one
//     butterfly, LB lines): per term r it loads QT twiddles and LB inputs for QT * LB complex FMAs.
#pragma once

#include <cuda_runtime.h>

struct FftPlan {
    int n;
    int nfac;
    int fac[16];
    int gemm_radix;   // odd radix handled by pass_odd_gemm (needs its cos/sin table in shared memory); 0 = none
    // per-pass constants, filled by the host (make_plan): Ns = product of the radices already applied, M = n / radix,
    // tstep = n / (Ns * radix), and multipliers for division by M and Ns without the integer-divide sequence
    int Ns[16], M[16], tstep[16];
    unsigned mM[16], mNs[16];
};

// n / d for 0 <= n, n * d < 2^32, with m = ceil(2^32 / d) precomputed (m == 0 encodes d == 1)
__host__ __device__ inline unsigned fastdiv_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)(((1ull << 32) + d - 1) / d); }
__device__ __forceinline__ int fastdiv(int n, unsigned m) { return m ? (int)__umulhi((unsigned)n, m) : n; }

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
template <typename T2>
__device__ __forceinline__ T2 cmul(const T2 a, const T2 w) {
    T2 r;
    r.x = a.x * w.x - a.y * w.y;
    r.y = a.x * w.y + a.y * w.x;
    return r;
}

template <typename T2, int LB>
__device__ __forceinline__ void pass_radix4(const T2* __restrict__ a, T2* __restrict__ b, const T2* __restrict__ tw, int N, int Ns,
                                            int M, int tstep, unsigned mM, unsigned mNs, int groups, bool inverse) {
    const int items = M * groups;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int g = fastdiv(it, mM), j = it - g * M;
        const int jq = fastdiv(j, mNs), k = j - jq * Ns;
        T2 w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
        if (inverse) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        const int dst = jq * Ns * 4 + k;
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const T2* in = a + (size_t)(g * LB + l) * N + j;
            const T2 x0 = in[0], x1 = cmul(in[M], w1), x2 = cmul(in[2 * M], w2), x3 = cmul(in[3 * M], w3);
            const T2 s02 = mk2<T2>(x0.x + x2.x, x0.y + x2.y), d02 = mk2<T2>(x0.x - x2.x, x0.y - x2.y);
            const T2 s13 = mk2<T2>(x1.x + x3.x, x1.y + x3.y), d13 = mk2<T2>(x1.x - x3.x, x1.y - x3.y);
            T2* out = b + (size_t)(g * LB + l) * N + dst;
            out[0] = mk2<T2>(s02.x + s13.x, s02.y + s13.y);
            out[2 * Ns] = mk2<T2>(s02.x - s13.x, s02.y - s13.y);
            // forward: q=1 -> d02 - i d13, q=3 -> d02 + i d13 ; inverse: the conjugates
            if (!inverse) {
                out[Ns] = mk2<T2>(d02.x + d13.y, d02.y - d13.x);
                out[3 * Ns] = mk2<T2>(d02.x - d13.y, d02.y + d13.x);
            } else {
                out[Ns] = mk2<T2>(d02.x - d13.y, d02.y + d13.x);
                out[3 * Ns] = mk2<T2>(d02.x + d13.y, d02.y - d13.x);
            }
        }
    }
}

template <typename T2, int LB>
__device__ __forceinline__ void pass_radix2(const T2* __restrict__ a, T2* __restrict__ b, const T2* __restrict__ tw, int N, int Ns,
                                            int M, int tstep, unsigned mM, unsigned mNs, int groups, bool inverse) {
    const int items = M * groups;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int g = fastdiv(it, mM), j = it - g * M;
        const int jq = fastdiv(j, mNs), k = j - jq * Ns;
        T2 w1 = tw[k * tstep];
        if (inverse) w1.y = -w1.y;
        const int dst = jq * Ns * 2 + k;
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const T2* in = a + (size_t)(g * LB + l) * N + j;
            const T2 x0 = in[0], x1 = cmul(in[M], w1);
            T2* out = b + (size_t)(g * LB + l) * N + dst;
            out[0] = mk2<T2>(x0.x + x1.x, x0.y + x1.y);
            out[Ns] = mk2<T2>(x0.x - x1.x, x0.y - x1.y);
        }
    }
}

// Radix 3 and radix 5 as straight butterflies (r2 call 32): 1500 = 5 * 5 * 5 * 4 * 3, the long axis of the strips of
// 3000 x 3000 tiles, went through the generic odd-radix pass -- a 64-bit modulo per twiddle and a fold pass per
// factor -- and cols_xpower_kernel took half of the registration time of BASELINE configs[4].  Same item mapping as
// pass_radix4: one thread per butterfly and group of LB lines; r * k * tstep < N for r < R, no reduction needed.
template <typename T2, int LB>
__device__ __forceinline__ void pass_radix3(const T2* __restrict__ a, T2* __restrict__ b, const T2* __restrict__ tw, int N, int Ns,
                                            int M, int tstep, unsigned mM, unsigned mNs, int groups, bool inverse) {
    using T = decltype(T2().x);
    const T kS = (T)0.86602540378443864676372317075294L;          // sin(2 pi / 3)
    const T sg = inverse ? (T)-1 : (T)1;
    const int items = M * groups;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int g = fastdiv(it, mM), j = it - g * M;
        const int jq = fastdiv(j, mNs), k = j - jq * Ns;
        T2 w1 = tw[k * tstep], w2 = tw[2 * k * tstep];
        if (inverse) { w1.y = -w1.y; w2.y = -w2.y; }
        const int dst = jq * Ns * 3 + k;
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const T2* in = a + (size_t)(g * LB + l) * N + j;
            const T2 x0 = in[0], x1 = cmul(in[M], w1), x2 = cmul(in[2 * M], w2);
            const T2 s = mk2<T2>(x1.x + x2.x, x1.y + x2.y), d = mk2<T2>(x1.x - x2.x, x1.y - x2.y);
            const T2 m = mk2<T2>(x0.x - (T)0.5 * s.x, x0.y - (T)0.5 * s.y);
            const T2 e = mk2<T2>(sg * kS * d.y, sg * kS * d.x);     // forward: X1 = m - i kS d = (m.x + kS d.y, m.y - kS d.x)
            T2* out = b + (size_t)(g * LB + l) * N + dst;
            out[0] = mk2<T2>(x0.x + s.x, x0.y + s.y);
            out[Ns] = mk2<T2>(m.x + e.x, m.y - e.y);
            out[2 * Ns] = mk2<T2>(m.x - e.x, m.y + e.y);
        }
    }
}

template <typename T2, int LB>
__device__ __forceinline__ void pass_radix5(const T2* __restrict__ a, T2* __restrict__ b, const T2* __restrict__ tw, int N, int Ns,
                                            int M, int tstep, unsigned mM, unsigned mNs, int groups, bool inverse) {
    using T = decltype(T2().x);
    const T c1 = (T)0.30901699437494742410229341718282L, c2 = (T)-0.80901699437494742410229341718282L;   // cos(2 pi / 5), cos(4 pi / 5)
    const T s1 = (T)0.95105651629515357211643933337938L, s2 = (T)0.58778525229247312916870595463907L;    // sin(2 pi / 5), sin(4 pi / 5)
    const T sg = inverse ? (T)-1 : (T)1;
    const int items = M * groups;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int g = fastdiv(it, mM), j = it - g * M;
        const int jq = fastdiv(j, mNs), k = j - jq * Ns;
        const int kt = k * tstep;
        T2 w1 = tw[kt], w2 = tw[2 * kt], w3 = tw[3 * kt], w4 = tw[4 * kt];
        if (inverse) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; w4.y = -w4.y; }
        const int dst = jq * Ns * 5 + k;
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const T2* in = a + (size_t)(g * LB + l) * N + j;
            const T2 x0 = in[0], x1 = cmul(in[M], w1), x2 = cmul(in[2 * M], w2), x3 = cmul(in[3 * M], w3), x4 = cmul(in[4 * M], w4);
            const T2 p1 = mk2<T2>(x1.x + x4.x, x1.y + x4.y), d1 = mk2<T2>(x1.x - x4.x, x1.y - x4.y);
            const T2 p2 = mk2<T2>(x2.x + x3.x, x2.y + x3.y), d2 = mk2<T2>(x2.x - x3.x, x2.y - x3.y);
            const T2 a1 = mk2<T2>(fma(c2, p2.x, fma(c1, p1.x, x0.x)), fma(c2, p2.y, fma(c1, p1.y, x0.y)));
            const T2 a2 = mk2<T2>(fma(c1, p2.x, fma(c2, p1.x, x0.x)), fma(c1, p2.y, fma(c2, p1.y, x0.y)));
            // forward: X1 = a1 - i b1, X4 = a1 + i b1, X2 = a2 - i b2, X3 = a2 + i b2; inverse: the conjugate signs
            const T2 b1 = mk2<T2>(sg * fma(s2, d2.x, s1 * d1.x), sg * fma(s2, d2.y, s1 * d1.y));
            const T2 b2 = mk2<T2>(sg * fma(-s1, d2.x, s2 * d1.x), sg * fma(-s1, d2.y, s2 * d1.y));
            T2* out = b + (size_t)(g * LB + l) * N + dst;
            out[0] = mk2<T2>(x0.x + p1.x + p2.x, x0.y + p1.y + p2.y);
            out[Ns] = mk2<T2>(a1.x + b1.y, a1.y - b1.x);
            out[4 * Ns] = mk2<T2>(a1.x - b1.y, a1.y + b1.x);
            out[2 * Ns] = mk2<T2>(a2.x + b2.y, a2.y - b2.x);
            out[3 * Ns] = mk2<T2>(a2.x - b2.y, a2.y + b2.x);
        }
    }
}

template <typename T2, int LB, int QT>
__device__ __forceinline__ void pass_generic(const T2* __restrict__ a, T2* __restrict__ b, const T2* __restrict__ tw, int N, int Ns,
                                             int R, int groups, bool inverse) {
    const int M = N / R;
    const int tstep = N / (Ns * R);
    const int qtiles = (R + QT - 1) / QT;
    const int items = M * qtiles * groups;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int g = it / (M * qtiles);
        const int rem = it - g * (M * qtiles);
        const int j = rem / qtiles, qt = rem - j * qtiles;      // q tile fastest: neighbouring lanes share j -> broadcast inputs
        const int k = j % Ns;
        int step[QT], idx[QT];
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const int q = min(qt * QT + t, R - 1);
            step[t] = (int)(((long long)k * tstep + (long long)q * M) % N);
            idx[t] = 0;
        }
        T2 acc[QT][LB];
#pragma unroll
        for (int t = 0; t < QT; ++t)
#pragma unroll
            for (int l = 0; l < LB; ++l) acc[t][l].x = acc[t][l].y = 0;
        const T2* in = a + (size_t)g * LB * N + j;
        for (int r = 0; r < R; ++r) {
            T2 x[LB];
#pragma unroll
            for (int l = 0; l < LB; ++l) x[l] = in[(size_t)l * N + (size_t)r * M];
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                T2 w = tw[idx[t]];
                if (inverse) w.y = -w.y;
                idx[t] += step[t];
                if (idx[t] >= N) idx[t] -= N;
#pragma unroll
                for (int l = 0; l < LB; ++l) cfma(acc[t][l], x[l], w);
            }
        }
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const int q = qt * QT + t;
            if (q < R) {
                const int dst = (j / Ns) * Ns * R + k + q * Ns;
#pragma unroll
                for (int l = 0; l < LB; ++l) b[(size_t)(g * LB + l) * N + dst] = acc[t][l];
            }
        }
    }
}

// Odd radix R (large primes included) exploiting W^{q(R-r)} = conj(W^{qr}):
//   with e_r = y_r + y_{R-r}, o_r = y_r - y_{R-r} (r = 1..h, h = (R-1)/2, y = pre-twiddled inputs)
//   X[q]   = y_0 + sum_r e_r cos(2 pi q r / R) -/+ i sum_r o_r sin(2 pi q r / R)
//   X[R-q] = y_0 + sum_r e_r cos(..)           +/- i sum_r o_r sin(..)
// i.e. 4 real FMAs per (q, r) pair and TWO outputs -- a quarter of the multiplies of the plain DFT.
// Step 1 rewrites the input buffer in place (e_r at slot r, o_r at slot R-r); step 2 is the tiled sum.
template <typename T2, int LB, int QT>
__device__ __forceinline__ void pass_odd_sym(T2* __restrict__ a, T2* __restrict__ b, const T2* __restrict__ tw, int N, int Ns, int R,
                                             int groups, bool inverse) {
    const int M = N / R;
    const int h = (R - 1) / 2;
    const int tstep = N / (Ns * R);
    const int nl = groups * LB;
    // ---- step 1: pre-twiddle + fold
    for (int it = threadIdx.x; it < M * h * nl; it += blockDim.x) {
        const int l = it / (M * h);
        const int rem = it - l * (M * h);
        const int r = rem / M + 1, j = rem - (r - 1) * M;
        const int k = j % Ns;
        T2* line = a + (size_t)l * N + j;
        T2 y1 = line[(size_t)r * M], y2 = line[(size_t)(R - r) * M];
        if (Ns > 1) {
            T2 w1 = tw[(int)(((long long)r * k * tstep) % N)], w2 = tw[(int)(((long long)(R - r) * k * tstep) % N)];
            if (inverse) { w1.y = -w1.y; w2.y = -w2.y; }
            y1 = cmul(y1, w1);
            y2 = cmul(y2, w2);
        }
        line[(size_t)r * M] = mk2<T2>(y1.x + y2.x, y1.y + y2.y);
        line[(size_t)(R - r) * M] = mk2<T2>(y1.x - y2.x, y1.y - y2.y);
    }
    __syncthreads();
    // ---- step 2: tiles of QT outputs q in [1, h] (plus one item per butterfly for q = 0)
    const int qtiles = (h + QT - 1) / QT + 1;
    const int items = M * qtiles * groups;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int g = it / (M * qtiles);
        const int rem = it - g * (M * qtiles);
        const int j = rem / qtiles, qt = rem - j * qtiles;
        const int k = j % Ns;
        const T2* in = a + (size_t)g * LB * N + j;
        const int dst0 = (j / Ns) * Ns * R + k;
        if (qt == qtiles - 1) {                       // X[0] = y_0 + sum_r e_r
#pragma unroll
            for (int l = 0; l < LB; ++l) {
                T2 acc = in[(size_t)l * N];
                for (int r = 1; r <= h; ++r) {
                    const T2 e = in[(size_t)l * N + (size_t)r * M];
                    acc.x += e.x;
                    acc.y += e.y;
                }
                b[(size_t)(g * LB + l) * N + dst0] = acc;
            }
            continue;
        }
        int step[QT], idx[QT];
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const int q = min(qt * QT + t + 1, h);
            step[t] = (int)(((long long)q * M) % N);
            idx[t] = step[t];                          // r starts at 1
        }
        T2 ce[QT][LB], so[QT][LB];
#pragma unroll
        for (int t = 0; t < QT; ++t)
#pragma unroll
            for (int l = 0; l < LB; ++l) { ce[t][l].x = ce[t][l].y = 0; so[t][l].x = so[t][l].y = 0; }
        for (int r = 1; r <= h; ++r) {
            T2 e[LB], o[LB];
#pragma unroll
            for (int l = 0; l < LB; ++l) {
                e[l] = in[(size_t)l * N + (size_t)r * M];
                o[l] = in[(size_t)l * N + (size_t)(R - r) * M];
            }
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                const T2 w = tw[idx[t]];               // (cos, -sin) of 2 pi q r / R
                idx[t] += step[t];
                if (idx[t] >= N) idx[t] -= N;
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    ce[t][l].x = fma(e[l].x, w.x, ce[t][l].x);
                    ce[t][l].y = fma(e[l].y, w.x, ce[t][l].y);
                    so[t][l].x = fma(o[l].x, -w.y, so[t][l].x);
                    so[t][l].y = fma(o[l].y, -w.y, so[t][l].y);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const int q = qt * QT + t + 1;
            if (q <= h) {
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    const T2 y0 = in[(size_t)l * N];
                    const T2 base = mk2<T2>(y0.x + ce[t][l].x, y0.y + ce[t][l].y);
                    // forward: X[q] = base - i So, X[R-q] = base + i So ; inverse: swapped
                    const T2 lo = mk2<T2>(base.x + so[t][l].y, base.y - so[t][l].x);
                    const T2 hi = mk2<T2>(base.x - so[t][l].y, base.y + so[t][l].x);
                    T2* out = b + (size_t)(g * LB + l) * N + dst0;
                    out[(size_t)q * Ns] = inverse ? hi : lo;
                    out[(size_t)(R - q) * Ns] = inverse ? lo : hi;
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Large odd radix (e.g. 107 of 214 = 2 * 107) as a register-tiled real matrix product on packed f32x2.
//
// Same conjugate-symmetric formulation as pass_odd_sym -- for q = 1..h
//     Ce[q] = sum_r e_r cos(2 pi q r / R),   So[q] = sum_r o_r sin(2 pi q r / R)        (e, o complex)
// i.e. two real (h x h) matrices applied to 2 * (lines * M) real columns -- but organised as a GEMM:
//   * step 1 folds the inputs and TRANSPOSES them into the other buffer as T[r][j][line] (line fastest), so the
//     four lines a thread owns are 32 contiguous bytes: two 128-bit loads per operand instead of four 64-bit ones;
//   * cos / sin come from a dedicated table ctab[2][h][QP] (QP = h rounded up to 4) in shared memory: one 128-bit
//     broadcast load gives the 4 outputs q of a thread, no per-term index arithmetic;
//   * a thread owns 4 (q) x 4 (lines) complex accumulators for Ce and for So; a complex value IS a packed f32x2,
//     so every term is one FFMA2 (fma.rn.f32x2) with the duplicated cos / sin: 32 FFMA2 per 6 LDS.128.
// The result is written back to the INPUT buffer in the standard [line][index] layout.
__device__ __forceinline__ unsigned long long fx2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fx2_dup(float v) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ float2 fx2_unpack(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

#ifndef SB_GEMM_LT
#define SB_GEMM_LT 2          // lines per thread of the GEMM tile (4 or 2): 4 q x LT lines complex accumulators for Ce and So
#endif
__device__ __forceinline__ void pass_odd_gemm(float2* __restrict__ a, float2* __restrict__ b, const float2* __restrict__ tw,
                                              const float* __restrict__ ctab, int N, int Ns, int R, int nlines, bool inverse) {
    const int M = N / R;
    const int h = (R - 1) / 2;
    const int QP = (h + 3) & ~3;
    const int tstep = N / (Ns * R);
    // ---- step 1: (pre-twiddle,) fold, transpose.  T[(r * M + j) * nlines + line]; r = 0 holds y_0.
    // Lanes run over lines (LP = lines rounded up to a power of two <= 32), the rest of the block over (r, j).
    int lpl = 2;
    while ((1 << lpl) < nlines && lpl < 5) ++lpl;
    const int LP = 1 << lpl;
    const int line0 = threadIdx.x & (LP - 1);
    const int slot = threadIdx.x >> lpl, nslots = blockDim.x >> lpl;
    for (int line = line0; line < nlines; line += LP) {
        const float2* al = a + (size_t)line * N;
        for (int idx = slot; idx < M; idx += nslots) b[(size_t)idx * nlines + line] = al[idx];
        for (int r = 1 + slot; r <= h; r += nslots) {
            for (int j = 0; j < M; ++j) {
                float2 y1 = al[r * M + j], y2 = al[(R - r) * M + j];
                if (Ns > 1) {
                    const int k = j % Ns;
                    float2 w1 = tw[(int)(((long long)r * k * tstep) % N)], w2 = tw[(int)(((long long)(R - r) * k * tstep) % N)];
                    if (inverse) { w1.y = -w1.y; w2.y = -w2.y; }
                    y1 = cmul(y1, w1);
                    y2 = cmul(y2, w2);
                }
                b[(size_t)(r * M + j) * nlines + line] = make_float2(y1.x + y2.x, y1.y + y2.y);
                b[(size_t)((R - r) * M + j) * nlines + line] = make_float2(y1.x - y2.x, y1.y - y2.y);
            }
        }
    }
    __syncthreads();
    // ---- step 2: the GEMM.  item = (q tile of 4, column group = (j, 4 lines)); lanes run over column groups.
    constexpr int LT = SB_GEMM_LT;
    const int lgs = nlines / LT;
    const int ncg = M * lgs;
    const int nqt = h >> 2;                                // FULL tiles of 4 outputs q; the h & 3 leftover outputs go below
    const float* cosb = ctab;
    const float* sinb = ctab + h * QP;
    const int es = (M * nlines) >> 1;                      // ulonglong2 stride of one r step in T
    const int n_gemm = nqt * ncg;
    // work list of the block: [GEMM items | leftover-q columns | X[0] columns]; the two cheap tails keep the threads
    // busy that the GEMM items do not cover (214 = 2 * 107, 24 lines: 156 + 48 + 48 of 256 threads)
    const int ncol = M * nlines;
    const int nleft = (h & 3) ? ncol : 0;
    for (int it = threadIdx.x; it < n_gemm + nleft + ncol; it += blockDim.x) {
        if (it < n_gemm) {
            const int qt = it / ncg, cg = it - qt * ncg;
            const int j = cg / lgs, lg = cg - j * lgs;
            const ulonglong2* ep = reinterpret_cast<const ulonglong2*>(b + (size_t)(M + j) * nlines + lg * LT);
            const ulonglong2* op = reinterpret_cast<const ulonglong2*>(b + (size_t)((R - 1) * M + j) * nlines + lg * LT);
            const float4* cp = reinterpret_cast<const float4*>(cosb + qt * 4);
            const float4* sp = reinterpret_cast<const float4*>(sinb + qt * 4);
            unsigned long long ce[4][LT], so[4][LT];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int l = 0; l < LT; ++l) ce[t][l] = so[t][l] = 0ull;
#pragma unroll 2
            for (int r = 1; r <= h; ++r) {
                const float4 c = *cp, s = *sp;
                unsigned long long e[LT], o[LT];
#pragma unroll
                for (int v = 0; v < LT / 2; ++v) {
                    const ulonglong2 ev = ep[v], ov = op[v];
                    e[2 * v] = ev.x; e[2 * v + 1] = ev.y;
                    o[2 * v] = ov.x; o[2 * v + 1] = ov.y;
                }
                cp += QP >> 2;
                sp += QP >> 2;
                ep += es;
                op -= es;
                const float cv[4] = {c.x, c.y, c.z, c.w}, sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const unsigned long long cc = fx2_dup(cv[t]), ss = fx2_dup(sv[t]);
#pragma unroll
                    for (int l = 0; l < LT; ++l) {
                        ce[t][l] = fx2_fma(e[l], cc, ce[t][l]);
                        so[t][l] = fx2_fma(o[l], ss, so[t][l]);
                    }
                }
            }
            const int k = j % Ns;
            const int dst0 = (j / Ns) * Ns * R + k;
            const float2* y0p = b + (size_t)j * nlines + lg * LT;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                const float2 y0 = y0p[l];
                float2* out = a + (size_t)(lg * LT + l) * N + dst0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int q = qt * 4 + t + 1;
                    const float2 cev = fx2_unpack(ce[t][l]), sov = fx2_unpack(so[t][l]);
                    const float bx = y0.x + cev.x, by = y0.y + cev.y;
                    // forward: X[q] = base - i So, X[R-q] = base + i So ; inverse: swapped
                    const float2 lo = make_float2(bx + sov.y, by - sov.x), hi = make_float2(bx - sov.y, by + sov.x);
                    out[(size_t)q * Ns] = inverse ? hi : lo;
                    out[(size_t)(R - q) * Ns] = inverse ? lo : hi;
                }
            }
        } else if (it < n_gemm + nleft) {
            // leftover outputs q = 4 * nqt + 1 .. h of one column (j, line)
            const int col = it - n_gemm;
            const int j = col / nlines, line = col - j * nlines;
            const int nq = h & 3;
            float2 ce[3], so[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) ce[t] = so[t] = make_float2(0.f, 0.f);
            const float* cp = cosb + 4 * nqt;
            const float* sp = sinb + 4 * nqt;
            for (int r = 1; r <= h; ++r) {
                const float2 e = b[(size_t)(r * M + j) * nlines + line], o = b[(size_t)((R - r) * M + j) * nlines + line];
#pragma unroll
                for (int t = 0; t < 3; ++t)
                    if (t < nq) {
                        const float c = cp[t], s = sp[t];
                        ce[t].x = fmaf(e.x, c, ce[t].x);
                        ce[t].y = fmaf(e.y, c, ce[t].y);
                        so[t].x = fmaf(o.x, s, so[t].x);
                        so[t].y = fmaf(o.y, s, so[t].y);
                    }
                cp += QP;
                sp += QP;
            }
            const float2 y0 = b[(size_t)j * nlines + line];
            float2* out = a + (size_t)line * N + (j / Ns) * Ns * R + (j % Ns);
#pragma unroll
            for (int t = 0; t < 3; ++t)
                if (t < nq) {
                    const int q = 4 * nqt + t + 1;
                    const float bx = y0.x + ce[t].x, by = y0.y + ce[t].y;
                    const float2 lo = make_float2(bx + so[t].y, by - so[t].x), hi = make_float2(bx - so[t].y, by + so[t].x);
                    out[(size_t)q * Ns] = inverse ? hi : lo;
                    out[(size_t)(R - q) * Ns] = inverse ? lo : hi;
                }
        } else {
            // X[0] = y_0 + sum_r e_r of one column
            const int col = it - n_gemm - nleft;
            const int j = col / nlines, line = col - j * nlines;
            float2 acc = b[(size_t)j * nlines + line];
            for (int r = 1; r <= h; ++r) {
                const float2 e = b[(size_t)(r * M + j) * nlines + line];
                acc.x += e.x;
                acc.y += e.y;
            }
            a[(size_t)line * N + (j / Ns) * Ns * R + (j % Ns)] = acc;
        }
    }
}

template <typename T2>
struct OddGemm {      // float64 lines have no packed-FMA path: fall through to pass_odd_sym
    static __device__ __forceinline__ bool run(T2*, T2*, const T2*, const float*, int, int, int, int, bool) { return false; }
};
template <>
struct OddGemm<float2> {
    static __device__ __forceinline__ bool run(float2* a, float2* b, const float2* tw, const float* ctab, int N, int Ns, int R,
                                               int nlines, bool inverse) {
        pass_odd_gemm(a, b, tw, ctab, N, Ns, R, nlines, inverse);
        return true;
    }
};

// In-place (ping-pong) FFT of `nlines` lines of length plan.n held in shared memory, line stride
// plan.n.  nlines must be a multiple of LB.  Returns the buffer that holds the result.  All
// threads of the block must call; ends with a __syncthreads().
template <typename T2, int LB>
__device__ T2* fft_lines(T2* buf0, T2* buf1, const T2* __restrict__ tw, const FftPlan& plan, int nlines, bool inverse,
                         const float* __restrict__ ctab = nullptr) {
    // NOTE: odd-radix passes fold their inputs in place, so the input buffer is clobbered.
    const int N = plan.n;
    T2* a = buf0;
    T2* b = buf1;
    int Ns = 1;
    const int groups = nlines / LB;
    for (int f = 0; f < plan.nfac; ++f) {
        const int R = plan.fac[f];
        bool in_place = false;
        if (R == 4) pass_radix4<T2, LB>(a, b, tw, N, Ns, plan.M[f], plan.tstep[f], plan.mM[f], plan.mNs[f], groups, inverse);
        else if (R == 2) pass_radix2<T2, LB>(a, b, tw, N, Ns, plan.M[f], plan.tstep[f], plan.mM[f], plan.mNs[f], groups, inverse);
        else if (R == 5) pass_radix5<T2, LB>(a, b, tw, N, Ns, plan.M[f], plan.tstep[f], plan.mM[f], plan.mNs[f], groups, inverse);
        else if (R == 3) pass_radix3<T2, LB>(a, b, tw, N, Ns, plan.M[f], plan.tstep[f], plan.mM[f], plan.mNs[f], groups, inverse);
        else if (ctab != nullptr && R == plan.gemm_radix && (nlines & 3) == 0 &&
                 OddGemm<T2>::run(a, b, tw, ctab, N, Ns, R, nlines, inverse)) in_place = true;
        else if (R & 1) pass_odd_sym<T2, LB, (sizeof(T2) == 8 ? (LB >= 4 ? 4 : 8) : (LB >= 4 ? 2 : 4))>(a, b, tw, N, Ns, R, groups, inverse);
        else pass_generic<T2, LB, (LB >= 4 ? 4 : 8)>(a, b, tw, N, Ns, R, groups, inverse);
        __syncthreads();
        if (!in_place) {                     // pass_odd_gemm leaves its result in the input buffer
            T2* t = a;
            a = b;
            b = t;
        }
        Ns *= R;
    }
    return a;
}
