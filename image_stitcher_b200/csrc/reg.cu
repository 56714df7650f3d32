// K0-K4: pairwise phase-cross-correlation registration of overlap strips (sm_100a).
//
// Replaces, for a batch of tile pairs, calculate_horizontal_shift / calculate_vertical_shift
// (stitcher_process.py:664-708), normalize_image (:844-855) and the un-vendored third-party
// skimage.registration.phase_cross_correlation(a, b, upsample_factor=uf) they call (algorithm
// restated in oracle/pcc_ref.py).  Chain per pair, all on the device:
//
//   K0 tile_minmax      whole-tile min/max (normalize_image stretches the WHOLE tile, :852)
//   K1 rows_fwd         strip crop + float64 stretch + truncating cast fused into the load;
//                       both strips packed as z = a + i b; FFT along x           -> Z
//   K2 cols_xpower      FFT along y of column pairs (kx, -kx); unpack the two real spectra,
//                       R = A conj(B) / max(|A conj(B)|, 100 eps); store R (for K4); inverse FFT
//                       along y                                                   -> Y (in place)
//   K3 rows_inv_argmax  rows packed two at a time (R is Hermitian, so the result is real), inverse
//                       FFT along x, |cc|, first-maximum argmax (warp shuffles)   -> coarse peak
//   K4 updft            skimage's matrix-multiply upsampled DFT in a ceil(1.5 uf)^2 window around
//                       the coarse peak: T = conj(R) Ex^T, out = Ey T, argmax     -> fine peak
//
// The kernels return INTEGER peak indices; the float64 shift and the reference's Python round()
// are rebuilt from them on the host, so the integer shifts are exact by construction.
// Arithmetic is float32 or float64 (template); SB_PREC_AUTO redoes low-confidence pairs in float64.
#include "sb_common.cuh"
#include "fft.cuh"
#include "reg_common.cuh"
#include "reg_tc.cuh"

#include <algorithm>
#include <cmath>
#include <map>

#ifndef SB_REG_CTAS
#define SB_REG_CTAS 3        // resident blocks per SM the FFT kernels are sized for (registers and shared memory)
#endif

namespace {

// ------------------------------------------------------------------------------------------ K0
__global__ void __launch_bounds__(256) minmax_init_kernel(int2* mm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mm[i] = make_int2(0x7fffffff, -1);
}

// grid = (blocks_per_tile, n_tiles).  128-bit loads, warp shuffles, one atomic pair per block.
__global__ void __launch_bounds__(256) tile_minmax_kernel(const uint16_t* const* __restrict__ tiles, int64_t px, int2* mm) {
    const uint16_t* t = tiles[blockIdx.y];
    unsigned lo = 0xffffu, hi = 0u;
    const int64_t nvec = ((reinterpret_cast<uintptr_t>(t) & 15) == 0) ? px / 8 : 0;
    const uint4* tv = reinterpret_cast<const uint4*>(t);
    // packed running min / max (two uint16 per register); four independent 128-bit loads in flight per thread
    unsigned lo2 = 0xffffffffu, hi2 = 0u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(tv + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            lo2 = __vminu2(__vminu2(lo2, v[u].x), __vminu2(v[u].y, __vminu2(v[u].z, v[u].w)));
            hi2 = __vmaxu2(__vmaxu2(hi2, v[u].x), __vmaxu2(v[u].y, __vmaxu2(v[u].z, v[u].w)));
        }
    }
    for (; i < nvec; i += stride) {
        const uint4 v = __ldcs(tv + i);
        lo2 = __vminu2(__vminu2(lo2, v.x), __vminu2(v.y, __vminu2(v.z, v.w)));
        hi2 = __vmaxu2(__vmaxu2(hi2, v.x), __vmaxu2(v.y, __vmaxu2(v.z, v.w)));
    }
    lo = min(lo2 & 0xffffu, lo2 >> 16);
    hi = max(hi2 & 0xffffu, hi2 >> 16);
    for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned v = t[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ unsigned slo[8], shi[8];
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        atomicMin(&mm[blockIdx.y].x, (int)lo);
        atomicMax(&mm[blockIdx.y].y, (int)hi);
    }
}

__global__ void __launch_bounds__(256) normalize_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                        const int2* __restrict__ mm, int64_t px, int maxval) {
    const int2 m = mm[blockIdx.y];
    const uint16_t* t = in + (int64_t)blockIdx.y * px;
    uint16_t* o = out + (int64_t)blockIdx.y * px;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = (uint16_t)stretch_px_f64(t[i], m.x, m.y, maxval);
}

// cos/sin table of the big odd radix (fft.cuh: pass_odd_gemm): global -> shared, right after the twiddles.
// All threads call; visibility is covered by the __syncthreads every kernel has before its first FFT pass.
template <typename T2>
__device__ __forceinline__ float* stage_ctab(T2* after_tw, const float* __restrict__ ctab_g, int ctab_n) {
    if (ctab_n == 0) return nullptr;
    float* dst = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(after_tw) + 15) & ~(uintptr_t)15);
    const float4* src = reinterpret_cast<const float4*>(ctab_g);
    for (int i = threadIdx.x; i < ctab_n / 4; i += blockDim.x) reinterpret_cast<float4*>(dst)[i] = __ldg(src + i);
    return dst;
}

// ------------------------------------------------------------------------------------------ K1
template <typename T, int LB>
__global__ void __launch_bounds__(256, SB_REG_CTAS) rows_fwd_kernel(const PairDesc* __restrict__ pairs, const int2* __restrict__ mm,
                                                       int tile_w, int Sh, int Sw, int lpb, int nrb, int swap, int maxval,
                                                       const typename Vec2<T>::type* __restrict__ tw_g, FftPlan plan,
                                                       const float* __restrict__ ctab_g, int ctab_n,
                                                       typename Vec2<T>::type* __restrict__ Z, int* __restrict__ nonzero,
                                                       unsigned long long* __restrict__ strip_sum) {
    using T2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    T2* buf0 = reinterpret_cast<T2*>(smem_raw);
    T2* buf1 = buf0 + (size_t)lpb * Sw;
    T2* tw = buf1 + (size_t)lpb * Sw;
    float* ctab = stage_ctab(tw + Sw, ctab_g, ctab_n);
    const int p = blockIdx.x / nrb, rb = blockIdx.x - p * nrb;
    const int y0 = rb * lpb;
    const PairDesc pd = pairs[p];
    int seen = 0;                                // bit 0: strip a has a non-zero pixel, bit 1: strip b
    unsigned sum_a = 0, sum_b = 0;               // strip sums of this thread (a few dozen 16-bit values each)
    const int2 ma = mm[pd.a_tile], mb = mm[pd.b_tile];
    const float inva = ma.y > ma.x ? (float)maxval / (float)(ma.y - ma.x) : 0.f;
    const float invb = mb.y > mb.x ? (float)maxval / (float)(mb.y - mb.x) : 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < Sw; i += blockDim.x) tw[i] = tw_g[i];
    if (!swap) {
        // strip crop + stretch fused into the load: a warp walks one strip row at a time (coalesced 64-byte reads of both
        // strips), four independent pixel pairs in flight per lane; no integer divisions in the index arithmetic
        for (int l = warp; l < lpb; l += nwarps) {
            T2* row = buf0 + (size_t)l * Sw;
            if (y0 + l >= Sh) {
                for (int x = lane; x < Sw; x += 32) row[x] = mk2<T2, T>(0, 0);
                continue;
            }
            const uint16_t* pa = pd.a + (size_t)(y0 + l) * tile_w;
            const uint16_t* pb = pd.b + (size_t)(y0 + l) * tile_w;
            for (int x0 = lane; x0 < Sw; x0 += 128) {
                unsigned av[4], bv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + 32 * u;
                    av[u] = x < Sw ? pa[x] : 0u;
                    bv[u] = x < Sw ? pb[x] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + 32 * u;
                    if (x < Sw) {
                        const int na = stretch_px(av[u], ma.x, ma.y, inva, maxval), nb = stretch_px(bv[u], mb.x, mb.y, invb, maxval);
                        seen |= (na != 0 ? 1 : 0) | (nb != 0 ? 2 : 0);
                        sum_a += (unsigned)na;
                        sum_b += (unsigned)nb;
                        row[x] = mk2<T2, T>((T)(na * kInScale), (T)(nb * kInScale));
                    }
                }
            }
        }
    } else {
        // Transposed frame (wide strips are processed as their transpose so that the long axis is the column pass):
        // frame row l is image column y0 + l, frame column x is image row x.  Lanes run over the frame rows -- adjacent
        // image columns, contiguous in memory -- and the rest of the block over the image rows.
        int lpl = 0;
        while ((1 << lpl) < lpb && lpl < 5) ++lpl;
        const int l = threadIdx.x & ((1 << lpl) - 1);
        const int xs = threadIdx.x >> lpl, nxs = blockDim.x >> lpl;
        for (int l2 = l; l2 < lpb; l2 += (1 << lpl)) {
            const bool live = y0 + l2 < Sh;
            const uint16_t* pa = pd.a + (y0 + l2);
            const uint16_t* pb = pd.b + (y0 + l2);
            T2* row = buf0 + (size_t)l2 * Sw;
            for (int x0 = xs; x0 < Sw; x0 += 4 * nxs) {
                unsigned av[4], bv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + u * nxs;
                    av[u] = (live && x < Sw) ? pa[(size_t)x * tile_w] : 0u;
                    bv[u] = (live && x < Sw) ? pb[(size_t)x * tile_w] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + u * nxs;
                    if (x < Sw) {
                        int na = 0, nb = 0;
                        if (live) {
                            na = stretch_px(av[u], ma.x, ma.y, inva, maxval);
                            nb = stretch_px(bv[u], mb.x, mb.y, invb, maxval);
                            seen |= (na != 0 ? 1 : 0) | (nb != 0 ? 2 : 0);
                            sum_a += (unsigned)na;
                            sum_b += (unsigned)nb;
                        }
                        row[x] = mk2<T2, T>((T)(na * kInScale), (T)(nb * kInScale));
                    }
                }
            }
        }
    }
    // An all-zero strip has an exactly zero spectrum in the reference (P == 0 -> cc == 0 -> argmax 0);
    // the packed transform only gets it to rounding noise, so record the fact instead.
    const int any_a = __syncthreads_or(seen & 1), any_b = __syncthreads_or(seen & 2);   // predicate OR, not bitwise
    if (threadIdx.x == 0 && (any_a || any_b)) atomicOr(&nonzero[p], (any_a ? 1 : 0) | (any_b ? 2 : 0));
    // Strip sums: the two strips share ONE complex transform (z = a + i b), so a strip much fainter than its partner is
    // recovered from the packed spectrum by cancellation and loses float32 digits in proportion to the ratio of their
    // magnitudes; SB_PREC_AUTO repeats such pairs in float64 (see reg_complete).
    sum_a = __reduce_add_sync(0xffffffffu, sum_a);
    sum_b = __reduce_add_sync(0xffffffffu, sum_b);
    if ((threadIdx.x & 31) == 0 && (sum_a | sum_b)) {
        if (sum_a) atomicAdd(&strip_sum[2 * p], (unsigned long long)sum_a);
        if (sum_b) atomicAdd(&strip_sum[2 * p + 1], (unsigned long long)sum_b);
    }
    T2* res = fft_lines<T2, LB>(buf0, buf1, tw, plan, lpb, false, ctab);
    // Z is kept TRANSPOSED (Zt[kx][y], y fastest) so that the column pass reads and writes whole contiguous lines;
    // here the lpb rows of this block are lpb consecutive y of every kx: runs of 8 * lpb contiguous bytes.
    // Lanes run over the rows (LP = lpb rounded up to a power of two), the rest of the block over kx.
    T2* zp = Z + (size_t)p * Sh * Sw;
    int lpl = 0;
    while ((1 << lpl) < lpb && lpl < 5) ++lpl;
    const int l = threadIdx.x & ((1 << lpl) - 1);
    const int xs = threadIdx.x >> lpl, nxs = blockDim.x >> lpl;
    for (int l2 = l; l2 < lpb; l2 += (1 << lpl))
        if (y0 + l2 < Sh)
            for (int x = xs; x < Sw; x += nxs) zp[(size_t)x * Sh + y0 + l2] = res[(size_t)l2 * Sw + x];
}

// ------------------------------------------------------------------------------------------ K2
// One block owns G columns kx and their mirrors (Sw - kx) % Sw: 2G lines of length Sh.
template <typename T, int G>
__global__ void __launch_bounds__(256, SB_REG_CTAS) cols_xpower_kernel(int Sh, int Sw, int ncg,
                                                          const typename Vec2<T>::type* __restrict__ tw_g, FftPlan plan,
                                                          const float* __restrict__ ctab_g, int ctab_n,
                                                          typename Vec2<T>::type* __restrict__ Z,
                                                          typename Vec2<T>::type* __restrict__ Rbuf) {
    using T2 = typename Vec2<T>::type;
    constexpr int NL = 2 * G;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    T2* buf0 = reinterpret_cast<T2*>(smem_raw);
    T2* buf1 = buf0 + (size_t)NL * Sh;
    T2* tw = buf1 + (size_t)NL * Sh;
    float* ctab = stage_ctab(tw + Sh, ctab_g, ctab_n);
    const int p = blockIdx.x / ncg, cg = blockIdx.x - p * ncg;
    T2* zp = Z + (size_t)p * Sh * Sw;
    T2* rp = Rbuf + (size_t)p * Sh * Sw;
    const int half = Sw / 2;                    // columns 0..half own their mirrors
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < Sh; i += blockDim.x) tw[i] = tw_g[i];
    // line l < G: column kx = cg*G + l ; line G + l: its mirror (unused when the column is self-mirrored).
    // A column of the strip is a contiguous line of Zt: one warp streams one line at a time.
    for (int l = warp; l < NL; l += nwarps) {
        const int kx = cg * G + (l < G ? l : l - G);
        const int kxm = kx == 0 ? 0 : Sw - kx;
        const bool live = kx <= half && (l < G || kxm != kx);
        const T2* src = zp + (size_t)(l < G ? kx : kxm) * Sh;
        T2* dst = buf0 + (size_t)l * Sh;
        for (int y = lane; y < Sh; y += 32) dst[y] = live ? src[y] : mk2<T2, T>(0, 0);
    }
    __syncthreads();
    constexpr int CLB = NL >= 4 ? 4 : NL;
    T2* f = fft_lines<T2, CLB>(buf0, buf1, tw, plan, NL, false, ctab);
    // unpack A = FFT(a), B = FFT(b) from Z = FFT(a + i b); R = A conj(B) / max(|A conj(B)|, clamp)
    for (int l = 0; l < G; ++l) {
        const int kx = cg * G + l;
        if (kx > half) break;
        const int kxm = kx == 0 ? 0 : Sw - kx;
        const bool self = (kxm == kx);
        T2* l1 = f + (size_t)l * Sh;
        T2* l2 = self ? l1 : f + (size_t)(G + l) * Sh;
        T2* r1 = rp + (size_t)kx * Sh;            // R is kept transposed as well (Rt[kx][ky])
        T2* r2 = rp + (size_t)kxm * Sh;
        for (int ky = threadIdx.x; ky < Sh; ky += blockDim.x) {
            const int kym = ky == 0 ? 0 : Sh - ky;
            if (self && ky > kym) continue;       // the partner (kym, kx) lives in the same line: handle each pair once
            const T2 z1 = l1[ky], z2 = l2[kym];
            const T ax = (T)0.5 * (z1.x + z2.x), ay = (T)0.5 * (z1.y - z2.y);
            const T bx = (T)0.5 * (z1.y + z2.y), by = (T)-0.5 * (z1.x - z2.x);
            T px = ax * bx + ay * by, py = ay * bx - ax * by;
            const T mag = sqrt(px * px + py * py);
            const T den = mag > (T)kClamp ? mag : (T)kClamp;
            px /= den;
            py /= den;
            if (self && ky == kym) py = 0;        // self-conjugate bin: exactly real
            l1[ky] = mk2<T2, T>(px, py);
            l2[kym] = mk2<T2, T>(px, -py);
            r1[ky] = mk2<T2, T>(px, py);
            r2[kym] = mk2<T2, T>(px, -py);
        }
    }
    __syncthreads();
    T2* other = (f == buf0) ? buf1 : buf0;
    T2* y = fft_lines<T2, CLB>(f, other, tw, plan, NL, true, ctab);
    for (int l = warp; l < NL; l += nwarps) {
        const int kx = cg * G + (l < G ? l : l - G);
        if (kx > half) continue;
        const int kxm = kx == 0 ? 0 : Sw - kx;
        if (l >= G && kxm == kx) continue;
        T2* dst = zp + (size_t)(l < G ? kx : kxm) * Sh;
        const T2* src = y + (size_t)l * Sh;
        for (int yy = lane; yy < Sh; yy += 32) dst[yy] = src[yy];
    }
}

// ------------------------------------------------------------------------------------------ K3
template <typename T, int LB>
__global__ void __launch_bounds__(256, SB_REG_CTAS) rows_inv_argmax_kernel(int Sh, int Sw, int lpb, int nrb, int swap,
                                                              const typename Vec2<T>::type* __restrict__ tw_g, FftPlan plan,
                                                              const float* __restrict__ ctab_g, int ctab_n,
                                                              const typename Vec2<T>::type* __restrict__ Y,
                                                              CtaBest* __restrict__ best, float* __restrict__ rowmax) {
    using T2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    T2* buf0 = reinterpret_cast<T2*>(smem_raw);
    T2* buf1 = buf0 + (size_t)lpb * Sw;
    T2* tw = buf1 + (size_t)lpb * Sw;
    float* ctab = stage_ctab(tw + Sw, ctab_g, ctab_n);
    const int p = blockIdx.x / nrb, rb = blockIdx.x - p * nrb;
    const int l0 = rb * lpb;                                 // first packed line of this block
    const T2* yp = Y + (size_t)p * Sh * Sw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < Sw; i += blockDim.x) tw[i] = tw_g[i];
    {   // lanes over the packed lines (2 * lpb consecutive y of one x are contiguous in Yt), the rest of the block over x
        int lpl = 0;
        while ((1 << lpl) < lpb && lpl < 5) ++lpl;
        const int lq = threadIdx.x & ((1 << lpl) - 1);
        const int xs = threadIdx.x >> lpl, nxs = blockDim.x >> lpl;
        for (int l = lq; l < lpb; l += (1 << lpl)) {
            const int y1 = 2 * (l0 + l), y2 = y1 + 1;
            const bool h1 = y1 < Sh, h2 = y2 < Sh;
            // four x positions (eight independent loads) in flight per thread: this phase is pure L2 latency
            for (int x0 = xs; x0 < Sw; x0 += 4 * nxs) {
                T2 r1[4], r2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + u * nxs;
                    r1[u] = (h1 && x < Sw) ? yp[(size_t)x * Sh + y1] : mk2<T2, T>(0, 0);
                    r2[u] = (h2 && x < Sw) ? yp[(size_t)x * Sh + y2] : mk2<T2, T>(0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + u * nxs;
                    if (x < Sw) buf0[(size_t)l * Sw + x] = mk2<T2, T>(r1[u].x - r2[u].y, r1[u].y + r2[u].x);   // row y1 + i * row y2
                }
            }
        }
    }
    __syncthreads();
    const T2* res = fft_lines<T2, LB>(buf0, buf1, tw, plan, lpb, true, ctab);
    T bv = (T)-1, b2 = (T)-1;
    int bi = 0x7fffffff;
    float* rmax = rowmax + (size_t)p * Sh;
    for (int l = warp; l < lpb; l += nwarps) {
        const int y1 = 2 * (l0 + l), y2 = y1 + 1;
        if (y1 >= Sh) break;
        const T2* row = res + (size_t)l * Sw;
        T m1 = (T)0, m2 = (T)0;                                  // maxima of the two rows of this line (|cc| >= 0)
        for (int x = lane; x < Sw; x += 32) {
            const T2 v = row[x];
            const T a1 = fabs(v.x), a2 = fabs(v.y);
            // first maximum in the C order of the ORIGINAL strip: in a transposed frame (y, x) is (col, row) there
            top2_update<T>(bv, bi, b2, a1, swap ? x * Sh + y1 : y1 * Sw + x);
            m1 = a1 > m1 ? a1 : m1;
            if (y2 < Sh) {
                top2_update<T>(bv, bi, b2, a2, swap ? x * Sh + y2 : y2 * Sw + x);
                m2 = a2 > m2 ? a2 : m2;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
            m1 = o1 > m1 ? o1 : m1;
            m2 = o2 > m2 ? o2 : m2;
        }
        if (lane == 0) {                                         // per-row maxima: the runner-up outside the peak's band
            const double sc = 1.0 / ((double)Sh * (double)Sw);
            rmax[y1] = (float)((double)m1 * sc);
            if (y2 < Sh) rmax[y2] = (float)((double)m2 * sc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const T o2 = __shfl_xor_sync(0xffffffffu, b2, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        top2_merge<T>(bv, bi, b2, ov, oi, o2);
    }
    __shared__ double sv[8], s2[8];
    __shared__ int si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = (double)bv; s2[threadIdx.x >> 5] = (double)b2; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = (double)bv, c = (double)b2;
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) top2_merge<double>(b, bi, c, sv[w], si[w], s2[w]);
        const double sc = 1.0 / ((double)Sh * (double)Sw);
        CtaBest& o = best[(size_t)p * nrb + rb];
        o.val = b * sc;
        o.second = c > 0.0 ? c * sc : 0.0;
        o.idx = bi;
    }
}

// One warp per pair: the global first maximum, the second-largest |cc| anywhere else (the near-tie test of SB_PREC_AUTO)
// and the runner-up = largest |cc| outside the band of three frame rows (circular) centred on the peak row -- the frame
// row is the strip's LONG axis position (image row for horizontal pairs, image column for vertical pairs in the
// transposed frame), so the band removes the peak's own line and its two neighbours.
__global__ void __launch_bounds__(32) peak_final_kernel(const CtaBest* __restrict__ best, int nrb, int Sh, int Sw, int swap,
                                                        const float* __restrict__ rowmax, const int* __restrict__ nonzero,
                                                        const int* __restrict__ fault,
                                                        const unsigned long long* __restrict__ strip_sum, PeakOut* out) {
    const int p = blockIdx.x;
    if (*fault) {                                // the tensor pipeline of an earlier kernel timed out: poison the result
        if (threadIdx.x == 0) {
            out[p].coarse_y = out[p].coarse_x = -999;
            out[p].fine_y = out[p].fine_x = -1;
            out[p].peak = out[p].second = out[p].runner_up = out[p].fine_peak = out[p].fine_second = 0.f;
            out[p].skew = 1.f;
        }
        return;
    }
    if (nonzero[p] != 3) {                       // a strip is identically zero: cc == 0 everywhere, first index wins
        if (threadIdx.x == 0) {
            out[p].coarse_y = out[p].coarse_x = 0;
            out[p].peak = out[p].second = out[p].runner_up = out[p].fine_peak = out[p].fine_second = 0.f;
            out[p].fine_y = out[p].fine_x = -1;
            out[p].skew = 1.f;
        }
        return;
    }
    double bv = -1.0, b2 = -1.0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < nrb; i += 32) {
        const CtaBest c = best[(size_t)p * nrb + i];
        top2_merge<double>(bv, bi, b2, c.val, c.idx, c.second);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const double o2 = __shfl_xor_sync(0xffffffffu, b2, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        top2_merge<double>(bv, bi, b2, ov, oi, o2);
    }
    // frame coordinates of the peak (the index was formed in the original strip's C order)
    const int cy = swap ? bi % Sh : bi / Sw, cx = swap ? bi / Sh : bi % Sw;
    float ru = 0.f;
    const float* rm = rowmax + (size_t)p * Sh;
    for (int y = threadIdx.x; y < Sh; y += 32) {
        int d = y - cy;
        if (d < 0) d = -d;
        if (d > 1 && d < Sh - 1) ru = fmaxf(ru, rm[y]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ru = fmaxf(ru, __shfl_xor_sync(0xffffffffu, ru, o));
    if (threadIdx.x == 0) {
        out[p].coarse_y = cy;
        out[p].coarse_x = cx;
        out[p].peak = (float)bv;
        out[p].second = (float)fmax(b2, 0.0);
        out[p].runner_up = ru;
        out[p].fine_y = out[p].fine_x = -1;
        out[p].fine_peak = out[p].fine_second = 0.f;
        float skew = 1.f;                               // ratio of the strip magnitudes (packed transform only)
        if (strip_sum != nullptr) {
            const double sa = (double)strip_sum[2 * p], sb = (double)strip_sum[2 * p + 1];
            skew = (float)(fmax(sa, sb) / fmax(fmin(sa, sb), 1.0));
        }
        out[p].skew = skew;
    }
}

// ------------------------------------------------------------------------------------------ K4
// Twiddles of skimage's _upsampled_dft: exp(-2 pi i (u - off) fftfreq(n, uf)[x]) with
// off = dftshift - shift*uf; (u - off) and n*uf*fftfreq are integers, so the phase is reduced
// exactly in integer arithmetic before sincospi.
template <typename T>
__global__ void __launch_bounds__(256) updft_twiddle_kernel(const PeakOut* __restrict__ peaks, int Sh, int Sw, int uf, int rs,
                                                            int dftshift, typename Vec2<T>::type* __restrict__ Ex,
                                                            typename Vec2<T>::type* __restrict__ Ey) {
    using T2 = typename Vec2<T>::type;
    const int p = blockIdx.y;
    const PeakOut pk = peaks[p];
    const int cy = pk.coarse_y > Sh / 2 ? pk.coarse_y - Sh : pk.coarse_y;     // shift[shift > fix(n/2)] -= n
    const int cx = pk.coarse_x > Sw / 2 ? pk.coarse_x - Sw : pk.coarse_x;
    const int total = rs * (Sw + Sh);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const bool isx = i < rs * Sw;
        const int n = isx ? Sw : Sh;
        const int ii = isx ? i : i - rs * Sw;
        const int u = ii / n, k = ii - u * n;
        const long long m = (long long)u - dftshift + (long long)(isx ? cx : cy) * uf;
        const int sk = (k < (n + 1) / 2) ? k : k - n;                        // n * fftfreq(n)[k]
        const long long den = (long long)n * uf;
        long long num = (m * sk) % den;
        if (num < 0) num += den;
        double s, c;
        sincospi(-2.0 * (double)num / (double)den, &s, &c);
        T2 w = mk2<T2, T>((T)c, (T)s);
        if (isx) Ex[((size_t)p * rs + u) * Sw + k] = w;
        else Ey[((size_t)p * rs + u) * Sh + k] = w;
    }
}

// T[u][y] = sum_x conj(R[y][x]) Ex[u][x] with R stored transposed (Rt[x][y]).  A block owns YT = 64 rows y and the
// whole x range, split over 256 / YT = 4 interleaved x slices; a thread keeps all rs (<= UMAX) outputs of its y in
// registers.  Lanes run over y, so every load of Rt is a contiguous 256-byte warp access; the twiddles Ex[.][x] are
// staged in shared memory in chunks of XC and read as broadcasts.  Slices are reduced through shared memory.
template <typename T, int UMAX>
__global__ void __launch_bounds__(256) updft_rows_kernel(int Sh, int Sw, int rs, int u0, int nyb,
                                                         const typename Vec2<T>::type* __restrict__ Rbuf,
                                                         const typename Vec2<T>::type* __restrict__ Ex,
                                                         typename Vec2<T>::type* __restrict__ Tm) {
    using T2 = typename Vec2<T>::type;
    constexpr int YT = 64, NS = 256 / YT;
    constexpr int XC = 128;                      // x chunk staged in shared memory
    static_assert(YT <= XC, "the reduction buffer [UMAX][YT] reuses the twiddle staging area [XC][UMAX]");
    __shared__ __align__(16) unsigned char ex_raw[UMAX * XC * sizeof(T2)];
    T2* exs = reinterpret_cast<T2*>(ex_raw);
    const int p = blockIdx.x / nyb, yb = blockIdx.x - p * nyb;
    const int ys = threadIdx.x % YT, xs = threadIdx.x / YT;
    const int y = yb * YT + ys;
    const T2* rp = Rbuf + (size_t)p * Sh * Sw;
    const T2* ex = Ex + (size_t)p * rs * Sw;
    T2 acc[UMAX];
#pragma unroll
    for (int u = 0; u < UMAX; ++u) acc[u].x = acc[u].y = 0;
    for (int xc = 0; xc < Sw; xc += XC) {
        const int nx = min(XC, Sw - xc);
        __syncthreads();
        for (int i = threadIdx.x; i < UMAX * nx; i += blockDim.x) {
            const int u = i / nx, x = i - u * nx;
            exs[x * UMAX + u] = (u0 + u < rs) ? ex[(size_t)(u0 + u) * Sw + xc + x] : mk2<T2, T>(0, 0);
        }
        __syncthreads();
        if (y < Sh) {
            // four independent loads of Rt in flight per thread: this loop is L2-latency bound otherwise
            for (int x0 = xs; x0 < nx; x0 += 4 * NS) {
                T2 r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = x0 + k * NS;
                    r[k] = x < nx ? rp[(size_t)(xc + x) * Sh + y] : mk2<T2, T>(0, 0);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = x0 + k * NS;
                    if (x < nx) {
                        const T2 rc = mk2<T2, T>(r[k].x, -r[k].y);
#pragma unroll
                        for (int u = 0; u < UMAX; ++u) cfma(acc[u], rc, exs[x * UMAX + u]);
                    }
                }
            }
        }
    }
    // reduce the NS x slices: slice k parks its sums in shared memory, slice 0 adds them
    for (int k = 1; k < NS; ++k) {
        __syncthreads();
        if (xs == k) {
#pragma unroll
            for (int u = 0; u < UMAX; ++u) exs[u * YT + ys] = acc[u];
        }
        __syncthreads();
        if (xs == 0) {
#pragma unroll
            for (int u = 0; u < UMAX; ++u) {
                const T2 o = exs[u * YT + ys];
                acc[u].x += o.x;
                acc[u].y += o.y;
            }
        }
    }
    if (xs == 0 && y < Sh) {
#pragma unroll
        for (int u = 0; u < UMAX; ++u)
            if (u0 + u < rs) Tm[(size_t)p * rs * Sh + (size_t)(u0 + u) * Sh + y] = acc[u];
    }
}

// out[v][u] = sum_y Ey[v][y] T[u][y].  grid = (pair, v): every warp reduces a few u over y; |out|^2 to global.
template <typename T>
__global__ void __launch_bounds__(256) updft_cols_kernel(int Sh, int rs, const typename Vec2<T>::type* __restrict__ Tm,
                                                         const typename Vec2<T>::type* __restrict__ Ey, double* __restrict__ mag2) {
    using T2 = typename Vec2<T>::type;
    const int p = blockIdx.x / rs, v = blockIdx.x - p * rs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const T2* tp = Tm + (size_t)p * rs * Sh;
    const T2* ey = Ey + ((size_t)p * rs + v) * Sh;
    for (int u = warp; u < rs; u += nw) {
        T2 acc = mk2<T2, T>(0, 0);
        for (int y = lane; y < Sh; y += 32) cfma(acc, ey[y], tp[(size_t)u * Sh + y]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
        }
        if (lane == 0) mag2[((size_t)p * rs + v) * rs + u] = (double)(acc.x * acc.x + acc.y * acc.y);   // monotone in |.|
    }
}

// first maximum (C order) of the rs x rs window and the second-largest value (near-tie test); one warp per pair
__global__ void __launch_bounds__(32) updft_final_kernel(int rs, int swap, const double* __restrict__ mag2, float inv_n,
                                                         const int* __restrict__ nonzero, PeakOut* __restrict__ peaks) {
    const int p = blockIdx.x;
    if (nonzero[p] != 3) {                       // zero cross-power: the upsampled window is all zero -> index 0
        if (threadIdx.x == 0) { peaks[p].fine_y = peaks[p].fine_x = 0; peaks[p].fine_peak = peaks[p].fine_second = 0.f; }
        return;
    }
    double bv = -1.0, b2 = -1.0;
    int bi = 0x7fffffff;
    for (int o = threadIdx.x; o < rs * rs; o += 32) {
        const int v = o / rs, u = o - v * rs;            // mag2 is [v][u] in the frame; original order is [u][v] when swapped
        top2_update<double>(bv, bi, b2, mag2[(size_t)p * rs * rs + o], swap ? u * rs + v : o);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, s);
        const double o2 = __shfl_xor_sync(0xffffffffu, b2, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        top2_merge<double>(bv, bi, b2, ov, oi, o2);
    }
    if (threadIdx.x == 0) {
        peaks[p].fine_y = swap ? bi % rs : bi / rs;      // frame coordinates again
        peaks[p].fine_x = swap ? bi / rs : bi % rs;
        peaks[p].fine_peak = (float)(sqrt(bv) * (double)inv_n);
        peaks[p].fine_second = (float)(sqrt(fmax(b2, 0.0)) * (double)inv_n);
    }
}

// ------------------------------------------------------------------------------------------ host side

FftPlan make_plan(int n) {
    // prime factors, twos merged into fours, largest radix first (the order is free for Stockham)
    std::vector<int> f;
    int m = n, twos = 0;
    while (m % 2 == 0) { ++twos; m /= 2; }
    for (int p = 3; (long long)p * p <= m; p += 2)
        while (m % p == 0) { f.push_back(p); m /= p; }
    if (m > 1) f.push_back(m);
    for (; twos >= 2; twos -= 2) f.push_back(4);
    if (twos) f.push_back(2);
    std::sort(f.begin(), f.end(), [](int a, int b) { return a > b; });
    FftPlan pl;
    pl.n = n;
    pl.nfac = 0;
    pl.gemm_radix = 0;
    for (int v : f) pl.fac[pl.nfac++] = v;
    int Ns = 1;
    for (int i = 0; i < 16; ++i) {
        const int R = i < pl.nfac ? pl.fac[i] : 1;
        pl.Ns[i] = Ns;
        pl.M[i] = n / R;
        pl.tstep[i] = n / (Ns * R);
        pl.mM[i] = fastdiv_magic((unsigned)pl.M[i]);
        pl.mNs[i] = fastdiv_magic((unsigned)Ns);
        if (i < pl.nfac) Ns *= R;
    }
    return pl;
}

// The largest radix, when it is a big odd one (107 of 214, 157 of 314): worth the register-tiled f32x2 pass.
int big_odd_radix(const FftPlan& pl) { return (pl.nfac > 0 && (pl.fac[0] & 1) && pl.fac[0] >= 13) ? pl.fac[0] : 0; }

// cos/sin table of pass_odd_gemm: [2][h][QP] floats, h = (R-1)/2, QP = h rounded up to 4 (zero padded);
// entry [0][r-1][q-1] = cos(2 pi q r / R), [1][r-1][q-1] = sin(2 pi q r / R), computed in long double.
int get_ctab(sb_ctx* ctx, int R, const float** out, int* n_floats) {
    const int h = (R - 1) / 2, QP = (h + 3) & ~3;
    const int n = 2 * h * QP;
    const uint64_t key = ((uint64_t)1 << 40) | (uint64_t)R;
    auto it = ctx->twiddle_cache.find(key);
    if (it == ctx->twiddle_cache.end()) {
        std::vector<float> t((size_t)n, 0.0f);
        const long double tau = 6.283185307179586476925286766559L;
        for (int r = 1; r <= h; ++r)
            for (int q = 1; q <= h; ++q) {
                const long double a = tau * (long double)(((long long)q * r) % R) / (long double)R;
                t[(size_t)(r - 1) * QP + (q - 1)] = (float)cosl(a);
                t[(size_t)h * QP + (size_t)(r - 1) * QP + (q - 1)] = (float)sinl(a);
            }
        DevBuf b;
        int rc = sb_reserve(ctx, b, (size_t)n * sizeof(float));
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpy(b.p, t.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
        it = ctx->twiddle_cache.emplace(key, b).first;
    }
    *out = reinterpret_cast<const float*>(it->second.p);
    *n_floats = n;
    return SB_OK;
}

template <typename T>
int get_twiddles(sb_ctx* ctx, int n, const typename Vec2<T>::type** out) {
    using T2 = typename Vec2<T>::type;
    const uint64_t key = ((uint64_t)n << 1) | (sizeof(T) == 8 ? 1 : 0);
    auto it = ctx->twiddle_cache.find(key);
    if (it == ctx->twiddle_cache.end()) {
        std::vector<T2> h(n);
        const long double tau = 6.283185307179586476925286766559L;
        for (int m = 0; m < n; ++m) {
            const long double a = -tau * (long double)m / (long double)n;
            h[m].x = (T)cosl(a);
            h[m].y = (T)sinl(a);
        }
        DevBuf b;
        int rc = sb_reserve(ctx, b, (size_t)n * sizeof(T2));
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpy(b.p, h.data(), (size_t)n * sizeof(T2), cudaMemcpyHostToDevice));
        it = ctx->twiddle_cache.emplace(key, b).first;
    }
    *out = reinterpret_cast<const T2*>(it->second.p);
    return SB_OK;
}

struct GroupGeom {          // strips of one direction
    int Sh, Sw;
    int a_y0, a_x0, b_y0, b_x0;
};

bool group_swapped(const GroupGeom& g) {
    static const bool off = getenv("SB_REG_NO_SWAP") != nullptr;
    return g.Sw > g.Sh && !off;
}

// lines per block so that 2 buffers + the twiddle table fit comfortably in shared memory
int pick_lines(int n, size_t elem, int lb, size_t budget) {
    if ((size_t)n * elem * 3 > budget) return 0;
    int l = (int)((budget - (size_t)n * elem) / (2 * (size_t)n * elem));
    l = std::min(l, 32);
    l = l / lb * lb;
    return l;
}

// Enqueues the whole chain for one direction group on the lane's stream; the per-pair PeakOut records are copied to
// `h_out` (pinned for asynchronous jobs).  With do_sync the call returns when they have arrived.
template <typename T>
int run_group(sb_ctx* ctx, Lane* lane, const std::vector<PairDesc>& pairs, const GroupGeom& g, int tile_w,
              const int2* d_mm, int uf, int maxval, PeakOut* h_out, bool do_sync, PairDesc* pinned_pairs = nullptr) {
    cudaStream_t st = lane->stream;
    using T2 = typename Vec2<T>::type;
    const int n = (int)pairs.size();
    // Wide strips (vertical pairs: 214 x 1024) run in a transposed frame so that the long axis is always the column pass
    // and every group has the H-like shape the kernels are tuned for; indices are mapped back in finish_pair.
    const int swap = group_swapped(g) ? 1 : 0;
    const int Sh = swap ? g.Sw : g.Sh, Sw = swap ? g.Sh : g.Sw;
    const size_t strip = (size_t)Sh * Sw;
    // float32 frames with a 1024-long axis: the forward short-axis transform runs on the tensor cores and the column pass
    // as warp-level register FFTs (reg_tc.cu); everything else -- and float64 -- stays on the radix engine below
    TcPlan tc;
    if (sizeof(T) == 4 && tile_w % 8 == 0) {             // (the bulk copies of the strip rows need a 16-byte tile pitch)
        const int rc_tc = sb_tc_plan(ctx, Sh, Sw, &tc);
        if (rc_tc) return rc_tc;
    }
    const FftPlan plan_x = make_plan(Sw), plan_y = make_plan(Sh);
    const T2 *tw_x = nullptr, *tw_y = nullptr;
    int rc = get_twiddles<T>(ctx, Sw, &tw_x);
    if (rc) return rc;
    rc = get_twiddles<T>(ctx, Sh, &tw_y);
    if (rc) return rc;

    const int rs = (3 * uf + 1) / 2;            // ceil(1.5 * uf)
    const int dftshift = rs / 2;                // fix(rs / 2)

    // Sub-batches: up to kWays of them run concurrently (the lane's stream and its auxiliary streams, each with its own
    // workspace slice) -- the chain is a sequence of short dependent kernels, and the neighbours' blocks fill the tails
    // and launch gaps.  Measured on B200 (1152 pairs): 64 MB of spectra in flight (L2 resident) 18.6 ms, 128 MB 17.1,
    // 256 MB 16.7, 384 MB over 3 streams 16.1 (15.3 with 3 blocks per SM), 768 MB over 4 streams 14.9: fewer, larger launches beat
    // strict L2 residency.  The persistent tensor-core kernels (r2) want MANY tiles per block -- at 54 pairs a block sees
    // 3-6 tiles and the pipeline fill / drain is a third of the launch -- so the budget is 4 GB over 2 streams: the 576
    // pairs of one direction of a 96-well plate in one sub-batch (7.8 -> 7.35 ms per 1152 pairs at 4.5 GB / 1 stream).
    // (the radix engine -- float64, or strips the tensor-core kernels do not take -- keeps its round-1 optimum: 768 MB / 4)
    static const int kWaysEnv = getenv("SB_REG_WAYS") ? std::max(1, std::min(4, atoi(getenv("SB_REG_WAYS")))) : 0;
    static const int kBudgetEnv = getenv("SB_REG_L2_MB") ? std::max(8, atoi(getenv("SB_REG_L2_MB"))) : 0;
    const int kWays = kWaysEnv ? kWaysEnv : (tc.ok ? 2 : 4);
    const size_t per_pair = 2 * strip * sizeof(T2);
    const size_t kL2Budget = (size_t)(kBudgetEnv ? kBudgetEnv : (tc.ok ? 4096 : 768)) << 20;
    int B = (int)std::max<size_t>(1, kL2Budget / per_pair);
    int ways = 1;
    while (ways < kWays && B / (ways + 1) >= 2 && n > B / (ways + 1)) ++ways;
    B = std::min(std::max(1, B / ways), n);

    // cos/sin tables of the big odd radices (float32 lines only)
    FftPlan px_plan = plan_x, py_plan = plan_y;
    const float *ctab_x = nullptr, *ctab_y = nullptr;
    int ctab_xn = 0, ctab_yn = 0;
    if (sizeof(T) == 4 && !getenv("SB_REG_NO_GEMM")) {
        if (big_odd_radix(plan_x)) {
            px_plan.gemm_radix = big_odd_radix(plan_x);
            rc = get_ctab(ctx, px_plan.gemm_radix, &ctab_x, &ctab_xn);
            if (rc) return rc;
        }
        if (big_odd_radix(plan_y)) {
            py_plan.gemm_radix = big_odd_radix(plan_y);
            rc = get_ctab(ctx, py_plan.gemm_radix, &ctab_y, &ctab_yn);
            if (rc) return rc;
        }
    }
    const size_t ctab_xb = ctab_xn ? (size_t)ctab_xn * 4 + 16 : 0, ctab_yb = ctab_yn ? (size_t)ctab_yn * 4 + 16 : 0;

    // launch geometry: two blocks per SM wherever the lines + tables allow it
    constexpr size_t kBudget = (size_t)(227 / SB_REG_CTAS - 1) * 1024 + 512;
    int lbx = 4, lpbx = pick_lines(Sw, sizeof(T2), 4, kBudget - ctab_xb);
    if (lpbx < 4) { lbx = 1; lpbx = pick_lines(Sw, sizeof(T2), 1, 200 * 1024 - ctab_xb); }
    if (lpbx < 1) return sb_fail(ctx, SB_ERR_UNSUPPORTED, "strip width %d too large for the shared-memory FFT", Sw);
    if (lbx != 4) { px_plan.gemm_radix = 0; }                    // pass_odd_gemm needs a multiple of 4 lines
    const size_t smem_x = (size_t)(2 * lpbx + 1) * Sw * sizeof(T2) + (px_plan.gemm_radix ? ctab_xb : 0);
    // columns per block: short lines (e.g. 214 = 2 * 107) take up to 16 columns + 16 mirrors so that the prime-radix
    // pass has enough independent work for 256 threads and global accesses are 64..128-byte segments
    static const int kG[] = {16, 12, 8, 4, 2, 1};
    auto smem_for = [&](int g) { return (size_t)(2 * 2 * g + 1) * Sh * sizeof(T2) + ctab_yb; };
    int G = 1;
    for (int g : kG)
        if (smem_for(g) <= kBudget) { G = g; break; }
    if (G == 1 && smem_for(2) <= 200 * 1024) G = 2;
    if (G < 2) py_plan.gemm_radix = 0;
    const size_t smem_y = smem_for(G);
    if (smem_y > 227 * 1024) return sb_fail(ctx, SB_ERR_UNSUPPORTED, "strip height %d too large for the shared-memory FFT", Sh);
    const int nrb_fwd = (Sh + lpbx - 1) / lpbx;
    const int nlines_inv = (Sh + 1) / 2;
    const int nrb_inv = (nlines_inv + lpbx - 1) / lpbx;
    const int ncg = (Sw / 2 + 1 + G - 1) / G;
    const int rows_per_block = 64;    // updft_rows_kernel: YT rows y per block
    const int nrb_up = (Sh + rows_per_block - 1) / rows_per_block;
    const float* cx = px_plan.gemm_radix ? ctab_x : nullptr;
    const int cxn = px_plan.gemm_radix ? ctab_xn : 0;
    const float* cy = py_plan.gemm_radix ? ctab_y : nullptr;
    const int cyn = py_plan.gemm_radix ? ctab_yn : 0;

    // workspace: Z | R | Ex | Ey | T | best | peaks | pair descriptors
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += round_up64((int64_t)bytes, 256); return o; };
    const size_t o_Z = carve((size_t)B * strip * sizeof(T2));
    const size_t o_R = carve((size_t)B * strip * sizeof(T2));
    const size_t o_Ex = carve((size_t)B * rs * Sw * sizeof(T2));
    const size_t o_Ey = carve((size_t)B * rs * Sh * sizeof(T2));
    const size_t o_T = carve((size_t)B * rs * Sh * sizeof(T2));
    const size_t o_best = carve((size_t)B * std::max(nrb_inv, Sh >> 7) * sizeof(CtaBest));
    const size_t o_mag = carve((size_t)B * rs * rs * sizeof(double));
    const size_t o_rmax = carve((size_t)B * Sh * sizeof(float));
    const size_t o_Zh = carve(tc.ok ? sb_tc_zh_bytes(tc, B) : 0);
    static const bool no_tc_updft = getenv("SB_REG_NO_TC_UPDFT") != nullptr;
    const bool tc_updft = tc.ok && uf > 1 && rs <= 16 && !no_tc_updft;      // upsampled-DFT rows stage on the tensor cores
    const size_t o_Bup = carve(tc_updft ? sb_tc_updft_table_bytes(tc, B) : 0);
    const bool half_lines = tc.ok && tc.inverse && (uf == 1 || tc_updft);   // no radix kernel reads R / Y: half arrays
    const int tc_lines = half_lines ? Sw / 2 + 1 : Sw;
    const size_t way_bytes = off;                           // everything above exists once per concurrent sub-batch
    off = way_bytes * ways;
    const size_t o_peaks = carve((size_t)n * sizeof(PeakOut));
    const size_t o_pairs = carve((size_t)n * sizeof(PairDesc));
    const size_t o_nz = carve((size_t)(n + 1) * sizeof(int));      // + the tensor pipeline's fault flag
    const size_t o_sums = carve((size_t)2 * n * sizeof(unsigned long long));
    rc = sb_reserve(ctx, lane->reg_work, off);
    if (rc) return rc;
    uint8_t* w = (uint8_t*)lane->reg_work.p;
    T2* Z = (T2*)(w + o_Z);
    T2* Rb = (T2*)(w + o_R);
    T2* Ex = (T2*)(w + o_Ex);
    T2* Ey = (T2*)(w + o_Ey);
    T2* Tm = (T2*)(w + o_T);
    CtaBest* best = (CtaBest*)(w + o_best);
    double* mag2 = (double*)(w + o_mag);
    float* rowmax = (float*)(w + o_rmax);
    PeakOut* peaks = (PeakOut*)(w + o_peaks);
    PairDesc* d_pairs = (PairDesc*)(w + o_pairs);
    int* d_nz = (int*)(w + o_nz);
    int* d_fault = d_nz + n;
    unsigned long long* d_sums = (unsigned long long*)(w + o_sums);
    SB_CUDA(ctx, cudaMemsetAsync(d_nz, 0, (o_sums - o_nz) + (size_t)2 * n * sizeof(unsigned long long), st));
    lane->dbg_ptr[0] = tc.ok ? w + o_Zh : nullptr;  lane->dbg_bytes[0] = tc.ok ? sb_tc_zh_bytes(tc, std::min(B, n)) : 0;
    const size_t dbg_strip = tc.ok ? (size_t)tc_lines * Sh : strip;          // (half arrays: n/2 + 1 lines per pair)
    lane->dbg_ptr[1] = w + o_Z;                     lane->dbg_bytes[1] = (size_t)std::min(B, n) * dbg_strip * sizeof(T2);
    lane->dbg_ptr[2] = w + o_R;                     lane->dbg_bytes[2] = (size_t)std::min(B, n) * dbg_strip * sizeof(T2);
    lane->dbg_ptr[3] = w + o_T;                     lane->dbg_bytes[3] = (size_t)std::min(B, n) * rs * Sh * sizeof(T2);
    // Descriptors go up from PINNED memory when the caller provides it: a copy from pageable memory synchronises the
    // stream first, i.e. the host would wait for everything already enqueued on this lane (uploads, earlier groups).
    const PairDesc* h_pairs = pairs.data();
    if (pinned_pairs) {
        memcpy(pinned_pairs, pairs.data(), (size_t)n * sizeof(PairDesc));
        h_pairs = pinned_pairs;
    }
    SB_CUDA(ctx, cudaMemcpyAsync(d_pairs, h_pairs, (size_t)n * sizeof(PairDesc), cudaMemcpyHostToDevice, st));

    auto k1 = lbx == 4 ? rows_fwd_kernel<T, 4> : rows_fwd_kernel<T, 1>;
    auto k3 = lbx == 4 ? rows_inv_argmax_kernel<T, 4> : rows_inv_argmax_kernel<T, 1>;
    auto k2 = G == 16 ? cols_xpower_kernel<T, 16>
              : G == 12 ? cols_xpower_kernel<T, 12>
              : G == 8  ? cols_xpower_kernel<T, 8>
              : G == 4  ? cols_xpower_kernel<T, 4>
              : G == 2  ? cols_xpower_kernel<T, 2>
                        : cols_xpower_kernel<T, 1>;
    SB_CUDA(ctx, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
    SB_CUDA(ctx, cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
    SB_CUDA(ctx, cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_y));

    const float inv_n = 1.0f / ((float)Sh * (float)Sw);
    cudaStream_t way_stream[4] = {st, st, st, st};
    if (ways > 1) {
        if (!lane->aux_fork) SB_CUDA(ctx, cudaEventCreateWithFlags(&lane->aux_fork, cudaEventDisableTiming));
        SB_CUDA(ctx, cudaEventRecord(lane->aux_fork, st));            // pair descriptors, min/max, earlier lane work
        for (int k = 1; k < ways; ++k) {
            if (!lane->aux[k - 1]) {
                SB_CUDA(ctx, cudaStreamCreateWithFlags(&lane->aux[k - 1], cudaStreamNonBlocking));
                SB_CUDA(ctx, cudaEventCreateWithFlags(&lane->aux_join[k - 1], cudaEventDisableTiming));
            }
            way_stream[k] = lane->aux[k - 1];
            SB_CUDA(ctx, cudaStreamWaitEvent(lane->aux[k - 1], lane->aux_fork, 0));
        }
    }
    int way = 0;
    for (int p0 = 0; p0 < n; p0 += B, way = (way + 1) % ways) {
        const int nb = std::min(B, n - p0);
        cudaStream_t ws = way_stream[way];
        const size_t wo = way_bytes * way;
        T2 *Zw = (T2*)((uint8_t*)Z + wo), *Rw = (T2*)((uint8_t*)Rb + wo), *Exw = (T2*)((uint8_t*)Ex + wo);
        T2 *Eyw = (T2*)((uint8_t*)Ey + wo), *Tw = (T2*)((uint8_t*)Tm + wo);
        CtaBest* bestw = (CtaBest*)((uint8_t*)best + wo);
        double* magw = (double*)((uint8_t*)mag2 + wo);
        float* rmaxw = (float*)((uint8_t*)rowmax + wo);
        if (tc.ok) {
            void* Zhw = w + o_Zh + wo;
            rc = sb_tc_forward(ctx, ws, tc, d_pairs + p0, nb, d_mm, tile_w, swap, maxval, Zhw, d_nz + p0, d_fault);
            if (rc) return rc;
            // full arrays with the mirrored (conjugate) lines only for the radix kernels downstream: R for updft_rows_kernel,
            // Y for rows_inv_argmax_kernel; the tensor-core stages read the half arrays
            rc = sb_tc_columns(ctx, ws, tc, nb, Zhw, Rw, Zw, tc_lines, half_lines ? 0 : (1 | (tc.inverse ? 0 : 2)));
            if (rc) return rc;
            ctx->launches -= 2;                          // (counted below with the other two)
        } else {
            k1<<<nb * nrb_fwd, 256, smem_x, ws>>>(d_pairs + p0, d_mm, tile_w, Sh, Sw, lpbx, nrb_fwd, swap, maxval, tw_x, px_plan, cx, cxn, Zw, d_nz + p0, d_sums + 2 * p0);
            k2<<<nb * ncg, 256, smem_y, ws>>>(Sh, Sw, ncg, tw_y, py_plan, cy, cyn, Zw, Rw);
        }
        if (tc.ok && tc.inverse) {
            rc = sb_tc_inverse(ctx, ws, tc, nb, Zw, tc_lines, swap, bestw, rmaxw, d_fault);
            if (rc) return rc;
            ctx->launches--;
        } else {
            k3<<<nb * nrb_inv, 256, smem_x, ws>>>(Sh, Sw, lpbx, nrb_inv, swap, tw_x, px_plan, cx, cxn, Zw, bestw, rmaxw);
        }
        peak_final_kernel<<<nb, 32, 0, ws>>>(bestw, (tc.ok && tc.inverse) ? (Sh >> 7) : nrb_inv, Sh, Sw, swap, rmaxw, d_nz + p0, d_fault, tc.ok ? nullptr : d_sums + 2 * p0, peaks + p0);
        ctx->launches += 4;
        if (uf > 1 && tc_updft) {
            rc = sb_tc_updft_rows(ctx, ws, tc, nb, peaks + p0, uf, rs, dftshift, Rw, tc_lines, w + o_Bup + wo, Eyw, Tw, d_fault);
            if (rc) return rc;
        } else if (uf > 1) {
            updft_twiddle_kernel<T><<<dim3(8, nb), 256, 0, ws>>>(peaks + p0, Sh, Sw, uf, rs, dftshift, Exw, Eyw);
            ctx->launches++;
            for (int u0 = 0; u0 < rs; u0 += 16) {            // rs = 15 for the reference's upsample_factor 10: one launch
                updft_rows_kernel<T, 16><<<nb * nrb_up, 256, 0, ws>>>(Sh, Sw, rs, u0, nrb_up, Rw, Exw, Tw);
                ctx->launches++;
            }
        }
        if (uf > 1) {
            updft_cols_kernel<T><<<nb * rs, 256, 0, ws>>>(Sh, rs, Tw, Eyw, magw);
            updft_final_kernel<<<nb, 32, 0, ws>>>(rs, swap, magw, inv_n, d_nz + p0, peaks + p0);
            ctx->launches += 2;
        }
        SB_CUDA(ctx, cudaGetLastError());
    }
    for (int k = 1; k < ways; ++k) {
        SB_CUDA(ctx, cudaEventRecord(lane->aux_join[k - 1], lane->aux[k - 1]));
        SB_CUDA(ctx, cudaStreamWaitEvent(st, lane->aux_join[k - 1], 0));
    }
    SB_CUDA(ctx, cudaMemcpyAsync(h_out, peaks, (size_t)n * sizeof(PeakOut), cudaMemcpyDeviceToHost, st));
    if (do_sync) SB_CUDA(ctx, cudaStreamSynchronize(st));
    return SB_OK;
}

// skimage's float64 shift from integer indices, then the reference's round() (half-to-even)
void finish_pair(const PeakOut& pk_frame, const GroupGeom& g, int dir, int uf, sb_pair_result* r) {
    PeakOut pk = pk_frame;
    if (group_swapped(g)) {                      // the kernels worked on the transposed strip: map the indices back
        std::swap(pk.coarse_y, pk.coarse_x);
        std::swap(pk.fine_y, pk.fine_x);
    }
    const int shape[2] = {g.Sh, g.Sw};
    const int coarse[2] = {pk.coarse_y, pk.coarse_x};
    const int fine[2] = {pk.fine_y, pk.fine_x};
    double shift[2];
    for (int d = 0; d < 2; ++d) {
        double s = (double)coarse[d];
        const double mid = std::trunc((double)shape[d] / 2.0);            // np.fix(axis_size / 2)
        if (s > mid) s -= (double)shape[d];
        if (uf > 1) {
            const double ufd = (double)uf;
            s = std::nearbyint(s * ufd) / ufd;                            // np.round(shift * uf) / uf
            const double dftshift = std::trunc(std::ceil(ufd * 1.5) / 2.0);
            s += ((double)fine[d] - dftshift) / ufd;
        }
        if (shape[d] == 1) s = 0.0;
        shift[d] = s;
    }
    r->shift[0] = shift[0];
    r->shift[1] = shift[1];
    // Python round(float) == round-half-to-even of the binary value == nearbyint in the default mode
    if (dir == SB_DIR_HORIZONTAL) {
        r->dy = (int32_t)std::nearbyint(shift[0]);
        r->dx = (int32_t)std::nearbyint(shift[1] - (double)g.Sw);
    } else {
        r->dy = (int32_t)std::nearbyint(shift[0] - (double)g.Sh);
        r->dx = (int32_t)std::nearbyint(shift[1]);
    }
    r->coarse[0] = pk.coarse_y;
    r->coarse[1] = pk.coarse_x;
    r->fine[0] = uf > 1 ? pk.fine_y : -1;
    r->fine[1] = uf > 1 ? pk.fine_x : -1;
    r->peak = pk.peak;
    r->second = pk.second;
    r->runner_up = pk.runner_up;
    r->fine_peak = pk.fine_peak;
    r->fine_second = uf > 1 ? pk.fine_second : 0.f;
}

// upload (host) or adopt (device) the unique tiles of a job and compute their min/max
struct TileSet {
    std::map<const void*, int> index;       // caller pointer -> tile index
    std::vector<const uint16_t*> dev;       // device address per tile index
};

int prepare_tiles(sb_ctx* ctx, Lane* lane, const std::vector<const void*>& ptrs, int H, int W, int mem, TileSet& ts,
                  int2** d_mm_out, int2* h_mm, const uint16_t** pinned_table = nullptr) {
    cudaStream_t st = lane->stream;
    for (const void* p : ptrs) {
        if (!p) return sb_fail(ctx, SB_ERR_INVALID, "NULL tile pointer");
        if (ts.index.emplace(p, (int)ts.index.size()).second) ts.dev.push_back(nullptr);
    }
    const int nt = (int)ts.index.size();
    const size_t tile_bytes = (size_t)H * W * 2;
    if (mem == SB_MEM_HOST) {
        int rc = sb_reserve(ctx, lane->reg_tiles, tile_bytes * nt);
        if (rc) return rc;
        for (auto& kv : ts.index) {
            uint8_t* d = (uint8_t*)lane->reg_tiles.p + tile_bytes * kv.second;
            SB_CUDA(ctx, cudaMemcpyAsync(d, kv.first, tile_bytes, cudaMemcpyHostToDevice, st));
            ts.dev[kv.second] = (const uint16_t*)d;
        }
    } else {
        for (auto& kv : ts.index) ts.dev[kv.second] = (const uint16_t*)kv.first;
    }
    // device table of tile pointers + min/max
    const size_t meta = round_up64((size_t)nt * sizeof(void*), 256) + (size_t)nt * sizeof(int2);
    int rc = sb_reserve(ctx, lane->reg_meta, meta);
    if (rc) return rc;
    const uint16_t* const* h_table = ts.dev.data();
    if (pinned_table) {                      // see run_group: pageable sources would stall the host on the lane's stream
        memcpy(pinned_table, ts.dev.data(), (size_t)nt * sizeof(void*));
        h_table = pinned_table;
    }
    SB_CUDA(ctx, cudaMemcpyAsync(lane->reg_meta.p, h_table, (size_t)nt * sizeof(void*), cudaMemcpyHostToDevice, st));
    int2* d_mm = (int2*)((uint8_t*)lane->reg_meta.p + round_up64((size_t)nt * sizeof(void*), 256));
    minmax_init_kernel<<<(nt + 255) / 256, 256, 0, st>>>(d_mm, nt);
    const int bpt = std::max(1, std::min(64, (ctx->sm_count * 16 + nt - 1) / nt));
    tile_minmax_kernel<<<dim3(bpt, nt), 256, 0, st>>>((const uint16_t* const*)lane->reg_meta.p, (int64_t)H * W, d_mm);
    ctx->launches += 2;
    SB_CUDA(ctx, cudaGetLastError());
    if (h_mm) SB_CUDA(ctx, cudaMemcpyAsync(h_mm, d_mm, (size_t)nt * sizeof(int2), cudaMemcpyDeviceToHost, st));
    *d_mm_out = d_mm;
    return SB_OK;
}

}  // namespace

// A registration job between its enqueue and its completion (asynchronous jobs park here until sb_sync(lane)).
struct RegGroup {
    int dir;
    GroupGeom g;
    std::vector<int> ids;          // indices into the job's pair list
    std::vector<PairDesc> pd;
    size_t first;                  // first PeakOut of the group in the pinned result block
};
struct RegPending {
    sb_register_job job;           // scalars only; `pairs` points into `pairs_copy`
    std::vector<sb_pair> pairs_copy;
    sb_pair_result* out = nullptr;
    int maxval = 65535;
    std::vector<RegGroup> groups;
    TileSet ts;
    int2* d_mm = nullptr;
    int2* h_mm = nullptr;          // pinned
    PeakOut* h_peaks = nullptr;    // pinned
};

static int reg_enqueue(sb_ctx* ctx, Lane* lane, const sb_register_job* job, sb_pair_result* out, int maxval, RegPending& pr) {
    const int H = job->tile_h, W = job->tile_w, n = job->n_pairs;
    pr.job = *job;
    pr.maxval = maxval;
    pr.pairs_copy.assign(job->pairs, job->pairs + n);
    pr.job.pairs = pr.pairs_copy.data();
    pr.out = out;

    std::vector<const void*> ptrs;
    for (int i = 0; i < n; ++i) {
        SB_CHECK(ctx, job->pairs[i].dir == SB_DIR_HORIZONTAL || job->pairs[i].dir == SB_DIR_VERTICAL, "pair %d: bad dir", i);
        ptrs.push_back(job->pairs[i].ref);
        ptrs.push_back(job->pairs[i].mov);
    }
    // pinned block of the lane: [min/max of every unique tile | PeakOut of every pair | tile pointer table | pair descriptors]
    const size_t mm_bytes = round_up64((size_t)2 * n * sizeof(int2), 64);
    const size_t pk_bytes = round_up64((size_t)n * sizeof(PeakOut), 64);
    const size_t tb_bytes = round_up64((size_t)2 * n * sizeof(void*), 64);
    int rc = sb_reserve_pinned(ctx, &lane->reg_host, &lane->reg_host_cap, mm_bytes + pk_bytes + tb_bytes + (size_t)n * sizeof(PairDesc));
    if (rc) return rc;
    pr.h_mm = reinterpret_cast<int2*>(lane->reg_host);
    pr.h_peaks = reinterpret_cast<PeakOut*>((uint8_t*)lane->reg_host + mm_bytes);
    const uint16_t** pinned_table = reinterpret_cast<const uint16_t**>((uint8_t*)lane->reg_host + mm_bytes + pk_bytes);
    PairDesc* pinned_pairs = reinterpret_cast<PairDesc*>((uint8_t*)lane->reg_host + mm_bytes + pk_bytes + tb_bytes);
    rc = prepare_tiles(ctx, lane, ptrs, H, W, job->mem, pr.ts, &pr.d_mm, pr.h_mm, pinned_table);
    if (rc) return rc;

    size_t first = 0;
    for (int dir = 0; dir < 2; ++dir) {
        RegGroup grp;
        grp.dir = dir;
        for (int i = 0; i < n; ++i)
            if (job->pairs[i].dir == dir) grp.ids.push_back(i);
        if (grp.ids.empty()) continue;
        GroupGeom& g = grp.g;
        if (dir == SB_DIR_HORIZONTAL) {
            // img_left[margin:-margin, -ov:], img_right[margin:-margin, :ov]   (:677-679)
            const int margin = (int)((double)H * 0.25), ov = job->max_overlap_x;
            SB_CHECK(ctx, margin >= 1 && H - 2 * margin >= 1, "tile height %d too small for the 25%% margin", H);
            SB_CHECK(ctx, ov >= 1 && ov <= W, "max_overlap_x %d outside [1, %d]", ov, W);
            g = {H - 2 * margin, ov, margin, W - ov, margin, 0};
        } else {
            // img_top[-ov:, margin:-margin], img_bot[:ov, margin:-margin]      (:700-702)
            const int margin = (int)((double)W * 0.25), ov = job->max_overlap_y;
            SB_CHECK(ctx, margin >= 1 && W - 2 * margin >= 1, "tile width %d too small for the 25%% margin", W);
            SB_CHECK(ctx, ov >= 1 && ov <= H, "max_overlap_y %d outside [1, %d]", ov, H);
            g = {ov, W - 2 * margin, H - ov, margin, 0, margin};
        }
        grp.pd.resize(grp.ids.size());
        for (size_t k = 0; k < grp.ids.size(); ++k) {
            const sb_pair& sp = job->pairs[grp.ids[k]];
            const int ia = pr.ts.index[sp.ref], ib = pr.ts.index[sp.mov];
            grp.pd[k].a = pr.ts.dev[ia] + (size_t)g.a_y0 * W + g.a_x0;
            grp.pd[k].b = pr.ts.dev[ib] + (size_t)g.b_y0 * W + g.b_x0;
            grp.pd[k].a_tile = ia;
            grp.pd[k].b_tile = ib;
        }
        grp.first = first;
        first += grp.ids.size();
        rc = job->precision == SB_PREC_F64
                 ? run_group<double>(ctx, lane, grp.pd, g, W, pr.d_mm, job->upsample_factor, maxval, pr.h_peaks + grp.first, false,
                                     pinned_pairs + grp.first)
                 : run_group<float>(ctx, lane, grp.pd, g, W, pr.d_mm, job->upsample_factor, maxval, pr.h_peaks + grp.first, false,
                                    pinned_pairs + grp.first);
        if (rc) return rc;
        pr.groups.push_back(std::move(grp));
    }
    return SB_OK;
}

// Waits for the lane, redoes low-confidence pairs in float64 (SB_PREC_AUTO) and writes the caller's results.
static int reg_complete(sb_ctx* ctx, Lane* lane, RegPending& pr) {
    SB_CUDA(ctx, cudaStreamSynchronize(lane->stream));
    const sb_register_job* job = &pr.job;
    const int W = job->tile_w;
    const int first_prec = job->precision == SB_PREC_F64 ? SB_PREC_F64 : SB_PREC_F32;
    std::vector<int2> mm(pr.h_mm, pr.h_mm + pr.ts.index.size());      // the pinned block is reused by the redo below
    for (RegGroup& grp : pr.groups) {
        const GroupGeom& g = grp.g;
        std::vector<PeakOut> res(pr.h_peaks + grp.first, pr.h_peaks + grp.first + grp.ids.size());
        std::vector<int> prec(grp.ids.size(), first_prec);
        for (const PeakOut& q : res)
            if (q.coarse_y == -999) return sb_fail(ctx, SB_ERR_CUDA, "registration: the tensor-core pipeline did not complete (timeout)");
        if (job->precision == SB_PREC_AUTO) {
            // Two reasons to repeat a pair in float64, the arithmetic the reference uses:
            //  (1) low confidence -- the peak does not stand clear of the correlation noise floor (rms 1/sqrt(N), expected
            //      maximum ~ sqrt(2 ln N / N)) or of the best value outside its own band of rows: an argmax among
            //      near-equal noise values;
            //  (2) a near-tie -- the second-largest |cc| (typically the neighbouring pixel of a half-pixel shift) or the
            //      second-largest value of the upsampled window is within kTie (relative) of the maximum.  The float32
            //      chain is accurate to ~1e-6 of the peak (FFT error eps * sqrt(log2 N) per element, averaged over N
            //      unit-magnitude bins), so a margin above 1e-4 cannot be reversed by the arithmetic; below it float32
            //      and complex128 may pick different indices, and after the reference's half-even round() a different
            //      integer shift.
            const double N = (double)g.Sh * g.Sw;
            const double floor_max = std::sqrt(2.0 * std::log(N) / N);
            constexpr float kTie = 1e-4f;
            std::vector<PairDesc> redo;
            std::vector<int> redo_k;
            for (size_t k = 0; k < grp.ids.size(); ++k) {
                const PeakOut& q = res[k];
                const bool zero = q.peak == 0.f && q.second == 0.f;      // identically zero strip (peak_final_kernel): cc == 0 exactly
                const bool low = !(q.peak > 4.0 * floor_max) || !(q.peak > 1.5f * q.runner_up);
                const bool tie_c = !(q.peak - q.second > kTie * q.peak);
                const bool tie_f = job->upsample_factor > 1 && !(q.fine_peak - q.fine_second > kTie * q.fine_peak);
                // (3) strips of very different magnitude through the PACKED transform of the radix path: float32 error grows
                //     with their ratio (4e-8 of the peak at 1:1, 2e-4 at 1:20000 -- measured); 32 keeps it below 1e-5
                const bool skewed = q.skew > 32.f;
                if (!zero && (low || tie_c || tie_f || skewed)) {
                    redo.push_back(grp.pd[k]);
                    redo_k.push_back((int)k);
                }
            }
            if (!redo.empty()) {
                std::vector<PeakOut> res2(redo.size());
                int rc = run_group<double>(ctx, lane, redo, g, W, pr.d_mm, job->upsample_factor, pr.maxval, res2.data(), true);
                if (rc) return rc;
                for (size_t j = 0; j < redo_k.size(); ++j) {
                    res[redo_k[j]] = res2[j];
                    prec[redo_k[j]] = SB_PREC_F64;
                }
            }
        }
        for (size_t k = 0; k < grp.ids.size(); ++k) {
            sb_pair_result* r = &pr.out[grp.ids[k]];
            memset(r, 0, sizeof(*r));
            finish_pair(res[k], g, grp.dir, job->upsample_factor, r);
            const sb_pair& sp = job->pairs[grp.ids[k]];
            const int2 ma = mm[pr.ts.index[sp.ref]], mb = mm[pr.ts.index[sp.mov]];
            r->ref_min = ma.x; r->ref_max = ma.y; r->mov_min = mb.x; r->mov_max = mb.y;
            r->precision = prec[k];
        }
    }
    return SB_OK;
}

// Completes the lane's parked asynchronous registration, if any (called from sb_sync and before new work on the lane).
int sb_register_complete(sb_ctx* ctx, int lane_idx) {
    Lane* lane = sb_lane(ctx, lane_idx);
    if (!lane || !lane->reg_pending) return SB_OK;
    RegPending* pr = static_cast<RegPending*>(lane->reg_pending);
    lane->reg_pending = nullptr;
    const int rc = reg_complete(ctx, lane, *pr);
    delete pr;
    return rc;
}

void sb_register_discard(sb_ctx* ctx, int lane_idx) {
    Lane* lane = sb_lane(ctx, lane_idx);
    if (lane && lane->reg_pending) {
        delete static_cast<RegPending*>(lane->reg_pending);
        lane->reg_pending = nullptr;
    }
}

int sb_register_pairs_impl(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out, bool async, int maxval) {
    SB_CHECK(ctx, job && out, "job/out is NULL");
    SB_CHECK(ctx, job->dtype == SB_U16, "only uint16 pixels are implemented");
    SB_CHECK(ctx, job->n_pairs >= 0 && (job->n_pairs == 0 || job->pairs), "bad pair list");
    SB_CHECK(ctx, job->upsample_factor >= 1 && job->upsample_factor <= 100, "upsample_factor %d out of range [1, 100]",
             job->upsample_factor);
    SB_CHECK(ctx, job->precision >= SB_PREC_F32 && job->precision <= SB_PREC_AUTO, "unknown precision %d", job->precision);
    SB_CHECK(ctx, job->tile_h > 0 && job->tile_w > 0, "bad tile shape");
    Lane* lane = sb_lane(ctx, job->lane);
    SB_CHECK(ctx, lane != nullptr, "lane %d out of range", job->lane);
    int rc = sb_register_complete(ctx, job->lane);        // one parked job per lane: finish the previous one first
    if (rc) return rc;
    if (job->n_pairs == 0) return SB_OK;
    RegPending* pr = new RegPending();
    rc = reg_enqueue(ctx, lane, job, out, maxval, *pr);
    if (rc) {
        cudaStreamSynchronize(lane->stream);              // nothing of the failed job may still use its buffers
        delete pr;
        return rc;
    }
    if (async) {
        lane->reg_pending = pr;
        return SB_OK;
    }
    rc = reg_complete(ctx, lane, *pr);
    delete pr;
    return rc;
}

int sb_normalize_impl(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int dtype, int mem,
                      int maxval) {
    SB_CHECK(ctx, dtype == SB_U16, "only uint16 pixels are implemented");
    SB_CHECK(ctx, tiles && out && n_tiles > 0 && tile_h > 0 && tile_w > 0, "bad arguments");
    Lane* lane = sb_lane(ctx, 0);
    cudaStream_t st = lane->stream;
    const size_t px = (size_t)tile_h * tile_w;
    std::vector<const void*> ptrs;
    for (int i = 0; i < n_tiles; ++i) ptrs.push_back((const uint8_t*)tiles + i * px * 2);
    TileSet ts;
    int2* d_mm = nullptr;
    int rc = sb_register_complete(ctx, 0);
    if (rc) return rc;
    rc = prepare_tiles(ctx, lane, ptrs, tile_h, tile_w, mem, ts, &d_mm, nullptr);
    if (rc) return rc;
    // tiles were given contiguously, so tile i has index i and (host case) sits at reg_tiles + i * px
    const uint16_t* d_in = mem == SB_MEM_HOST ? (const uint16_t*)lane->reg_tiles.p : (const uint16_t*)tiles;
    uint16_t* d_out = (uint16_t*)out;
    if (mem == SB_MEM_HOST) {
        rc = sb_reserve(ctx, lane->canvas, px * 2 * n_tiles);
        if (rc) return rc;
        d_out = (uint16_t*)lane->canvas.p;
    }
    normalize_kernel<<<dim3(std::max(1, ctx->sm_count * 4 / n_tiles), n_tiles), 256, 0, st>>>(d_in, d_out, d_mm, (int64_t)px, maxval);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    if (mem == SB_MEM_HOST) SB_CUDA(ctx, cudaMemcpyAsync(out, d_out, px * 2 * n_tiles, cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    return SB_OK;
}

// ------------------------------------------------------------------------------------------ self-test (test hook)
// Exhaustive proof that the integer stretches fused into the strip loads (stretch_px of the radix kernels; stretch_mulhi and
// stretch_core of the tensor-core converters -- and stretch_bits, the form stretch_mulhi replaced -- with their hand-over to
// stretch_px) equal the float64 expression of
// normalize_image (stretch_px_f64, stitcher_process.py:844-855) for EVERY pixel value and tile range: all pairs
// a = v - min, b = max - min with 0 <= a <= b, 1 <= b <= maxval (2^31 pairs for uint16).
namespace {
__global__ void __launch_bounds__(256) selftest_stretch_kernel(int maxval, unsigned long long* __restrict__ res) {
    const int b = blockIdx.x + 1;
    const float inv = (float)maxval / (float)b;                         // as rows_fwd_kernel forms it
    const float inv_lo = stretch_inv_lo((unsigned)b, (unsigned)maxval);   // as the tensor-core converters form it
    const unsigned magic_b = kStretchMagic * (unsigned)b;
    const StretchMagic smagic = stretch_magic((unsigned)b, (unsigned)maxval);
    unsigned long long bad = 0, first = ~0ull, fallbacks = 0;
    for (int a = threadIdx.x; a <= b; a += blockDim.x) {
        const int fast = stretch_px((unsigned)a, 0, b, inv, maxval), ref = stretch_px_f64((unsigned)a, 0, b, maxval);
        // the converters' form: magic-number bits, one-sided repair; `exact` hands the value to stretch_px
        bool ex = false, ex2 = false;
        const unsigned bits = stretch_bits((unsigned)a, (unsigned)b, inv_lo, (unsigned)maxval, magic_b, ex);
        const int fast_tc = ex ? fast : (int)(bits - kStretchMagic);
        const int core = stretch_core((unsigned)a, (unsigned)b, inv, (unsigned)maxval, ex2);      // (edge chunks, b < maxval)
        const int fast_core = (ex2 || b == maxval) ? fast : core;
        bool ex3 = false;                                                                         // interior chunks: multiply-high
        const unsigned mh = stretch_mulhi((unsigned)a, (unsigned)b, smagic, (unsigned)maxval, ex3);
        const int fast_mh = (ex3 || b == maxval) ? fast : (int)mh;
        fallbacks += ex3 ? 1 : 0;
        if (fast != ref || fast_tc != ref || fast_core != ref || fast_mh != ref || ex3 != ex) {
            ++bad;
            const unsigned long long key = ((unsigned long long)b << 32) | (unsigned)a;
            first = key < first ? key : first;
        }
    }
    if (bad) {
        atomicAdd(res + 1, bad);
        atomicMin(res + 2, first);
    }
    if (fallbacks) atomicAdd(res + 3, fallbacks);
}
}  // namespace

int sb_selftest_stretch_impl(sb_ctx* ctx, int maxval, uint64_t* out) {
    SB_CHECK(ctx, maxval == 255 || maxval == 65535, "selftest: maxval must be 255 or 65535");
    Lane* lane = sb_lane(ctx, 0);
    int rc = sb_reserve(ctx, lane->work, 64);
    if (rc) return rc;
    unsigned long long init[4] = {0ull, 0ull, ~0ull, 0ull};
    SB_CUDA(ctx, cudaMemcpyAsync(lane->work.p, init, sizeof(init), cudaMemcpyHostToDevice, lane->stream));
    selftest_stretch_kernel<<<maxval, 256, 0, lane->stream>>>(maxval, (unsigned long long*)lane->work.p);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    SB_CUDA(ctx, cudaMemcpyAsync(out, lane->work.p, 32, cudaMemcpyDeviceToHost, lane->stream));   // out[3]: exact-quotient fallbacks
    SB_CUDA(ctx, cudaStreamSynchronize(lane->stream));
    out[0] = (uint64_t)maxval * ((uint64_t)maxval + 3) / 2;            // sum over b of (b + 1)
    return SB_OK;
}
