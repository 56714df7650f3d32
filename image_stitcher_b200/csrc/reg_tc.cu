// Tensor-core (tcgen05) stages of the registration chain -- see umma.cuh for the operand layout and the 3-term tf32
// split, reg.cu for the chain itself.
#include "sb_common.cuh"
#include "reg_common.cuh"
#include "reg_tc.cuh"
#include "umma.cuh"
#include "fft_warp.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

// ==========================================================================================================
// T1 / T3  xdft_tc_kernel: the transforms along the SHORT strip axis (length n, ANY n) on the tensor cores.
//
// MODE 0 (forward, T1).  For a real row x[0..n) the half spectrum X[k], k = 0..n/2, is two real matrix products of the
// row folded about its centre -- e[j] = x[j] + x[n-j], o[j] = x[j] - x[n-j] (e[0] = x[0], e[n/2] = x[n/2] for even n):
//         Re X[k] =  sum_j e[j] cos(2 pi j k / n)            Im X[k] = -sum_j o[j] sin(2 pi j k / n)
// MODE 1 (inverse + argmax, T3).  The correlation row cc[y][x] from the half spectrum Y[y][k] of a REAL signal, folded
// on the output side: with c_0 = c_{n/2} = 1, c_k = 2 otherwise,
//         P[x] = sum_k Re Y[k] c_k cos(2 pi k x / n)         Q[x] = -sum_k Im Y[k] c_k sin(2 pi k x / n)
//         cc[x] = P[x] + Q[x],   cc[n - x] = P[x] - Q[x]     (x = 0 .. n/2)
//
// MODE 2 (first stage of the upsampled-DFT refinement, T4).  skimage's _upsampled_dft contracts the cross-power with a
// 15 x n twiddle matrix per pair: T[u][y] = sum_x conj(R[y][x]) Ex[u][x].  As one real product per 128 rows: A = (Re R,
// Im R) over ALL n columns, B = the pair's twiddles arranged as [Ex_re | Ex_im] for the real part and [Ex_im | -Ex_re]
// for the imaginary part (N = 32 accumulator columns: 16 real + 16 imaginary outputs), both parts into ONE accumulator.
//
// A block owns 128 rows (M) of one strip image / one pair; K runs in chunks of 8 through a kTcStages-deep ring of
// shared-memory operand stages: the A sub-tiles (part 0 / part 1, tf32 hi / lo) are produced by 8 warps -- MODE 0 from
// the stretched strip rows staged in shared memory (crop + normalize_image's stretch fused into the coalesced load, as
// in the radix path), MODE 1 straight from the column pass's output (one row per lane: coalesced) -- the B sub-tiles
// (cos / -sin tables, hi / lo, laid out on the host exactly as the tensor core reads them) arrive by 1-D bulk copies.
// One thread issues 6 tcgen05.mma.kind::tf32 per chunk (3-term split x 2 parts) into TMEM columns [0, NP) / [NP, 2 NP).
// Epilogue, one row per thread (TMEM lane == strip row): MODE 0 stores the spectrum TRANSPOSED, Zh[img][k][y] -- a whole
// 256-byte warp store per k; MODE 1 folds P, Q into |cc| and keeps the first maximum, the second-largest value and the
// row maximum (what rows_inv_argmax_kernel of the radix path emits).
// ==========================================================================================================
// The kernel is PERSISTENT (one block per SM) and warp-specialised, every hand-over an mbarrier:
//     warps 0-15  converters   MODE 0: raw strip rows in shared memory -> stretch (normalize_image, :844-855) -> fold ->
//                              tf32 split -> A operand stages; MODE 1 / 2: Y or R -> A stages straight from global
//                              (coalesced).  Two groups of 8 warps take alternate chunks: the conversion is the
//                              longest stage (r2 cycle counters: 12 k of 18 k cycles per tile with one group)
//     warp  16    one thread   tcgen05.mma issue, tcgen05.commit
//     warps 17-20 epilogue     TMEM accumulators of tile i-1 (2 accumulator buffers) -> global
//     warp  21                 (MODE 0) the raw strip rows of tile i+1 by 1-D bulk copies (2 staged tiles), 4 rows per lane
//     warp  22    one thread   B stages (table slices) by bulk copy -- off the MMA thread, whose loop bounds modes 1 / 2
// so the load of one tile, the operand conversion and the MMAs of the next and the read-back of the previous overlap
// (the first version ran the phases one after the other in one-tile blocks: tensor pipe 6 % busy, issue 22 %; the second
// loaded the rows with ordinary loads in four warps and was bound by their latency).
constexpr int kCvWarps = 16;             // converter warps: two groups of 8, group g converts the chunks c = g (mod 2)
constexpr int kMmaWarp = kCvWarps;       // warp 16
constexpr int kEpiWarp0 = kCvWarps + 1;  // warps 17-20 (warp % 4 = 1, 2, 3, 0: one TMEM lane quadrant each)
constexpr int kLoadWarp = kCvWarps + 5;  // warp 21
constexpr int kBWarp = kCvWarps + 6;     // warp 22
constexpr int kTcThreads = (kCvWarps + 7) * 32;
constexpr int kTcSh = 1024;              // the long strip axis these kernels are built for (sb_tc_plan admits no other)
constexpr int kTcStages = 3;             // operand ring depth (A + B)
constexpr int kCvDepth = 6;              // MODE 1 / 2: chunks each converter warp keeps in flight (cp.async into its own ring)
constexpr int kCvRingBytes = kCvDepth * kCvWarps * 8 * 144;      // 108 KB: per warp and chunk 8 lines x (16 rows x 8 B + 16)
constexpr int kAStage = 4 * 128 * 32;    // part0_hi | part0_lo | part1_hi | part1_lo, each 128 rows x 8 k (K-major, LBO 2048, SBO 128)

struct TcSmem {                          // offsets into dynamic shared memory
    int a_off, b_off, stg_off, bar_off, total;
    int b_stage;                         // bytes of one B stage: 4 sub-tiles of NP rows x 8 k
};
__host__ __device__ inline TcSmem tc_smem_layout(int NP, int stg_bytes, int stg_bufs) {   // stg_bufs = 0: MODE 1
    TcSmem L;
    L.b_stage = 4 * NP * 32;
    L.a_off = 0;
    L.b_off = kTcStages * kAStage;
    L.stg_off = L.b_off + kTcStages * L.b_stage;
    L.bar_off = L.stg_off + stg_bufs * stg_bytes;
    L.total = L.bar_off + 256;
    return L;
}
// row pitch of a staged tile: a multiple of 16 bytes (bulk-copy destination) with an ODD number of 16-byte units, so that
// the rows a warp reads side by side spread over the banks (at most 2-way conflicts)
__host__ __device__ inline int tc_pitch(int payload_bytes) {
    int p = (payload_bytes + 14 + 15) & ~15;             // + up to 14 bytes of leading misalignment
    if (((p >> 4) & 1) == 0) p += 16;
    return p;
}

struct TcArgs {
    // geometry
    int Sh, n, NP, nchunks, swap;
    int n_tiles;                         // tiles of 128 rows in this launch
    int stg_bufs, acc_bufs;              // staged-tile buffers (MODE 0: 1 or 2), TMEM accumulator buffers (1 or 2)
    int stg_bytes;                       // bytes of one staged tile
    const uint8_t* Bmat;                 // operand images of this mode's tables
    int* fault;
    // MODE 0
    const PairDesc* pairs;
    const int2* mm;
    int tile_w, maxval;
    float2* Zh;
    int* nonzero;
    // MODE 1
    const float2* Y;                     // [pair][line][y], the first n/2 + 1 lines of every pair are read
    int lines_in;                        // lines per pair in Y
    CtaBest* best;                       // [pair][Sh / 128]
    float* rowmax;                       // [pair][Sh]
    // MODE 2 (upsampled-DFT rows): Y = the full cross-power R [pair][x][y], Bmat = per-pair twiddle images
    float2* Tm;                          // [pair][u][y], u < rs
    int rs;
    int half;                            // R holds the lines 0 .. n/2 only (lines_in = n/2 + 1)
};

#ifdef SB_TC_PROFILE
__device__ long long g_tc_prof[3][16];      // per mode: cycles block 0 spent per role / wait (see scratch/tc_profile.py)
#define TC_T0() const long long t0__ = clock64()
#define TC_ACC(slot) atomicAdd((unsigned long long*)&g_tc_prof[MODE][slot], (unsigned long long)(clock64() - t0__))
#else
#define TC_T0()
#define TC_ACC(slot)
#endif

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) xdft_tc_kernel(const TcArgs g) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int Sh = kTcSh;                   // (compile-time: every row / tile index below is a shift or a mask)
    const int n = g.n, NP = g.NP, nchunks = g.nchunks;
    const TcSmem L = tc_smem_layout(NP, g.stg_bytes, g.stg_bufs);      // (MODE 1 / 2: one "staged tile" = the converters' prefetch ring)
    uint8_t* a_st = smem + L.a_off;
    uint8_t* b_st = smem + L.b_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* stg_full = bars;                  // [2] bulk copies -> converters
    uint64_t* stg_empty = bars + 2;             // [2] converters -> loader
    uint64_t* ab_full = bars + 4;               // [kTcStages] converters + bulk copy -> MMA
    uint64_t* ab_empty = bars + 4 + kTcStages;  // [kTcStages] tcgen05.commit -> converters, bulk copy
    uint64_t* acc_full = bars + 4 + 2 * kTcStages;   // [2] tcgen05.commit -> epilogue
    uint64_t* acc_empty = acc_full + 2;         // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int4* tile_info = reinterpret_cast<int4*>(smem + L.bar_off + 128);     // [2] MODE 0: (min, max, row misalignment | stretch magic hi, stretch magic lo) of a staged tile
    __shared__ double s_val[4], s_sec[4];       // MODE 1: block reduction of the epilogue warps
    __shared__ int s_idx[4];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tiles = Sh >> 7;                  // tiles per strip image
    const int nb = n / 2 + 1, nh = n / 2, no = (n - 1) / 2;
    const int pitch_ns = tc_pitch(2 * n), pitch_sw = tc_pitch(256);      // staged row pitch: plain / transposed frame
    const int acc_cols = MODE == 2 ? NP : 2 * NP;
    const int kdim = MODE == 2 ? n : nb;                       // valid K indices of the converters' source (MODE 1 / 2)
    const uint32_t want_cols = (uint32_t)(g.acc_bufs * acc_cols);
    const uint32_t tmem_cols = want_cols <= 32 ? 32 : want_cols <= 64 ? 64 : want_cols <= 128 ? 128 : want_cols <= 256 ? 256 : 512;
    const int my_tiles = blockIdx.x < g.n_tiles ? (g.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int my_chunks = my_tiles * nchunks;

    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                umma::mbar_init(stg_full + i, 2);         // the loader's expect_tx arrival (+ the copies' bytes) and the tile record
                umma::mbar_init(stg_empty + i, kCvWarps); // every converter warp
                umma::mbar_init(acc_full + i, 1);
                umma::mbar_init(acc_empty + i, 4);
            }
            for (int s = 0; s < kTcStages; ++s) {
                umma::mbar_init(ab_full + s, 9);          // the 8 warps of one converter group + the bulk copy's expect_tx arrival
                umma::mbar_init(ab_empty + s, 1);         // tcgen05.commit
            }
            umma::mbar_init_fence();
        }
        __syncwarp();
        umma::tmem_alloc(tmem_slot, tmem_cols);
        umma::tmem_relinquish();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = *tmem_slot;
    bool ok = true;
#ifdef SB_TC_PROFILE
    const long long t_kernel0 = clock64();
#endif

    if (warp == kLoadWarp) {
        // =========================================================================== loader (MODE 0): one warp
        // (issuing a bulk copy costs the issuing thread ~130 cycles: one thread issuing the 128 rows of a tile bounded the
        // whole kernel at 17 k cycles per tile -- r2 cycle counters -- so the rows are spread over the 32 lanes)
        if (MODE == 0) {
            for (int it = 0; it < my_tiles && ok; ++it) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int mt = tile % tiles, img = (tile / tiles) & 1, p = tile / (2 * tiles);
                const int y0 = mt << 7;
                const int b = it % g.stg_bufs, u = it / g.stg_bufs;
                { TC_T0(); if (u >= 1) ok = umma::mbar_wait(stg_empty + b, (u - 1) & 1); if (blockIdx.x == 0 && lane == 0) TC_ACC(0); }
                uint8_t* stg = smem + L.stg_off + b * g.stg_bytes;
                const PairDesc pd = g.pairs[p];
                const uint16_t* src = img ? pd.b : pd.a;
                // rows start 16-byte aligned in shared memory; the copies start at the 16-byte boundary below the strip
                // pixel (the same offset a0 for every row: the tile pitch is a multiple of 16 bytes)
                const uint8_t* row0 = reinterpret_cast<const uint8_t*>(g.swap ? src + y0 : src + (size_t)y0 * g.tile_w);
                const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(row0) & 15);
                if (!g.swap) {
                    const uint32_t bytes = (a0 + 2u * (uint32_t)n + 15u) & ~15u;
                    if (lane == 0) umma::mbar_expect_tx(stg_full + b, bytes * 128u);
                    for (int r = lane; r < 128; r += 32)
                        umma::bulk_g2s(stg + r * pitch_ns, row0 - a0 + (size_t)r * g.tile_w * 2, bytes, stg_full + b);
                } else {
                    const uint32_t bytes = (a0 + 256u + 15u) & ~15u;
                    if (lane == 0) umma::mbar_expect_tx(stg_full + b, bytes * (uint32_t)n);
                    for (int x = lane; x < n; x += 32)
                        umma::bulk_g2s(stg + x * pitch_sw, row0 - a0 + (size_t)x * g.tile_w * 2, bytes, stg_full + b);
                }
                // the tile's record for the converters -- two DEPENDENT global loads (descriptor -> min/max table) that used
                // to open every tile of every converter thread; here they ride on the loader, a tile ahead
                if (lane == 1) {
                    const int2 m = g.mm[img ? pd.b_tile : pd.a_tile];
                    // (+ the 48-bit stretch magic of the tile, one 64-bit divide here instead of one per converter thread)
                    const StretchMagic sm = stretch_magic(m.y > m.x ? (unsigned)(m.y - m.x) : 1u, (unsigned)g.maxval);
                    tile_info[b] = make_int4(m.x, m.y, (int)(a0 | (sm.hi << 4)), (int)sm.lo);
                    umma::mbar_arrive(stg_full + b);               // (release: the record is visible to whoever sees the phase)
                }
            }
        }
    } else if (warp == kBWarp) {
        // =========================================================================== B stages (whole warp, one elected lane issues)
        int s = 0;
        uint32_t ph = 1;                                           // parity of the PREVIOUS use of the stage
        int c = 0, it = 0;
        for (int gc = 0; gc < my_chunks && ok; ++gc) {
            { TC_T0(); if (gc >= kTcStages) ok = umma::mbar_wait(ab_empty + s, ph); if (blockIdx.x == 0 && lane == 0) TC_ACC(5); }
            if (umma::elect_one()) {
                const int t2 = blockIdx.x + it * gridDim.x;                      // MODE 2: the tables belong to the tile's pair
                const size_t img = (MODE == 2 ? (size_t)(t2 / tiles) * nchunks : 0) + (size_t)c;
                umma::mbar_expect_tx(ab_full + s, (uint32_t)L.b_stage);
                umma::bulk_g2s(b_st + s * L.b_stage, g.Bmat + img * L.b_stage, (uint32_t)L.b_stage, ab_full + s);
            }
            __syncwarp();
            if (++s == kTcStages) { s = 0; ph ^= 1u; }
            if (++c == nchunks) { c = 0; ++it; }
        }
    } else if (warp < kCvWarps) {
        // =========================================================================== converters: thread = (row, 4 k)
        const int grp = warp >> 3;                                 // this group converts the chunks c = grp (mod 2)
        const int row = 16 * (warp & 7) + (lane >> 1), kg = lane & 1;
        // MODE 1 / 2: the source elements of a warp -- its 16 rows of the 8 lines of a chunk, Y or R stored [line][y] -- are
        // fetched kCvDepth chunks ahead by cp.async into a ring only this warp reads back (wait_group + __syncwarp, no
        // block barrier): 16-byte copies that BYPASS L1 (two adjacent rows of one line each; lane = line + 8 * row pair).
        // History (r2): held in registers, 2 chunks ahead, the loads bounded the upsampled-DFT stage (63 % of its stall
        // samples on the long scoreboard); 8-byte cp.async.ca did not help either -- with 220 KB of the SM's 256 KB carved
        // out as shared memory the L1 holds only ~28 KB of lines in flight, both forms ran at 7.5 B / cycle / SM.
        // Mirrored lines (MODE 2 on a HALF array, k > n/2: R[y][n-kx] = conj(R[-y][kx])) run backwards in memory, so their
        // row pairs are not 16-byte aligned: those stay 8-byte copies of each thread's own elements.
        constexpr int kLinePitch = 144;                            // bytes: 16 rows x 8 + 16 (bank spread for the read-back)
        constexpr int kWarpRing = 8 * kLinePitch;                  // one chunk of one warp
        uint8_t* ring = smem + L.stg_off + (size_t)warp * (kCvDepth * kWarpRing);
        const int cl = lane & 7, crp = lane >> 3;                  // copy role: line of the chunk, row pair (and row pair + 4)
        int is_it = 0, is_c = grp, is_m = 0;                       // the next chunk to fetch: tile, chunk, ring slot
        const float2 *is_p = nullptr, *is_pm = nullptr;            // direct: (line 8 c + cl, rows 2 crp ..); mirrored: this thread's own
        auto issue_tile = [&]() {                                  // pointers for the first chunk of tile is_it
            const int tile2 = blockIdx.x + is_it * gridDim.x;
            const int yw = ((tile2 % tiles) << 7) + 16 * (warp & 7);
            const float2* base = g.Y + (size_t)(tile2 / tiles) * g.lines_in * Sh;
            is_p = base + (size_t)(8 * grp + cl) * Sh + yw + 2 * crp;
            is_pm = base + (ptrdiff_t)(n - (8 * grp + 4 * kg)) * Sh + ((Sh - (yw + (lane >> 1))) & (Sh - 1));
        };
        auto issue_next = [&]() {
            if (is_it < my_tiles) {
                uint8_t* dst = ring + is_m * kWarpRing;
                const int kd = 8 * is_c + cl;                      // direct copy: two 16-byte pieces of line kd
                if (kd < kdim && !(MODE == 2 && g.half && kd >= nb)) {
                    umma::cp_async16(dst + cl * kLinePitch + crp * 16, is_p);
                    umma::cp_async16(dst + cl * kLinePitch + (crp + 4) * 16, is_p + 8);
                }
                if (MODE == 2 && g.half) {
                    const int k0 = 8 * is_c + 4 * kg;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k0 + i >= nb && k0 + i < kdim)
                            umma::cp_async8(dst + (4 * kg + i) * kLinePitch + (lane >> 1) * 8, is_pm - i * Sh);
                }
                if (++is_m == kCvDepth) is_m = 0;
                is_c += 2;
                is_p += 16 * Sh;
                is_pm -= 16 * Sh;
                if (is_c >= nchunks) {
                    is_c = grp;
                    if (++is_it < my_tiles) issue_tile();
                }
            }
            umma::cp_async_commit();                               // (an empty group keeps the count in step at the tail)
        };
        int rd_m = 0;                                              // ring slot of the chunk being converted
        if (MODE != 0) {
            if (my_tiles > 0) issue_tile();
#pragma unroll 1
            for (int d = 0; d < kCvDepth; ++d) issue_next();
        }
        for (int it = 0; it < my_tiles && ok; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int mt = tile % tiles;
            const int img = MODE == 0 ? (tile / tiles) & 1 : 0;
            const int p = MODE == 0 ? tile / (2 * tiles) : tile / tiles;
            const int y0 = mt << 7;
            const int b = MODE == 0 ? it % g.stg_bufs : 0, u = MODE == 0 ? it / g.stg_bufs : 0;
            const uint16_t* sbase = nullptr;                       // MODE 0: first staged pixel of this thread's row
            int sstep = 0;                                         // ... and the distance (uint16 elements) between its columns
            int mn = 0, mx = 0, seen = 0, smode = 0;
            unsigned sb_ = 1u;
            float inv = 0.f;
            StretchMagic smagic = {0u, 0u};
            if (MODE == 0) {
                { TC_T0(); ok = umma::mbar_wait(stg_full + b, u & 1); if (blockIdx.x == 0 && t == 0) TC_ACC(1); }
                const int4 ti = tile_info[b];
                mn = ti.x;
                mx = ti.y;
                inv = mx > mn ? (float)g.maxval / (float)(mx - mn) : 0.f;
                sb_ = mx > mn ? (unsigned)(mx - mn) : 1u;
                smagic = {(unsigned)ti.w, (unsigned)ti.z >> 4};
                smode = mx <= mn ? 0 : (sb_ == (unsigned)g.maxval ? 1 : 2);     // constant tile / full range (identity) / general
                const uint8_t* stg = smem + L.stg_off + b * g.stg_bytes;
                const uint32_t a0 = (uint32_t)ti.z & 15u;
                if (!g.swap) {
                    sbase = reinterpret_cast<const uint16_t*>(stg + row * pitch_ns + a0);
                    sstep = 1;
                } else {
                    sbase = reinterpret_cast<const uint16_t*>(stg + a0) + row;
                    sstep = pitch_sw >> 1;
                }
            }
            for (int c = grp; c < nchunks && ok; c += 2) {
                const int gc = it * nchunks + c;                   // position of this chunk in the operand ring
                const int s = gc % kTcStages, use = gc / kTcStages;
                float ph[4], pl[4], qh[4], ql[4];
#ifdef SB_TC_PROFILE
                const long long t_conv0 = clock64();
#endif
#ifdef SB_TC_EXP_NOCV                                                          // (timing experiment: no operand conversion)
                if (true) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) ph[i] = pl[i] = qh[i] = ql[i] = 0.f;
                } else
#endif
                if (MODE == 0 && smode == 2 && c >= 1 && 8 * c + 7 <= no) {
                    // interior chunk of a general tile (12 of the 14 chunks of a 214-wide strip): every j has both fold
                    // partners, no guards; stretch by a 48-bit multiply-high (stretch_mulhi: 2 instructions per pixel + the
                    // exact-quotient check; the float-estimate forms before it cost 16-19)
                    unsigned ra[4], rb[4], ba[4], bb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = 8 * c + 4 * kg + i;
                        ra[i] = sbase[j * sstep];
                        rb[i] = sbase[(n - j) * sstep];
                    }
                    bool exact = false;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ba[i] = stretch_mulhi(ra[i] - (unsigned)mn, sb_, smagic, (unsigned)g.maxval, exact);
                        bb[i] = stretch_mulhi(rb[i] - (unsigned)mn, sb_, smagic, (unsigned)g.maxval, exact);
                    }
                    if (exact) {                                   // (rare) an exact quotient somewhere: the float64 expression decides
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            ba[i] = (unsigned)stretch_px(ra[i], mn, mx, inv, g.maxval);
                            bb[i] = (unsigned)stretch_px(rb[i], mn, mx, inv, g.maxval);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const unsigned es = ba[i] + bb[i];                               // a + b       in [0, 2^17)
                        const unsigned os = ba[i] - bb[i] + 0x400000u;                   // a - b + 2^22 in (0, 2^23)
                        seen |= (int)es;
                        const float e = __uint_as_float(kStretchMagic | es) - 8388608.0f;
                        const float o = __uint_as_float(kStretchMagic | os) - 12582912.0f;
                        umma::split_tf32(e, ph[i], pl[i]);
                        umma::split_tf32(o, qh[i], ql[i]);
                    }
                } else if (MODE == 0) {
                    unsigned ra[4], rb[4];
                    bool va[4], vb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                  // raw pixels first (shared-memory latency overlaps)
                        const int j = 8 * c + 4 * kg + i;
                        va[i] = j <= nh;
                        vb[i] = j >= 1 && j <= no;
                        ra[i] = va[i] ? sbase[j * sstep] : (unsigned)mn;
                        rb[i] = vb[i] ? sbase[(n - j) * sstep] : (unsigned)mn;
                    }
                    // normalize_image's stretch (:844-855), branch-free for the whole vector; tile-uniform special cases
                    int xa[4], xb[4];
                    if (smode == 2) {
                        bool exact = false;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            xa[i] = stretch_core(ra[i] - (unsigned)mn, sb_, inv, (unsigned)g.maxval, exact);
                            xb[i] = stretch_core(rb[i] - (unsigned)mn, sb_, inv, (unsigned)g.maxval, exact);
                        }
                        if (exact) {                               // (rare) an exact quotient somewhere: the float64 expression decides
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                xa[i] = stretch_px(ra[i], mn, mx, inv, g.maxval);
                                xb[i] = stretch_px(rb[i], mn, mx, inv, g.maxval);
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {              // full-range tile: identity; constant tile: zero
                            xa[i] = smode == 1 ? (int)ra[i] - mn : 0;
                            xb[i] = smode == 1 ? (int)rb[i] - mn : 0;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int a_ = va[i] ? xa[i] : 0, b_ = vb[i] ? xb[i] : 0;
                        seen |= a_ | b_;
                        const float e = (float)(a_ + b_), o = vb[i] ? (float)(a_ - b_) : 0.f;      // (the 2^-16 input scaling is in the tables)
                        umma::split_tf32(e, ph[i], pl[i]);
                        umma::split_tf32(o, qh[i], ql[i]);
                    }
                } else {
                    umma::cp_async_wait<kCvDepth - 1>();             // every lane's copies for this chunk have landed ...
                    __syncwarp();                                  // ... and are visible to the whole warp
                    const uint8_t* src = ring + rd_m * kWarpRing + (4 * kg) * kLinePitch + (lane >> 1) * 8;
                    float2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = 8 * c + 4 * kg + i;
                        v[i] = k < kdim ? *reinterpret_cast<const float2*>(src + i * kLinePitch) : make_float2(0.f, 0.f);
                        if (MODE == 2 && g.half && k >= nb) v[i].y = -v[i].y;
                    }
                    __syncwarp();                                  // all reads of the slot done before any lane refills it
                    if (++rd_m == kCvDepth) rd_m = 0;
                    issue_next();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        umma::split_tf32(v[i].x, ph[i], pl[i]);
                        umma::split_tf32(v[i].y, qh[i], ql[i]);
                    }
                }
#ifdef SB_TC_PROFILE
                if (blockIdx.x == 0 && t == 0) atomicAdd((unsigned long long*)&g_tc_prof[MODE][10], (unsigned long long)(clock64() - t_conv0));
#endif
                { TC_T0(); if (use >= 1) ok = umma::mbar_wait(ab_empty + s, (use - 1) & 1); if (blockIdx.x == 0 && t == 0) TC_ACC(2); }
                uint8_t* dst = a_st + s * kAStage + kg * 2048 + row * 16;
                *reinterpret_cast<float4*>(dst) = make_float4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<float4*>(dst + 4096) = make_float4(pl[0], pl[1], pl[2], pl[3]);
                *reinterpret_cast<float4*>(dst + 8192) = make_float4(qh[0], qh[1], qh[2], qh[3]);
                *reinterpret_cast<float4*>(dst + 12288) = make_float4(ql[0], ql[1], ql[2], ql[3]);
#ifdef SB_TC_PROFILE
                const long long t_f0 = clock64();
#endif
                umma::fence_smem_to_async();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(ab_full + s);
#ifdef SB_TC_PROFILE
                if (blockIdx.x == 0 && t == 0) atomicAdd((unsigned long long*)&g_tc_prof[MODE][11], (unsigned long long)(clock64() - t_f0));
#endif
            }
            if (MODE == 0) {
                // an all-zero strip has an exactly zero spectrum in the reference: record whether this one has a non-zero pixel
                seen = __reduce_or_sync(0xffffffffu, (unsigned)seen);
                if (lane == 0) {
                    if (seen) atomicOr(&g.nonzero[p], img ? 2 : 1);
                    umma::mbar_arrive(stg_empty + b);              // every read of the staged tile is done
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // =========================================================================== MMA issue
        // The WHOLE warp runs the loop (uniform control flow, descriptors in uniform registers) and one elected lane
        // issues: with the loop inside `if (lane == 0)` the compiler wrapped every tcgen05 instruction in an election
        // loop and moved each descriptor through R2UR -- ~100 dependent instructions, ~900 cycles per K chunk, the floor
        // of all three modes (r2 cycle counters; the MMAs themselves take 6 x 60 cycles).
        const uint32_t idesc = umma::idesc_tf32(128, NP);
        const uint32_t bsub16 = ((uint32_t)NP * 32) >> 4, blbo = (uint32_t)NP * 16;
        // descriptors of stage 0; the 14-bit address field (16-byte units) advances by the stage size
        const uint64_t dA0 = umma::desc_kmajor(umma::smem_addr(a_st), 2048, 128);
        const uint64_t dB0 = umma::desc_kmajor(umma::smem_addr(b_st), blbo, 128);
        const uint32_t a_step = kAStage >> 4, b_step = (uint32_t)L.b_stage >> 4;
        int s = 0;
        uint32_t ph = 0;                                           // stage and its phase parity in the operand ring
        for (int it = 0; it < my_tiles && ok; ++it) {
            const int acc = it % g.acc_bufs, ua = it / g.acc_bufs;
            if (ua >= 1) {
                { TC_T0(); ok = umma::mbar_wait(acc_empty + acc, (ua - 1) & 1); if (blockIdx.x == 0 && lane == 0) TC_ACC(3); }
                umma::fence_after_sync();
            }
            const uint32_t d0 = tb + (uint32_t)(acc * acc_cols);
            const uint32_t d1 = MODE == 2 ? d0 : d0 + NP;          // MODE 2: both parts feed one accumulator
            for (int c = 0; c < nchunks && ok; ++c) {
                { TC_T0(); ok = umma::mbar_wait(ab_full + s, ph); if (blockIdx.x == 0 && lane == 0) TC_ACC(4); }
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint64_t eh = dA0 + (uint64_t)(s * a_step), ch = dB0 + (uint64_t)(s * b_step);
                    const uint64_t el = eh + (4096 >> 4), oh = eh + (8192 >> 4), ol = eh + (12288 >> 4);
                    const uint64_t cl = ch + bsub16, sh = ch + 2 * bsub16, sl = ch + 3 * bsub16;
                    const uint32_t accum = c > 0 ? 1u : 0u;
                    umma::mma_tf32(d0, eh, ch, idesc, accum);
#ifndef SB_TC_EXP_MMA1                                                         // (timing experiment: one product per part)
                    umma::mma_tf32(d0, el, ch, idesc, 1);
                    umma::mma_tf32(d0, eh, cl, idesc, 1);
#endif
                    umma::mma_tf32(d1, oh, sh, idesc, MODE == 2 ? 1u : accum);
#ifndef SB_TC_EXP_MMA1
                    umma::mma_tf32(d1, ol, sh, idesc, 1);
                    umma::mma_tf32(d1, oh, sl, idesc, 1);
#endif
                    umma::mma_commit(ab_empty + s);
                    if (c == nchunks - 1) umma::mma_commit(acc_full + acc);
                }
                __syncwarp();
                if (++s == kTcStages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // =========================================================================== epilogue: one row per thread
        const int q = warp & 3;                                    // TMEM lane quadrant this warp may read
        const int ew = warp - kEpiWarp0;
        for (int it = 0; it < my_tiles && ok; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int mt = tile % tiles;
            const int img = MODE == 0 ? (tile / tiles) & 1 : 0;
            const int p = MODE == 0 ? tile / (2 * tiles) : tile / tiles;
            const int y0 = mt << 7;
            const int acc = it % g.acc_bufs, ua = it / g.acc_bufs;
            { TC_T0(); ok = umma::mbar_wait(acc_full + acc, ua & 1); if (blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0) TC_ACC(6); }
            umma::fence_after_sync();
            const uint32_t trow = tb + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * acc_cols);
            const int y = y0 + 32 * q + lane;
#ifdef SB_TC_EXP_NOEPI                                                         // (timing experiment: no read-back)
            if (false) {
#else
            if (MODE == 0) {
#endif
                if (ok) {
                    float2* zp = g.Zh + ((size_t)(p * 2 + img) * nb) * Sh + y;
                    for (int k0 = 0; k0 < NP; k0 += 16) {            // (NP is a multiple of 16)
                        uint32_t re[16], im[16];
                        umma::tmem_ld16(trow + k0, re);
                        umma::tmem_ld16(trow + NP + k0, im);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (k0 + i < nb) zp[(size_t)(k0 + i) * Sh] = make_float2(__uint_as_float(re[i]), __uint_as_float(im[i]));
                    }
                }
#ifdef SB_TC_EXP_NOEPI
            } else if (true) {
#endif
            } else if (MODE == 2) {
                if (ok) {
                    float2* tp = g.Tm + (size_t)p * g.rs * Sh + y;
                    uint32_t re[16], im[16];
                    umma::tmem_ld8(trow, *reinterpret_cast<uint32_t(*)[8]>(re));
                    umma::tmem_ld8(trow + 8, *reinterpret_cast<uint32_t(*)[8]>(re + 8));
                    umma::tmem_ld8(trow + 16, *reinterpret_cast<uint32_t(*)[8]>(im));
                    umma::tmem_ld8(trow + 24, *reinterpret_cast<uint32_t(*)[8]>(im + 8));
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (u < g.rs) tp[(size_t)u * Sh] = make_float2(__uint_as_float(re[u]), __uint_as_float(im[u]));
                }
            } else {
                // |cc| of this thread's row: first maximum (ties -> lowest index in the C order of the ORIGINAL strip),
                // second-largest value, row maximum -- what rows_inv_argmax_kernel of the radix path emits
                // Branch-free, two independent chains (the data-dependent branches of a running top-2 cost ~30 k cycles per
                // tile and made this epilogue the bound of the whole kernel -- r2 cycle counters): the pixels x = 0 .. n/2 come
                // from P + Q in ascending order (a strict > keeps the first of equal maxima), the pixels n - x from P - Q in
                // DESCENDING order (>= keeps the lowest); every pixel of the first chain precedes every pixel of the second.
                // s = max(s, min(m, v)) before m = max(m, v) keeps the second-largest value of the multiset.
                float mA = -1.f, sA = -1.f, mB = -1.f, sB = -1.f;
                int xA = 0, xB = 0;
                if (ok) {
                    for (int x0 = 0; x0 < NP; x0 += 16) {
                        uint32_t pr[16], qr[16];
                        umma::tmem_ld16(trow + x0, pr);
                        umma::tmem_ld16(trow + NP + x0, qr);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int x = x0 + i;
                            const float P = __uint_as_float(pr[i]), Q = __uint_as_float(qr[i]);
                            const float va = x <= nh ? fabsf(P + Q) : -2.f;
                            const float vb = (x >= 1 && x <= no) ? fabsf(P - Q) : -2.f;
                            sA = fmaxf(sA, fminf(mA, va));
                            xA = va > mA ? x : xA;
                            mA = fmaxf(mA, va);
                            sB = fmaxf(sB, fminf(mB, vb));
                            xB = vb >= mB ? x : xB;
                            mB = fmaxf(mB, vb);
                        }
                    }
                }
                const bool takeB = mB > mA;                            // equal maxima: the first chain holds the lower pixel
                float bv = takeB ? mB : mA;
                float b2 = fmaxf(fmaxf(sA, sB), fminf(mA, mB));
                const int xw = takeB ? n - xB : xA;
                int bi = ok ? (g.swap ? xw * Sh + y : y * n + xw) : 0x7fffffff;
                const float rm = fmaxf(bv, 0.f);
                const double sc = 1.0 / ((double)Sh * (double)n);
                g.rowmax[(size_t)p * Sh + y] = (float)((double)rm * sc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const float o2 = __shfl_xor_sync(0xffffffffu, b2, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    top2_merge<float>(bv, bi, b2, ov, oi, o2);
                }
                if (lane == 0) { s_val[ew] = (double)bv; s_sec[ew] = (double)b2; s_idx[ew] = bi; }
                asm volatile("bar.sync 2, 128;" ::: "memory");         // the four epilogue warps
                if (ew == 0 && lane == 0) {
                    double bb = s_val[0], c2 = s_sec[0];
                    int i0 = s_idx[0];
                    for (int w = 1; w < 4; ++w) top2_merge<double>(bb, i0, c2, s_val[w], s_idx[w], s_sec[w]);
                    CtaBest& o = g.best[(size_t)p * tiles + mt];
                    o.val = bb * sc;
                    o.second = c2 > 0.0 ? c2 * sc : 0.0;
                    o.idx = i0;
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");         // s_val is reused by the next tile
            }
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(acc_empty + acc);
        }
    }
    if (!ok && lane == 0) atomicExch(g.fault, 1);
    umma::fence_before_sync();
    __syncthreads();
#ifdef SB_TC_PROFILE
    if (blockIdx.x == 0 && t == 0) {
        atomicAdd((unsigned long long*)&g_tc_prof[MODE][7], (unsigned long long)(clock64() - t_kernel0));
        atomicAdd((unsigned long long*)&g_tc_prof[MODE][8], (unsigned long long)my_tiles);
        atomicAdd((unsigned long long*)&g_tc_prof[MODE][9], 1ull);
    }
#endif
    if (warp == kMmaWarp) umma::tmem_dealloc(tb, tmem_cols);
}

// Twiddles of skimage's _upsampled_dft for the tensor-core rows stage (MODE 2) and the column stage:
//   Ex[u][x] = exp(-2 pi i (u - off_x) fftfreq(n, uf)[x]),  off = dftshift - shift * uf   (shift = wrapped coarse peak)
// (u - off) and n * uf * fftfreq are integers, so the phase is reduced exactly in integer arithmetic before sincospi.
// Ex goes straight into the per-pair B operand image of xdft_tc_kernel<2> -- per chunk of 8 x: [B0_hi | B0_lo | B1_hi |
// B1_lo], 32 rows (output columns: u = real part, 16 + u = imaginary part) x 8, K-major; B0 = [Ex_re | Ex_im] multiplies
// Re R, B1 = [Ex_im | -Ex_re] multiplies Im R.  Ey[v][y] is written as plain complex rows for updft_cols_kernel.
__global__ void __launch_bounds__(256) updft_tables_tc_kernel(const PeakOut* __restrict__ peaks, int Sh, int n, int uf, int rs,
                                                              int dftshift, int nchunks, float* __restrict__ Bimg,
                                                              float2* __restrict__ Ey) {
    const int p = blockIdx.y;
    const PeakOut pk = peaks[p];
    const int cy = pk.coarse_y > Sh / 2 ? pk.coarse_y - Sh : pk.coarse_y;     // shift[shift > fix(n/2)] -= n
    const int cx = pk.coarse_x > n / 2 ? pk.coarse_x - n : pk.coarse_x;
    const int npad = nchunks * 8;
    const int total = 16 * npad + rs * Sh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const bool isx = i < 16 * npad;
        const int ii = isx ? i : i - 16 * npad;
        const int len = isx ? npad : Sh, nn = isx ? n : Sh;
        const int u = ii / len, k = ii - u * len;
        double s = 0.0, c = 0.0;
        if (u < rs && k < nn) {
            const long long m = (long long)u - dftshift + (long long)(isx ? cx : cy) * uf;
            const int sk = (k < (nn + 1) / 2) ? k : k - nn;                  // n * fftfreq(n)[k]
            const long long den = (long long)nn * uf;
            long long num = (m * sk) % den;
            if (num < 0) num += den;
            sincospi(-2.0 * (double)num / (double)den, &s, &c);
        }
        if (!isx) {
            Ey[((size_t)p * rs + u) * Sh + k] = make_float2((float)c, (float)s);
            continue;
        }
        float rh, rl, ih, il, nh_, nl_;
        umma::split_tf32((float)c, rh, rl);
        umma::split_tf32((float)s, ih, il);
        umma::split_tf32(-(float)c, nh_, nl_);
        float* img = Bimg + ((size_t)p * nchunks + (k >> 3)) * (4 * 32 * 8);      // 4 sub-tiles of 32 rows x 8 k
        const int e_re = ((k & 7) >> 2) * (32 * 4) + u * 4 + (k & 3), e_im = e_re + 16 * 4;
        img[e_re] = rh;               img[32 * 8 + e_re] = rl;                    // B0: column u      <- Ex_re
        img[e_im] = ih;               img[32 * 8 + e_im] = il;                    //     column 16 + u <- Ex_im
        img[2 * 32 * 8 + e_re] = ih;  img[3 * 32 * 8 + e_re] = il;                // B1: column u      <- Ex_im
        img[2 * 32 * 8 + e_im] = nh_; img[3 * 32 * 8 + e_im] = nl_;               //     column 16 + u <- -Ex_re
    }
}

// ==========================================================================================================
// T2  cols_warp_kernel: the column pass on lines of 1024 -- FFT of the two half spectra, normalised cross-power,
// inverse FFT -- one warp per column, everything between the global load and the global stores in registers
// (fft_warp.cuh).  Only the n/2 + 1 columns of the half spectrum are transformed; the mirrored columns of the full
// arrays the radix kernels downstream read (R for the upsampled DFT, Y for the inverse row pass) are their
// conjugates: R[ky][n-kx] = conj(R[-ky][kx]), Y[y][n-kx] = conj(Y[y][kx]).
// ==========================================================================================================
struct ColsSmem {
    wfft::WarpBuf buf;
    float park_re[1024];
    float park_im[1024];
};

__global__ void __launch_bounds__(128) cols_warp_kernel(int n_cols, int nb, int n, int lines_out, int mirror,
                                                         const float2* __restrict__ tw_g, const float2* __restrict__ Zh,
                                                         float2* __restrict__ Rout, float2* __restrict__ Yout) {
    extern __shared__ __align__(128) uint8_t smem[];
    float2* tw = reinterpret_cast<float2*>(smem);                                     // [k2][l]: W1024^(l k2)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ColsSmem& sm = *reinterpret_cast<ColsSmem*>(smem + 1024 * sizeof(float2) + (size_t)warp * sizeof(ColsSmem));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tw[i] = tw_g[i];
    __syncthreads();
    constexpr int N = 1024;
    // max(|P|, 100 eps) with the inputs scaled by 2^-16: the clamp (5.2e-24) squared underflows float32, so it only
    // catches |P|^2 == 0 -- an exactly zero product, which the division by the clamp leaves at zero
    const float clamp2 = 0.f;
    for (int col = blockIdx.x * 4 + warp; col < n_cols; col += gridDim.x * 4) {
        const int p = col / nb, kx = col - p * nb;
        const float2* la = Zh + ((size_t)(p * 2) * nb + kx) * N;
        const float2* lb = Zh + ((size_t)(p * 2 + 1) * nb + kx) * N;
        float2 x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = la[lane + 32 * j];
        wfft::fft1024<false>(x, sm.buf, tw, lane);
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            sm.park_re[s * 32 + lane] = x[s].x;
            sm.park_im[s * 32 + lane] = x[s].y;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = lb[lane + 32 * j];
        wfft::fft1024<false>(x, sm.buf, tw, lane);
        // R = A conj(B) / max(|A conj(B)|, 100 eps): slot s holds ky = lane + 32 brev5(s) in both transforms
        float2* rl = Rout + ((size_t)p * lines_out + kx) * N;
        const bool inner = kx >= 1 && kx <= (n - 1) / 2;
        const bool mir = (mirror & 1) && inner;                                      // R: read by the upsampled-DFT stage
        const bool mir_y = (mirror & 2) && inner;                                    // Y: only the radix inverse reads the mirrored lines
        float2* rm = Rout + ((size_t)p * lines_out + (n - kx)) * N;
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            const float ax = sm.park_re[s * 32 + lane], ay = sm.park_im[s * 32 + lane];
            const float bx = x[s].x, by = x[s].y;
            float px = ax * bx + ay * by, py = ay * bx - ax * by;
            const float m2 = px * px + py * py;
            const float inv = m2 > clamp2 ? rsqrtf(m2) : (float)(1.0 / kClamp);
            px *= inv;
            py *= inv;
            x[s] = make_float2(px, py);
            const int ky = lane + 32 * wfft::brev5(s);
            rl[ky] = x[s];
            if (mir) rm[(N - ky) & (N - 1)] = make_float2(px, -py);
        }
        float2 y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = x[wfft::brev5(j)];                      // natural order for the inverse
        wfft::fft1024<true>(y, sm.buf, tw, lane);
        float2* yl = Yout + ((size_t)p * lines_out + kx) * N;
        float2* ym = Yout + ((size_t)p * lines_out + (n - kx)) * N;
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            const int yy = lane + 32 * wfft::brev5(s);
            yl[yy] = y[s];
            if (mir_y) ym[yy] = make_float2(y[s].x, -y[s].y);
        }
    }
}


// ------------------------------------------------------------------------------------------ self-test (test hook)
// D[128 x N] = A[128 x K] . B[N x K]^T through the exact device functions the registration kernels use (descriptor
// encoding, operand layout, 3-term split, TMEM read-back), checked on the device against a float64 product.
__device__ __forceinline__ float st_val(unsigned r, unsigned k, unsigned salt) {
    unsigned h = (r * 2654435761u) ^ (k * 40503u + salt * 2246822519u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return (float)(int)(h & 0xffffff) * (1.0f / 8388608.0f) - 1.0f;          // [-1, 1), 24 significant bits
}

template <int N, int K>
__global__ void __launch_bounds__(128) selftest_umma_kernel(int variant, float* __restrict__ D, unsigned long long* __restrict__ res) {
    constexpr int KC = K / 4;                               // K chunks of 4 elements (16 bytes)
    constexpr uint32_t LBO_A = 128 * 16 + 16, LBO_B = N * 16 + 16, SBO = 128;   // padded chunk pitch: conflict-free column stores too
    extern __shared__ __align__(128) uint8_t smem[];
    float* a_hi = reinterpret_cast<float*>(smem);
    float* a_lo = reinterpret_cast<float*>(smem + KC * LBO_A);
    float* b_hi = reinterpret_cast<float*>(smem + 2 * KC * LBO_A);
    float* b_lo = reinterpret_cast<float*>(smem + 2 * KC * LBO_A + KC * LBO_B);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int k = 0; k < K; ++k) {
        float hi, lo;
        umma::split_tf32(st_val(t, k, 1), hi, lo);
        const uint32_t off = (k >> 2) * LBO_A + t * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_hi) + off) = hi;
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_lo) + off) = lo;
        if (t < N) {
            umma::split_tf32(st_val(t, k, 2), hi, lo);
            const uint32_t offb = (k >> 2) * LBO_B + t * 16 + (k & 3) * 4;
            *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(b_hi) + offb) = hi;
            *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(b_lo) + offb) = lo;
        }
    }
    if (t == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_init_fence();
    }
    if (warp == 0) {
        umma::tmem_alloc(&tmem_base, 128);
        umma::tmem_relinquish();
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = tmem_base;
    if (t == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t lboa = variant == 1 ? SBO : LBO_A, sboa = variant == 1 ? LBO_A : SBO;     // variant 1: swapped roles (diagnosis)
        const uint32_t lbob = variant == 1 ? SBO : LBO_B, sbob = variant == 1 ? LBO_B : SBO;
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t ao = ks * 2 * LBO_A, bo = ks * 2 * LBO_B;
            const uint64_t ah = umma::desc_kmajor(umma::smem_addr(a_hi) + ao, lboa, sboa), al = umma::desc_kmajor(umma::smem_addr(a_lo) + ao, lboa, sboa);
            const uint64_t bh = umma::desc_kmajor(umma::smem_addr(b_hi) + bo, lbob, sbob), bl = umma::desc_kmajor(umma::smem_addr(b_lo) + bo, lbob, sbob);
            umma::mma_tf32(tb, ah, bh, idesc, acc);
            acc = 1;
            if (variant != 2) {                              // variant 2: single tf32 product (shows what the split buys)
                umma::mma_tf32(tb, al, bh, idesc, 1);
                umma::mma_tf32(tb, ah, bl, idesc, 1);
            }
        }
        umma::mma_commit(&bar);
    }
    const bool ok = umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (!ok) {
        if (t == 0) atomicExch(res + 3, 0xDEADull);
    } else {
        for (int c = 0; c < N; c += 16) {
            uint32_t v[16];
            umma::tmem_ld16(tb + ((uint32_t)(32 * warp) << 16) + c, v);
            umma::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) D[(size_t)(32 * warp + lane) * N + c + i] = __uint_as_float(v[i]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 128);
}

template <int N, int K>
__global__ void __launch_bounds__(128) selftest_umma_check_kernel(const float* __restrict__ D, double tol, unsigned long long* __restrict__ res) {
    const int m = blockIdx.x, n = threadIdx.x;
    if (n >= N) return;
    double ref = 0.0;
    for (int k = 0; k < K; ++k) ref += (double)st_val(m, k, 1) * (double)st_val(n, k, 2);
    const double err = fabs((double)D[(size_t)m * N + n] - ref);
    atomicMax(res + 2, (unsigned long long)(err * 1e12));           // largest error in units of 1e-12
    if (!(err <= tol)) atomicAdd(res + 1, 1ull);
}

// Issue-rate probe (tuning hook, timing only -- the operands are zeros): `rounds` x 6 tcgen05.mma in the product's pattern
// (two accumulators, A / B sub-tiles of a K chunk of 8) from one thread, one commit at the end; cycles from the first issue
// to the arrival.  layout 0 = the product's K-major no-swizzle form (LBO = rows * 16, SBO = 128), 1 = the same with the
// 128-byte-swizzle bit set in the descriptors (row pitch 128 B, SBO = 1024).
__global__ void __launch_bounds__(128) umma_rate_kernel(int N, int layout, int rounds, int per_round, int commit_mode,
                                                        unsigned long long* __restrict__ res) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, dummy;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (64 + 64) * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (t == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_init(&dummy, 1u << 20);                 // commit_mode 1: a commit per round lands here (never completes)
        umma::mbar_init_fence();
    }
    if (warp == 0) {
        umma::tmem_alloc(&tmem_base, 512);
        umma::tmem_relinquish();
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = tmem_base;
    if (t == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t a0 = umma::smem_addr(smem), b0 = a0 + 64 * 1024;
        uint64_t ad[4], bd[4];
        for (int i = 0; i < 4; ++i) {
            if (layout == 0) {
                ad[i] = umma::desc_kmajor(a0 + i * 4096, 2048, 128);
                bd[i] = umma::desc_kmajor(b0 + i * N * 32, N * 16, 128);
            } else {
                ad[i] = umma::desc_kmajor(a0 + i * 16384, 16, 1024) | (2ull << 61);
                bd[i] = umma::desc_kmajor(b0 + i * 16384, 16, 1024) | (2ull << 61);
            }
        }
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            umma::mma_tf32(tb, ad[0], bd[0], idesc, 1);
            if (per_round >= 6) {
                umma::mma_tf32(tb, ad[1], bd[0], idesc, 1);
                umma::mma_tf32(tb, ad[0], bd[1], idesc, 1);
            }
            umma::mma_tf32(tb + N, ad[2], bd[2], idesc, 1);
            if (per_round >= 6) {
                umma::mma_tf32(tb + N, ad[3], bd[2], idesc, 1);
                umma::mma_tf32(tb + N, ad[2], bd[3], idesc, 1);
            }
            if (commit_mode == 1) umma::mma_commit(&dummy);
        }
        const long long t1 = clock64();
        umma::mma_commit(&bar);
        const bool ok = umma::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        res[0] = (unsigned long long)rounds * (per_round >= 6 ? 6 : 2);
        res[1] = (unsigned long long)(t2 - t0);
        res[2] = (unsigned long long)(t1 - t0);
        res[3] = ok ? 0ull : 0xDEADull;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
static void split_tf32_host(float v, float& hi, float& lo) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;                 // round to nearest, ties away (cvt.rna.tf32.f32)
    memcpy(&hi, &u, 4);
    lo = v - hi;
}

int sb_tc_profile_read(long long* out48) {          // test / tuning hook: the cycle counters of a -DSB_TC_PROFILE build
#ifdef SB_TC_PROFILE
    long long h[3][16];
    if (cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h)) != cudaSuccess) return -1;
    memcpy(out48, h, sizeof(h));
    long long z[3][16] = {};
    cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
    return 48;
#else
    (void)out48;
    return 0;
#endif
}

size_t sb_tc_zh_bytes(const TcPlan& plan, int n_pairs) { return (size_t)n_pairs * 2 * plan.nb * plan.Sh * sizeof(float2); }

int sb_tc_plan(sb_ctx* ctx, int Sh, int n, TcPlan* plan) {
    *plan = TcPlan();
    static const bool off = getenv("SB_REG_NO_TC") != nullptr;
    if (off || Sh != kTcSh || n < 8 || n > 400) return SB_OK;
    plan->Sh = Sh;
    plan->n = n;
    plan->nb = n / 2 + 1;
    plan->NP = (plan->nb + 15) & ~15;
    plan->nchunks = (plan->nb + 7) / 8;
    plan->pitch_w = 0;
    plan->stg_bytes = (std::max(128 * tc_pitch(2 * n), n * tc_pitch(256)) + 127) & ~127;
    plan->acc_bufs = 4 * plan->NP <= 512 ? 2 : 1;
    plan->stg_bufs = tc_smem_layout(plan->NP, plan->stg_bytes, 2).total <= 226 * 1024 ? 2 : 1;
    const TcSmem L = tc_smem_layout(plan->NP, plan->stg_bytes, plan->stg_bufs);
    plan->smem_fwd = L.total;
    if (L.total > 226 * 1024) return SB_OK;
    const int nb = plan->nb, NP = plan->NP, nh = n / 2, no = (n - 1) / 2;
    // forward tables: per chunk [cos_hi | cos_lo | -sin_hi | -sin_lo], each NP rows (output bin k) x 8 (input j), K-major
    const uint64_t key = ((uint64_t)2 << 40) | (uint64_t)n;                 // (forward tables, scaled by 2^-16)
    auto it = ctx->twiddle_cache.find(key);
    if (it == ctx->twiddle_cache.end()) {
        std::vector<float> img((size_t)plan->nchunks * 4 * NP * 8, 0.0f);
        const long double tau = 6.283185307179586476925286766559L;
        for (int c = 0; c < plan->nchunks; ++c)
            for (int k = 0; k < nb; ++k)
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = 8 * c + jj;
                    const long double ang = tau * (long double)(((long long)j * k) % n) / (long double)n;
                    // x 2^-16 (exact): the input scaling of the chain (kInScale), folded into the table so that the converters
                    // feed the stretched integers as they are
                    const float cv = j <= nh ? (float)cosl(ang) * (float)kInScale : 0.0f;
                    const float sv = (j >= 1 && j <= no) ? (float)(-sinl(ang)) * (float)kInScale : 0.0f;
                    float ch, cl, sh, sl;
                    split_tf32_host(cv, ch, cl);
                    split_tf32_host(sv, sh, sl);
                    const size_t base = (size_t)c * 4 * NP * 8, e = (size_t)(jj / 4) * NP * 4 + (size_t)k * 4 + (jj % 4);
                    img[base + e] = ch;
                    img[base + (size_t)NP * 8 + e] = cl;
                    img[base + (size_t)2 * NP * 8 + e] = sh;
                    img[base + (size_t)3 * NP * 8 + e] = sl;
                }
        DevBuf b;
        int rc = sb_reserve(ctx, b, img.size() * sizeof(float));
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpy(b.p, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        it = ctx->twiddle_cache.emplace(key, b).first;
    }
    plan->Bfwd = reinterpret_cast<const uint8_t*>(it->second.p);
    // inverse tables: per chunk [c cos_hi | c cos_lo | -c sin_hi | -c sin_lo], NP rows (output column x) x 8 (bin k);
    // c_k = 1 for k = 0 and the Nyquist bin, 2 otherwise (the conjugate half of the spectrum folded in)
    const uint64_t keyi = ((uint64_t)4 << 40) | (uint64_t)n;
    auto iti = ctx->twiddle_cache.find(keyi);
    if (iti == ctx->twiddle_cache.end()) {
        std::vector<float> img((size_t)plan->nchunks * 4 * NP * 8, 0.0f);
        const long double tau = 6.283185307179586476925286766559L;
        for (int c = 0; c < plan->nchunks; ++c)
            for (int x = 0; x <= nh; ++x)
                for (int kk = 0; kk < 8; ++kk) {
                    const int k = 8 * c + kk;
                    if (k > nh) continue;
                    const long double ck = (k == 0 || 2 * k == n) ? 1.0L : 2.0L;
                    const long double ang = tau * (long double)(((long long)k * x) % n) / (long double)n;
                    const float cv = (float)(ck * cosl(ang));
                    const float sv = (x >= 1 && x <= no && k >= 1 && k <= no) ? (float)(-ck * sinl(ang)) : 0.0f;
                    float ch, cl, sh, sl;
                    split_tf32_host(cv, ch, cl);
                    split_tf32_host(sv, sh, sl);
                    const size_t base = (size_t)c * 4 * NP * 8, e = (size_t)(kk / 4) * NP * 4 + (size_t)x * 4 + (kk % 4);
                    img[base + e] = ch;
                    img[base + (size_t)NP * 8 + e] = cl;
                    img[base + (size_t)2 * NP * 8 + e] = sh;
                    img[base + (size_t)3 * NP * 8 + e] = sl;
                }
        DevBuf b;
        int rc = sb_reserve(ctx, b, img.size() * sizeof(float));
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpy(b.p, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        iti = ctx->twiddle_cache.emplace(keyi, b).first;
    }
    plan->Binv = reinterpret_cast<const uint8_t*>(iti->second.p);
    plan->smem_inv = tc_smem_layout(NP, kCvRingBytes, 1).total;
    static const bool no_inv = getenv("SB_REG_NO_TC_INV") != nullptr;
    plan->inverse = !no_inv && plan->smem_inv <= 226 * 1024;      // (n > 268: the tables + the prefetch ring no longer fit -- radix inverse)
    // column-pass twiddles: tw[k2][l] = exp(-2 pi i l k2 / 1024)
    const uint64_t key2 = ((uint64_t)3 << 40) | 1024u;
    auto it2 = ctx->twiddle_cache.find(key2);
    if (it2 == ctx->twiddle_cache.end()) {
        std::vector<float2> tw(1024);
        const long double tau = 6.283185307179586476925286766559L;
        for (int k2 = 0; k2 < 32; ++k2)
            for (int l = 0; l < 32; ++l) {
                const long double a = -tau * (long double)(l * k2) / 1024.0L;
                tw[(size_t)k2 * 32 + l] = make_float2((float)cosl(a), (float)sinl(a));
            }
        DevBuf b;
        int rc = sb_reserve(ctx, b, tw.size() * sizeof(float2));
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpy(b.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        it2 = ctx->twiddle_cache.emplace(key2, b).first;
    }
    plan->tw1024 = it2->second.p;
    plan->ok = true;
    return SB_OK;
}

int sb_tc_forward(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, const void* d_pairs, int n_pairs, const int2* d_mm, int tile_w,
                  int swap, int maxval, void* Zh, int* d_nonzero, int* d_fault) {
    static int configured = 0;                       // largest dynamic shared-memory size granted so far
    if (plan.smem_fwd > configured) {
        SB_CUDA(ctx, cudaFuncSetAttribute(xdft_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_fwd));
        configured = plan.smem_fwd;
    }
    TcArgs g = {};
    g.Sh = plan.Sh; g.n = plan.n; g.NP = plan.NP; g.nchunks = plan.nchunks; g.swap = swap; g.stg_bytes = plan.stg_bytes;
    g.Bmat = plan.Bfwd; g.fault = d_fault;
    g.pairs = static_cast<const PairDesc*>(d_pairs); g.mm = d_mm; g.tile_w = tile_w; g.maxval = maxval;
    g.Zh = static_cast<float2*>(Zh); g.nonzero = d_nonzero;
    g.n_tiles = n_pairs * 2 * (plan.Sh >> 7); g.stg_bufs = plan.stg_bufs; g.acc_bufs = plan.acc_bufs;
    const int grid = std::min(g.n_tiles, ctx->sm_count);
    xdft_tc_kernel<0><<<grid, kTcThreads, plan.smem_fwd, st>>>(g);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

int sb_tc_inverse(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, int n_pairs, const void* Y, int lines_in, int swap, void* best,
                  float* rowmax, int* d_fault) {
    static int configured = 0;
    if (plan.smem_inv > configured) {
        SB_CUDA(ctx, cudaFuncSetAttribute(xdft_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_inv));
        configured = plan.smem_inv;
    }
    TcArgs g = {};
    g.Sh = plan.Sh; g.n = plan.n; g.NP = plan.NP; g.nchunks = plan.nchunks; g.swap = swap; g.stg_bytes = kCvRingBytes;
    g.Bmat = plan.Binv; g.fault = d_fault;
    g.Y = static_cast<const float2*>(Y); g.lines_in = lines_in;
    g.best = static_cast<CtaBest*>(best); g.rowmax = rowmax;
    g.n_tiles = n_pairs * (plan.Sh >> 7); g.stg_bufs = 1; g.acc_bufs = plan.acc_bufs;
    const int grid = std::min(g.n_tiles, ctx->sm_count);
    xdft_tc_kernel<1><<<grid, kTcThreads, plan.smem_inv, st>>>(g);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

size_t sb_tc_updft_table_bytes(const TcPlan& plan, int n_pairs) {
    return (size_t)n_pairs * ((plan.n + 7) / 8) * (4 * 32 * 8) * sizeof(float);
}

// First stage of the upsampled-DFT refinement on the tensor cores: tables (B operand images + Ey) from the coarse
// peaks, then T[pair][u][y] = sum_x conj(R[y][x]) Ex[u][x].  rs <= 16 (upsample factors up to 10).
int sb_tc_updft_rows(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, int n_pairs, const void* d_peaks, int uf, int rs,
                     int dftshift, const void* R, int lines_in, void* Bimg, void* Ey, void* Tm, int* d_fault) {
    const int nch = (plan.n + 7) / 8;
    updft_tables_tc_kernel<<<dim3(8, n_pairs), 256, 0, st>>>(static_cast<const PeakOut*>(d_peaks), plan.Sh, plan.n, uf, rs, dftshift, nch,
                                                             static_cast<float*>(Bimg), static_cast<float2*>(Ey));
    const int smem = tc_smem_layout(32, kCvRingBytes, 1).total;
    static int configured = 0;
    if (smem > configured) {
        SB_CUDA(ctx, cudaFuncSetAttribute(xdft_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    TcArgs g = {};
    g.Sh = plan.Sh; g.n = plan.n; g.NP = 32; g.nchunks = nch; g.swap = 0; g.stg_bytes = kCvRingBytes;
    g.Bmat = static_cast<const uint8_t*>(Bimg); g.fault = d_fault;
    g.Y = static_cast<const float2*>(R); g.lines_in = lines_in; g.half = lines_in < plan.n ? 1 : 0;
    g.Tm = static_cast<float2*>(Tm); g.rs = rs;
    g.n_tiles = n_pairs * (plan.Sh >> 7); g.stg_bufs = 1; g.acc_bufs = 2;
    const int grid = std::min(g.n_tiles, ctx->sm_count);
    xdft_tc_kernel<2><<<grid, kTcThreads, smem, st>>>(g);
    ctx->launches += 2;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

int sb_tc_columns(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, int n_pairs, const void* Zh, void* R, void* Y, int lines_out,
                  int mirror) {
    const int smem = (int)(1024 * sizeof(float2) + 4 * sizeof(ColsSmem));
    static bool configured = false;
    if (!configured) {
        SB_CUDA(ctx, cudaFuncSetAttribute(cols_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int n_cols = n_pairs * plan.nb;
    const int grid = std::max(1, std::min((n_cols + 3) / 4, ctx->sm_count * 3));
    cols_warp_kernel<<<grid, 128, smem, st>>>(n_cols, plan.nb, plan.n, lines_out, mirror, static_cast<const float2*>(plan.tw1024),
                                              static_cast<const float2*>(Zh), static_cast<float2*>(R), static_cast<float2*>(Y));
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

// out[0] = elements checked, out[1] = elements off by more than the tolerance, out[2] = largest |error| * 1e12,
// out[3] = 0xDEAD if the MMA pipeline never signalled completion (bounded wait).
int sb_selftest_umma_impl(sb_ctx* ctx, int variant, uint64_t* out) {
    constexpr int N = 112, K = 56;
    Lane* lane = sb_lane(ctx, 0);
    if (variant >= 1000) {          // issue-rate probe: 1000 + N + 1000 * layout + 10000 * (1 = two MMAs per round instead of six)
        const int v = variant - 1000, n_ = v % 1000, layout = (v / 1000) % 10, two = (v / 10000) % 10, cm = (v / 100000) % 10;
        SB_CHECK(ctx, n_ >= 8 && n_ <= 248 && n_ % 8 == 0 && layout <= 1, "selftest: bad rate-probe request %d", variant);
        int rc = sb_reserve(ctx, lane->work, 64);
        if (rc) return rc;
        SB_CUDA(ctx, cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        umma_rate_kernel<<<1, 128, 128 * 1024, lane->stream>>>(n_, layout, 256, two ? 2 : 6, cm, (unsigned long long*)lane->work.p);
        ctx->launches++;
        SB_CUDA(ctx, cudaGetLastError());
        SB_CUDA(ctx, cudaMemcpyAsync(out, lane->work.p, 32, cudaMemcpyDeviceToHost, lane->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(lane->stream));
        return SB_OK;
    }
    SB_CHECK(ctx, variant >= 0 && variant <= 2, "selftest: unknown tensor-core variant %d", variant);
    int rc = sb_reserve(ctx, lane->work, 64 + (size_t)128 * N * 4);
    if (rc) return rc;
    unsigned long long* res = (unsigned long long*)lane->work.p;
    float* D = (float*)((uint8_t*)lane->work.p + 64);
    SB_CUDA(ctx, cudaMemsetAsync(lane->work.p, 0, 64 + (size_t)128 * N * 4, lane->stream));
    const size_t smem = 2 * (K / 4) * (128 * 16 + 16) + 2 * (K / 4) * (N * 16 + 16) + 128;
    auto kern = selftest_umma_kernel<N, K>;
    SB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<1, 128, smem, lane->stream>>>(variant, D, res);
    // float32-grade: 3 x tf32 keeps ~2^-21 per product; a single tf32 product (variant 2) is ~2^-11
    const double tol = variant == 2 ? 2e-2 : 2e-5;
    selftest_umma_check_kernel<N, K><<<128, 128, 0, lane->stream>>>(D, tol, res);
    ctx->launches += 2;
    SB_CUDA(ctx, cudaGetLastError());
    SB_CUDA(ctx, cudaMemcpyAsync(out, res, 32, cudaMemcpyDeviceToHost, lane->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(lane->stream));
    out[0] = (uint64_t)128 * N;
    return SB_OK;
}
