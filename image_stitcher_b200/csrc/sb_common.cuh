// Shared host/device plumbing for libstitchb200: context, lanes, error handling, PTX wrappers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/stitchb200.h"

#define SB_NUM_LANES 3

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct FieldPool {            // flat- or dark-fields of all channels, one contiguous device array
    void* dev = nullptr;      // [n_slots][h][w] of float or double
    int dtype = SB_FIELD_F32;
    int h = 0, w = 0;
    int n_slots = 0;
    bool fast_ok = true;      // every value inside the range where the branch-free float32 divide is exact
    std::vector<int> slot_of_channel;   // channel -> slot or -1
    bool any() const { for (int s : slot_of_channel) if (s >= 0) return true; return false; }
    int slot(int c) const { return (c >= 0 && c < (int)slot_of_channel.size()) ? slot_of_channel[c] : -1; }
};

struct ResidentCanvas {       // the row-major canvas the lane's last sb_fuse_region left on the device (sb_pyramid, src == NULL)
    const void* p = nullptr;
    int64_t pitch = 0;        // elements per row
    int planes = 0, h = 0, w = 0, dtype = 0;
};

struct Lane {
    cudaStream_t own = nullptr;
    cudaStream_t stream = nullptr;
    DevBuf tiles, canvas, meta, work;
    DevBuf u8_stage, u8_tiles, u8_canvas16, u8_canvas8, u8_reg_stage, u8_reg_tiles;   // uint8 pixel path (u8.cu)
    void* meta_host = nullptr;      // pinned staging for per-job metadata
    size_t meta_host_cap = 0;
    cudaEvent_t meta_free = nullptr; // the previous job's metadata H2D has been consumed
    cudaEvent_t mark = nullptr;      // sb_lane_mark / sb_lane_wait_mark
    bool marked = false;
    DevBuf pyr_src, pyr_out;        // sb_pyramid: uploaded level 0 (host sources only) and the levels before their download
    ResidentCanvas resident;
    DevBuf reg_tiles, reg_work, reg_meta;   // registration workspace of this lane (grown on demand)
    void* reg_host = nullptr;       // pinned block the registration results are copied into
    size_t reg_host_cap = 0;
    void* reg_pending = nullptr;    // parked asynchronous registration job (RegPending in reg.cu)
    void* dbg_ptr[4] = {nullptr, nullptr, nullptr, nullptr};   // last registration group's Zh / Y / R / T of sub-batch 0 (sb_debug_read, tests)
    size_t dbg_bytes[4] = {0, 0, 0, 0};
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // extra streams of the lane: registration sub-batches rotate over them
    cudaEvent_t aux_fork = nullptr, aux_join[3] = {nullptr, nullptr, nullptr};
    std::vector<int32_t> rect_pieces;   // cached rectangle decomposition of the rectangle-streaming paste kernel ...
    std::vector<int32_t> rect_key;      // ... and the geometry it was computed for (compared in full)
    std::vector<int32_t> blend_cells;   // cached cell decomposition of the blend modes (10 ints per cell) ...
    std::vector<int32_t> blend_key;     // ... and its geometry
    std::vector<int32_t> perm;      // cached block-row order of the paste kernel (fuse.cu) ...
    uint64_t perm_sig = 0;          // ... and the geometry signature it was computed for
};

struct sb_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    int64_t launches = 0;
    Lane lanes[SB_NUM_LANES];
    FieldPool flat, dark;
    // driver entry point for TMA descriptors (resolved through the runtime: no libcuda link dependency)
    CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    std::unordered_map<uint64_t, DevBuf> twiddle_cache;   // key: (n << 1 | is_double)
};

int sb_fail(sb_ctx* ctx, int code, const char* fmt, ...);
int sb_reserve(sb_ctx* ctx, DevBuf& b, size_t bytes);
int sb_reserve_pinned(sb_ctx* ctx, void** p, size_t* cap, size_t bytes);
Lane* sb_lane(sb_ctx* ctx, int lane);

#define SB_CUDA(ctx, call)                                                                       \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return sb_fail((ctx), SB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                                  \
    } while (0)

#define SB_CHECK(ctx, cond, ...)                                        \
    do {                                                                \
        if (!(cond)) return sb_fail((ctx), SB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

static inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// entry points implemented in fuse.cu / reg.cu, called from api.cu
int sb_fuse_region_impl(sb_ctx* ctx, const sb_fuse_job* job, int lane);
int sb_fuse_regions_impl(sb_ctx* ctx, const sb_fuse_job* jobs, int n_jobs, int lane, bool* batched);
int sb_flatfield_apply_impl(sb_ctx* ctx, int channel, const void* tiles, void* out, int n_tiles, int tile_h,
                            int tile_w, int dtype, int mem);
// maxval: iinfo(dtype).max of the caller's pixels (65535, or 255 when the tiles were widened from uint8)
int sb_register_pairs_impl(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out, bool async, int maxval = 65535);
int sb_register_complete(sb_ctx* ctx, int lane);     // finish the lane's parked asynchronous registration, if any
void sb_register_discard(sb_ctx* ctx, int lane);
int sb_normalize_impl(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int dtype,
                      int mem, int maxval = 65535);
// uint8 pixels (u8.cu): widen -> uint16 kernels -> narrow with saturation
int sb_fuse_region_u8(sb_ctx* ctx, const sb_fuse_job* job, int lane);
int sb_register_pairs_u8(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out, bool async);
int sb_flatfield_apply_u8(sb_ctx* ctx, int channel, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int mem);
int sb_normalize_u8(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int mem);
// test hooks (sb_selftest): exhaustive device-side proofs of the two "exact by construction" primitives
int sb_selftest_div_impl(sb_ctx* ctx, int expo, uint64_t* out);        // fuse.cu
int sb_selftest_stretch_impl(sb_ctx* ctx, int maxval, uint64_t* out);  // reg.cu
int sb_selftest_umma_impl(sb_ctx* ctx, int variant, uint64_t* out);    // reg_tc.cu
// flatfield.cu
int sb_estimate_flatfield_impl(sb_ctx* ctx, const void* const* tiles, int n_tiles, int tile_h, int tile_w, int dtype, int mem,
                               int grid, double sigma, float* field_out, int out_mem);
// pyramid.cu
int64_t sb_pyramid_elems_impl(int64_t n_planes, int height, int width, int n_levels);
int sb_pyramid_impl(sb_ctx* ctx, const void* src, int src_mem, int n_planes, int height, int width, int64_t src_row_pitch,
                    int dtype, int n_levels, void* out, int out_mem, int lane);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------- device-side PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SB_DONE;\n"
        "bra SB_WAIT;\n"
        "SB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier, with an L2 eviction hint.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// bulk prefetch of [p, p + bytes) into L2 (p 16-byte aligned, bytes a multiple of 16); no destination, no completion
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk_hint(const void* p, unsigned bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(policy) : "memory");
}
// read-only 128-bit load with an L2 eviction-priority hint (createpolicy result)
__device__ __forceinline__ float4 ldg_f4_hint(const float4* p, uint64_t policy) {
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(policy));
    return v;
}
// read-only 256-bit load (sm_100: LDG.E.256): 8 floats of one lane from a 32-byte aligned address.  A warp reading 32 x 32
// contiguous bytes this way costs the L1 data pipe 8 wavefronts (1 KB / 128 B); the same bytes as two 128-bit loads at a
// lane stride of 32 bytes cost 8 wavefronts EACH (every load touches all 8 lines).
__device__ __forceinline__ void ldg_f8(const float* p, float4& lo, float4& hi) {
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
        : "l"(p));
}
// streaming 128-bit store: written once, never re-read by this kernel
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
#endif  // __CUDACC__
