// uint8 pixels (SB_U8): 8-bit Squid acquisitions (BMP, RGB planes).  The reference derives everything from the tile
// dtype (stitcher_process.py:340): normalize_image scales to iinfo(dtype).max (:854), apply_flatfield_correction clips
// to it (:838-841) and the canvas has the tile dtype (:503).  The uint16 kernels are reused: tiles are widened on the
// device, the uint16 result is narrowed with saturation.  trunc(clip(x, 0, 255)) == min(trunc(clip(x, 0, 65535)), 255),
// so the narrowing reproduces the reference's clip exactly; registration takes the 8-bit scale as a parameter.
#include "sb_common.cuh"

#include <map>
#include <vector>

namespace {

// rows of `w` pixels: in pitch `ip` bytes, out pitch `op` elements
__global__ void __launch_bounds__(256) widen_u8_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out, int64_t n) {
    const int64_t i8 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i8 + 8 <= n && ((reinterpret_cast<uintptr_t>(in + i8) & 7) == 0) && ((reinterpret_cast<uintptr_t>(out + i8) & 15) == 0)) {
        const uint2 v = *reinterpret_cast<const uint2*>(in + i8);
        uint4 o;
        o.x = __byte_perm(v.x, 0, 0x4140);
        o.y = __byte_perm(v.x, 0, 0x4342);
        o.z = __byte_perm(v.y, 0, 0x4140);
        o.w = __byte_perm(v.y, 0, 0x4342);
        *reinterpret_cast<uint4*>(out + i8) = o;
    } else {
        for (int64_t i = i8; i < i8 + 8 && i < n; ++i) out[i] = in[i];
    }
}

// rows x width elements, source pitch sp (elements), destination pitch dp (elements); saturates at 255
__global__ void __launch_bounds__(256) narrow_u16_kernel(const uint16_t* __restrict__ in, int64_t sp, uint8_t* __restrict__ out,
                                                         int64_t dp, int64_t rows, int64_t width) {
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const uint16_t* src = in + r * sp;
        uint8_t* dst = out + r * dp;
        for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < width; x += (int64_t)gridDim.x * blockDim.x)
            dst[x] = (uint8_t)min((unsigned)src[x], 255u);
    }
}

int widen(sb_ctx* ctx, cudaStream_t st, const uint8_t* in, uint16_t* out, int64_t n) {
    if (n <= 0) return SB_OK;
    widen_u8_kernel<<<(unsigned)((n + 2047) / 2048), 256, 0, st>>>(in, out, n);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

int narrow(sb_ctx* ctx, cudaStream_t st, const uint16_t* in, int64_t sp, uint8_t* out, int64_t dp, int64_t rows, int64_t width) {
    if (rows <= 0 || width <= 0) return SB_OK;
    const unsigned gx = (unsigned)std::min<int64_t>((width + 255) / 256, 64);
    const unsigned gy = (unsigned)std::min<int64_t>(rows, 32768);
    narrow_u16_kernel<<<dim3(gx, gy), 256, 0, st>>>(in, sp, out, dp, rows, width);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

// widened copies of the unique 8-bit tiles of a job in `pool` (tight rows); returns caller pointer -> widened pointer
int widen_tiles(sb_ctx* ctx, Lane* lane, DevBuf& stage, DevBuf& pool, const std::vector<const void*>& ptrs, int H, int W, int mem,
                std::map<const void*, const uint16_t*>& wide) {
    for (const void* p : ptrs) {
        if (!p) return sb_fail(ctx, SB_ERR_INVALID, "NULL tile pointer");
        wide.emplace(p, nullptr);
    }
    const size_t px = (size_t)H * W;
    int rc = sb_reserve(ctx, pool, px * 2 * wide.size());
    if (rc) return rc;
    if (mem == SB_MEM_HOST) {
        rc = sb_reserve(ctx, stage, px * wide.size());
        if (rc) return rc;
    }
    size_t i = 0;
    for (auto& kv : wide) {
        uint16_t* dst = (uint16_t*)pool.p + px * i;
        const uint8_t* src = (const uint8_t*)kv.first;
        if (mem == SB_MEM_HOST) {
            uint8_t* d8 = (uint8_t*)stage.p + px * i;
            SB_CUDA(ctx, cudaMemcpyAsync(d8, kv.first, px, cudaMemcpyHostToDevice, lane->stream));
            src = d8;
        }
        rc = widen(ctx, lane->stream, src, dst, (int64_t)px);
        if (rc) return rc;
        kv.second = dst;
        ++i;
    }
    return SB_OK;
}

}  // namespace

int sb_fuse_region_u8(sb_ctx* ctx, const sb_fuse_job* job, int lane_idx) {
    SB_CHECK(ctx, job->blend == SB_BLEND_PASTE, "uint8 pixels are implemented for the reference's paste mode only");
    SB_CHECK(ctx, job->n_tiles >= 0 && (job->n_tiles == 0 || job->tiles), "bad tile list");
    SB_CHECK(ctx, job->tile_w % 8 == 0, "uint8 tiles need tile_w %% 8 == 0, got %d", job->tile_w);
    SB_CHECK(ctx, job->out != nullptr, "out is NULL");
    const bool sync_call = lane_idx < 0;
    Lane* lane = sb_lane(ctx, sync_call ? 0 : lane_idx);
    SB_CHECK(ctx, lane != nullptr, "lane %d out of range", lane_idx);
    const int H = job->tile_h, W = job->tile_w, n = job->n_tiles;
    std::vector<const void*> ptrs;
    for (int i = 0; i < n; ++i) ptrs.push_back(job->tiles[i].px);
    std::map<const void*, const uint16_t*> wide;
    int rc = widen_tiles(ctx, lane, lane->u8_stage, lane->u8_tiles, ptrs, H, W, job->tile_mem, wide);
    if (rc) return rc;
    std::vector<sb_tile> tiles(job->tiles, job->tiles + n);
    for (auto& t : tiles) t.px = wide[t.px];

    // uint16 canvas on the device in the library's own pitch / chunk order, then narrowed into the caller's buffer
    const int planes = job->num_c * job->num_z;
    const bool chunked = job->out_layout == SB_LAYOUT_CHUNKED;
    const int64_t pitch16 = sb_canvas_pitch(job->width);
    const int64_t elems = chunked ? sb_chunked_plane_elems(job->height, job->width, job->chunk_h, job->chunk_w) * planes
                                  : pitch16 * job->height * planes;
    SB_CHECK(ctx, elems > 0, "bad canvas / chunk shape");
    rc = sb_reserve(ctx, lane->u8_canvas16, (size_t)elems * 2);
    if (rc) return rc;
    sb_fuse_job j16 = *job;
    j16.tiles = tiles.data();
    j16.dtype = SB_U16;
    j16.tile_mem = SB_MEM_DEVICE;
    j16.out = lane->u8_canvas16.p;
    j16.out_mem = SB_MEM_DEVICE;
    j16.out_row_pitch = 0;
    rc = sb_fuse_region_impl(ctx, &j16, sync_call ? 0 : lane_idx);    // enqueues on the lane's stream; no host wait
    if (rc) return rc;
    cudaStream_t st = lane->stream;
    const uint16_t* c16 = (const uint16_t*)lane->u8_canvas16.p;
    const int64_t rows = (int64_t)job->height * planes;
    if (job->out_mem == SB_MEM_DEVICE) {
        if (chunked) rc = narrow(ctx, st, c16, elems, (uint8_t*)job->out, elems, 1, elems);
        else {
            const int64_t dp = job->out_row_pitch ? job->out_row_pitch : pitch16;
            SB_CHECK(ctx, dp >= job->width, "device out_row_pitch < width");
            rc = narrow(ctx, st, c16, pitch16, (uint8_t*)job->out, dp, rows, job->width);
        }
        if (rc) return rc;
    } else {
        rc = sb_reserve(ctx, lane->u8_canvas8, (size_t)elems);
        if (rc) return rc;
        uint8_t* c8 = (uint8_t*)lane->u8_canvas8.p;
        if (chunked) {
            rc = narrow(ctx, st, c16, elems, c8, elems, 1, elems);
            if (rc) return rc;
            SB_CUDA(ctx, cudaMemcpyAsync(job->out, c8, (size_t)elems, cudaMemcpyDeviceToHost, st));
        } else {
            rc = narrow(ctx, st, c16, pitch16, c8, pitch16, rows, job->width);
            if (rc) return rc;
            const int64_t hp = job->out_row_pitch ? job->out_row_pitch : job->width;
            SB_CHECK(ctx, hp >= job->width, "host out_row_pitch < width");
            SB_CUDA(ctx, cudaMemcpy2DAsync(job->out, (size_t)hp, c8, (size_t)pitch16, (size_t)job->width, (size_t)rows,
                                           cudaMemcpyDeviceToHost, st));
        }
    }
    if (sync_call) SB_CUDA(ctx, cudaStreamSynchronize(st));
    return SB_OK;
}

int sb_register_pairs_u8(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out, bool async) {
    SB_CHECK(ctx, job && out, "job/out is NULL");
    SB_CHECK(ctx, job->n_pairs >= 0 && (job->n_pairs == 0 || job->pairs), "bad pair list");
    Lane* lane = sb_lane(ctx, job->lane);
    SB_CHECK(ctx, lane != nullptr, "lane %d out of range", job->lane);
    int rc = sb_register_complete(ctx, job->lane);      // a parked job still reads the lane's widened tiles
    if (rc) return rc;
    std::vector<const void*> ptrs;
    for (int i = 0; i < job->n_pairs; ++i) {
        ptrs.push_back(job->pairs[i].ref);
        ptrs.push_back(job->pairs[i].mov);
    }
    std::map<const void*, const uint16_t*> wide;
    rc = widen_tiles(ctx, lane, lane->u8_reg_stage, lane->u8_reg_tiles, ptrs, job->tile_h, job->tile_w, job->mem, wide);
    if (rc) return rc;
    std::vector<sb_pair> pairs(job->pairs, job->pairs + job->n_pairs);
    for (auto& p : pairs) {
        p.ref = wide[p.ref];
        p.mov = wide[p.mov];
    }
    sb_register_job j16 = *job;
    j16.pairs = pairs.data();
    j16.dtype = SB_U16;
    j16.mem = SB_MEM_DEVICE;
    return sb_register_pairs_impl(ctx, &j16, out, async, 255);
}

// tiles -> uint16 on the device, `body` on uint16 device buffers, result narrowed back
template <typename F>
static int u8_elementwise(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int mem, F body) {
    SB_CHECK(ctx, tiles && out && n_tiles > 0 && tile_h > 0 && tile_w > 0, "bad arguments");
    Lane* lane = sb_lane(ctx, 0);
    cudaStream_t st = lane->stream;
    const int64_t total = (int64_t)n_tiles * tile_h * tile_w;
    int rc = sb_reserve(ctx, lane->u8_tiles, (size_t)total * 2);
    if (rc) return rc;
    rc = sb_reserve(ctx, lane->u8_canvas16, (size_t)total * 2);
    if (rc) return rc;
    const uint8_t* src = (const uint8_t*)tiles;
    if (mem == SB_MEM_HOST) {
        rc = sb_reserve(ctx, lane->u8_stage, (size_t)total);
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpyAsync(lane->u8_stage.p, tiles, (size_t)total, cudaMemcpyHostToDevice, st));
        src = (const uint8_t*)lane->u8_stage.p;
    }
    rc = widen(ctx, st, src, (uint16_t*)lane->u8_tiles.p, total);
    if (rc) return rc;
    rc = body((const uint16_t*)lane->u8_tiles.p, (uint16_t*)lane->u8_canvas16.p);
    if (rc) return rc;
    uint8_t* dst = (uint8_t*)out;
    if (mem == SB_MEM_HOST) {
        rc = sb_reserve(ctx, lane->u8_canvas8, (size_t)total);
        if (rc) return rc;
        dst = (uint8_t*)lane->u8_canvas8.p;
    }
    rc = narrow(ctx, st, (const uint16_t*)lane->u8_canvas16.p, total, dst, total, 1, total);
    if (rc) return rc;
    if (mem == SB_MEM_HOST) SB_CUDA(ctx, cudaMemcpyAsync(out, dst, (size_t)total, cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    return SB_OK;
}

int sb_flatfield_apply_u8(sb_ctx* ctx, int channel, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int mem) {
    return u8_elementwise(ctx, tiles, out, n_tiles, tile_h, tile_w, mem, [&](const uint16_t* in16, uint16_t* out16) {
        return sb_flatfield_apply_impl(ctx, channel, in16, out16, n_tiles, tile_h, tile_w, SB_U16, SB_MEM_DEVICE);
    });
}

int sb_normalize_u8(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int mem) {
    return u8_elementwise(ctx, tiles, out, n_tiles, tile_h, tile_w, mem, [&](const uint16_t* in16, uint16_t* out16) {
        return sb_normalize_impl(ctx, in16, out16, n_tiles, tile_h, tile_w, SB_U16, SB_MEM_DEVICE, 255);
    });
}
