// C-ABI entry points of libstitchb200 (include/stitchb200.h): context lifecycle, memory helpers,
// flat/dark-field storage.  The compute entry points forward to fuse.cu / reg.cu.
#include "sb_common.cuh"
#include "reg_tc.cuh"

#include <algorithm>
#include <mutex>

static thread_local std::string g_create_error;

int sb_fail(sb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    if (code == SB_ERR_CUDA) cudaGetLastError();   // clear the sticky-less error state
    return code;
}

int sb_reserve(sb_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return SB_OK;
    if (b.p) {
        // growing a buffer that earlier asynchronous work may still use: drain first
        cudaDeviceSynchronize();
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = round_up64((int64_t)bytes, 1 << 20);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return sb_fail(ctx, SB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return SB_OK;
}

int sb_reserve_pinned(sb_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return SB_OK;
    if (*p) {
        cudaDeviceSynchronize();
        cudaFreeHost(*p);
        *p = nullptr;
        *cap = 0;
    }
    const size_t want = round_up64((int64_t)bytes, 1 << 16);
    cudaError_t e = cudaMallocHost(p, want);
    if (e != cudaSuccess) {
        *p = nullptr;
        return sb_fail(ctx, SB_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    *cap = want;
    return SB_OK;
}

// Every entry point runs on the context's device, whatever the calling thread's current device is (several contexts
// on different GPUs may live in one process); the caller's device is restored on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const sb_ctx* ctx) {
        int cur = -1;
        if (ctx && cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device) {
            prev = cur;
            cudaSetDevice(ctx->device);
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define SB_ENTER(ctx) DeviceGuard sb_device_guard__(ctx)

Lane* sb_lane(sb_ctx* ctx, int lane) {
    if (!ctx || lane < 0 || lane >= SB_NUM_LANES) return nullptr;
    return &ctx->lanes[lane];
}

extern "C" {

int sb_version(void) { return SB_ABI_VERSION; }

int sb_create(int device, sb_ctx** out) {
    if (!out) return sb_fail(nullptr, SB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return sb_fail(nullptr, SB_ERR_NODEVICE, "no CUDA device available (%s); libstitchb200 has no CPU path",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return sb_fail(nullptr, SB_ERR_INVALID, "device %d out of range [0, %d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return sb_fail(nullptr, SB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return sb_fail(nullptr, SB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return sb_fail(nullptr, SB_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
    sb_ctx* ctx = new sb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < SB_NUM_LANES; ++i) {
        Lane& l = ctx->lanes[i];
        if (cudaStreamCreateWithFlags(&l.own, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.meta_free, cudaEventDisableTiming) != cudaSuccess) {
            sb_fail(nullptr, SB_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            delete ctx;
            return SB_ERR_CUDA;
        }
        l.stream = l.own;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        sb_fail(nullptr, SB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        delete ctx;
        return SB_ERR_CUDA;
    }
    ctx->encode_tiled = reinterpret_cast<decltype(ctx->encode_tiled)>(fn);
    *out = ctx;
    return SB_OK;
}

static void free_field(FieldPool& f) {
    if (f.dev) cudaFree(f.dev);
    f = FieldPool();
}

void sb_destroy(sb_ctx* ctx) {
    if (!ctx) return;
    SB_ENTER(ctx);
    cudaDeviceSynchronize();
    for (int i = 0; i < SB_NUM_LANES; ++i) {
        Lane& l = ctx->lanes[i];
        sb_register_discard(ctx, i);
        for (DevBuf* b : {&l.tiles, &l.canvas, &l.meta, &l.work, &l.reg_tiles, &l.reg_work, &l.reg_meta, &l.u8_stage, &l.u8_tiles,
                          &l.u8_canvas16, &l.u8_canvas8, &l.u8_reg_stage, &l.u8_reg_tiles, &l.pyr_src, &l.pyr_out})
            if (b->p) cudaFree(b->p);
        if (l.meta_host) cudaFreeHost(l.meta_host);
        if (l.reg_host) cudaFreeHost(l.reg_host);
        if (l.meta_free) cudaEventDestroy(l.meta_free);
        if (l.mark) cudaEventDestroy(l.mark);
        if (l.aux_fork) cudaEventDestroy(l.aux_fork);
        for (int k = 0; k < 3; ++k) {
            if (l.aux_join[k]) cudaEventDestroy(l.aux_join[k]);
            if (l.aux[k]) cudaStreamDestroy(l.aux[k]);
        }
        if (l.own) cudaStreamDestroy(l.own);
    }
    for (auto& kv : ctx->twiddle_cache)
        if (kv.second.p) cudaFree(kv.second.p);
    free_field(ctx->flat);
    free_field(ctx->dark);
    delete ctx;
}

const char* sb_last_error(const sb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int64_t sb_kernel_launches(const sb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int sb_num_lanes(const sb_ctx*) { return SB_NUM_LANES; }
int sb_device_sm_count(const sb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

void* sb_host_alloc(sb_ctx* ctx, size_t bytes) {
    SB_ENTER(ctx);
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        sb_fail(ctx, SB_ERR_NOMEM, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void sb_host_free(sb_ctx* ctx, void* p) {
    SB_ENTER(ctx);
    if (p) cudaFreeHost(p);
}
void* sb_device_alloc(sb_ctx* ctx, size_t bytes) {
    SB_ENTER(ctx);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        sb_fail(ctx, SB_ERR_NOMEM, "cudaMalloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void sb_device_free(sb_ctx* ctx, void* p) {
    SB_ENTER(ctx);
    if (p) cudaFree(p);
}
int sb_memcpy_h2d(sb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    SB_ENTER(ctx);
    SB_CUDA(ctx, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return SB_OK;
}
int sb_memcpy_d2h(sb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    SB_ENTER(ctx);
    SB_CUDA(ctx, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return SB_OK;
}

int sb_memcpy_async(sb_ctx* ctx, int lane, void* dst, const void* src, size_t bytes, int kind) {
    SB_ENTER(ctx);
    Lane* l = sb_lane(ctx, lane);
    if (!l) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    SB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, k, l->stream));
    return SB_OK;
}

int sb_memcpy2d_async(sb_ctx* ctx, int lane, void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                      size_t width_bytes, size_t height, int kind) {
    SB_ENTER(ctx);
    Lane* l = sb_lane(ctx, lane);
    if (!l) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    SB_CHECK(ctx, dst != nullptr && src != nullptr && width_bytes <= dst_pitch && width_bytes <= src_pitch, "bad 2-D copy");
    if (width_bytes == 0 || height == 0) return SB_OK;
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    SB_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, k, l->stream));
    return SB_OK;
}

static int set_field(sb_ctx* ctx, FieldPool& pool, const char* what, int channel, const void* field, int dtype, int mem,
                     int h, int w) {
    SB_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    SB_CHECK(ctx, channel >= 0 && channel < 4096, "%s: channel %d out of range", what, channel);
    SB_CHECK(ctx, field != nullptr && h > 0 && w > 0, "%s: bad field", what);
    SB_CHECK(ctx, dtype == SB_FIELD_F32 || dtype == SB_FIELD_F64, "%s: unknown dtype %d", what, dtype);
    const size_t eb = dtype == SB_FIELD_F64 ? 8 : 4;
    if (pool.dev && (pool.h != h || pool.w != w || pool.dtype != dtype))
        return sb_fail(ctx, SB_ERR_INVALID, "%s: shape/dtype differs from fields already set; call sb_clear_fields first",
                       what);
    if ((int)pool.slot_of_channel.size() <= channel) pool.slot_of_channel.resize(channel + 1, -1);
    int slot = pool.slot_of_channel[channel];
    const size_t one = (size_t)h * w * eb;
    if (slot < 0) {
        // grow the contiguous pool by one slot (fields are set once per run: simplicity over speed)
        void* nd = nullptr;
        cudaDeviceSynchronize();
        if (cudaMalloc(&nd, one * (pool.n_slots + 1)) != cudaSuccess)
            return sb_fail(ctx, SB_ERR_NOMEM, "%s: cudaMalloc failed", what);
        if (pool.dev) {
            cudaMemcpy(nd, pool.dev, one * pool.n_slots, cudaMemcpyDeviceToDevice);
            cudaFree(pool.dev);
        }
        pool.dev = nd;
        slot = pool.n_slots++;
        pool.slot_of_channel[channel] = slot;
        pool.h = h;
        pool.w = w;
        pool.dtype = dtype;
    }
    SB_CUDA(ctx, cudaMemcpy((uint8_t*)pool.dev + (size_t)slot * one, field, one,
                            mem == SB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    // Range check, once per field: the paste fast path divides with the branch-free div.rn.f32 sequence,
    // which is exact only while no operand exponent is extreme (what FCHK guards in the compiler's code).
    if (dtype == SB_FIELD_F32) {
        std::vector<float> tmp;
        const float* h_field = (const float*)field;
        if (mem == SB_MEM_DEVICE) {
            tmp.resize((size_t)h * w);
            SB_CUDA(ctx, cudaMemcpy(tmp.data(), field, one, cudaMemcpyDeviceToHost));
            h_field = tmp.data();
        }
        const bool is_flat = (&pool == &ctx->flat);
        bool ok = true;
        for (size_t i = 0; i < (size_t)h * w && ok; ++i) {
            const float v = h_field[i];
            // flat >= 2^-5 and |dark| <= 2^16 also keep |quotient| < 2^22, which the magic-number truncation of the
            // paste kernel needs (fuse.cu: trunc_sat_pack); fields outside take the generic kernel
            ok = is_flat ? (v >= 0.03125f && v <= 1048576.0f) : (v >= -65536.0f && v <= 65536.0f);
        }
        if (!ok) pool.fast_ok = false;
    } else {
        pool.fast_ok = false;      // float64 fields take the generic kernel (float64 divide, as the reference does)
    }
    return SB_OK;
}

int sb_set_flatfield(sb_ctx* ctx, int channel, const void* field, int dtype, int mem, int h, int w) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    return set_field(ctx, ctx->flat, "sb_set_flatfield", channel, field, dtype, mem, h, w);
}
int sb_set_darkfield(sb_ctx* ctx, int channel, const void* field, int dtype, int mem, int h, int w) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    return set_field(ctx, ctx->dark, "sb_set_darkfield", channel, field, dtype, mem, h, w);
}
int sb_clear_fields(sb_ctx* ctx) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    cudaDeviceSynchronize();
    free_field(ctx->flat);
    free_field(ctx->dark);
    return SB_OK;
}

int64_t sb_canvas_pitch(int32_t width) { return round_up64(width, 64); }
int64_t sb_chunked_plane_elems(int32_t height, int32_t width, int32_t chunk_h, int32_t chunk_w) {
    if (chunk_h <= 0 || chunk_w <= 0) return 0;
    return (int64_t)((height + chunk_h - 1) / chunk_h) * ((width + chunk_w - 1) / chunk_w) * chunk_h * chunk_w;
}

int sb_fuse_region(sb_ctx* ctx, const sb_fuse_job* job, int lane) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    if (lane >= SB_NUM_LANES) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    Lane* l = sb_lane(ctx, lane < 0 ? 0 : lane);
    l->resident = ResidentCanvas();
    const bool u8 = job && job->dtype == SB_U8;
    const int rc = u8 ? sb_fuse_region_u8(ctx, job, lane) : sb_fuse_region_impl(ctx, job, lane);
    if (rc == SB_OK && job->out_layout == SB_LAYOUT_ROWMAJOR) {
        // remember where the canvas lives on the device: sb_pyramid(src == NULL) builds the multiscale levels from it
        ResidentCanvas& r = l->resident;
        if (job->out_mem == SB_MEM_DEVICE) {
            r.p = job->out;
            r.pitch = job->out_row_pitch ? job->out_row_pitch : sb_canvas_pitch(job->width);
        } else {
            r.p = u8 ? l->u8_canvas8.p : l->canvas.p;
            r.pitch = sb_canvas_pitch(job->width);
        }
        r.planes = job->num_c * job->num_z;
        r.h = job->height;
        r.w = job->width;
        r.dtype = job->dtype;
    }
    return rc;
}

int sb_fuse_regions(sb_ctx* ctx, const sb_fuse_job* jobs, int32_t n_jobs, int lane) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    if (lane >= SB_NUM_LANES) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return sb_fail(ctx, SB_ERR_INVALID, "bad job list");
    if (n_jobs == 0) return SB_OK;
    if (n_jobs > 1) {
        bool batched = false;
        const int rc = sb_fuse_regions_impl(ctx, jobs, n_jobs, lane, &batched);
        if (rc || batched) {
            sb_lane(ctx, lane < 0 ? 0 : lane)->resident = ResidentCanvas();
            return rc;
        }
    }
    for (int j = 0; j < n_jobs; ++j) {                     // not one geometry / not device-resident: region by region
        const int rc = sb_fuse_region(ctx, &jobs[j], lane);
        if (rc) return rc;
    }
    return SB_OK;
}

int sb_estimate_flatfield(sb_ctx* ctx, const void* const* tiles, int32_t n_tiles, int32_t tile_h, int32_t tile_w, int dtype,
                          int mem, int32_t grid, double sigma, float* field_out, int out_mem) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    return sb_estimate_flatfield_impl(ctx, tiles, n_tiles, tile_h, tile_w, dtype, mem, grid, sigma, field_out, out_mem);
}

int64_t sb_pyramid_elems(int32_t n_planes, int32_t height, int32_t width, int32_t n_levels) {
    if (n_planes <= 0 || height <= 0 || width <= 0 || n_levels < 1) return -1;
    return sb_pyramid_elems_impl(n_planes, height, width, n_levels);
}

int sb_pyramid(sb_ctx* ctx, const void* src, int src_mem, int32_t n_planes, int32_t height, int32_t width,
               int64_t src_row_pitch, int dtype, int32_t n_levels, void* out, int out_mem, int lane) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    return sb_pyramid_impl(ctx, src, src_mem, n_planes, height, width, src_row_pitch, dtype, n_levels, out, out_mem, lane);
}

int sb_selftest(sb_ctx* ctx, int which, int64_t arg, uint64_t* out) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    SB_CHECK(ctx, out != nullptr, "out is NULL");
    if (which == SB_SELFTEST_STRETCH) return sb_selftest_stretch_impl(ctx, (int)arg, out);
    if (which == SB_SELFTEST_DIVIDE) return sb_selftest_div_impl(ctx, (int)arg, out);
    if (which == SB_SELFTEST_UMMA) return sb_selftest_umma_impl(ctx, (int)arg, out);
    return sb_fail(ctx, SB_ERR_INVALID, "unknown self-test %d", which);
}

int64_t sb_debug_read(sb_ctx* ctx, int lane, int which, void* out, int64_t max_bytes) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    Lane* l = sb_lane(ctx, lane);
    if (!l || which < 0 || which > 3 || !out) return sb_fail(ctx, SB_ERR_INVALID, "sb_debug_read: bad arguments");
    cudaDeviceSynchronize();
    const size_t nbytes = std::min<size_t>(l->dbg_bytes[which], max_bytes < 0 ? 0 : (size_t)max_bytes);
    if (nbytes && cudaMemcpy(out, l->dbg_ptr[which], nbytes, cudaMemcpyDeviceToHost) != cudaSuccess)
        return sb_fail(ctx, SB_ERR_CUDA, "sb_debug_read: copy failed");
    return (int64_t)nbytes;
}

int sb_debug_tc_profile(sb_ctx* ctx, long long* out48) {
    if (!ctx || !out48) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    cudaDeviceSynchronize();
    return sb_tc_profile_read(out48);
}

int sb_sync(sb_ctx* ctx, int lane) {
    SB_ENTER(ctx);
    if (!ctx) return SB_ERR_INVALID;
    if (lane >= SB_NUM_LANES) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    for (int i = 0; i < SB_NUM_LANES; ++i)
        if (lane < 0 || lane == i) {
            SB_CUDA(ctx, cudaStreamSynchronize(ctx->lanes[i].stream));
            const int rc = sb_register_complete(ctx, i);   // results of a parked sb_register_pairs_async land now
            if (rc) return rc;
        }
    return SB_OK;
}

int sb_lane_mark(sb_ctx* ctx, int lane) {
    SB_ENTER(ctx);
    Lane* l = sb_lane(ctx, lane);
    if (!l) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    if (!l->mark) SB_CUDA(ctx, cudaEventCreateWithFlags(&l->mark, cudaEventDisableTiming));
    SB_CUDA(ctx, cudaEventRecord(l->mark, l->stream));
    l->marked = true;
    return SB_OK;
}

int sb_lane_wait_mark(sb_ctx* ctx, int lane, int other) {
    SB_ENTER(ctx);
    Lane* l = sb_lane(ctx, lane);
    Lane* o = sb_lane(ctx, other);
    if (!l || !o) return sb_fail(ctx, SB_ERR_INVALID, "lane %d / %d out of range", lane, other);
    if (o->marked && l != o) SB_CUDA(ctx, cudaStreamWaitEvent(l->stream, o->mark, 0));
    return SB_OK;
}

int sb_set_lane_stream(sb_ctx* ctx, int lane, void* cuda_stream) {
    SB_ENTER(ctx);
    Lane* l = sb_lane(ctx, lane);
    if (!l) return sb_fail(ctx, SB_ERR_INVALID, "lane %d out of range", lane);
    SB_CUDA(ctx, cudaStreamSynchronize(l->stream));
    l->stream = cuda_stream ? (cudaStream_t)cuda_stream : l->own;
    return SB_OK;
}

int sb_flatfield_apply(sb_ctx* ctx, int channel, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w,
                       int dtype, int mem) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    ctx->lanes[0].resident = ResidentCanvas();   // these helpers stage through lane 0's canvas buffers
    if (dtype == SB_U8) return sb_flatfield_apply_u8(ctx, channel, tiles, out, n_tiles, tile_h, tile_w, mem);
    return sb_flatfield_apply_impl(ctx, channel, tiles, out, n_tiles, tile_h, tile_w, dtype, mem);
}

int sb_register_pairs(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    if (job && job->dtype == SB_U8) return sb_register_pairs_u8(ctx, job, out, false);
    return sb_register_pairs_impl(ctx, job, out, false);
}

int sb_register_pairs_async(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    if (job && job->dtype == SB_U8) return sb_register_pairs_u8(ctx, job, out, true);
    return sb_register_pairs_impl(ctx, job, out, true);
}

int sb_normalize(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w, int dtype, int mem) {
    if (!ctx) return SB_ERR_INVALID;
    SB_ENTER(ctx);
    ctx->lanes[0].resident = ResidentCanvas();   // these helpers stage through lane 0's canvas buffers
    if (dtype == SB_U8) return sb_normalize_u8(ctx, tiles, out, n_tiles, tile_h, tile_w, mem);
    return sb_normalize_impl(ctx, tiles, out, n_tiles, tile_h, tile_w, dtype, mem);
}

}  // extern "C"
