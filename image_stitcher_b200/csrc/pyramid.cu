// Nearest-neighbour x2 multiscale levels of a fused canvas.
//
// The reference hands the stitched region to ome_zarr's writer with ``Scaler(method="nearest")``
// (stitcher_process.py:1061-1062, stitcher.py:797-798): level l+1 keeps every second row and column of level l,
// i.e. ``level[..., ::2, ::2]`` with sizes ceil(h/2) x ceil(w/2).  Pure byte movement, HBM-bound: level 1 touches
// the even rows of level 0 (every sector of them) and writes a quarter of its pixels:
//     algorithmic bytes of level l+1 = esize * (h_l/2 * w_l  +  h_l/2 * w_l/2) = 0.375 x the bytes of level l.
// The levels are produced while the canvas is still resident in the lane's device buffer, so the host never
// re-reads the canvas with a strided NumPy copy and nothing is uploaded again.
#include "sb_common.cuh"

namespace {

// One thread per group of V = 8 / sizeof(T) output pixels (one 64-bit store); the group boundaries follow the 8-byte
// alignment of the DENSE destination, so the ragged head / tail of a row (width not a multiple of V) are the only
// narrow stores.  Sources are read with V independent loads at stride 2: a warp covers 32 * V * 2 contiguous pixels.
template <typename T>
__global__ void __launch_bounds__(256)
pyramid_down2_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows_total, int dh, int dw, int sh,
                     int64_t spitch) {
    constexpr int V = 8 / (int)sizeof(T);
    for (int64_t r = blockIdx.y; r < rows_total; r += gridDim.y) {
        const int64_t p = r / dh;
        const int y = (int)(r - p * dh);
        const T* __restrict__ s = src + (p * sh + 2 * (int64_t)y) * spitch;
        T* __restrict__ d = dst + r * dw;
        int head = (int)(((8 - (reinterpret_cast<uintptr_t>(d) & 7)) & 7) / sizeof(T));
        if (head > dw) head = dw;
        const int n_groups = (dw - head) / V;
        const int tail0 = head + n_groups * V;
        for (int g = blockIdx.x * blockDim.x + threadIdx.x - 1; g <= n_groups; g += gridDim.x * blockDim.x) {
            if (g >= 0 && g < n_groups) {
                const int x = head + g * V;
                T v[V];
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = __ldg(s + 2 * (int64_t)(x + i));
                uint64_t pack = 0;
#pragma unroll
                for (int i = 0; i < V; ++i) pack |= (uint64_t)v[i] << (8 * (int)sizeof(T) * i);
                __stcs(reinterpret_cast<unsigned long long*>(d + x), (unsigned long long)pack);
            } else {
                const int x0 = g < 0 ? 0 : tail0, x1 = g < 0 ? head : dw;
                for (int x = x0; x < x1; ++x) d[x] = __ldg(s + 2 * (int64_t)x);
            }
        }
    }
}

template <typename T>
int launch_level(sb_ctx* ctx, cudaStream_t st, const T* src, T* dst, int64_t n_planes, int sh, int sw, int64_t spitch) {
    constexpr int V = 8 / (int)sizeof(T);
    const int dh = (sh + 1) / 2, dw = (sw + 1) / 2;
    const int64_t rows_total = n_planes * dh;
    const int groups = dw / V + 2;
    const int bx = (groups + 255) / 256;
    const int64_t by = rows_total < 65535 ? rows_total : 65535;
    dim3 grid((unsigned)bx, (unsigned)by);
    pyramid_down2_kernel<T><<<grid, 256, 0, st>>>(src, dst, rows_total, dh, dw, sh, spitch);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

}  // namespace

int64_t sb_pyramid_elems_impl(int64_t n_planes, int height, int width, int n_levels) {
    int64_t total = 0;
    int h = height, w = width;
    for (int l = 1; l < n_levels; ++l) {
        h = (h + 1) / 2;
        w = (w + 1) / 2;
        total += n_planes * (int64_t)h * w;
    }
    return total;
}

int sb_pyramid_impl(sb_ctx* ctx, const void* src, int src_mem, int n_planes, int height, int width, int64_t src_row_pitch,
                    int dtype, int n_levels, void* out, int out_mem, int lane_idx) {
    SB_CHECK(ctx, dtype == SB_U16 || dtype == SB_U8, "unknown pixel dtype %d", dtype);
    SB_CHECK(ctx, n_planes > 0 && height > 0 && width > 0 && n_levels >= 1, "bad canvas shape / level count");
    SB_CHECK(ctx, lane_idx < SB_NUM_LANES, "lane %d out of range", lane_idx);
    const bool sync_call = lane_idx < 0;
    Lane* lane = sb_lane(ctx, sync_call ? 0 : lane_idx);
    const size_t es = dtype == SB_U8 ? 1 : 2;
    cudaStream_t st = lane->stream;
    if (n_levels == 1) return SB_OK;
    SB_CHECK(ctx, out != nullptr, "out is NULL");

    const void* d_src = src;
    int64_t spitch = src_row_pitch ? src_row_pitch : width;
    if (!src) {
        // the canvas the lane's last row-major sb_fuse_region left on the device
        const ResidentCanvas& rc = lane->resident;
        SB_CHECK(ctx, rc.p != nullptr, "no resident canvas on this lane (run a row-major sb_fuse_region first)");
        SB_CHECK(ctx, rc.planes == n_planes && rc.h == height && rc.w == width && rc.dtype == dtype,
                 "resident canvas is %dx%dx%d dtype %d, asked for %dx%dx%d dtype %d", rc.planes, rc.h, rc.w, rc.dtype,
                 n_planes, height, width, dtype);
        d_src = rc.p;
        spitch = rc.pitch;
    } else if (src_mem == SB_MEM_HOST) {
        SB_CHECK(ctx, spitch >= width, "src_row_pitch < width");
        const size_t bytes = (size_t)n_planes * height * width * es;
        int rc = sb_reserve(ctx, lane->pyr_src, bytes);
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpy2DAsync(lane->pyr_src.p, (size_t)width * es, src, (size_t)spitch * es, (size_t)width * es,
                                       (size_t)n_planes * height, cudaMemcpyHostToDevice, st));
        d_src = lane->pyr_src.p;
        spitch = width;
    } else {
        SB_CHECK(ctx, spitch >= width, "src_row_pitch < width");
    }

    const int64_t total = sb_pyramid_elems_impl(n_planes, height, width, n_levels);
    char* d_out = (char*)out;
    if (out_mem == SB_MEM_HOST) {
        int rc = sb_reserve(ctx, lane->pyr_out, (size_t)total * es);
        if (rc) return rc;
        d_out = (char*)lane->pyr_out.p;
    }
    const char* cur = (const char*)d_src;
    int64_t cur_pitch = spitch;
    int h = height, w = width;
    int64_t off = 0;
    for (int l = 1; l < n_levels; ++l) {
        const int dh = (h + 1) / 2, dw = (w + 1) / 2;
        char* dst = d_out + (size_t)off * es;
        const int64_t lvl = (int64_t)n_planes * dh * dw;
        int rc = dtype == SB_U8
                     ? launch_level<uint8_t>(ctx, st, (const uint8_t*)cur, (uint8_t*)dst, n_planes, h, w, cur_pitch)
                     : launch_level<uint16_t>(ctx, st, (const uint16_t*)cur, (uint16_t*)dst, n_planes, h, w, cur_pitch);
        if (rc) return rc;
        cur = dst;
        cur_pitch = dw;
        h = dh;
        w = dw;
        off += lvl;
    }
    if (out_mem == SB_MEM_HOST) SB_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)total * es, cudaMemcpyDeviceToHost, st));
    if (sync_call) SB_CUDA(ctx, cudaStreamSynchronize(st));
    return SB_OK;
}
