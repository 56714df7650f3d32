// Flat-field ESTIMATION from a sample of tiles -- an extension, not BaSiC.
//
// The reference fits its fields with the third-party BaSiCPy (get_flatfields, stitcher_process.py:505-571), which is
// neither in the reference tree nor in this image; when it cannot be imported the host mirror falls back to this
// robust estimator (definition and oracle: oracle/flatfield_ref.py):
//   cell means on a g x g grid per tile -> divide by the tile mean -> per-cell MEDIAN over the tiles -> separable
//   Gaussian on the grid -> mean 1 -> bilinear interpolation to H x W float32.
// Only the first step touches the pixels: one pass over the sample (n tiles x H x W x esize algorithmic bytes,
// HBM-bound, 128-bit loads, integer sums -- exact); everything after it works on n x g x g doubles.
#include "sb_common.cuh"

namespace {

constexpr int kMaxGrid = 512;
constexpr int kMaxTiles = 128;

__device__ __forceinline__ int cell_edge(int c, int n, int g) { return (int)(((long long)c * n + g - 1) / g); }

// grid (g, n_tiles): block (cy, t) sums the rows of cell row cy of tile t into g column cells.
template <typename T>
__global__ void __launch_bounds__(256) ff_cell_sums_kernel(const T* const* __restrict__ tiles, int h, int w, int g,
                                                           unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long cell[kMaxGrid];
    const int cy = blockIdx.x, t = blockIdx.y;
    for (int i = threadIdx.x; i < g; i += blockDim.x) cell[i] = 0ull;
    __syncthreads();
    const T* __restrict__ src = tiles[t];
    const int y0 = cell_edge(cy, h, g), y1 = cell_edge(cy + 1, h, g);
    constexpr int V = 16 / (int)sizeof(T);
    const bool vec = (w % V) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (vec) {
        const int nv = w / V;
        for (int j = threadIdx.x; j < nv; j += blockDim.x) {
            unsigned int col[V];
#pragma unroll
            for (int i = 0; i < V; ++i) col[i] = 0u;
            for (int y = y0; y < y1; ++y) {
                const uint4 q = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)y * w) + j);
                const unsigned int wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (sizeof(T) == 2) {
                        col[2 * k] += wd[k] & 0xffffu;
                        col[2 * k + 1] += wd[k] >> 16;
                    } else {
                        col[(4 * k) % V] += wd[k] & 0xffu;
                        col[(4 * k + 1) % V] += (wd[k] >> 8) & 0xffu;
                        col[(4 * k + 2) % V] += (wd[k] >> 16) & 0xffu;
                        col[(4 * k + 3) % V] += wd[k] >> 24;
                    }
                }
            }
            // the V pixels of a vector usually share one cell: merge runs before touching shared memory
            int cx_run = (int)(((long long)(j * V) * g) / w);
            unsigned long long run = 0ull;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int cx = (int)(((long long)(j * V + i) * g) / w);
                if (cx != cx_run) {
                    atomicAdd(&cell[cx_run], run);
                    cx_run = cx;
                    run = 0ull;
                }
                run += col[i];
            }
            atomicAdd(&cell[cx_run], run);
        }
    } else {
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            unsigned long long s = 0ull;
            for (int y = y0; y < y1; ++y) s += src[(size_t)y * w + x];
            atomicAdd(&cell[(int)(((long long)x * g) / w)], s);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g; i += blockDim.x) sums[((size_t)t * g + cy) * g + i] = cell[i];
}

// one block per tile: mean intensity = sum of the cell sums / (h * w), fixed summation order
__global__ void __launch_bounds__(256) ff_tile_mean_kernel(const unsigned long long* __restrict__ sums, int g, double inv_px,
                                                           double* __restrict__ tile_mean) {
    __shared__ unsigned long long part[256];
    const unsigned long long* s = sums + (size_t)blockIdx.x * g * g;
    unsigned long long acc = 0ull;
    for (int i = threadIdx.x; i < g * g; i += blockDim.x) acc += s[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_mean[blockIdx.x] = (double)part[0] * inv_px;
}

// one thread per cell: median over the usable tiles of (cell mean / tile mean); numpy's rule for even counts
__global__ void __launch_bounds__(128) ff_median_kernel(const unsigned long long* __restrict__ sums, const double* __restrict__ tile_mean,
                                                        int n, int h, int w, int g, double* __restrict__ med) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g * g) return;
    const int cy = c / g, cx = c - cy * g;
    const double count = (double)(cell_edge(cy + 1, h, g) - cell_edge(cy, h, g)) * (double)(cell_edge(cx + 1, w, g) - cell_edge(cx, w, g));
    double v[kMaxTiles];
    int m = 0;
    for (int t = 0; t < n; ++t) {
        const double tm = tile_mean[t];
        if (!(tm > 0.0)) continue;
        const double x = ((double)sums[(size_t)t * g * g + c] / count) / tm;
        int i = m++;
        while (i > 0 && v[i - 1] > x) { v[i] = v[i - 1]; --i; }      // insertion sort: n <= 128
        v[i] = x;
    }
    med[c] = m == 0 ? 1.0 : (m & 1) ? v[m >> 1] : 0.5 * (v[(m >> 1) - 1] + v[m >> 1]);
}

__device__ __forceinline__ int fold_symmetric(int j, int n) {       // d c b a | a b c d, any distance
    const int period = 2 * n;
    j %= period;
    if (j < 0) j += period;
    return j >= n ? period - 1 - j : j;
}

// separable Gaussian along x (axis = 1) or y (axis = 0); weights[2 r + 1] in constant-sized global memory
__global__ void __launch_bounds__(128) ff_smooth_kernel(const double* __restrict__ in, double* __restrict__ out, int g, int axis,
                                                        const double* __restrict__ weights, int r) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g * g) return;
    const int cy = c / g, cx = c - cy * g;
    double acc = 0.0;
    for (int k = -r; k <= r; ++k) {
        const int j = fold_symmetric((axis ? cx : cy) + k, g);
        acc += weights[k + r] * (axis ? in[cy * g + j] : in[j * g + cx]);
    }
    out[c] = acc;
}

// single block: divide the grid by its mean (fixed summation order)
__global__ void __launch_bounds__(256) ff_normalise_kernel(double* __restrict__ f, int cells) {
    __shared__ double part[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) acc += f[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    const double mean = part[0] / (double)cells;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) f[i] = mean > 0.0 ? f[i] / mean : 1.0;
}

// bilinear interpolation from the cell centres; block row y, threads along x (coalesced float stores)
__global__ void __launch_bounds__(256) ff_upsample_kernel(const double* __restrict__ f, int g, int h, int w, float* __restrict__ out) {
    const int y = blockIdx.y;
    double u = ((double)y + 0.5) * (double)g / (double)h - 0.5;
    u = fmin(fmax(u, 0.0), (double)g - 1.0);
    const int j0 = min((int)floor(u), max(g - 2, 0)), j1 = min(j0 + 1, g - 1);
    const double fy = u - (double)j0;
    const double* __restrict__ r0 = f + (size_t)j0 * g;
    const double* __restrict__ r1 = f + (size_t)j1 * g;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
        double ux = ((double)x + 0.5) * (double)g / (double)w - 0.5;
        ux = fmin(fmax(ux, 0.0), (double)g - 1.0);
        const int i0 = min((int)floor(ux), max(g - 2, 0)), i1 = min(i0 + 1, g - 1);
        const double fx = ux - (double)i0;
        const double top = r0[i0] * (1.0 - fx) + r0[i1] * fx;
        const double bot = r1[i0] * (1.0 - fx) + r1[i1] * fx;
        out[(size_t)y * w + x] = (float)(top * (1.0 - fy) + bot * fy);
    }
}

}  // namespace

int sb_estimate_flatfield_impl(sb_ctx* ctx, const void* const* tiles, int n_tiles, int tile_h, int tile_w, int dtype, int mem,
                               int grid, double sigma, float* field_out, int out_mem) {
    SB_CHECK(ctx, dtype == SB_U16 || dtype == SB_U8, "unknown pixel dtype %d", dtype);
    SB_CHECK(ctx, tiles && field_out && n_tiles > 0 && tile_h > 0 && tile_w > 0, "bad arguments");
    SB_CHECK(ctx, n_tiles <= kMaxTiles, "at most %d tiles per estimate (the reference samples at most 80), got %d", kMaxTiles, n_tiles);
    SB_CHECK(ctx, tile_h <= 65535 && tile_w <= 65535, "tile larger than 65535 pixels on a side");
    SB_CHECK(ctx, sigma >= 0.0 && sigma <= 64.0, "sigma out of range");
    for (int i = 0; i < n_tiles; ++i) SB_CHECK(ctx, tiles[i] != nullptr, "tile %d is NULL", i);
    int g = grid > 0 ? grid : 128;
    g = std::min(std::min(g, kMaxGrid), std::min(tile_h, tile_w));
    Lane* lane = sb_lane(ctx, 0);
    cudaStream_t st = lane->stream;
    const size_t es = dtype == SB_U8 ? 1 : 2;
    const size_t tile_bytes = (size_t)tile_h * tile_w * es;
    const size_t tile_stride = (tile_bytes + 255) & ~(size_t)255;

    // workspace: [pointer table][sums n g g u64][tile means][3 grids of doubles][weights][field H W f32 when the output is host]
    const int r = (int)std::ceil(3.0 * sigma);
    const size_t cells = (size_t)g * g;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_ptr = take((size_t)n_tiles * sizeof(void*));
    const size_t o_sums = take((size_t)n_tiles * cells * 8);
    const size_t o_mean = take((size_t)n_tiles * 8);
    const size_t o_g0 = take(cells * 8), o_g1 = take(cells * 8);
    const size_t o_w = take((size_t)(2 * r + 1) * 8);
    const size_t o_field = out_mem == SB_MEM_HOST ? take((size_t)tile_h * tile_w * 4) : 0;
    int rc = sb_reserve(ctx, lane->work, off);
    if (rc) return rc;
    uint8_t* wk = (uint8_t*)lane->work.p;

    std::vector<const void*> ptrs(tiles, tiles + n_tiles);
    if (mem == SB_MEM_HOST) {
        rc = sb_reserve(ctx, lane->tiles, tile_stride * n_tiles);
        if (rc) return rc;
        for (int i = 0; i < n_tiles; ++i) {
            void* d = (uint8_t*)lane->tiles.p + tile_stride * i;
            SB_CUDA(ctx, cudaMemcpyAsync(d, tiles[i], tile_bytes, cudaMemcpyHostToDevice, st));
            ptrs[i] = d;
        }
    }
    std::vector<double> wts((size_t)2 * r + 1, 1.0);
    double wsum = 0.0;
    for (int k = -r; k <= r; ++k) {
        wts[k + r] = sigma > 0.0 ? std::exp(-((double)k * k) / (2.0 * sigma * sigma)) : 1.0;
        wsum += wts[k + r];
    }
    for (double& v : wts) v /= wsum;
    // small pageable uploads: the call is synchronous, and the vectors outlive the stream sync below
    SB_CUDA(ctx, cudaMemcpyAsync(wk + o_ptr, ptrs.data(), (size_t)n_tiles * sizeof(void*), cudaMemcpyHostToDevice, st));
    SB_CUDA(ctx, cudaMemcpyAsync(wk + o_w, wts.data(), wts.size() * 8, cudaMemcpyHostToDevice, st));

    unsigned long long* sums = (unsigned long long*)(wk + o_sums);
    double* tmean = (double*)(wk + o_mean);
    double *g0 = (double*)(wk + o_g0), *g1 = (double*)(wk + o_g1);
    const double* dw = (const double*)(wk + o_w);
    float* d_field = out_mem == SB_MEM_HOST ? (float*)(wk + o_field) : field_out;

    const dim3 sgrid((unsigned)g, (unsigned)n_tiles);
    if (dtype == SB_U8) ff_cell_sums_kernel<uint8_t><<<sgrid, 256, 0, st>>>((const uint8_t* const*)(wk + o_ptr), tile_h, tile_w, g, sums);
    else ff_cell_sums_kernel<uint16_t><<<sgrid, 256, 0, st>>>((const uint16_t* const*)(wk + o_ptr), tile_h, tile_w, g, sums);
    ff_tile_mean_kernel<<<n_tiles, 256, 0, st>>>(sums, g, 1.0 / ((double)tile_h * (double)tile_w), tmean);
    const int cb = (int)((cells + 127) / 128);
    ff_median_kernel<<<cb, 128, 0, st>>>(sums, tmean, n_tiles, tile_h, tile_w, g, g0);
    ff_smooth_kernel<<<cb, 128, 0, st>>>(g0, g1, g, 1, dw, r);
    ff_smooth_kernel<<<cb, 128, 0, st>>>(g1, g0, g, 0, dw, r);
    ff_normalise_kernel<<<1, 256, 0, st>>>(g0, (int)cells);
    const dim3 ugrid((unsigned)std::min(64, (tile_w + 255) / 256), (unsigned)tile_h);
    ff_upsample_kernel<<<ugrid, 256, 0, st>>>(g0, g, tile_h, tile_w, d_field);
    ctx->launches += 7;
    SB_CUDA(ctx, cudaGetLastError());
    if (out_mem == SB_MEM_HOST)
        SB_CUDA(ctx, cudaMemcpyAsync(field_out, d_field, (size_t)tile_h * tile_w * 4, cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    return SB_OK;
}
