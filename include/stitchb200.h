/*
 * stitchb200.h -- C ABI of the B200-native registration + fusion hot path.
 *
 * This is the drop-in boundary for sohamazing/image-stitcher's L1 methods
 * (SURVEY.md section 8b).  The reference has no FFI of its own: the seam is the
 * method surface of its orchestrator classes, so every entry point below names
 * the reference method(s) whose arithmetic it replaces.  All citations are into
 * /root/reference/stitcher_process.py (stitcher.py holds identical copies).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in any signature
 *     (streams are passed as an opaque void*, NULL = the context's own stream);
 *   - the caller owns every buffer; the library never frees or keeps caller
 *     pointers after a synchronous call returns (asynchronous lane calls keep
 *     them until sb_sync(lane));
 *   - every function returns SB_OK (0) or a negative sb_status and records a
 *     message retrievable with sb_last_error(); the Python wrapper raises
 *     RuntimeError, which flows into the reference's existing handlers
 *     (stitch_region :954-956, run :2034-2037);
 *   - one sb_ctx per process and device, created lazily inside the forked
 *     worker (StitcherProcess.run), calls on one context serialised by the caller;
 *   - there is NO CPU fallback: without a CUDA device sb_create fails.
 */
#ifndef STITCHB200_H
#define STITCHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 2

typedef struct sb_ctx sb_ctx;

typedef enum sb_status {
    SB_OK = 0,
    SB_ERR_INVALID = -1,      /* bad argument / inconsistent geometry            */
    SB_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed             */
    SB_ERR_NOMEM = -3,        /* device or pinned-host allocation failed         */
    SB_ERR_UNSUPPORTED = -4,  /* valid request this build does not implement     */
    SB_ERR_NODEVICE = -5      /* no usable CUDA device (there is no CPU path)    */
} sb_status;

enum { SB_MEM_HOST = 0, SB_MEM_DEVICE = 1 };
enum { SB_U16 = 0, SB_U8 = 1 };                               /* pixel dtype (self.dtype, :160,:340).  SB_U8: every
                                                                 range follows the dtype like the reference's
                                                                 iinfo(self.dtype) -- stretch to 255 (:854), clip at
                                                                 255 (:838-841), uint8 canvas (:503); paste mode
                                                                 only, tile_w % 8 == 0                           */
enum { SB_FIELD_F32 = 0, SB_FIELD_F64 = 1 };                  /* flat/dark-field dtype (a12: result_type rule)  */
enum { SB_BLEND_PASTE = 0,                                    /* reference: crop-to-seam + overwrite (:789-817) */
       SB_BLEND_LINEAR = 1, SB_BLEND_FEATHER = 2 };           /* extensions, defined by oracle/blend_ref.py     */
enum { SB_LAYOUT_ROWMAJOR = 0,                                /* (1,C,Z,Hc,Wc) C-order, what save_region_* take */
       SB_LAYOUT_CHUNKED = 1 };                               /* zarr-v2 chunk order, edge chunks zero-padded   */
enum { SB_PREC_F32 = 0, SB_PREC_F64 = 1, SB_PREC_AUTO = 2 };  /* registration arithmetic                        */
enum { SB_DIR_HORIZONTAL = 0, SB_DIR_VERTICAL = 1 };

/* ------------------------------------------------------------------ lifecycle */

int sb_version(void);
/* Creates the context on CUDA device `device` (cudaSetDevice + streams + lanes). */
int sb_create(int device, sb_ctx** out);
void sb_destroy(sb_ctx* ctx);
/* Last error message of `ctx` (or of the failed sb_create when ctx == NULL). */
const char* sb_last_error(const sb_ctx* ctx);
/* Number of CUDA kernels this context has launched so far (bench.py's gpu_launches). */
int64_t sb_kernel_launches(const sb_ctx* ctx);
int sb_num_lanes(const sb_ctx* ctx);
int sb_device_sm_count(const sb_ctx* ctx);

/* ------------------------------------------------------------------ memory helpers
 * Pinned host buffers make the H2D/D2H legs of the host-memory entry points
 * asynchronous; device buffers let a caller without torch keep a plate resident. */
void* sb_host_alloc(sb_ctx* ctx, size_t bytes);
void sb_host_free(sb_ctx* ctx, void* p);
void* sb_device_alloc(sb_ctx* ctx, size_t bytes);
void sb_device_free(sb_ctx* ctx, void* p);
int sb_memcpy_h2d(sb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int sb_memcpy_d2h(sb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* Asynchronous copy on a lane's stream (ordered with that lane's kernels); kind: 0 = host->device,
 * 1 = device->host, 2 = device->device.  Host memory should come from sb_host_alloc. */
int sb_memcpy_async(sb_ctx* ctx, int lane, void* dst, const void* src, size_t bytes, int kind);
/* The same for a rectangle of `height` rows of `width_bytes` bytes inside pitched buffers (cudaMemcpy2DAsync): the
 * upload of the part of a tile that can reach the canvas -- in paste mode the pixels a later tile overwrites (:817) are
 * never read by sb_fuse_region, so a host pipeline need not send them (WellPipeline). */
int sb_memcpy2d_async(sb_ctx* ctx, int lane, void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                      size_t width_bytes, size_t height, int kind);

/* ------------------------------------------------------------------ flat / dark fields
 * Replaces the *storage* of self.flatfields (:158, :524): one H x W field per
 * monochrome channel index, kept on the device until cleared.  A channel without
 * a field passes through unchanged (:837 `if channel_idx in self.flatfields`).
 * Darkfield is an extension (the reference fits BaSiC with get_darkfield=False, :521). */
int sb_set_flatfield(sb_ctx* ctx, int channel, const void* field, int field_dtype, int mem,
                     int tile_h, int tile_w);
int sb_set_darkfield(sb_ctx* ctx, int channel, const void* field, int field_dtype, int mem,
                     int tile_h, int tile_w);
int sb_clear_fields(sb_ctx* ctx);

/* Standalone apply_flatfield_correction(tile, channel_idx) (:828-842):
 * out = trunc(clip(tile / flat[channel], 0, dtype_max)); n_tiles tiles of one channel. */
int sb_flatfield_apply(sb_ctx* ctx, int channel, const void* tiles, void* out, int n_tiles,
                       int tile_h, int tile_w, int dtype, int mem);

/* ------------------------------------------------------------------ fusion
 * Replaces stitch_region (:883-956) + init_output (:489-503) + place_tile (:739-769) +
 * place_single_channel_tile (:771-826) + apply_flatfield_correction (:828-842).
 * Geometry (a7 canvas size, a9 tile positions, a11 seam crops) is computed by the
 * caller with the reference's own integer/float64 rules and passed in as ints. */
typedef struct sb_tile {
    const void* px;       /* H x W pixels, row-major, tight rows; host or device per job.tile_mem   */
    int32_t x, y;         /* canvas position of the tile's UNCROPPED origin (x_pixel, y_pixel :928-942) */
    int32_t c, z;         /* destination plane: channel index, z level                               */
    int32_t crop_t, crop_b, crop_l, crop_r;   /* seam crops (:796-799); 0 in coordinate mode         */
} sb_tile;                /* array order == paste order: a later tile overwrites an earlier one (:817) */

typedef struct sb_fuse_job {
    const sb_tile* tiles;
    int32_t n_tiles;
    int32_t tile_h, tile_w;
    int32_t dtype;            /* SB_U16 | SB_U8                                                    */
    int32_t tile_mem;         /* SB_MEM_HOST | SB_MEM_DEVICE.  Device tiles must be 16-byte aligned,
                                 tile_w*elem % 16 == 0, and all pointers congruent modulo one row
                                 (true for any pool of equally sized tiles)                        */
    int32_t num_c, num_z;     /* canvas planes (1, C, Z, height, width)                            */
    int32_t height, width;    /* a7                                                                */
    int32_t apply_flatfield;  /* self.apply_flatfield                                              */
    int32_t blend;            /* SB_BLEND_*                                                        */
    int32_t blend_ov_x, blend_ov_y;   /* nominal overlap (px) = ramp width of SB_BLEND_LINEAR      */
    void* out;                /* canvas                                                            */
    int32_t out_mem;
    int32_t out_layout;       /* SB_LAYOUT_*                                                       */
    int64_t out_row_pitch;    /* elements between rows (ROWMAJOR).  Host: any >= width.  Device:
                                 multiple of 64.  0 = default (width on host, round_up(width,64) on device) */
    int32_t chunk_h, chunk_w; /* CHUNKED: chunk shape, multiples of 64 (reference: 2048 or 512)    */
    int32_t field_c0;         /* flat / dark field of tile channel c = the context's field of channel
                                 field_c0 + c.  0 for a whole region; a job that fuses a WINDOW of a
                                 region's planes (one (channel, z) band of a large mosaic on one of
                                 several GPUs) numbers its planes from 0 and names the
                                 window's first channel here                                        */
    int32_t reserved0;        /* 0                                                                  */
} sb_fuse_job;

/* lane < 0: run and wait.  lane in [0, sb_num_lanes): enqueue H2D -> kernels -> D2H on that
 * lane's stream and return; sb_sync(lane) waits.  Buffers must stay valid until then. */
int sb_fuse_region(sb_ctx* ctx, const sb_fuse_job* job, int lane);
/* Batched form for plates: n_jobs regions (wells) in one call.  When the regions share one geometry (same tile
 * lattice, crops, canvas and layout -- true for the wells of a plate) and tiles and canvases are device memory, they
 * are pasted by ONE kernel launch, channel by channel across the regions, so that a single flat-field is live in L2
 * at a time (stitch_region is called once per region by the reference, :1990; the result per region is identical).
 * Anything else is fused region by region, exactly like n_jobs calls of sb_fuse_region on that lane. */
int sb_fuse_regions(sb_ctx* ctx, const sb_fuse_job* jobs, int32_t n_jobs, int lane);
int sb_sync(sb_ctx* ctx, int lane);      /* lane < 0: all lanes */
/* Cross-lane ordering for pipelines: sb_lane_mark records a marker at the current end of `lane`'s stream;
 * sb_lane_wait_mark makes everything enqueued on `lane` afterwards wait for `other`'s latest marker (no host wait).
 * WellPipeline uses it to keep the uploads of consecutive regions back to back on the bus while each region's
 * kernels and download overlap the next upload. */
int sb_lane_mark(sb_ctx* ctx, int lane);
int sb_lane_wait_mark(sb_ctx* ctx, int lane, int other);
/* Use the caller's stream (e.g. torch's current stream) for a lane; NULL restores the lane's own. */
int sb_set_lane_stream(sb_ctx* ctx, int lane, void* cuda_stream);
/* Device-canvas row pitch (elements) the library uses for a given width: round_up(width, 64). */
int64_t sb_canvas_pitch(int32_t width);
/* Number of elements of a CHUNKED plane: ceil(h/chunk_h)*ceil(w/chunk_w)*chunk_h*chunk_w. */
int64_t sb_chunked_plane_elems(int32_t height, int32_t width, int32_t chunk_h, int32_t chunk_w);

/* ------------------------------------------------------------------ registration
 * Replaces calculate_horizontal_shift (:664-685) / calculate_vertical_shift (:687-708),
 * normalize_image (:844-855) and the call to skimage.registration.phase_cross_correlation
 * (un-vendored third-party; algorithm restated in oracle/pcc_ref.py) for a batch of pairs.
 * A batch of 2-3 pairs is the reference's calculate_shifts (:573-662); larger batches are
 * the all-pairs mode behind the declared-but-unused dynamic_registration flag. */
typedef struct sb_pair {
    const void* ref;      /* img_left / img_top   : full H x W tile                       */
    const void* mov;      /* img_right / img_bot                                          */
    int32_t dir;          /* SB_DIR_HORIZONTAL | SB_DIR_VERTICAL                          */
    int32_t reserved;
} sb_pair;

typedef struct sb_pair_result {
    int32_t dy, dx;           /* the reference's return value: (round(s0), round(s1 - Sw)) for H (:685),
                                 (round(s0 - Sh), round(s1)) for V (:708); Python half-even round  */
    double shift[2];          /* skimage's float64 sub-pixel shift (row, col), rebuilt from indices */
    int32_t coarse[2];        /* argmax |ifft2(P)| (row, col), first maximum in C order             */
    int32_t fine[2];          /* argmax of the upsampled-DFT window (row, col); -1 if upsample == 1 */
    float peak;               /* |cc| at the coarse peak                                                        */
    float second;             /* second-largest |cc| at any other pixel (for a half-pixel shift: the neighbour) */
    float runner_up;          /* largest |cc| outside the band of three lines centred on the peak, the lines
                                 being positions along the strip's LONG axis (image rows for horizontal pairs,
                                 image columns for vertical pairs), wrapping around: a confidence measure      */
    float fine_peak;          /* |.| at the maximum of the upsampled-DFT window (0 if upsample == 1)             */
    float fine_second;        /* second-largest value of that window                                            */
    int32_t ref_min, ref_max, mov_min, mov_max;   /* whole-tile min/max used by normalize_image     */
    int32_t precision;        /* SB_PREC_F32 or SB_PREC_F64: arithmetic that produced this result   */
} sb_pair_result;

typedef struct sb_register_job {
    const sb_pair* pairs;
    int32_t n_pairs;
    int32_t tile_h, tile_w;
    int32_t dtype;            /* SB_U16 | SB_U8 */
    int32_t mem;              /* where ref/mov live */
    int32_t max_overlap_x;    /* strip width of horizontal pairs (max_x_overlap, :608) */
    int32_t max_overlap_y;    /* strip height of vertical pairs (max_y_overlap, :609)  */
    int32_t upsample_factor;  /* reference: 10 (:684, :707) */
    int32_t precision;        /* SB_PREC_* ; AUTO = f32, then pairs are repeated in f64 (the reference's arithmetic)
                                 when (a) peak <= 4 x the expected noise maximum sqrt(2 ln N / N) or peak <=
                                 1.5 x runner_up (low confidence), or (b) peak - second <= 1e-4 peak, or
                                 fine_peak - fine_second <= 1e-4 fine_peak (near-tie: float32 could pick another
                                 index than complex128), or (c) the radix path packed two strips whose sums differ
                                 by more than 32 x into one transform (float32 digits are lost in proportion).
                                 sb_pair_result.precision says which arithmetic won. */
    int32_t lane;             /* stream to run on (ordered after that lane's earlier copies); the call
                                 still returns only when the results are on the host                  */
} sb_register_job;

int sb_register_pairs(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out);
/* Asynchronous form for pipelines (calculate_shifts of region i overlapping the upload of region i+1): enqueues the
 * chain on job->lane and returns; `out` is filled when sb_sync(job->lane) returns (or when the next registration
 * on that lane starts).  The pair list is copied; tiles and `out` must stay valid until then.  One parked job per lane. */
int sb_register_pairs_async(sb_ctx* ctx, const sb_register_job* job, sb_pair_result* out);

/* Standalone normalize_image(img) (:844-855) for n_tiles tiles (whole-tile min/max stretch,
 * float64 arithmetic, truncating cast). */
int sb_normalize(sb_ctx* ctx, const void* tiles, void* out, int n_tiles, int tile_h, int tile_w,
                 int dtype, int mem);

/* Flat-field ESTIMATE from a sample of n_tiles tiles of one channel (EXTENSION).  The reference fits its fields with
 * the third-party BaSiCPy (get_flatfields, stitcher_process.py:505-571: <= 32 random tiles per timepoint, stop above
 * 48); BaSiC is not reproduced here.  When it is not importable the host mirror calls this robust estimator instead:
 * g x g cell means per tile (g = grid, 0 = 128, at most 512) / tile mean -> per-cell median over the tiles -> separable
 * Gaussian (sigma in cells, symmetric boundary) -> mean 1 -> bilinear interpolation to tile_h x tile_w float32
 * (definition and oracle: oracle/flatfield_ref.py).  The result is what sb_set_flatfield() takes.  At most 128 tiles.
 * tiles: array of n_tiles pointers (host array; the pixels are `mem`).  Synchronous. */
int sb_estimate_flatfield(sb_ctx* ctx, const void* const* tiles, int32_t n_tiles, int32_t tile_h, int32_t tile_w, int dtype,
                          int mem, int32_t grid, double sigma, float* field_out, int out_mem);

/* Multiscale levels of a fused canvas: the reference saves regions through ome_zarr's Scaler(method="nearest")
 * (stitcher_process.py:1061-1062, stitcher.py:797-798), i.e. level l+1 = level l[..., ::2, ::2] with
 * ceil(h/2) x ceil(w/2) pixels.  Levels 1 .. n_levels-1 are written to `out` back to back, each a dense
 * (n_planes, h_l, w_l) array; sb_pyramid_elems() is their total element count (-1 on bad arguments).
 * `src` is level 0: n_planes planes of height x width, rows src_row_pitch elements apart (0 = width), planes
 * height rows apart -- or NULL for the canvas the lane's last row-major sb_fuse_region left on the device
 * (host or device output), which costs no upload; shape and dtype must match that job.  Any later sb_fuse_region on
 * the lane, and sb_flatfield_apply / sb_normalize (which stage through lane 0's buffers), invalidate it.
 * lane < 0: synchronous on lane 0; lane >= 0: enqueued, valid after sb_sync(lane). */
int64_t sb_pyramid_elems(int32_t n_planes, int32_t height, int32_t width, int32_t n_levels);
int sb_pyramid(sb_ctx* ctx, const void* src, int src_mem, int32_t n_planes, int32_t height, int32_t width,
               int64_t src_row_pitch, int dtype, int32_t n_levels, void* out, int out_mem, int lane);

/* ------------------------------------------------------------------ test hook
 * Exhaustive on-device proofs of the two primitives whose exactness the parity claims rest on (tests/test_exhaustive_gpu.py):
 *   SB_SELFTEST_STRETCH, arg = 255 | 65535: the integer forms of normalize_image (:844-855) fused into the strip loads
 *       (radix kernels and tensor-core converters, each with its hand-over of exact quotients to the float64 sequence)
 *       equal the float64 expression for every (v - min, max - min) pair of that pixel range; out[3] = how many pairs
 *       took the exact-quotient hand-over;
 *   SB_SELFTEST_DIVIDE, arg = e in [-5, 19]: the packed float32 divide + truncation / rounding of the paste kernels
 *       equals IEEE a / b and trunc(clip(.)) (:838-841) for every uint16 a and every float32 b in [2^e, 2^(e+1)).
 *   SB_SELFTEST_UMMA, arg = 0: a 128 x 112 x 56 product through the tcgen05 building blocks of the registration
 *       kernels (shared-memory descriptors, 3-term tf32 split, TMEM read-back) against a float64 product on the device;
 *       out[2] = largest |error| * 1e12, out[3] = 0xDEAD if the tensor pipeline never signalled completion.
 *       (arg >= 1000: tcgen05.mma issue-rate probe, a tuning hook -- scratch/umma_rate.py.)
 * out[0] = cases checked, out[1] = mismatches, out[2] = smallest mismatching case key (or ~0).  out holds 4 values. */
enum { SB_SELFTEST_STRETCH = 0, SB_SELFTEST_DIVIDE = 1, SB_SELFTEST_UMMA = 2 };
int sb_selftest(sb_ctx* ctx, int which, int64_t arg, uint64_t* out);
/* Test hook: intermediates of the lane's last float32 registration group (first sub-batch), as left in its workspace:
 * which = 0: half spectra along the short strip axis Zh[pair][image][k][y] (tensor-core path only), 1: the inverse
 * column transform Y[pair][x][y], 2: the normalised cross-power R[pair][x][y], 3: the first stage of the upsampled
 * DFT T[pair][u][y]; complex64, y (the strip's long axis) fastest.  Returns the bytes copied (<= max_bytes) or a negative status. */
int64_t sb_debug_read(sb_ctx* ctx, int lane, int which, void* out, int64_t max_bytes);
/* Tuning hook: per-role wait / work cycle counters of the tensor-core kernels, filled only by a -DSB_TC_PROFILE build
 * (3 modes x 16 counters of block 0; read-and-clear).  Returns the number of values written (0 in a normal build). */
int sb_debug_tc_profile(sb_ctx* ctx, long long* out48);

#ifdef __cplusplus
}
#endif
#endif /* STITCHB200_H */
