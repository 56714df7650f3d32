// Interface between reg.cu (the registration chain) and reg_tc.cu (its tensor-core / warp-FFT stages).
#pragma once

#include "sb_common.cuh"



struct TcPlan {
    bool ok = false;       // this strip shape takes the tensor-core path
    int Sh = 0, n = 0;     // frame: long axis (1024), short axis
    int nb = 0;            // n / 2 + 1 bins of the half spectrum
    int NP = 0;            // nb rounded up to 16: accumulator columns per part
    int nchunks = 0;       // K chunks of 8 (nb rounded up to 8, / 8)
    int pitch_w = 0;       // (unused)
    int stg_bytes = 0;     // bytes of one staged tile of raw strip rows (forward transform)
    int smem_fwd = 0;      // dynamic shared memory of fwd_x_tc_kernel
    const uint8_t* Bfwd = nullptr;   // cos / -sin operand images of the forward transform (device)
    const uint8_t* Binv = nullptr;   // ... and of the inverse transform
    int smem_inv = 0;
    int stg_bufs = 1, acc_bufs = 1;  // staged-tile buffers (forward), TMEM accumulator buffers
    bool inverse = false;            // the inverse short-axis transform + argmax runs on the tensor cores too
    const void* tw1024 = nullptr;    // [k2][l] table of W1024^(l k2) (device, float2)
};

// Decides whether a frame of Sh x n (after the transposed-frame swap) takes the tensor-core path and builds / caches
// its tables.  Returns SB_OK with plan->ok == false when it does not.
int sb_tc_plan(sb_ctx* ctx, int Sh, int n, TcPlan* plan);
// Zh[p][img][k][y] <- half spectra along the short axis of both strips of `n_pairs` pairs (d_pairs: PairDesc[]).
int sb_tc_forward(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, const void* d_pairs, int n_pairs, const int2* d_mm, int tile_w,
                  int swap, int maxval, void* Zh, int* d_nonzero, int* d_fault);
// Column pass: FFT along the long axis, cross-power, inverse FFT.  R / Y are the full [x][y] arrays of the radix chain
// (mirror = 1: the conjugate columns are written too) or half arrays of nb lines (mirror = 0, lines_out = nb).
int sb_tc_columns(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, int n_pairs, const void* Zh, void* R, void* Y, int lines_out,
                  int mirror);
// Inverse short-axis transform of the half spectrum Y (first nb of `lines_in` lines per pair) fused with the argmax:
// best[pair][Sh / 128] (CtaBest) and rowmax[pair][Sh], the outputs of rows_inv_argmax_kernel.
int sb_tc_inverse(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, int n_pairs, const void* Y, int lines_in, int swap, void* best,
                  float* rowmax, int* d_fault);
size_t sb_tc_zh_bytes(const TcPlan& plan, int n_pairs);
// Upsampled-DFT refinement, first stage, on the tensor cores (rs <= 16): twiddle tables from the coarse peaks (B operand
// images into `Bimg`, sb_tc_updft_table_bytes() bytes; Ey[pair][v][y]) and T[pair][u][y] = sum_x conj(R[y][x]) Ex[u][x].
// R holds `lines_in` lines per pair: all n, or the half array of n/2 + 1 (the rest follows from R[y][n-kx] = conj(R[-y][kx])).
size_t sb_tc_updft_table_bytes(const TcPlan& plan, int n_pairs);
int sb_tc_updft_rows(sb_ctx* ctx, cudaStream_t st, const TcPlan& plan, int n_pairs, const void* d_peaks, int uf, int rs,
                     int dftshift, const void* R, int lines_in, void* Bimg, void* Ey, void* Tm, int* d_fault);

int sb_tc_profile_read(long long* out48);   // cycle counters of a -DSB_TC_PROFILE build (0 = not such a build)
