// conv_gemm.cuh -- implicit-GEMM convolution / dense layer shared by the CNN (K2)
// and ResNet101 (K5) operators.  NHWC float32 activations; B = [K][N] weights.
#pragma once
#include "iss_common.cuh"

#define ISS_F_RESIDUAL 64     /* internal: + residual[m][n] after the pre-affine, before ReLU */

struct ConvArgs {
    const float *in;        // NHWC activations, or the log-mel rows when `first`
    const float *w;         // [K][N]
    const float *bias, *pre_scale, *pre_shift, *post_scale, *post_shift;
    const float *residual;  // [M][N] or nullptr
    float *out;             // [M][N]
    int64_t M;              // n_img * OH * OW
    int N, K;
    int H, W, C;            // input dims
    int OH, OW;
    int KH, KW, SH, SW, PT, PL;
    int flags;
    // `first` (sliding z-normalised patches of the segmenter CNNs) only
    int ld;
    const int32_t *row0; const float *mu; const float *sigma;
    // tensor-core path: weights transposed + split, [N][Kp] each (nullptr => fp32 CUDA-core kernel)
    const float *wt_hi; const float *wt_lo; const float *wt_tiled; int Kp;
    int debug_same_addr;    // timing experiment only (ISS_B200_TC_DEBUG=1): every gather hits the same 128 bytes
    // slab kernel (conv_gemm_tc3_kernel) only, filled in by iss_launch_conv_tc
    int slab_R;             // output rows (of width OW) per 128-row GEMM tile
    int slab_rows;          // input rows the slab is sized for
    int64_t in_elems;       // floats in `in` (loads past the end are zero-filled)
    unsigned long long *prof;   // ISS_B200_TC_PROF=1: per-role wait-cycle counters (experiments only)
    // fused 2x2/2 'valid' max-pooling of the INPUT (ISS_B200_FUSE_POOL=1, slab kernel only): `in` is the un-pooled
    // NHWC tensor [n_img][inH][inW][C]; H = inH / 2 and W = inW / 2 are the pooled dims the convolution sees
    int pool_in, inH, inW;
};

#define ISS_GEMM_FP32  0      /* fp32 CUDA cores (conv_gemm.cu) */
#define ISS_GEMM_TC_SS 1      /* tcgen05 3xTF32, A and B from shared memory */
#define ISS_GEMM_TC_TS 2      /* tcgen05 3xTF32, A from tensor memory, B from shared memory */
#define ISS_GEMM_TC_F16 3     /* EXPERIMENTAL (not yet run on hardware): fp16 hi/lo split, kind::f16, conv_gemm_tc_f16.cu; other layers as engine 2 */
#ifndef ISS_GEMM_DEFAULT
#define ISS_GEMM_DEFAULT ISS_GEMM_TC_TS
#endif

extern "C" int iss_get_gemm_mode(void);
bool iss_conv_tc_eligible(const ConvArgs &a);
int iss_launch_conv_tc(const ConvArgs &a, int mode, cudaStream_t st);
// true if a convolution described by `a` (pooled dims, weights set) can take its input through the fused pooling path
bool iss_conv_poolin_supported(const ConvArgs &a, int mode);
// W[K][N] -> device buffer: [2][N][Kp] row-major (hi, lo) then the tiled/pre-swizzled image [2*N*Kp]; Kp = K rounded up to 32
int iss_prepare_tc_weights(const float *h_w, int K, int N, float **d_out, int *Kp_out);

// Launches the layer on `st`.  first = gather from log-mel rows with (x - mu) / sigma.
int iss_launch_conv(const ConvArgs &a, bool first, cudaStream_t st);
