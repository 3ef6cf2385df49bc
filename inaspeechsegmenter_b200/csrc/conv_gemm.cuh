// conv_gemm.cuh -- implicit-GEMM convolution / dense layer shared by the CNN (K2)
// and ResNet101 (K5) operators.  NHWC float32 activations; B = [K][N] weights.
#pragma once
#include <cuda_fp16.h>

#include "iss_common.cuh"

// split-half word of the fp16-split engine: lo16 << 16 | hi16 with hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ uint32_t iss_pack_split(float y)
{
    const __half hh = __float2half_rn(y);
    const __half ll = __float2half_rn(y - __half2float(hh));
    return (uint32_t)__half_as_ushort(hh) | ((uint32_t)__half_as_ushort(ll) << 16);
}
// two values at once: the pair conversions map to F2FP (full-rate ALU) instead of the scalar F2F (conversion unit)
__device__ __forceinline__ void iss_pack_split2(float y0, float y1, uint32_t &w0, uint32_t &w1)
{
    const __half2 hh = __floats2half2_rn(y0, y1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
    const uint32_t h = *reinterpret_cast<const uint32_t *>(&hh), l = *reinterpret_cast<const uint32_t *>(&ll);
    w0 = __byte_perm(h, l, 0x5410);
    w1 = __byte_perm(h, l, 0x7632);
}
__device__ __forceinline__ float iss_unpack_split(uint32_t w)
{
    return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))) + __half2float(__ushort_as_half((unsigned short)(w >> 16)));
}

#define ISS_F_RESIDUAL 64     /* internal: + residual[m][n] after the pre-affine, before ReLU */

// First layer folded into the slab fill of the convolution behind it (conv_gemm_tc_f16d.cu, FIRST mode): the
// first Conv2D of the segmenter CNNs has ONE input channel and its input is the z-normalised patch
// (x - mu_j) / sigma_j (segmenter.py:82), so by linearity
//     conv(x^)[t, f, c] = (Y[row0_j + t, f, c] - mu_j * S_c) / sigma_j,   Y = conv(raw log-mel), S_c = sum of the filter
// and Y is shared by all the patches that overlap a frame (97 % overlap: hop 2 of 68 frames).  Y is computed once per
// batch in float64 and stored as a two-float number Yh + Yl (first_linear_kernel); bias and the BatchNorm affine
// (scale s_c, shift t_c) fold with the normalisation into ONE multiply-add per value,
//     a[t, f, c] = Y * alpha_jc + beta_jc,   alpha_jc = s_c / sigma_j,   beta_jc = (b_c s_c + t_c) - mu_j S_c alpha_jc,
// with alpha rounded to float32 FIRST and beta (float64, stored as two floats) computed from the rounded alpha, so the
// cancellation between Y alpha and beta is exact to ~1e-7 relative (first_coef_kernel).  The slab fill evaluates
// fma(Yh, alpha, beta_h) + fma(Yl, alpha, beta_l), ReLU, the optional second affine, and writes fp16 hi / lo planes
// straight into shared memory: the first layer's output (283 KB per patch, 47 % of all activation traffic) never
// exists in HBM.
struct FirstFuse {
    const float *Yh, *Yl;           // [y_rows][W * C] each, row r = conv of log-mel frames y_f0 + r ..; Y = Yh + Yl
    int64_t y_f0;                   // log-mel frame of Y row 0
    int64_t y_rows;
    const int32_t *row0;            // per patch of the batch: first log-mel frame
    const float *coef;              // per patch [3][C]: alpha, beta_hi, beta_lo
    const float *post_scale, *post_shift;   // first layer's affine behind the ReLU (nullptr = absent)
    int flags;                      // ISS_F_* of the first layer
    int64_t n_img;                  // patches in the batch
};

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
    // fp16-split engine (conv_gemm_tc_f16.cu): tiled fp16 hi/lo weight image (nullptr => not prepared) and its scale
    const void *wt_f16; float wt_f16_inv_scale;
    // activation formats of that engine: fp32 values (0) or split-half words lo16 << 16 | hi16 (1)
    int in_packed, out_packed;
    int residual_packed;    // `residual` holds split-half words (fp16-split kernels only)
    int pool_h, pool_w;     // direct kernel only, > 0: `in` is the un-pooled [n][pool_h][pool_w][C] tensor, 2x2 / stride-2 max taken in the slab fill
    const FirstFuse *first;         // host pointer, non-null: FIRST mode of the direct kernel (copied into the kernel parameters)
    // slab kernel (conv_gemm_tc_f16.cu) only, filled in by iss_launch_conv_tc_f16
    int slab_R;             // output rows (of width OW) per 128-row GEMM tile
    int slab_rows;          // input rows the slab is sized for
    int slab_tpi;           // > 0: tiles per image (tiles never straddle images); 0: tiles over the global row sequence
    int64_t in_elems;       // floats in `in` (loads past the end are zero-filled)
};

#define ISS_GEMM_FP32  0      /* fp32 CUDA cores (conv_gemm.cu) */
#define ISS_GEMM_TC_TS 2      /* tcgen05 3xTF32 (kind::tf32), A from tensor memory, B from shared memory: every layer shape */
#define ISS_GEMM_TC_F16 3     /* fp16 hi/lo split on kind::f16 for the un-padded stride-1 slab convolutions (conv_gemm_tc_f16.cu); other layers as engine 2 */
#ifndef ISS_GEMM_DEFAULT
#define ISS_GEMM_DEFAULT ISS_GEMM_TC_F16
#endif

extern "C" int iss_get_gemm_mode(void);
bool iss_conv_tc_eligible(const ConvArgs &a);
int iss_launch_conv_tc(const ConvArgs &a, int mode, cudaStream_t st);
// W[K][N] -> device buffer: [2][N][Kp] row-major (hi, lo) then the tiled/pre-swizzled image [2*N*Kp]; Kp = K rounded up to 32
int iss_prepare_tc_weights(const float *h_w, int K, int N, float **d_out, int *Kp_out);
// W[K][N] -> fp16 hi/lo image of engine 3 (*d_out stays nullptr when the shape is not one that engine takes)
int iss_prepare_f16_weights(const float *h_w, int K, int N, void **d_out, float *inv_scale);
// does engine 3's slab kernel cover this layer (geometry + prepared image + shared memory)?
bool iss_conv_f16_slab_covers(const ConvArgs &a);
// does engine 3's direct kernel (conv_gemm_tc_f16d.cu: both operands from shared memory) cover this layer?  It reads
// split-half words (a.in_packed) or, with a.first set, evaluates the one-channel first layer in its slab fill
bool iss_conv_f16_direct_covers(const ConvArgs &a);
// does engine 3's gather kernel (conv_gemm_tc_f16g.cu: any stride / padding / 1x1) cover this layer?
bool iss_conv_f16_gather_covers(const ConvArgs &a);
// n-tile width of engine 3 for N output channels (must agree between the weight image and the launch)
int iss_f16_bn_for(int N);

// Launches the layer on `st`.  first = gather from log-mel rows with (x - mu) / sigma.
int iss_launch_conv(const ConvArgs &a, bool first, cudaStream_t st);
