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
};

// Launches the layer on `st`.  first = gather from log-mel rows with (x - mu) / sigma.
int iss_launch_conv(const ConvArgs &a, bool first, cudaStream_t st);
