// cnn.cu -- K2: sliding-patch z-normalisation + CNN forward (generic Keras-style
// Sequential: Conv2D / MaxPooling2D / Dense with fused bias + BatchNorm affine +
// ReLU epilogues, softmax head, non-finite override).
//
// Reference semantics: _get_patches (inaSpeechSegmenter/segmenter.py:76-88),
// DnnSegmenter.__call__ (:135-179: band select :146-147, gather of `inlabel`
// segments :156-162, keras predict :163, r[~finite] = 0.5 :175).  The network
// itself lives only in keras_*_cnn.hdf5 (not in the reference tree), so layers
// arrive as iss_layer_desc records (Keras channels_last semantics).
//
// Data layout: activations are NHWC float32 ([patch][row=time][col=mel][chan]),
// i.e. the implicit-GEMM A operand has K = (kh, kw, cin) contiguous in cin, and
// Keras kernels [kh][kw][cin][cout] are the row-major B operand [K][N] verbatim.
// The 68 x nmel patches are NEVER materialised (the reference builds ~1 GB per
// audio-hour per network): the first conv layer gathers straight from the
// log-mel rows and applies (x - mean) / std of its patch on the fly.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "conv_gemm.cuh"

namespace {

constexpr int PATCH_H = 68;            // frames per patch (segmenter.py:149)
constexpr int PATCH_HOP = 2;           // patch hop in frames (segmenter.py:149)
constexpr int PATCH_LFILL = PATCH_H / (2 * PATCH_HOP);     // 17 left replicas (segmenter.py:83)
// patches per layer-by-layer sweep: 8192 measured 6 % faster than 2048 (fewer partial waves of the persistent kernels);
// ISS_B200_CNN_BATCH: experiments; read once -- the workspace is sized with it
static int cnn_batch()
{
    static const int b = [] { const char *e = getenv("ISS_B200_CNN_BATCH"); const int v = e ? atoi(e) : 0; return v >= 64 && v <= 16384 ? v : 8192; }();
    return b;
}
#define CNN_BATCH cnn_batch()

struct Layer {
    iss_layer_desc d;
    int in_h, in_w, in_c;
    int out_h, out_w, out_c;
    float *d_wt = nullptr;      // tensor-core weights [2][N][Kp] (eligible layers only)
    int Kp = 0;
    void *d_wt_f16 = nullptr;   // fp16 hi/lo image of engine 3 (KHxKW > 1 convolutions with K % 64 == 0, N % 64 == 0)
    float f16_inv_scale = 1.f;
};

}  // namespace

struct iss_cnn {
    iss_ctx *ctx;
    std::vector<Layer> layers;
    float *d_blob;
    int64_t blob_len;
    int in_h, in_w;
    int n_classes;
    int64_t max_act;       // largest per-patch activation (floats) over all layer outputs
    double flops;
    double *d_first_S = nullptr;   // [cout of layer 0] float64 sums of the first layer's filter taps (FirstFuse)
    // live profiling of one layer (bench.py roofline)
    int prof_layer = -1;
    std::vector<cudaEvent_t> prof_ev;      // pairs: [2i] start, [2i+1] stop
    size_t prof_used = 0;
    double prof_flops = 0;
};

namespace {

// ------------------------------------------------------------------ patch bookkeeping
struct PatchArrays {
    int32_t *row0;         // first log-mel row of the patch
    float *mu, *sigma;     // float32 statistics (population std)
    uint8_t *finite;
};

__global__ void patch_index_kernel(const int32_t *__restrict__ seg_start, const int64_t *__restrict__ seg_off,
                                   int n_seg, int64_t n, int64_t U, int edge_left, int edge_right,
                                   int32_t *__restrict__ row0)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = n_seg;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= i) lo = mid; else hi = mid; }
    const int64_t p = seg_start[lo] + (i - seg_off[lo]);        // padded patch index
    int64_t j = p - (edge_left ? PATCH_LFILL : 0);              // un-replicated window index
    if (j < 0) j = 0;
    if (edge_right && j > U - 1) j = U - 1;
    row0[i] = (int32_t)(j * PATCH_HOP);
}

// one warp per patch: mean / population std over the 68 x w values (np.mean / np.std, segmenter.py:82)
__global__ void __launch_bounds__(256)
patch_stats_kernel(const float *__restrict__ mspec, int ld, int w, int64_t n, PatchArrays pa)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const float *base = mspec + (int64_t)pa.row0[i] * ld;
    const int cnt = PATCH_H * w;
    double s1 = 0.0, s2 = 0.0;
    for (int e = lane; e < cnt; e += 32) {
        const int r = e / w, c = e - r * w;
        const double v = (double)base[r * ld + c];
        s1 += v; s2 += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) {
        const double m = s1 / cnt;
        double var = s2 / cnt - m * m;
        if (var < 0.0) var = 0.0;
        const float mu = (float)m, sg = (float)sqrt(var);
        // data = (data - mean) / std: finite iff every input is finite and std > 0
        const bool fin = isfinite(mu) && isfinite(sg) && sg > 0.0f;
        pa.mu[i] = mu; pa.sigma[i] = sg; pa.finite[i] = fin ? 1 : 0;
    }
}

// ------------------------------------------------------------------ first layer: direct conv on the z-normalised patch
// The first Conv2D has one input channel (K = kh*kw ~ 20), far too thin for a GEMM tile.  One CTA
// stages a whole patch in shared memory, normalising each log-mel value exactly once with the
// reference's arithmetic ((x - mean) / std in float32, segmenter.py:82), zero-padded for 'same'
// convolutions; a thread owns 4 output channels x 4 consecutive output columns, so a filter tap is
// one LDS.128 (weights) + 4 broadcast LDS.32 (inputs) for 16 FMAs.  Taps are accumulated in the
// same (kh, kw) order as the implicit-GEMM kernel, so both produce identical bits.
struct FirstArgs {
    const float *mspec; int ld;
    const int32_t *row0; const float *mu; const float *sigma;
    const float *w, *bias, *pre_scale, *pre_shift, *post_scale, *post_shift;
    float *out;
    int64_t n;
    int H, W, OH, OW, KH, KW, SH, SW, PT, PL, Hp, Wp, Cout, flags;
    int out_packed;        // 1: store split-half words (lo16 << 16 | hi16) for the fp16-split slab convolution behind this layer
};

constexpr int FIRST_P = 4;

__global__ void __launch_bounds__(256)
conv_first_direct_kernel(const FirstArgs a)
{
    extern __shared__ __align__(16) float fsm[];
    float *ws = fsm;                                   // [K][Cout]
    float *xs = fsm + a.KH * a.KW * a.Cout;            // [Hp][Wp] (+ slack), zero border
    const int tid = threadIdx.x;
    const int CQ = a.Cout >> 2, streams = 256 / CQ;
    const int c4 = tid % CQ, stream = tid / CQ;
    const int K = a.KH * a.KW;
    for (int i = tid; i < K * a.Cout; i += 256) ws[i] = a.w[i];
    const int xs_len = a.Hp * a.Wp + FIRST_P * a.SW + a.KW;
    float eb[4], es1[4], et1[4], es2[4], et2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = c4 * 4 + q;
        eb[q] = (a.flags & ISS_F_BIAS) ? a.bias[c] : 0.f;
        es1[q] = (a.flags & ISS_F_AFFINE_PRE) ? a.pre_scale[c] : 1.f;  et1[q] = (a.flags & ISS_F_AFFINE_PRE) ? a.pre_shift[c] : 0.f;
        es2[q] = (a.flags & ISS_F_AFFINE_POST) ? a.post_scale[c] : 1.f; et2[q] = (a.flags & ISS_F_AFFINE_POST) ? a.post_shift[c] : 0.f;
    }
    const int nblk = (a.OW + FIRST_P - 1) / FIRST_P;
    const int G = a.OH * nblk;
    for (int64_t img = blockIdx.x; img < a.n; img += gridDim.x) {
        __syncthreads();                               // previous patch fully consumed (and ws visible)
        for (int i = tid; i < xs_len; i += 256) xs[i] = 0.f;
        __syncthreads();
        const float mu = a.mu[img], sg = a.sigma[img];
        const float *src = a.mspec + (int64_t)a.row0[img] * a.ld;
        for (int e = tid; e < a.H * a.W; e += 256) {
            const int r = e / a.W, c = e - r * a.W;
            xs[(r + a.PT) * a.Wp + c + a.PL] = __fdiv_rn(__fsub_rn(__ldg(src + (int64_t)r * a.ld + c), mu), sg);
        }
        __syncthreads();
        float *out_img = a.out + img * ((int64_t)a.OH * a.OW * a.Cout);
        for (int g = stream; g < G; g += streams) {
            const int oh = g / nblk, ow0 = (g - oh * nblk) * FIRST_P;
            float acc[FIRST_P][4];
#pragma unroll
            for (int p = 0; p < FIRST_P; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
            for (int r = 0; r < a.KH; ++r) {
                const float *xrow = xs + (oh * a.SH + r) * a.Wp + ow0 * a.SW;
                const float *wrow = ws + (r * a.KW) * a.Cout + c4 * 4;
                for (int t = 0; t < a.KW; ++t) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(wrow + t * a.Cout);
#pragma unroll
                    for (int p = 0; p < FIRST_P; ++p) {
                        const float x = xrow[p * a.SW + t];
                        acc[p][0] = fmaf(x, w4.x, acc[p][0]); acc[p][1] = fmaf(x, w4.y, acc[p][1]);
                        acc[p][2] = fmaf(x, w4.z, acc[p][2]); acc[p][3] = fmaf(x, w4.w, acc[p][3]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < FIRST_P; ++p) {
                if (ow0 + p >= a.OW) continue;
                float y[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v = acc[p][q] + eb[q];
                    if (a.flags & ISS_F_AFFINE_PRE) v = fmaf(v, es1[q], et1[q]);
                    if (a.flags & ISS_F_RELU) v = fmaxf(v, 0.f);
                    if (a.flags & ISS_F_SIGMOID) v = 1.f / (1.f + expf(-v));
                    if (a.flags & ISS_F_AFFINE_POST) v = fmaf(v, es2[q], et2[q]);
                    y[q] = v;
                }
                float *dst = out_img + ((int64_t)oh * a.OW + ow0 + p) * a.Cout + c4 * 4;
                if (a.out_packed)
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(iss_pack_split(y[0]), iss_pack_split(y[1]), iss_pack_split(y[2]), iss_pack_split(y[3]));
                else
                    *reinterpret_cast<float4 *>(dst) = make_float4(y[0], y[1], y[2], y[3]);
            }
        }
    }
}

// ------------------------------------------------------------------ first layer by linearity (FirstFuse, conv_gemm.cuh)
// Y[r][x][c] = sum_{kh,kw} w[kh][kw][c] * mspec[f0 + r + kh][x + kw] in float64, for the frames a batch of patches
// covers, stored as the two-float number Yh + Yl (FirstFuse, conv_gemm.cuh); one thread = one (row, column) position
// and 4 channels, filter in shared memory.
__global__ void __launch_bounds__(256)
first_linear_kernel(const float *__restrict__ mspec, int ld, int64_t f0, int64_t rows, int64_t n_frames, int OW, int KH, int KW, int C,
                    const float *__restrict__ w, float *__restrict__ Yh, float *__restrict__ Yl)
{
    extern __shared__ __align__(16) float wsm[];               // [KH * KW][C]  (float64 weights in shared memory measured slower: 61 vs 47 us)
    for (int i = threadIdx.x; i < KH * KW * C; i += 256) wsm[i] = w[i];
    __syncthreads();
    const int CQ = C >> 2;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;  // over [row][x][c / 4]
    if (i >= rows * OW * CQ) return;
    const int c4 = (int)(i % CQ);
    const int64_t rx = i / CQ;
    const int x = (int)(rx % OW);
    const int64_t r = rx / OW;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int kh = 0; kh < KH; ++kh) {
        const int64_t f = f0 + r + kh;
        if (f >= n_frames) continue;                            // rows past the last frame are never read by a patch
        const float *src = mspec + f * ld + x;
        for (int kw = 0; kw < KW; ++kw) {
            const double xv = (double)__ldg(src + kw);
            const float4 w4 = *reinterpret_cast<const float4 *>(wsm + (kh * KW + kw) * C + c4 * 4);
            acc[0] = fma(xv, (double)w4.x, acc[0]); acc[1] = fma(xv, (double)w4.y, acc[1]);
            acc[2] = fma(xv, (double)w4.z, acc[2]); acc[3] = fma(xv, (double)w4.w, acc[3]);
        }
    }
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { h[q] = (float)acc[q]; l[q] = (float)(acc[q] - (double)h[q]); }
    const int64_t o = (r * OW + x) * (int64_t)C + c4 * 4;
    *reinterpret_cast<float4 *>(Yh + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(Yl + o) = make_float4(l[0], l[1], l[2], l[3]);
}

// Per (patch, channel) coefficients of the fused first layer (FirstFuse): alpha = s_c / sigma_j rounded to float32,
// beta = (b_c s_c + t_c) - mu_j S_c alpha in float64 (from the ROUNDED alpha), stored as two floats.  coef[j][3][C].
__global__ void __launch_bounds__(256)
first_coef_kernel(const float *__restrict__ mu, const float *__restrict__ sigma, const double *__restrict__ S,
                  const float *__restrict__ bias, const float *__restrict__ pre_scale, const float *__restrict__ pre_shift,
                  int64_t n, int C, float *__restrict__ coef)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n * C) return;
    const int c = (int)(i % C);
    const int64_t j = i / C;
    const double s1 = pre_scale ? (double)pre_scale[c] : 1.0, t1 = pre_shift ? (double)pre_shift[c] : 0.0;
    const double b = bias ? (double)bias[c] : 0.0;
    const float alpha = (float)(s1 / (double)sigma[j]);
    const double beta = (b * s1 + t1) - (double)mu[j] * S[c] * (double)alpha;
    const float bh = (float)beta;
    float *dst = coef + j * 3 * C + c;
    dst[0] = alpha; dst[C] = bh; dst[2 * C] = (float)(beta - (double)bh);
}

// ------------------------------------------------------------------ max pooling (NHWC)
__global__ void __launch_bounds__(256)
maxpool_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t total, int H, int W, int C,
                    int OH, int OW, int KH, int KW, int SH, int SW, int PT, int PL)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // over [img][oh][ow][c]
    if (i >= total) return;
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int64_t img = r / OH;
    float best = -INFINITY;
    for (int y = 0; y < KH; ++y) {
        const int ih = oh * SH - PT + y;
        if (ih < 0 || ih >= H) continue;
        for (int x = 0; x < KW; ++x) {
            const int iw = ow * SW - PL + x;
            if (iw < 0 || iw >= W) continue;
            const float v = in[((img * H + ih) * W + iw) * C + c];
            best = (v > best || v != v) ? v : best;        // NaN propagates like TF's max
        }
    }
    out[i] = best;
}

// Four channels per thread (C % 4 == 0): 16-byte loads / stores, same window order and NaN rule as above.
// The pooling layers are pure HBM traffic (218 KB in / 53 KB out per patch for the first one).
__global__ void __launch_bounds__(256)
maxpool_nhwc_vec4_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int64_t total4, int H, int W, int C4,
                         int OH, int OW, int KH, int KW, int SH, int SW, int PT, int PL)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // over [img][oh][ow][c / 4]
    if (i >= total4) return;
    const int c = (int)(i % C4);
    int64_t r = i / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int64_t img = r / OH;
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    auto mx = [](float b, float v) { return (v > b || v != v) ? v : b; };
    for (int y = 0; y < KH; ++y) {
        const int ih = oh * SH - PT + y;
        if (ih < 0 || ih >= H) continue;
        for (int x = 0; x < KW; ++x) {
            const int iw = ow * SW - PL + x;
            if (iw < 0 || iw >= W) continue;
            const float4 v = __ldg(in + ((img * H + ih) * W + iw) * C4 + c);
            best.x = mx(best.x, v.x); best.y = mx(best.y, v.y); best.z = mx(best.z, v.z); best.w = mx(best.w, v.w);
        }
    }
    out[i] = best;
}

// Pooling over split-half words (the tensor between two fp16-split slab convolutions): the window maximum is taken on
// the decoded values hi + lo (exact in fp32) and the winning WORD is stored, so the value itself is unchanged.
__global__ void __launch_bounds__(256)
maxpool_nhwc_packed_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int64_t total4, int H, int W, int C4,
                           int OH, int OW, int KH, int KW, int SH, int SW, int PT, int PL)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // over [img][oh][ow][c / 4]
    if (i >= total4) return;
    const int c = (int)(i % C4);
    int64_t r = i / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int64_t img = r / OH;
    float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    uint32_t bw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bw[q] = iss_pack_split(-INFINITY);
    for (int y = 0; y < KH; ++y) {
        const int ih = oh * SH - PT + y;
        if (ih < 0 || ih >= H) continue;
        for (int x = 0; x < KW; ++x) {
            const int iw = ow * SW - PL + x;
            if (iw < 0 || iw >= W) continue;
            const uint4 v = __ldg(in + ((img * H + ih) * W + iw) * C4 + c);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float f = iss_unpack_split(w4[q]);
                if (f > bv[q] || f != f) { bv[q] = f; bw[q] = w4[q]; }
            }
        }
    }
    out[i] = make_uint4(bw[0], bw[1], bw[2], bw[3]);
}

// ------------------------------------------------------------------ softmax head + non-finite override
__global__ void __launch_bounds__(256)
softmax_head_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ finite, int64_t n, int K,
                    int do_softmax, float *__restrict__ probs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *l = logits + i * K;
    float *p = probs + i * K;
    if (!finite[i]) {                                   // r[finite == False, :] = 0.5 (segmenter.py:175)
        for (int k = 0; k < K; ++k) p[k] = 0.5f;
        return;
    }
    if (!do_softmax) { for (int k = 0; k < K; ++k) p[k] = l[k]; return; }
    float mx = l[0];
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, l[k]);
    float s = 0.f;
    float e[8];
    for (int k = 0; k < K; ++k) { e[k] = expf(l[k] - mx); s += e[k]; }
    for (int k = 0; k < K; ++k) p[k] = e[k] / s;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int iss_cnn_destroy(iss_cnn *cnn);

extern "C" int iss_cnn_create(iss_ctx *ctx, const iss_layer_desc *layers, int n_layers,
                              const float *h_blob, int64_t blob_len, int in_h, int in_w, iss_cnn **out)
{
    ISS_REQUIRE(ctx && layers && h_blob && out && n_layers > 0, ISS_ERR_INVALID, "iss_cnn_create: bad argument");
    ISS_REQUIRE(in_h == PATCH_H, ISS_ERR_UNSUPPORTED, "iss_cnn_create: patch height must be %d (segmenter.py:149)", PATCH_H);
    ISS_REQUIRE(in_w >= 1 && in_w <= ISS_NMEL, ISS_ERR_INVALID, "iss_cnn_create: in_w=%d", in_w);
    ISS_CUDA_OK(cudaSetDevice(ctx->device));
    iss_cnn *m = new iss_cnn();
    m->ctx = ctx; m->d_blob = nullptr; m->blob_len = blob_len; m->in_h = in_h; m->in_w = in_w;
    int h = in_h, w = in_w, c = 1;
    int64_t max_act = 0;
    double flops = 0;
    auto off_ok = [&](int64_t off, int64_t len) { return off >= 0 && off + len <= blob_len; };
    for (int i = 0; i < n_layers; ++i) {
        Layer L; L.d = layers[i]; L.in_h = h; L.in_w = w; L.in_c = c;
        const iss_layer_desc &d = L.d;
        bool ok = true;
        if (d.kind == ISS_LAYER_CONV2D || d.kind == ISS_LAYER_MAXPOOL) {
            ok = d.kh >= 1 && d.kw >= 1 && d.sh >= 1 && d.sw >= 1 && d.pad_top >= 0 && d.pad_left >= 0 &&
                 d.pad_bottom >= 0 && d.pad_right >= 0;
            if (ok) {
                L.out_h = (h + d.pad_top + d.pad_bottom - d.kh) / d.sh + 1;
                L.out_w = (w + d.pad_left + d.pad_right - d.kw) / d.sw + 1;
                ok = L.out_h >= 1 && L.out_w >= 1;
            }
            if (d.kind == ISS_LAYER_CONV2D) {
                ok = ok && d.cin == c && d.cout >= 1 && off_ok(d.w_off, (int64_t)d.kh * d.kw * d.cin * d.cout);
                L.out_c = d.cout;
                flops += 2.0 * L.out_h * L.out_w * (double)d.kh * d.kw * d.cin * d.cout;
            } else {
                L.out_c = c;
            }
        } else if (d.kind == ISS_LAYER_DENSE) {
            ok = d.cin == h * w * c && d.cout >= 1 && off_ok(d.w_off, (int64_t)d.cin * d.cout);
            L.out_h = 1; L.out_w = 1; L.out_c = d.cout;
            flops += 2.0 * (double)d.cin * d.cout;
        } else {
            ok = false;
        }
        if (ok && d.kind != ISS_LAYER_MAXPOOL) {
            if (d.flags & ISS_F_BIAS) ok = ok && off_ok(d.bias_off, d.cout);
            if (d.flags & ISS_F_AFFINE_PRE) ok = ok && off_ok(d.pre_scale_off, d.cout) && off_ok(d.pre_shift_off, d.cout);
            if (d.flags & ISS_F_AFFINE_POST) ok = ok && off_ok(d.post_scale_off, d.cout) && off_ok(d.post_shift_off, d.cout);
            if (d.flags & ISS_F_SOFTMAX) ok = ok && (i == n_layers - 1) && d.cout <= 8;
        }
        if (!ok) {
            delete m;
            iss_set_error("iss_cnn_create: layer %d (kind %d) inconsistent with input %dx%dx%d or blob", i, d.kind, h, w, c);
            return ISS_ERR_UNSUPPORTED;
        }
        h = L.out_h; w = L.out_w; c = L.out_c;
        max_act = std::max<int64_t>(max_act, (int64_t)h * w * c);
        m->layers.push_back(L);
    }
    if (!(h == 1 && w == 1 && c <= 8)) {
        delete m;
        iss_set_error("iss_cnn_create: network must end in a Dense head with <= 8 classes (got %dx%dx%d)", h, w, c);
        return ISS_ERR_UNSUPPORTED;
    }
    m->n_classes = c; m->max_act = max_act; m->flops = flops;
    cudaError_t e = cudaMalloc(&m->d_blob, (size_t)blob_len * sizeof(float));
    if (e != cudaSuccess) { delete m; iss_set_error("cudaMalloc blob: %s", cudaGetErrorString(e)); return ISS_ERR_NOMEM; }
    e = cudaMemcpy(m->d_blob, h_blob, (size_t)blob_len * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(m->d_blob); delete m; iss_set_error("cudaMemcpy blob: %s", cudaGetErrorString(e)); return ISS_ERR_CUDA; }
    for (size_t i = 1; i < m->layers.size(); ++i) {          // layer 0 gathers from the log-mel rows (fp32 kernel)
        Layer &L = m->layers[i];
        const iss_layer_desc &d = L.d;
        if (d.kind == ISS_LAYER_MAXPOOL) continue;
        const int K = (d.kind == ISS_LAYER_DENSE) ? d.cin : d.kh * d.kw * d.cin;
        const int C = (d.kind == ISS_LAYER_DENSE) ? d.cin : d.cin;
        if (C % 32 != 0 || K % 32 != 0 || d.cout % 32 != 0) continue;
        int rc = iss_prepare_tc_weights(h_blob + d.w_off, K, d.cout, &L.d_wt, &L.Kp);
        if (rc != ISS_OK) { iss_cnn_destroy(m); return rc; }
        if (d.cin % 64 == 0) {                                    // fp16-split engine (slab or gather kernel)
            rc = iss_prepare_f16_weights(h_blob + d.w_off, K, d.cout, &L.d_wt_f16, &L.f16_inv_scale);
            if (rc != ISS_OK) { iss_cnn_destroy(m); return rc; }
        }
    }
    {   // S_c = sum of the first layer's taps (float64), for the fused first layer
        const iss_layer_desc &d0 = m->layers[0].d;
        if (d0.kind == ISS_LAYER_CONV2D && d0.cin == 1) {
            std::vector<double> S(d0.cout, 0.0);
            for (int t = 0; t < d0.kh * d0.kw; ++t)
                for (int c = 0; c < d0.cout; ++c) S[c] += (double)h_blob[d0.w_off + (int64_t)t * d0.cout + c];
            e = cudaMalloc(&m->d_first_S, sizeof(double) * d0.cout);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_first_S, S.data(), sizeof(double) * d0.cout, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { iss_cnn_destroy(m); iss_set_error("first-layer sums: %s", cudaGetErrorString(e)); return ISS_ERR_CUDA; }
        }
    }
    *out = m;
    return ISS_OK;
}

extern "C" int iss_cnn_destroy(iss_cnn *cnn)
{
    if (!cnn) return ISS_OK;
    cudaSetDevice(cnn->ctx->device);
    for (cudaEvent_t ev : cnn->prof_ev) cudaEventDestroy(ev);
    for (Layer &L : cnn->layers) { if (L.d_wt) cudaFree(L.d_wt); if (L.d_wt_f16) cudaFree(L.d_wt_f16); }
    if (cnn->d_blob) cudaFree(cnn->d_blob);
    if (cnn->d_first_S) cudaFree(cnn->d_first_S);
    delete cnn;
    return ISS_OK;
}

extern "C" int iss_cnn_num_classes(const iss_cnn *cnn) { return cnn ? cnn->n_classes : -1; }
extern "C" double iss_cnn_flops_per_patch(const iss_cnn *cnn) { return cnn ? cnn->flops : 0.0; }

namespace {
// rows of the two-float first-layer map a batch of B patches may need (3x the contiguous case; batches whose
// patches are spread wider fall back to the un-fused first layer) and its size in bytes (0: no fusion possible)
int64_t first_y_rows(int64_t B) { return 3 * (PATCH_HOP * B + PATCH_H); }
size_t first_y_bytes(const iss_cnn *cnn, int64_t B)
{
    if (!cnn->d_first_S || cnn->layers.size() < 2) return 0;
    const Layer &L0 = cnn->layers[0];
    return 2 * align_up((size_t)first_y_rows(B) * L0.out_w * L0.out_c * sizeof(float), 256) +        // Yh, Yl
           align_up((size_t)B * 3 * L0.out_c * sizeof(float), 256);                                   // per-patch coefficients
}
}  // namespace

extern "C" int64_t iss_cnn_workspace_bytes(const iss_cnn *cnn, int64_t n, int n_seg)
{
    if (!cnn || n < 0 || n_seg < 0) return -1;
    const int64_t B = std::min<int64_t>(std::max<int64_t>(n, 1), CNN_BATCH);
    int64_t bytes = 0;
    bytes += align_up((size_t)n * 4, 256) * 3;            // row0, mu, sigma
    bytes += align_up((size_t)n, 256);                     // finite
    bytes += align_up((size_t)n * 8 * 4, 256);             // logits [n][<=8]
    bytes += 2 * align_up((size_t)B * cnn->max_act * 4, 256);
    bytes += align_up((size_t)(n_seg + 1) * 4, 256) + align_up((size_t)(n_seg + 1) * 8, 256);   // segment tables
    bytes += first_y_bytes(cnn, B);                                                               // float64 map of the fused first layer
    return bytes;
}

namespace {

// geometry + weight pointers of a Conv2D / Dense layer for `nb` patches (the fields iss_launch_conv dispatches on)
void fill_conv_args(const Layer &Lr, int64_t nb, ConvArgs &a)
{
    const iss_layer_desc &d = Lr.d;
    if (Lr.d_wt) { a.wt_hi = Lr.d_wt; a.wt_lo = Lr.d_wt + (size_t)d.cout * Lr.Kp; a.wt_tiled = Lr.d_wt + 2 * (size_t)d.cout * Lr.Kp; a.Kp = Lr.Kp; }
    a.wt_f16 = Lr.d_wt_f16; a.wt_f16_inv_scale = Lr.f16_inv_scale;
    a.N = d.cout;
    if (d.kind == ISS_LAYER_DENSE) {
        a.M = nb; a.K = d.cin; a.H = 1; a.W = 1; a.C = d.cin; a.OH = 1; a.OW = 1;
        a.KH = 1; a.KW = 1; a.SH = 1; a.SW = 1; a.PT = 0; a.PL = 0;
    } else {
        a.M = nb * Lr.out_h * Lr.out_w; a.K = d.kh * d.kw * d.cin;
        a.H = Lr.in_h; a.W = Lr.in_w; a.C = Lr.in_c; a.OH = Lr.out_h; a.OW = Lr.out_w;
        a.KH = d.kh; a.KW = d.kw; a.SH = d.sh; a.SW = d.sw; a.PT = d.pad_top; a.PL = d.pad_left;
    }
}

// Will this Conv2D / Dense layer run on the fp16-split engine (slab kernel for un-padded stride-1 KHxKW convolutions,
// gather kernel for everything else it has a weight image for)?  Only those kernels read / write split-half words.
bool f16_engine_covers(const Layer &L, const ConvArgs &a)
{
    if (L.d.kind == ISS_LAYER_MAXPOOL) return false;
    if (L.d.kind == ISS_LAYER_CONV2D && L.d.pad_bottom == 0 && L.d.pad_right == 0 && iss_conv_f16_slab_covers(a)) return true;
    // the gather kernel takes symmetric padding only through PT / PL + the bounds check: 'same' padding with an extra
    // bottom / right row is covered too (rows past the input are zero-filled)
    return iss_conv_f16_gather_covers(a);
}

// Should the tensor produced by layer `li` be stored as split-half words?  Yes iff the next compute layer
// (pooling layers in between keep the format) is a convolution the fp16-split slab kernel takes.
bool wants_packed_output(const iss_cnn *cnn, size_t li, int64_t nb)
{
    if (iss_get_gemm_mode() != ISS_GEMM_TC_F16) return false;
    size_t j = li + 1;
    while (j < cnn->layers.size() && cnn->layers[j].d.kind == ISS_LAYER_MAXPOOL) {
        if (cnn->layers[j].in_c % 4 != 0) return false;          // the packed pooling kernel moves 4 channels per thread
        ++j;
    }
    if (j >= cnn->layers.size()) return false;
    const Layer &Nx = cnn->layers[j];
    ConvArgs probe = {};
    fill_conv_args(Nx, nb, probe);
    return f16_engine_covers(Nx, probe);
}

}  // namespace

extern "C" int iss_cnn_forward(iss_ctx *ctx, iss_cnn *cnn, const float *d_mspec, int64_t L, int ld,
                               int edge_left, int edge_right,
                               const int32_t *h_seg_start, const int32_t *h_seg_stop, int n_seg,
                               float *d_probs, void *d_work, int64_t work_bytes, void *stream)
{
    ISS_REQUIRE(ctx && cnn, ISS_ERR_INVALID, "iss_cnn_forward: NULL handle");
    ISS_REQUIRE(ld >= cnn->in_w, ISS_ERR_INVALID, "iss_cnn_forward: ld=%d < nmel=%d", ld, cnn->in_w);
    if (n_seg <= 0) return ISS_OK;
    ISS_REQUIRE(h_seg_start && h_seg_stop, ISS_ERR_INVALID, "iss_cnn_forward: NULL segment arrays");
    ISS_REQUIRE(L >= PATCH_H, ISS_ERR_INVALID, "iss_cnn_forward: L=%lld < %d frames (pad short inputs first, segmenter.py:60-65)", (long long)L, PATCH_H);
    ISS_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = iss_stream(stream);
    const int64_t U = (L - PATCH_H) / PATCH_HOP + 1;                      // un-replicated windows
    const int64_t P = (edge_left ? PATCH_LFILL : 0) + U + (edge_right ? (PATCH_LFILL - 1 + (L % 2)) : 0);
    std::vector<int64_t> off(n_seg + 1);
    int64_t n = 0;
    for (int s = 0; s < n_seg; ++s) {
        ISS_REQUIRE(h_seg_start[s] >= 0 && h_seg_stop[s] >= h_seg_start[s] && h_seg_stop[s] <= P, ISS_ERR_INVALID,
                    "iss_cnn_forward: range %d = [%d,%d) outside [0,%lld)", s, h_seg_start[s], h_seg_stop[s], (long long)P);
        off[s] = n; n += h_seg_stop[s] - h_seg_start[s];
    }
    off[n_seg] = n;
    if (n == 0) return ISS_OK;
    ISS_REQUIRE(d_mspec && d_probs && d_work, ISS_ERR_INVALID, "iss_cnn_forward: NULL buffer");
    ISS_REQUIRE(work_bytes >= iss_cnn_workspace_bytes(cnn, n, n_seg), ISS_ERR_INVALID,
                "iss_cnn_forward: workspace %lld < required %lld", (long long)work_bytes, (long long)iss_cnn_workspace_bytes(cnn, n, n_seg));

    // ---- carve the workspace ----
    const int64_t B = std::min<int64_t>(n, CNN_BATCH);
    uint8_t *p = reinterpret_cast<uint8_t *>(d_work);
    size_t o = 0;
    PatchArrays pa;
    pa.row0 = reinterpret_cast<int32_t *>(p + o); o += align_up((size_t)n * 4, 256);
    pa.mu = reinterpret_cast<float *>(p + o);     o += align_up((size_t)n * 4, 256);
    pa.sigma = reinterpret_cast<float *>(p + o);  o += align_up((size_t)n * 4, 256);
    pa.finite = p + o;                            o += align_up((size_t)n, 256);
    float *logits = reinterpret_cast<float *>(p + o); o += align_up((size_t)n * 8 * 4, 256);
    float *act[2];
    act[0] = reinterpret_cast<float *>(p + o);    o += align_up((size_t)B * cnn->max_act * 4, 256);
    act[1] = reinterpret_cast<float *>(p + o);    o += align_up((size_t)B * cnn->max_act * 4, 256);
    int32_t *d_seg_start = reinterpret_cast<int32_t *>(p + o); o += align_up((size_t)(n_seg + 1) * 4, 256);
    int64_t *d_seg_off = reinterpret_cast<int64_t *>(p + o);   o += align_up((size_t)(n_seg + 1) * 8, 256);
    float *d_first_yh = nullptr, *d_first_yl = nullptr, *d_first_coef = nullptr;                // fused first layer (FirstFuse)
    if (first_y_bytes(cnn, B)) {
        const Layer &L0 = cnn->layers[0];
        const size_t plane = align_up((size_t)first_y_rows(B) * L0.out_w * L0.out_c * sizeof(float), 256);
        d_first_yh = reinterpret_cast<float *>(p + o); d_first_yl = reinterpret_cast<float *>(p + o + plane);
        d_first_coef = reinterpret_cast<float *>(p + o + 2 * plane);
        o += first_y_bytes(cnn, B);
    }
    // host copy of patch_index_kernel's arithmetic: first log-mel frame of patch i (for the span of a batch)
    auto host_row0 = [&](int64_t i) -> int64_t {
        int lo = 0, hi = n_seg;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
        int64_t j = h_seg_start[lo] + (i - off[lo]) - (edge_left ? PATCH_LFILL : 0);
        if (j < 0) j = 0;
        if (edge_right && j > U - 1) j = U - 1;
        return j * PATCH_HOP;
    };
    bool ranges_ascending = true;
    for (int s = 1; s < n_seg; ++s) ranges_ascending = ranges_ascending && h_seg_start[s] >= h_seg_stop[s - 1];
    ISS_CUDA_OK(cudaMemcpyAsync(d_seg_start, h_seg_start, sizeof(int32_t) * n_seg, cudaMemcpyHostToDevice, st));
    ISS_CUDA_OK(cudaMemcpyAsync(d_seg_off, off.data(), sizeof(int64_t) * (n_seg + 1), cudaMemcpyHostToDevice, st));

    patch_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_seg_start, d_seg_off, n_seg, n, U, edge_left, edge_right, pa.row0);
    patch_stats_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(d_mspec, ld, cnn->in_w, n, pa);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch(2);

    const float *blob = cnn->d_blob;
    const int K = cnn->n_classes;
    for (int64_t b0 = 0; b0 < n; b0 += B) {
        const int64_t nb = std::min<int64_t>(B, n - b0);
        const float *cur = nullptr;
        int which = 0;
        bool cur_packed = false;                                 // format of `cur`: fp32 values or split-half words
        FirstFuse ffuse = {};                                    // set by layer 0 when it is folded into layer 1's slab fill
        bool first_fused = false;
        int pend_pool_h = 0, pend_pool_w = 0;                    // un-pooled dims of a MaxPooling2D folded into the next convolution
        for (size_t li = 0; li < cnn->layers.size(); ++li) {
            const Layer &Lr = cnn->layers[li];
            const iss_layer_desc &d = Lr.d;
            const bool last = (li + 1 == cnn->layers.size());
            float *dst = last ? (logits + b0 * K) : act[which];
            const bool prof = ((int)li == cnn->prof_layer);
            if (prof) {
                if (cnn->prof_used + 2 > cnn->prof_ev.size()) {
                    for (int q = 0; q < 2; ++q) {
                        cudaEvent_t ev;
                        ISS_CUDA_OK(cudaEventCreate(&ev));
                        cnn->prof_ev.push_back(ev);
                    }
                }
                ISS_CUDA_OK(cudaEventRecord(cnn->prof_ev[cnn->prof_used], st));
            }
            if (d.kind == ISS_LAYER_MAXPOOL && cur_packed && li > 0 && li + 1 < cnn->layers.size()) {
                // 2x2 / stride-2 'valid' pooling in front of a convolution the direct kernel takes: the maximum is taken in that
                // kernel's slab fill (conv_gemm_tc_f16d.cu, IN_POOL) and this layer never runs.  ISS_B200_FUSE_POOL=0: A/B tests.
                const char *pool_env = getenv("ISS_B200_FUSE_POOL");
                const bool pool_off = pool_env && pool_env[0] == '0';
                const Layer &Nx = cnn->layers[li + 1];
                if (!pool_off && !prof && d.kh == 2 && d.kw == 2 && d.sh == 2 && d.sw == 2 && d.pad_top == 0 && d.pad_left == 0 &&
                    d.pad_bottom == 0 && d.pad_right == 0 && Nx.d.kind == ISS_LAYER_CONV2D && Nx.d.pad_bottom == 0 && Nx.d.pad_right == 0 &&
                    iss_get_gemm_mode() == ISS_GEMM_TC_F16) {
                    ConvArgs pn = {};
                    fill_conv_args(Nx, nb, pn);
                    pn.in_packed = 1;
                    pn.flags = Nx.d.flags & ~ISS_F_SOFTMAX;
                    pn.pool_h = Lr.in_h; pn.pool_w = Lr.in_w;
                    if (iss_conv_f16_direct_covers(pn)) { pend_pool_h = Lr.in_h; pend_pool_w = Lr.in_w; continue; }
                }
            }
            if (d.kind == ISS_LAYER_MAXPOOL) {
                ISS_REQUIRE(li > 0, ISS_ERR_UNSUPPORTED, "iss_cnn_forward: pooling as first layer is not supported");
                const int64_t total = nb * Lr.out_h * Lr.out_w * Lr.out_c;
                if (cur_packed)                                   // (format kept: the consumer behind the pooling asked for words)
                    maxpool_nhwc_packed_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, st>>>(
                        reinterpret_cast<const uint4 *>(cur), reinterpret_cast<uint4 *>(dst), total / 4, Lr.in_h, Lr.in_w, Lr.in_c / 4,
                        Lr.out_h, Lr.out_w, d.kh, d.kw, d.sh, d.sw, d.pad_top, d.pad_left);
                else if (Lr.in_c % 4 == 0)
                    maxpool_nhwc_vec4_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, st>>>(
                        reinterpret_cast<const float4 *>(cur), reinterpret_cast<float4 *>(dst), total / 4, Lr.in_h, Lr.in_w, Lr.in_c / 4,
                        Lr.out_h, Lr.out_w, d.kh, d.kw, d.sh, d.sw, d.pad_top, d.pad_left);
                else
                    maxpool_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cur, dst, total, Lr.in_h, Lr.in_w, Lr.in_c,
                        Lr.out_h, Lr.out_w, d.kh, d.kw, d.sh, d.sw, d.pad_top, d.pad_left);
                ISS_CUDA_OK(cudaGetLastError());
                iss_count_launch();
            } else {
                ConvArgs a = {};
                a.w = blob + d.w_off;
                a.bias = (d.flags & ISS_F_BIAS) ? blob + d.bias_off : nullptr;
                a.pre_scale = (d.flags & ISS_F_AFFINE_PRE) ? blob + d.pre_scale_off : nullptr;
                a.pre_shift = (d.flags & ISS_F_AFFINE_PRE) ? blob + d.pre_shift_off : nullptr;
                a.post_scale = (d.flags & ISS_F_AFFINE_POST) ? blob + d.post_scale_off : nullptr;
                a.post_shift = (d.flags & ISS_F_AFFINE_POST) ? blob + d.post_shift_off : nullptr;
                a.out = dst;
                a.flags = d.flags & ~ISS_F_SOFTMAX;
                fill_conv_args(Lr, nb, a);
                a.in_packed = cur_packed ? 1 : 0;
                const bool direct = li == 0 && d.kind == ISS_LAYER_CONV2D && d.cin == 1 &&
                                    (d.cout == 16 || d.cout == 32 || d.cout == 64 || d.cout == 128) && !(d.flags & ISS_F_SOFTMAX);
                // only the direct first-layer kernel and the fp16-split slab kernel can emit split-half words
                const bool can_pack = direct || (li > 0 && iss_get_gemm_mode() == ISS_GEMM_TC_F16 && f16_engine_covers(Lr, a));
                a.out_packed = (!last && can_pack && wants_packed_output(cnn, li, nb)) ? 1 : 0;
                ISS_REQUIRE(!a.in_packed || (li > 0 && f16_engine_covers(Lr, a)), ISS_ERR_UNSUPPORTED,
                            "iss_cnn_forward: layer %d was handed split-half words it cannot read", (int)li);
                int rc;
                // Fold the first layer into the next convolution's slab fill (FirstFuse, conv_gemm.cuh)?  Needs: the direct
                // first-layer shape without padding / stride, a packed-input slab convolution right behind it, and a batch
                // whose patches span few enough frames for the float64 map.
                const char *fuse_env = getenv("ISS_B200_FUSE_FIRST");      // "0" = keep the stand-alone first-layer kernel (A/B tests)
                const bool fuse_off = fuse_env && fuse_env[0] == '0';
                bool next_is_direct = false;                     // the FIRST mode lives in the direct kernel only
                if (li == 0 && cnn->layers.size() > 1 && cnn->layers[1].d.kind == ISS_LAYER_CONV2D &&
                    cnn->layers[1].d.pad_bottom == 0 && cnn->layers[1].d.pad_right == 0 && iss_get_gemm_mode() == ISS_GEMM_TC_F16) {
                    ConvArgs p1 = {};
                    fill_conv_args(cnn->layers[1], nb, p1);
                    p1.in_packed = 1;
                    p1.flags = cnn->layers[1].d.flags & ~ISS_F_SOFTMAX;
                    next_is_direct = iss_conv_f16_direct_covers(p1);
                }
                if (li == 0 && direct && a.out_packed && !fuse_off && d_first_yh && ranges_ascending && next_is_direct && d.sh == 1 && d.sw == 1 && d.pad_top == 0 && d.pad_left == 0 &&
                    d.pad_bottom == 0 && d.pad_right == 0 && !(d.flags & (ISS_F_SOFTMAX | ISS_F_SIGMOID))) {
                    const int64_t f_first = host_row0(b0), f_last = host_row0(b0 + nb - 1);
                    const int64_t rows = f_last - f_first + Lr.out_h;
                    if (rows <= first_y_rows(B)) {
                        const int64_t work = rows * Lr.out_w * (d.cout >> 2);
                        first_linear_kernel<<<(unsigned)((work + 255) / 256), 256, (size_t)d.kh * d.kw * d.cout * sizeof(float), st>>>(
                            d_mspec, ld, f_first, rows, L, Lr.out_w, d.kh, d.kw, d.cout, a.w, d_first_yh, d_first_yl);
                        first_coef_kernel<<<(unsigned)((nb * d.cout + 255) / 256), 256, 0, st>>>(
                            pa.mu + b0, pa.sigma + b0, cnn->d_first_S, a.bias, a.pre_scale, a.pre_shift, nb, d.cout, d_first_coef);
                        ISS_CUDA_OK(cudaGetLastError());
                        iss_count_launch(2);
                        ffuse.Yh = d_first_yh; ffuse.Yl = d_first_yl; ffuse.y_f0 = f_first; ffuse.y_rows = rows;
                        ffuse.row0 = pa.row0 + b0; ffuse.coef = d_first_coef;
                        ffuse.post_scale = a.post_scale; ffuse.post_shift = a.post_shift; ffuse.flags = a.flags; ffuse.n_img = nb;
                        first_fused = true;
                    }
                }
                if (li == 0 && first_fused) {
                    rc = ISS_OK;                                  // nothing stored: layer 1 reads the map
                } else if (li == 0) {
                    ISS_REQUIRE(d.kind == ISS_LAYER_CONV2D, ISS_ERR_UNSUPPORTED, "iss_cnn_forward: first layer must be Conv2D");
                    if (direct) {
                        FirstArgs f = {};
                        f.mspec = d_mspec; f.ld = ld; f.row0 = pa.row0 + b0; f.mu = pa.mu + b0; f.sigma = pa.sigma + b0;
                        f.w = a.w; f.bias = a.bias; f.pre_scale = a.pre_scale; f.pre_shift = a.pre_shift;
                        f.post_scale = a.post_scale; f.post_shift = a.post_shift; f.out = dst; f.n = nb;
                        f.H = Lr.in_h; f.W = Lr.in_w; f.OH = Lr.out_h; f.OW = Lr.out_w; f.KH = d.kh; f.KW = d.kw;
                        f.SH = d.sh; f.SW = d.sw; f.PT = d.pad_top; f.PL = d.pad_left;
                        f.Hp = Lr.in_h + d.pad_top + d.pad_bottom; f.Wp = Lr.in_w + d.pad_left + d.pad_right;
                        f.Cout = d.cout; f.flags = a.flags; f.out_packed = a.out_packed;
                        const size_t smem = ((size_t)d.kh * d.kw * d.cout + (size_t)f.Hp * f.Wp + FIRST_P * d.sw + d.kw + 8) * sizeof(float);
                        ISS_REQUIRE(smem <= 200 * 1024, ISS_ERR_UNSUPPORTED, "iss_cnn_forward: first layer too large for the direct kernel");
                        ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(conv_first_direct_kernel), 200 * 1024));
                        const unsigned grid = (unsigned)std::min<int64_t>(nb, (int64_t)ctx->sm_count * 8);
                        conv_first_direct_kernel<<<grid, 256, smem, st>>>(f);
                        ISS_CUDA_OK(cudaGetLastError());
                        iss_count_launch();
                        rc = ISS_OK;
                    } else {
                        a.in = d_mspec; a.ld = ld; a.row0 = pa.row0 + b0; a.mu = pa.mu + b0; a.sigma = pa.sigma + b0;
                        a.out_packed = 0;                         // the generic first-layer kernel writes fp32
                        rc = iss_launch_conv(a, true, st);
                    }
                } else {
                    a.in = cur;
                    a.pool_h = pend_pool_h; a.pool_w = pend_pool_w;
                    pend_pool_h = pend_pool_w = 0;
                    if (li == 1 && first_fused) { a.first = &ffuse; a.in = nullptr; a.in_packed = 1; }
                    rc = iss_launch_conv(a, false, st);
                }
                if (rc != ISS_OK) return rc;
                cur_packed = a.out_packed != 0;
            }
            if (prof) {
                ISS_CUDA_OK(cudaEventRecord(cnn->prof_ev[cnn->prof_used + 1], st));
                cnn->prof_used += 2;
                if (d.kind == ISS_LAYER_CONV2D) cnn->prof_flops += 2.0 * nb * Lr.out_h * Lr.out_w * (double)d.kh * d.kw * d.cin * d.cout;
                else if (d.kind == ISS_LAYER_DENSE) cnn->prof_flops += 2.0 * nb * (double)d.cin * d.cout;
            }
            cur = dst;
            which ^= 1;
        }
    }
    const Layer &head = cnn->layers.back();
    softmax_head_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, pa.finite, n, K,
        (head.d.flags & ISS_F_SOFTMAX) ? 1 : 0, d_probs);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

extern "C" int iss_cnn_profile(iss_cnn *cnn, int layer)
{
    ISS_REQUIRE(cnn, ISS_ERR_INVALID, "iss_cnn_profile: NULL");
    ISS_REQUIRE(layer >= -1 && layer < (int)cnn->layers.size(), ISS_ERR_INVALID, "iss_cnn_profile: layer %d", layer);
    cnn->prof_layer = layer;
    cnn->prof_used = 0;
    cnn->prof_flops = 0;
    return ISS_OK;
}

extern "C" int iss_cnn_profile_read(iss_cnn *cnn, double *total_ms, int64_t *launches, double *flops)
{
    ISS_REQUIRE(cnn && total_ms && launches && flops, ISS_ERR_INVALID, "iss_cnn_profile_read: NULL");
    ISS_CUDA_OK(cudaSetDevice(cnn->ctx->device));
    double t = 0;
    for (size_t i = 0; i + 1 < cnn->prof_used; i += 2) {
        ISS_CUDA_OK(cudaEventSynchronize(cnn->prof_ev[i + 1]));
        float ms = 0;
        ISS_CUDA_OK(cudaEventElapsedTime(&ms, cnn->prof_ev[i], cnn->prof_ev[i + 1]));
        t += ms;
    }
    *total_ms = t; *launches = (int64_t)(cnn->prof_used / 2); *flops = cnn->prof_flops;
    cnn->prof_used = 0; cnn->prof_flops = 0;
    return ISS_OK;
}

extern "C" int iss_cnn_num_layers(const iss_cnn *cnn) { return cnn ? (int)cnn->layers.size() : -1; }

extern "C" double iss_cnn_layer_flops(const iss_cnn *cnn, int layer)
{
    if (!cnn || layer < 0 || layer >= (int)cnn->layers.size()) return 0.0;
    const Layer &L = cnn->layers[layer];
    if (L.d.kind == ISS_LAYER_CONV2D) return 2.0 * L.out_h * L.out_w * (double)L.d.kh * L.d.kw * L.d.cin * L.d.cout;
    if (L.d.kind == ISS_LAYER_DENSE) return 2.0 * (double)L.d.cin * L.d.cout;
    return 0.0;
}
