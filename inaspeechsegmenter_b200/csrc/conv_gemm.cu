// conv_gemm.cu -- implicit-GEMM conv / dense on fp32 CUDA cores (see conv_gemm.cuh).
// 128 x {64,128} x 16 tiles, 8 x {4,8} register tiles, double-buffered shared memory,
// fused epilogue: bias -> BatchNorm affine -> (+residual) -> ReLU -> BatchNorm affine.
#include "conv_gemm.cuh"

namespace {

// ------------------------------------------------------------------ implicit-GEMM conv / dense (fp32 CUDA cores)
constexpr int BM = 128, BK = 16;

template <int BN, bool FIRST>
__global__ void __launch_bounds__(256, 2)
conv_gemm_f32_kernel(const ConvArgs a)
{
    constexpr int TM = 8, TN = BN / 16;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- per-thread A-gather coordinates: one output position, 8 consecutive k per tile ----
    const int ml = tid & (BM - 1), kh2 = tid >> 7;
    const int64_t m = m0 + ml;
    const bool m_ok = m < a.M;
    int ih0 = 0, iw0 = 0;
    const float *in_img = a.in;
    float mu = 0.f, sg = 1.f;
    {
        const int64_t mm = m_ok ? m : 0;
        const int ohw = a.OH * a.OW;
        const int64_t img = mm / ohw;
        const int rem = (int)(mm - img * ohw);
        const int oh = rem / a.OW, ow = rem - oh * a.OW;
        ih0 = oh * a.SH - a.PT; iw0 = ow * a.SW - a.PL;
        if (FIRST) { in_img = a.in + (int64_t)a.row0[img] * a.ld; mu = a.mu[img]; sg = a.sigma[img]; }
        else in_img = a.in + img * ((int64_t)a.H * a.W * a.C);
    }
    const bool vecA = !FIRST && (a.C % 8 == 0);
    const bool vecB = (a.N % 4 == 0);

    float ra[8];
    float rb[TN];

    auto load_tiles = [&](int kt) {
        const int kb = kt * BK + kh2 * 8;
        if (vecA) {
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (m_ok && kb < a.K) {
                const int tap = kb / a.C, c = kb - tap * a.C;
                const int r = tap / a.KW, s = tap - r * a.KW;
                const int ih = ih0 + r, iw = iw0 + s;
                if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) {
                    const float4 *p = reinterpret_cast<const float4 *>(in_img + ((int64_t)ih * a.W + iw) * a.C + c);
                    v0 = __ldg(p); v1 = __ldg(p + 1);
                }
            }
            ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
            ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = kb + q;
                float v = 0.f;
                if (m_ok && k < a.K) {
                    const int tap = k / a.C, c = k - tap * a.C;
                    const int r = tap / a.KW, s = tap - r * a.KW;
                    const int ih = ih0 + r, iw = iw0 + s;
                    if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) {
                        if (FIRST) v = __fdiv_rn(__fsub_rn(__ldg(in_img + (int64_t)ih * a.ld + iw), mu), sg);
                        else v = __ldg(in_img + ((int64_t)ih * a.W + iw) * a.C + c);
                    }
                }
                ra[q] = v;
            }
        }
        // B tile: BK x BN floats, 256 threads * TN
        {
            const int e = tid * TN;                  // element index in the tile
            const int kk = e / BN, nn = e - kk * BN;
            const int k = kt * BK + kk, n = n0 + nn;
            if (vecB && k < a.K && n + TN <= a.N) {
#pragma unroll
                for (int q = 0; q < TN; q += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(a.w + (int64_t)k * a.N + n + q));
                    rb[q] = v.x; rb[q + 1] = v.y; rb[q + 2] = v.z; rb[q + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < TN; ++q) rb[q] = (k < a.K && n + q < a.N) ? __ldg(a.w + (int64_t)k * a.N + n + q) : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) As[buf][kh2 * 8 + q][ml] = ra[q];
        const int e = tid * TN;
        const int kk = e / BN, nn = e - kk * BN;
#pragma unroll
        for (int q = 0; q < TN; ++q) Bs[buf][kk][nn + q] = rb[q];
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nkt = (a.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_tiles(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * TM + 4]);
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
            for (int q = 0; q < TN; q += 4) {
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * TN + q]);
                bv[q] = b.x; bv[q + 1] = b.y; bv[q + 2] = b.z; bv[q + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nkt) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias -> affine(pre) -> relu -> affine(post) ----
    float eb[TN], es1[TN], et1[TN], es2[TN], et2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        const bool ok = n < a.N;
        eb[j] = (ok && (a.flags & ISS_F_BIAS)) ? a.bias[n] : 0.f;
        es1[j] = (ok && (a.flags & ISS_F_AFFINE_PRE)) ? a.pre_scale[n] : 1.f;
        et1[j] = (ok && (a.flags & ISS_F_AFFINE_PRE)) ? a.pre_shift[n] : 0.f;
        es2[j] = (ok && (a.flags & ISS_F_AFFINE_POST)) ? a.post_scale[n] : 1.f;
        et2[j] = (ok && (a.flags & ISS_F_AFFINE_POST)) ? a.post_shift[n] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t mr = m0 + ty * TM + i;
        if (mr >= a.M) continue;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float x = acc[i][j] + eb[j];
            if (a.flags & ISS_F_AFFINE_PRE) x = fmaf(x, es1[j], et1[j]);
            if (a.flags & ISS_F_RESIDUAL) { const int n = n0 + tx * TN + j; if (n < a.N) x += __ldg(a.residual + mr * a.N + n); }
            if (a.flags & ISS_F_RELU) x = fmaxf(x, 0.f);
            if (a.flags & ISS_F_SIGMOID) x = 1.f / (1.f + expf(-x));
            if (a.flags & ISS_F_AFFINE_POST) x = fmaf(x, es2[j], et2[j]);
            v[j] = x;
        }
        float *o = a.out + mr * a.N + n0 + tx * TN;
        if (vecB && n0 + tx * TN + TN <= a.N) {
#pragma unroll
            for (int q = 0; q < TN; q += 4) *reinterpret_cast<float4 *>(o + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j) if (n0 + tx * TN + j < a.N) o[j] = v[j];
        }
    }
}


// Dense / 1x1 layer with a handful of outputs (the softmax heads: 512 -> 2 or 3).  The generic 128 x 64 tile kernel spends
// 48 us on this 6 MFLOP layer (one column of CTAs, 3 useful columns of 64); here one warp owns a GEMM row, the K dimension
// is split over the lanes with 16-byte loads, the weights sit transposed [N][K] in shared memory (conflict-free float4
// reads) and the lane partials are combined in a fixed order (xor tree): deterministic, independent of the grid.
constexpr int SMALL_N = 8;
__global__ void __launch_bounds__(256)
dense_small_n_kernel(const ConvArgs a)
{
    extern __shared__ __align__(16) float wsm_t[];                 // [N][K]
    for (int i = threadIdx.x; i < a.K * a.N; i += 256) { const int k = i / a.N, n = i - k * a.N; wsm_t[n * a.K + k] = a.w[i]; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K4 = a.K >> 2;
    for (int64_t m = (int64_t)blockIdx.x * 8 + warp; m < a.M; m += (int64_t)gridDim.x * 8) {
        const float4 *x = reinterpret_cast<const float4 *>(a.in + m * a.K);
        float acc[SMALL_N];
#pragma unroll
        for (int n = 0; n < SMALL_N; ++n) acc[n] = 0.f;
        for (int k4 = lane; k4 < K4; k4 += 32) {
            const float4 v = __ldg(x + k4);
#pragma unroll
            for (int n = 0; n < SMALL_N; ++n) {
                if (n < a.N) {
                    const float4 w = *reinterpret_cast<const float4 *>(wsm_t + n * a.K + 4 * k4);
                    acc[n] = fmaf(v.x, w.x, acc[n]); acc[n] = fmaf(v.y, w.y, acc[n]);
                    acc[n] = fmaf(v.z, w.z, acc[n]); acc[n] = fmaf(v.w, w.w, acc[n]);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < SMALL_N; ++n)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
        if (lane < a.N) {
            float xv = 0.f;
#pragma unroll
            for (int n = 0; n < SMALL_N; ++n) if (n == lane) xv = acc[n];
            const int n = lane;
            if (a.flags & ISS_F_BIAS) xv += a.bias[n];
            if (a.flags & ISS_F_AFFINE_PRE) xv = fmaf(xv, a.pre_scale[n], a.pre_shift[n]);
            if (a.flags & ISS_F_RELU) xv = fmaxf(xv, 0.f);
            if (a.flags & ISS_F_SIGMOID) xv = 1.f / (1.f + expf(-xv));
            if (a.flags & ISS_F_AFFINE_POST) xv = fmaf(xv, a.post_scale[n], a.post_shift[n]);
            a.out[m * a.N + n] = xv;
        }
    }
}

bool small_n_covers(const ConvArgs &a)
{
    return a.KH == 1 && a.KW == 1 && a.SH == 1 && a.SW == 1 && a.PT == 0 && a.PL == 0 && a.N <= SMALL_N && a.K % 4 == 0 && a.K == a.C &&
           (size_t)a.K * a.N * sizeof(float) <= 48 * 1024 && !a.in_packed && !a.out_packed && !(a.flags & ISS_F_RESIDUAL);
}

template <bool FIRST>
int launch_conv_t(const ConvArgs &a, cudaStream_t st)
{
    const int64_t gm = (a.M + BM - 1) / BM;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv: M too large");
    if (a.N <= 64) {
        dim3 grid((unsigned)gm, (unsigned)((a.N + 63) / 64));
        conv_gemm_f32_kernel<64, FIRST><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid((unsigned)gm, (unsigned)((a.N + 127) / 128));
        conv_gemm_f32_kernel<128, FIRST><<<grid, 256, 0, st>>>(a);
    }
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

}  // namespace

int iss_launch_conv(const ConvArgs &a, bool first, cudaStream_t st)
{
    if (!first && small_n_covers(a)) {
        const int64_t gm = (a.M + 7) / 8;
        dense_small_n_kernel<<<(unsigned)(gm < 148 * 8 ? gm : 148 * 8), 256, (size_t)a.K * a.N * sizeof(float), st>>>(a);
        ISS_CUDA_OK(cudaGetLastError());
        iss_count_launch();
        return ISS_OK;
    }
    if (!first) {
        const int mode = iss_get_gemm_mode();
        if (mode != ISS_GEMM_FP32 && iss_conv_tc_eligible(a)) return iss_launch_conv_tc(a, mode, st);
    }
    return first ? launch_conv_t<true>(a, st) : launch_conv_t<false>(a, st);
}
