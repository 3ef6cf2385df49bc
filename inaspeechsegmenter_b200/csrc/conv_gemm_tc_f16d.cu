// conv_gemm_tc_f16d.cu -- engine 3, DIRECT variant: the fp16-split convolution with BOTH tensor-core operands read
// from shared memory (tcgen05.mma SS form); there are no A-operand producer warps at all.
//
// Idea.  Lay the NHWC input of a whole batch out as ONE tall image of n_img * H rows of W pixels and give every pixel
// p = (img * H + ih) * W + iw a 128-byte shared-memory row per 64-channel plane (fp16 hi plane, fp16 lo plane,
// SWIZZLE_128B K-major).  Number the outputs the same way -- "slot" s = (img * H + oh) * W + ow -- and the A tile of
// filter tap (kh, kw) for the 128 slots s0 .. s0 + 127 is simply the 128 pixel rows starting at s0 + kh * W + kw:
// a contiguous, row-shifted window of the slab, which a UMMA shared-memory descriptor can address directly (the
// 128-byte swizzle is a function of the absolute shared-memory address, so a descriptor may start at any row:
// tools/umma_desc_offset_test.cu).  The price: slots with ow >= OW (KW - 1 per row) and oh >= OH (KH - 1 rows per
// image) are computed and thrown away -- (OW / W) * (OH / H) = 77 % useful rows for the 5x4 layer of the segmenter
// CNNs -- in exchange for deleting the producers' 32 KB of LDS + PRMT + tcgen05.st per k-block, which bounded the
// TMEM-operand kernel (conv_gemm_tc_f16.cu) at ~36 % of the tensor pipe.
//
// One persistent CTA per SM, 448 threads, warp-specialised:
//   warp 0      MMA issuer (warp-uniform loop, tcgen05 under elect.sync)
//   warp 1      weight loader: one cp.async.bulk per 16 KB stage [Bh | Bl], 3-deep ring
//   warps 2-5   epilogue (one TMEM lane quadrant each): accumulators -> bias/BN/ReLU -> split-half words or fp32 -> HBM
//   warps 6-13  slab fill: either de-interleaves split-half words from HBM (IN_PACKED) or EVALUATES the one-channel
//               first convolution from the shared float64 map Y (IN_FIRST, FirstFuse in conv_gemm.cuh)
// A tile is DT = 2 sub-tiles of 128 slots that share the slab and every weight stage (half the weight traffic of a
// 128-row tile); slab and accumulators are double-buffered, so fill(i+1), MMA(i) and epilogue(i-1) overlap.
// Per k-block (= one filter tap, C = 64) and sub-tile: Ah.[Bh | Bl] (N = 128) and Al.Bh (N = 64), as in the other
// fp16-split kernels; TMEM: 2 buffers x 2 sub-tiles x 128 columns = 512.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

int iss_launch_conv_tc_f16d(ConvArgs &a, cudaStream_t st);
bool iss_conv_f16_direct_covers(const ConvArgs &a);

namespace {

constexpr int HBK = 64;
constexpr int DT = 2;                                   // 128-slot sub-tiles per tile
constexpr int DSB = 3;                                  // weight stages
constexpr int DBN = 64;                                 // output channels (one n-tile)
constexpr int D_FILL_WARPS = 8;
constexpr int D_FILL_THREADS = 32 * D_FILL_WARPS;
constexpr int D_FIRST_FILL = 6;                         // first fill warp
constexpr int D_THREADS = 32 * (D_FIRST_FILL + D_FILL_WARPS);
constexpr int D_B_STAGE = 2 * DBN * 128;                // [hi rows | lo rows]
constexpr int D_TAB = 8;                                // images a slab may touch (IN_FIRST table)
constexpr int D_SMEM_MAX = 232448;

constexpr int DIN_PACKED = 1, DIN_FIRST = 2;

struct DirectArgs {
    const unsigned char *wt;    // tiled fp16 image [k-block][hi | lo][64 rows x 128 B, SWIZZLE_128B]
    float inv_scale;
    int npix;                   // slab rows (pixels), multiple of 8
    int n_tiles;
    int n_img;
    int64_t total_pix;          // n_img * H * W
    int desc_base_offset;       // 1: put (addr >> 7) & 7 into the descriptor's base-offset field
};

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

struct DSmem {                                           // everything behind the 1024-aligned operand buffers
    float k1[DBN], k0[DBN], es2[DBN], et2[DBN];         // epilogue: y = acc * k1 + k0, ReLU, y * es2 + et2
    double tab_mu[2][D_TAB], tab_inv[2][D_TAB];
    long long tab_row[2][D_TAB];
    uint64_t slab_full[2], slab_empty[2], acc_full[2], acc_empty[2], b_full[DSB], b_empty[DSB];
    uint32_t tmem_slot;
};

template <int MODE>
__global__ void __launch_bounds__(D_THREADS, 1)
conv_gemm_tc4h_kernel(const ConvArgs a, const FirstFuse ff, const DirectArgs d)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const uint32_t plane = (uint32_t)d.npix * 128u;       // bytes of one fp16 plane of the slab (multiple of 1024)
    unsigned char *b_ring = smem;
    unsigned char *slab = smem + DSB * D_B_STAGE;         // [2 buffers][hi plane | lo plane]
    DSmem *sm = reinterpret_cast<DSmem *>(slab + 4 * (size_t)plane);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = a.KH * a.KW;                          // one k-block per filter tap (C == 64)
    const uint32_t HW = (uint32_t)(a.H * a.W);

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm->slab_full[b], D_FILL_WARPS); mbar_init(&sm->slab_empty[b], 1);
            mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], 4);
        }
        for (int s = 0; s < DSB; ++s) { mbar_init(&sm->b_full[s], 1); mbar_init(&sm->b_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid >= 64 && tid < 64 + DBN) {                    // epilogue constants, once per CTA
        const int n = tid - 64;
        const bool has_bias = a.flags & ISS_F_BIAS, pre = a.flags & ISS_F_AFFINE_PRE, post = a.flags & ISS_F_AFFINE_POST;
        const float eb = has_bias ? __ldg(a.bias + n) : 0.f;
        const float s1 = pre ? __ldg(a.pre_scale + n) : 1.f, t1 = pre ? __ldg(a.pre_shift + n) : 0.f;
        sm->k1[n] = d.inv_scale * s1; sm->k0[n] = fmaf(eb, s1, t1);
        sm->es2[n] = post ? __ldg(a.post_scale + n) : 1.f; sm->et2[n] = post ? __ldg(a.post_shift + n) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_slot;
    const uint32_t slab_u32 = smem_u32(slab);

    if (warp == 0) {
        // ============================ MMA issuer ============================
        // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(DBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
        constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * DBN) >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
        const uint32_t tb = __reduce_or_sync(0xffffffffu, tmem_base);
        const uint32_t b_u32 = smem_u32(b_ring);
        uint32_t g = 0;                                   // running k-block count of this CTA (weight ring position)
        int i = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            const uint32_t use = (uint32_t)(i >> 1) & 1u;
            mbar_wait(&sm->slab_full[buf], use, 1);
            mbar_wait(&sm->acc_empty[buf], use ^ 1u, 2);
            tc_fence_after();
            const uint32_t slab_b = slab_u32 + (uint32_t)buf * 2u * plane;
            int kh = 0, kw = 0;
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const uint32_t sl = g % DSB;
                mbar_wait(&sm->b_full[sl], (g / DSB) & 1u, 3);
                tc_fence_after();
                const uint64_t db = make_sw128_desc(b_u32 + sl * D_B_STAGE);
                const uint32_t tap = slab_b + (uint32_t)(kh * a.W + kw) * 128u;
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < DT; ++t) {
                        const uint32_t arow = tap + (uint32_t)t * (128u * 128u);
                        uint64_t dah = make_sw128_desc(arow), dal = make_sw128_desc(arow + plane);
                        if (d.desc_base_offset) {
                            dah |= (uint64_t)((arow >> 7) & 7u) << 49;
                            dal |= (uint64_t)(((arow + plane) >> 7) & 7u) << 49;
                        }
                        const uint32_t dm = tb + (uint32_t)buf * 256u + (uint32_t)t * 128u;
#pragma unroll
                        for (int kk = 0; kk < HBK / 16; ++kk) {
                            umma_f16_ss(dm, dah + 2 * kk, db + 2 * kk, idesc2, (kb > 0 || kk > 0) ? 1u : 0u);     // Ah.[Bh | Bl]
                            umma_f16_ss(dm + DBN, dal + 2 * kk, db + 2 * kk, idesc, 1u);                          // Al.Bh
                        }
                    }
                    umma_commit(&sm->b_empty[sl]);
                    if (kb == nkb - 1) { umma_commit(&sm->slab_empty[buf]); umma_commit(&sm->acc_full[buf]); }
                }
                __syncwarp();
                if (++kw == a.KW) { kw = 0; ++kh; }
            }
        }
        tc_fence_before();
    } else if (warp == 1) {
        // ============================ weight loader ============================
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const uint32_t sl = g % DSB;
                mbar_wait(&sm->b_empty[sl], ((g / DSB) & 1u) ^ 1u, 4);
                if (elect_one()) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->b_full[sl])), "r"((uint32_t)D_B_STAGE) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(b_ring + sl * D_B_STAGE)), "l"(d.wt + (size_t)kb * D_B_STAGE),
                                   "r"((uint32_t)D_B_STAGE), "r"(smem_u32(&sm->b_full[sl])) : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp < D_FIRST_FILL) {
        // ============================ epilogue ============================
        const int quad = warp & 3;                        // the TMEM lanes this warp may read
        const uint32_t lane_addr = ((uint32_t)(quad * 32)) << 16;
        const bool relu = a.flags & ISS_F_RELU, post = a.flags & ISS_F_AFFINE_POST;
        int i = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            const uint32_t use = (uint32_t)(i >> 1) & 1u;
            if (lane == 0) mbar_wait(&sm->acc_full[buf], use, 5);
            __syncwarp();
            tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < DT; ++t) {
                const uint32_t slot = (uint32_t)tile * (DT * 128u) + (uint32_t)t * 128u + (uint32_t)(quad * 32 + lane);
                const uint32_t img = slot / HW, rem = slot - img * HW;
                const uint32_t oh = rem / (uint32_t)a.W, ow = rem - oh * (uint32_t)a.W;
                const bool valid = img < (uint32_t)d.n_img && oh < (uint32_t)a.OH && ow < (uint32_t)a.OW;
                float *dst = a.out + (((int64_t)img * a.OH + oh) * a.OW + ow) * DBN;
#pragma unroll 1
                for (int c = 0; c < DBN; c += 32) {
                    uint32_t acc[32], corr[32];
                    const uint32_t col = tmem_base + lane_addr + (uint32_t)buf * 256u + (uint32_t)t * 128u + (uint32_t)c;
                    tmem_ld32(col, acc);
                    tmem_ld32(col + DBN, corr);
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float y[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int n = c + 4 * j + q;
                                float v = fmaf(__uint_as_float(acc[4 * j + q]) + __uint_as_float(corr[4 * j + q]), sm->k1[n], sm->k0[n]);
                                if (relu) v = fmaxf(v, 0.f);
                                if (post) v = fmaf(v, sm->es2[n], sm->et2[n]);
                                y[q] = v;
                            }
                            if (a.out_packed)
                                *reinterpret_cast<uint4 *>(dst + c + 4 * j) = make_uint4(iss_pack_split(y[0]), iss_pack_split(y[1]), iss_pack_split(y[2]), iss_pack_split(y[3]));
                            else
                                *reinterpret_cast<float4 *>(dst + c + 4 * j) = make_float4(y[0], y[1], y[2], y[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->acc_empty[buf]);
        }
    } else {
        // ============================ slab fill ============================
        const int ftid = tid - D_FIRST_FILL * 32;
        const int j = ftid & 7;                           // 16-byte chunk = channels 8j .. 8j+7
        // per-thread channel constants of the fused first layer
        double Sc[8];
        float fk1[8], fk0[8], fs2[8], ft2[8];
        if constexpr (MODE == DIN_FIRST) {
            const int f_flags = ff.flags;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = j * 8 + e;
                Sc[e] = ff.S[c];
                const float eb = (f_flags & ISS_F_BIAS) ? ff.bias[c] : 0.f;
                const float s1 = (f_flags & ISS_F_AFFINE_PRE) ? ff.pre_scale[c] : 1.f, t1 = (f_flags & ISS_F_AFFINE_PRE) ? ff.pre_shift[c] : 0.f;
                fk1[e] = s1; fk0[e] = fmaf(eb, s1, t1);
                fs2[e] = (f_flags & ISS_F_AFFINE_POST) ? ff.post_scale[c] : 1.f; ft2[e] = (f_flags & ISS_F_AFFINE_POST) ? ff.post_shift[c] : 0.f;
            }
        }
        int i = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            const uint32_t use = (uint32_t)(i >> 1) & 1u;
            if (lane == 0) mbar_wait(&sm->slab_empty[buf], use ^ 1u, 6);
            __syncwarp();
            const uint32_t s0 = (uint32_t)tile * (DT * 128u);          // global pixel of slab row 0
            const uint32_t hi_base = slab_u32 + (uint32_t)buf * 2u * plane;
            uint32_t img0 = 0;
            if constexpr (MODE == DIN_FIRST) {
                img0 = s0 / HW;
                if (ftid < D_TAB) {
                    const int64_t img = (int64_t)img0 + ftid;
                    double mu = 0.0, inv = 0.0;
                    long long yr = -1;
                    if (img < ff.n_img) {
                        mu = (double)ff.mu[img];
                        inv = 1.0 / (double)ff.sigma[img];
                        yr = (long long)ff.row0[img] - ff.y_f0;
                    }
                    sm->tab_mu[buf][ftid] = mu; sm->tab_inv[buf][ftid] = inv; sm->tab_row[buf][ftid] = yr;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(D_FILL_THREADS) : "memory");
            }
            if constexpr (MODE == DIN_FIRST) {
                const int f_flags = ff.flags;
                const int64_t row_len = (int64_t)a.W * HBK;            // doubles per Y row
                constexpr int FB = 2;                                   // pixels whose loads are in flight together
                for (int pl0 = ftid >> 3; pl0 < d.npix; pl0 += FB * (D_FILL_THREADS / 8)) {
                    double2 yv[FB][4];
                    int ti[FB];
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        const int pl = pl0 + f * (D_FILL_THREADS / 8);
                        ti[f] = -1;                                     // -1: nothing to write, D_TAB: zero fill
                        if (pl < d.npix) {
                            const uint32_t gp = s0 + (uint32_t)pl;
                            const uint32_t img = gp / HW, rem = gp - img * HW;
                            const uint32_t ih = rem / (uint32_t)a.W, x = rem - ih * (uint32_t)a.W;
                            const uint32_t t = img - img0;
                            ti[f] = D_TAB;
                            if (t < D_TAB && sm->tab_row[buf][t] >= 0) {
                                ti[f] = (int)t;
                                const double2 *yp = reinterpret_cast<const double2 *>(ff.Y + (sm->tab_row[buf][t] + ih) * row_len + x * HBK + j * 8);
                                yv[f][0] = __ldg(yp); yv[f][1] = __ldg(yp + 1); yv[f][2] = __ldg(yp + 2); yv[f][3] = __ldg(yp + 3);
                            }
                        }
                    }
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        if (ti[f] < 0) continue;
                        const int pl = pl0 + f * (D_FILL_THREADS / 8);
                        uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                        if (ti[f] < D_TAB) {
                            const double mu = sm->tab_mu[buf][ti[f]], inv = sm->tab_inv[buf][ti[f]];
                            const double y8[8] = {yv[f][0].x, yv[f][0].y, yv[f][1].x, yv[f][1].y, yv[f][2].x, yv[f][2].y, yv[f][3].x, yv[f][3].y};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float v = fmaf((float)((y8[e] - mu * Sc[e]) * inv), fk1[e], fk0[e]);
                                if (f_flags & ISS_F_RELU) v = fmaxf(v, 0.f);
                                if (f_flags & ISS_F_AFFINE_POST) v = fmaf(v, fs2[e], ft2[e]);
                                w[e] = iss_pack_split(v);
                            }
                        }
                        const uint32_t dst = hi_base + (uint32_t)pl * 128u + (uint32_t)((j ^ (pl & 7)) << 4);
                        sts128(dst, __byte_perm(w[0], w[1], 0x5410), __byte_perm(w[2], w[3], 0x5410), __byte_perm(w[4], w[5], 0x5410), __byte_perm(w[6], w[7], 0x5410));
                        sts128(dst + plane, __byte_perm(w[0], w[1], 0x7632), __byte_perm(w[2], w[3], 0x7632), __byte_perm(w[4], w[5], 0x7632), __byte_perm(w[6], w[7], 0x7632));
                    }
                }
            } else {
                const uint4 *src = reinterpret_cast<const uint4 *>(a.in);
                for (int pl = ftid >> 3; pl < d.npix; pl += D_FILL_THREADS / 8) {
                    const int64_t gp = (int64_t)s0 + pl;
                    uint4 u0 = make_uint4(0u, 0u, 0u, 0u), u1 = u0;
                    if (gp < d.total_pix) {
                        const uint4 *p = src + gp * (HBK / 4) + j * 2;
                        u0 = __ldg(p); u1 = __ldg(p + 1);
                    }
                    const uint32_t dst = hi_base + (uint32_t)pl * 128u + (uint32_t)((j ^ (pl & 7)) << 4);
                    sts128(dst, __byte_perm(u0.x, u0.y, 0x5410), __byte_perm(u0.z, u0.w, 0x5410), __byte_perm(u1.x, u1.y, 0x5410), __byte_perm(u1.z, u1.w, 0x5410));
                    sts128(dst + plane, __byte_perm(u0.x, u0.y, 0x7632), __byte_perm(u0.z, u0.w, 0x7632), __byte_perm(u1.x, u1.y, 0x7632), __byte_perm(u1.z, u1.w, 0x7632));
                }
            }
            fence_proxy_async();                          // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->slab_full[buf]);
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int direct_npix(const ConvArgs &a)
{
    const int npix = DT * 128 + (a.KH - 1) * a.W + a.KW - 1;
    return (npix + 7) & ~7;
}

size_t direct_smem(const ConvArgs &a)
{
    return (size_t)DSB * D_B_STAGE + 4 * (size_t)direct_npix(a) * 128 + sizeof(DSmem) + 1024;
}

}  // namespace

// Does the direct kernel take this layer?  Un-padded stride-1 KHxKW convolution, 64 -> 64 channels (one k-block per
// filter tap, one n-tile), input either split-half words or the fused first layer.  ISS_B200_F16_DIRECT=0 turns it off
// (A/B runs against the TMEM-operand slab kernel).
bool iss_conv_f16_direct_covers(const ConvArgs &a)
{
    static const bool off = [] { const char *e = getenv("ISS_B200_F16_DIRECT"); return e && e[0] == '0'; }();
    if (off) return false;
    if (!a.wt_f16 || a.SH != 1 || a.SW != 1 || a.PT != 0 || a.PL != 0 || a.KH * a.KW <= 1) return false;
    if (a.OH != a.H - a.KH + 1 || a.OW != a.W - a.KW + 1 || a.Kp != a.K) return false;
    if (a.N != DBN || a.C != HBK || a.K != a.KH * a.KW * HBK) return false;
    if (a.flags & (ISS_F_SIGMOID | ISS_F_RESIDUAL)) return false;
    if (!a.first && !a.in_packed) return false;
    if (iss_f16_bn_for(a.N) != DBN) return false;                       // weight image tiled for 64-wide n-tiles
    const int64_t n_img = a.M / ((int64_t)a.OH * a.OW);
    if (n_img * a.H * a.W >= (1ll << 31) - 4096) return false;
    // the IN_FIRST table covers D_TAB images per slab
    if ((direct_npix(a) - 1) / (a.H * a.W) + 2 > D_TAB) return false;
    return direct_smem(a) <= (size_t)D_SMEM_MAX;
}

// Returns 1 when the layer is not covered (caller continues with the other fp16-split kernels).
int iss_launch_conv_tc_f16d(ConvArgs &a, cudaStream_t st)
{
    if (!iss_conv_f16_direct_covers(a)) return 1;
    static const int base_off = [] { const char *e = getenv("ISS_B200_DESC_BASE_OFFSET"); return e ? atoi(e) : 0; }();
    DirectArgs d = {};
    d.wt = reinterpret_cast<const unsigned char *>(a.wt_f16);
    d.inv_scale = a.wt_f16_inv_scale;
    d.npix = direct_npix(a);
    d.n_img = (int)(a.M / ((int64_t)a.OH * a.OW));
    d.total_pix = (int64_t)d.n_img * a.H * a.W;
    const int64_t total_slots = ((int64_t)(d.n_img - 1) * a.H + a.OH - 1) * a.W + a.OW;
    d.n_tiles = (int)((total_slots + DT * 128 - 1) / (DT * 128));
    d.desc_base_offset = base_off;
    int dev = 0, sms = 0;
    ISS_CUDA_OK(cudaGetDevice(&dev));
    ISS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(d.n_tiles < sms ? d.n_tiles : sms);
    const size_t smem = direct_smem(a);
    FirstFuse ff = {};
    if (a.first) {
        ff = *a.first;
        auto kern = conv_gemm_tc4h_kernel<DIN_FIRST>;
        ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), D_SMEM_MAX));
        kern<<<grid, D_THREADS, smem, st>>>(a, ff, d);
    } else {
        auto kern = conv_gemm_tc4h_kernel<DIN_PACKED>;
        ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), D_SMEM_MAX));
        kern<<<grid, D_THREADS, smem, st>>>(a, ff, d);
    }
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}
