// conv_gemm_tc_f16.cu -- engine 3 (the default, ISS_B200_GEMM=tc_f16): slab convolution on tcgen05 kind::f16
// with every fp32 operand split into TWO fp16 numbers.
//
//      x = hi + lo,  hi = fp16(x),  lo = fp16(x - hi)          (22 significant bits, like the TF32 split)
//      A.B ~= Ah.Bh + Ah.Bl + Al.Bh                            (three kind::f16 MMAs, fp32 accumulation)
//
// kind::f16 runs at twice the kind::tf32 rate and its operands are half as wide, so one 128-byte swizzle
// row holds 64 k-elements: per unit of K the tensor time, the weight-tile bytes, the TMEM-store bytes and --
// what bounds the TF32 kernel today -- the number of producer/issuer hand-shakes all halve.
// The weights of a layer are pre-scaled by an exact power of two so that max|w| lands in [2^12, 2^13) (fp16
// keeps 11 bits down to 6e-5 and is exact to 6e-8 below; the scale is undone in the epilogue); activations
// are used as they are, so |activation| must stay below 65504 (true after BatchNorm/z-normalisation; a
// network that violates it produces inf and must use engine 2).
// Algorithmic accuracy of the split (exact products and sums, tests/test_split_accuracy.py): 8e-7 on the
// stand-in VAD softmax vs 1.3e-6 for the TF32 split -- both at the fp32 oracle's own rounding noise.
//
// Validated on B200 in round 2 (profiles/r02_bringup_variants.txt): softmax error vs the fp32 CUDA-core engine
// 4.5e-6 (VAD) / 7.3e-6 (gender), i.e. below the TF32 split's 7-8e-6, and 1.74x faster on the slab layers.
// Dense packing of 16-bit A operands in tensor memory: 32-bit column c of lane m holds k = 2c in its low half,
// k = 2c+1 in its high half.
//
// PACKED activations ("split-half words"): a layer whose consumer is another slab convolution of this engine
// stores every activation as ONE 32-bit word  lo16 << 16 | hi16  (hi = fp16(x), lo = fp16(x - hi)) instead of the
// fp32 value -- same bytes in HBM, but the hi/lo split is then done once per element in the producing epilogue
// instead of KH*KW times in the consumers' A-operand producers, whose per-k-block work drops from 16 LDS.128 +
// ~190 conversions/subtractions to 16 LDS.128 + 64 PRMT (they were the bottleneck of the first version).  The value
// carried between layers is hi + lo, exactly what the tensor cores would have used anyway.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "f16_image.cuh"
#include "tc_common.cuh"

int iss_launch_conv_tc_f16(ConvArgs &a, cudaStream_t st);

namespace {

constexpr int HBK = 64;                     // k-elements per block = one 128-byte swizzle row of halves

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// Explicit shared-space load: through a generic pointer ptxas emits LD.E.128 with 64-bit address arithmetic for
// the slab reads (seen in conv_gemm_tc3_kernel's SASS; to be changed there too once this variant is validated).
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int BN, int SB, int ST>
struct TcHCfg {
    static constexpr int THREADS = 160;
    static constexpr int B_TILE = BN * 128;                              // BN rows x 64 halves
    static constexpr int B_STAGE = 2 * B_TILE;                           // hi | lo
    static constexpr uint32_t ACC_COLS = 2 * BN;                         // D_main | D_lo
    static constexpr uint32_t A_COLS = 64;                               // 32 packed columns hi + 32 lo per stage
    static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(ACC_COLS + ST * A_COLS);
    static constexpr int FIXED = SB * B_STAGE + 1024 + 256;              // + slab bytes
};

struct F16Args {
    const unsigned char *wt;    // tiled fp16 image [n-tile][k-block][hi|lo][BN rows x 128 B, SWIZZLE_128B]
    float inv_scale;
};

__device__ __forceinline__ uint4 lds128u(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// MODE: format of the input tensor -- IN_F32: fp32 activations, split in the A producers; IN_PACKED: split-half
// words (see the file header).  a.out_packed selects the output format.  (The fused first layer lives in the
// direct kernel, conv_gemm_tc_f16d.cu.)
constexpr int IN_F32 = 0, IN_PACKED = 1;

template <int BN, int SB, int ST, int MODE>
__global__ void __launch_bounds__(160, (TcHCfg<BN, SB, ST>::TMEM_COLS <= 256 ? 2 : 1))
conv_gemm_tc3h_kernel(const ConvArgs a, const F16Args h)
{
    constexpr bool PACKED = MODE != IN_F32;
    using Cfg = TcHCfg<BN, SB, ST>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *b_ring = smem;
    unsigned char *slab = smem + SB * Cfg::B_STAGE;
    const int slab_bytes = a.slab_rows * a.W * a.C * 4 < 32768 ? 32768 : a.slab_rows * a.W * a.C * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(slab + slab_bytes);
    uint64_t *fullA = bars, *emptyA = bars + ST, *emptyB = bars + 2 * ST, *fullB = bars + 2 * ST + SB, *accum = bars + 2 * ST + 2 * SB;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * ST + 2 * SB + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * BN;
    const int nkb = a.K / HBK;
    const int R = a.slab_R, KH1 = a.KH - 1;
    const int64_t Q = a.M / a.OW;
    // tile -> output rows [q0, q0 + nq) of the global row sequence q = img * OH + oh.  slab_tpi > 0: tiles never
    // straddle images (tile t of image img covers rows t*R ..), which keeps the slab at R + KH - 1 rows; otherwise
    // tiles are R consecutive rows of the whole sequence and may straddle an image boundary.
    int64_t q0, img0;
    int nq;
    if (a.slab_tpi > 0) {
        img0 = blockIdx.x / a.slab_tpi;
        const int oh0 = (int)(blockIdx.x - img0 * a.slab_tpi) * R;
        q0 = img0 * a.OH + oh0;
        nq = (a.OH - oh0) < R ? (a.OH - oh0) : R;
    } else {
        q0 = (int64_t)blockIdx.x * R;
        nq = (int)((Q - q0) < (int64_t)R ? (Q - q0) : (int64_t)R);
        img0 = q0 / a.OH;
    }
    const int64_t mbase = q0 * a.OW;
    const int valid = nq * a.OW;

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&fullA[s], 4); mbar_init(&emptyA[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&emptyB[s], 1); mbar_init(&fullB[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp != 4) {
        // ============================ slab fill (as conv_gemm_tc3_kernel) ============================
        const int quad = warp;
        const uint32_t slab_u32 = smem_u32(slab);
        const uint32_t pix_bytes = (uint32_t)a.C * 4;
        {
            const int cpp = a.C >> 2;
            const int64_t g0 = q0 + img0 * KH1;
            const int64_t img1 = (q0 + nq - 1) / a.OH;
            const int rows = nq + KH1 * (int)(img1 - img0 + 1);
            const int64_t first = g0 * a.W * a.C;
            const float *src0 = a.in + first;
            const int64_t avail = (a.in_elems - first) >> 2;
            const int total = rows * a.W * cpp;
            int p = tid / cpp, j = tid - p * cpp;
            const int dp = 128 / cpp, dj = 128 - dp * cpp;
            {
                for (int q = tid; q < total; q += 128) {
                    const uint32_t dst = slab_u32 + (uint32_t)p * pix_bytes + (uint32_t)(((j & ~7) | ((j ^ p) & 7)) << 4);
                    const bool ok = q < avail;
                    cp_async16_u32(dst, src0 + (ok ? (size_t)q * 4 : 0), ok ? 16 : 0);
                    p += dp; j += dj;
                    if (j >= cpp) { j -= cpp; ++p; }
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        // ============================ A producers ============================
        const int r = quad * 32 + lane;
        int pix0 = 0;
        if (r < valid) {
            const int dq = r / a.OW, ow = r - dq * a.OW;
            pix0 = (dq + KH1 * (int)((q0 + dq) / a.OH - img0)) * a.W + ow;
        }
        int is_c0 = 0, is_ss = 0, is_poff = 0;
        const uint32_t lane_addr = ((uint32_t)(quad * 32)) << 16;
        uint32_t hi[32], lo[32];                                 // 64 k-elements, two halves per register
        auto load_split = [&]() {
            const int p = pix0 + is_poff;
            const uint32_t base = slab_u32 + (uint32_t)p * pix_bytes + (uint32_t)is_c0 * 4;
            const uint32_t x = (uint32_t)(p & 7) << 4;
#pragma unroll
            for (int half = 0; half < 2; ++half) {               // 2 x 32 words keeps the register peak down
                if constexpr (PACKED) {
                    uint32_t v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 q = lds128u(base + ((((uint32_t)(half * 8 + j)) << 4) ^ x));
                        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {                // word = lo << 16 | hi; operand column = k even (low) | k odd (high)
                        hi[half * 16 + i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x5410);
                        lo[half * 16 + i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x7632);
                    }
                } else {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q = lds128(base + ((((uint32_t)(half * 8 + j)) << 4) ^ x));
                        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);      // low half = even k
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
                        hi[half * 16 + i] = *reinterpret_cast<const uint32_t *>(&hh);
                        lo[half * 16 + i] = *reinterpret_cast<const uint32_t *>(&ll);
                    }
                }
            }
            is_c0 += HBK;
            if (is_c0 == a.C) {
                is_c0 = 0; ++is_poff;
                if (++is_ss == a.KW) { is_ss = 0; is_poff += a.W - a.KW; }
            }
        };
        load_split();
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % ST;
            if (lane == 0) mbar_wait(&emptyA[st], ((kb / ST) & 1) ^ 1, 1);
            __syncwarp();
            tc_fence_after();
            const uint32_t ta = tmem_base + lane_addr + Cfg::ACC_COLS + st * Cfg::A_COLS;
            tmem_st32(ta, hi);
            tmem_st32(ta + 32, lo);
            if (kb + 1 < nkb) load_split();
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA[st]);
        }

        // ============================ epilogue ============================
        // (28 % of a CTA's life in the first version: fully unrolled 1700-instruction body -> instruction-cache misses, two
        // store paths per position.)  Bias + BatchNorm affine are folded into one FMA per value (inv_scale is a power of two),
        // the output format is a template of the store lambda, the row loop is unrolled by 2 only.
        if (lane == 0) mbar_wait(accum, 0, 5);
        __syncwarp();
        tc_fence_after();
        unsigned char *stage_buf = slab + quad * 4096;
        const int sub = lane >> 3, chunk = lane & 7;
        const bool has_bias = a.flags & ISS_F_BIAS, pre = a.flags & ISS_F_AFFINE_PRE, post = a.flags & ISS_F_AFFINE_POST;
        const bool relu = a.flags & ISS_F_RELU, resid = a.flags & ISS_F_RESIDUAL;
        const float inv_s = h.inv_scale;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            {
                uint32_t acc[32], corr[32];
                tmem_ld32(tmem_base + lane_addr + c, acc);
                tmem_ld32(tmem_base + lane_addr + BN + c, corr);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_float4(__uint_as_float(acc[4 * j]) + __uint_as_float(corr[4 * j]), __uint_as_float(acc[4 * j + 1]) + __uint_as_float(corr[4 * j + 1]),
                                    __uint_as_float(acc[4 * j + 2]) + __uint_as_float(corr[4 * j + 2]), __uint_as_float(acc[4 * j + 3]) + __uint_as_float(corr[4 * j + 3]));
            }
            __syncwarp();
            const int nb = n0 + c + chunk * 4;
            float k1[4], k0[4], es2[4], et2[4];                  // y = acc * k1 + k0  ==  ((acc * inv_s) + bias) * pre_scale + pre_shift
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float eb = has_bias ? __ldg(a.bias + nb + q) : 0.f;
                const float s1 = pre ? __ldg(a.pre_scale + nb + q) : 1.f, t1 = pre ? __ldg(a.pre_shift + nb + q) : 0.f;
                k1[q] = inv_s * s1; k0[q] = fmaf(eb, s1, t1);
                es2[q] = post ? __ldg(a.post_scale + nb + q) : 1.f; et2[q] = post ? __ldg(a.post_shift + nb + q) : 0.f;
            }
            auto rows = [&](auto packed_tag, auto resid_tag) {
                constexpr bool OUT_PACKED = decltype(packed_tag)::value, RESID = decltype(resid_tag)::value;
#pragma unroll 2
                for (int i = 0; i < 8; ++i) {
                    const int rl = 4 * i + sub;
                    const int rr = quad * 32 + rl;
                    if (rr >= valid) continue;
                    const float4 q4 = *reinterpret_cast<const float4 *>(stage_buf + rl * 128 + ((chunk ^ (rl & 7)) << 4));
                    float y[4] = {q4.x, q4.y, q4.z, q4.w};
                    float *dst = a.out + (mbase + rr) * a.N + nb;
                    float rv[4] = {0.f, 0.f, 0.f, 0.f};
                    if constexpr (RESID) {
                        const uint4 rw = __ldg(reinterpret_cast<const uint4 *>(a.residual + (mbase + rr) * a.N + nb));
                        if (a.residual_packed) { rv[0] = iss_unpack_split(rw.x); rv[1] = iss_unpack_split(rw.y); rv[2] = iss_unpack_split(rw.z); rv[3] = iss_unpack_split(rw.w); }
                        else { rv[0] = __uint_as_float(rw.x); rv[1] = __uint_as_float(rw.y); rv[2] = __uint_as_float(rw.z); rv[3] = __uint_as_float(rw.w); }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float t = fmaf(y[q], k1[q], k0[q]);
                        if constexpr (RESID) t += rv[q];
                        if (relu) t = fmaxf(t, 0.f);
                        if (post) t = fmaf(t, es2[q], et2[q]);
                        y[q] = t;
                    }
                    if constexpr (OUT_PACKED)
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(iss_pack_split(y[0]), iss_pack_split(y[1]), iss_pack_split(y[2]), iss_pack_split(y[3]));
                    else
                        *reinterpret_cast<float4 *>(dst) = make_float4(y[0], y[1], y[2], y[3]);
                }
            };
            if (resid) { if (a.out_packed) rows(std::true_type{}, std::true_type{}); else rows(std::false_type{}, std::true_type{}); }
            else { if (a.out_packed) rows(std::true_type{}, std::false_type{}); else rows(std::false_type{}, std::false_type{}); }
            __syncwarp();
        }
        tc_fence_before();
    } else {
        // ============================ B loader + MMA issuer (warp 4, warp-uniform; see tc_issuer_warp) ============================
        // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (bits 7-9, 10-12 = 0), both K-major, N >> 3, M >> 4
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
        constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
        const uint32_t tb = __reduce_or_sync(0xffffffffu, tmem_base);
        const unsigned char *wt = h.wt + (size_t)blockIdx.y * nkb * Cfg::B_STAGE;
        auto issue_b = [&](int kb) {
            if (kb < nkb) {
                const int sl = kb % SB;
                if (elect_one()) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fullB[sl])), "r"((uint32_t)Cfg::B_STAGE) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(b_ring + sl * Cfg::B_STAGE)), "l"(wt + (size_t)kb * Cfg::B_STAGE),
                                   "r"((uint32_t)Cfg::B_STAGE), "r"(smem_u32(&fullB[sl])) : "memory");
                }
                __syncwarp();
            }
        };
        for (int p = 0; p < SB - 1; ++p) issue_b(p);
        const uint32_t d_main = tb, d_lo = tb + BN;
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % ST, sl = kb % SB;
            mbar_wait(&fullB[sl], (kb / SB) & 1, 2);
            mbar_wait(&fullA[st], (kb / ST) & 1, 3);
            tc_fence_after();
            const uint64_t dbh = make_sw128_desc(smem_u32(b_ring + sl * Cfg::B_STAGE));
            const uint32_t ta = tb + Cfg::ACC_COLS + st * Cfg::A_COLS;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < HBK / 16; ++kk) {          // K = 16 per kind::f16 MMA = 8 packed TMEM columns = 32 smem bytes
                    const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                    umma_f16_ts(d_main, ta + kk * 8, dbh + 2 * kk, idesc2, first);           // Ah.[Bh | Bl]
                    umma_f16_ts(d_lo, ta + 32 + kk * 8, dbh + 2 * kk, idesc, 1u);            // Al.Bh
                }
                umma_commit(&emptyA[st]);
                umma_commit(&emptyB[sl]);
                if (kb == nkb - 1) umma_commit(accum);
            }
            __syncwarp();
            if (kb + SB - 1 < nkb) {
                if (kb >= 1) mbar_wait(&emptyB[(kb - 1) % SB], ((kb - 1) / SB) & 1, 4);
                issue_b(kb + SB - 1);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

constexpr int SMEM_CTA_MAX = 232448;
constexpr int SMEM_HALF_SM = 115712;

template <int BN, int SB, int ST, int MODE>
int launch_tc3h_p(const ConvArgs &a, const F16Args &h, int slab_bytes, cudaStream_t st)
{
    using Cfg = TcHCfg<BN, SB, ST>;
    auto kern = conv_gemm_tc3h_kernel<BN, SB, ST, MODE>;
    ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), SMEM_CTA_MAX));
    const int64_t Q = a.M / a.OW;
    const int64_t gm = a.slab_tpi > 0 ? (Q / a.OH) * a.slab_tpi : (Q + a.slab_R - 1) / a.slab_R;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv_tc_f16: M too large");
    dim3 grid((unsigned)gm, (unsigned)(a.N / BN));
    kern<<<grid, Cfg::THREADS, Cfg::FIXED + slab_bytes, st>>>(a, h);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

template <int BN, int SB, int ST>
int launch_tc3h(const ConvArgs &a, const F16Args &h, int slab_bytes, cudaStream_t st)
{
    ISS_REQUIRE(!a.first, ISS_ERR_UNSUPPORTED, "conv_tc_f16: the fused first layer needs the direct kernel");
    return a.in_packed ? launch_tc3h_p<BN, SB, ST, IN_PACKED>(a, h, slab_bytes, st) : launch_tc3h_p<BN, SB, ST, IN_F32>(a, h, slab_bytes, st);
}

}  // namespace

// Tile plan: R output rows per tile and the rows the slab must hold.  Two tilings: tiles that may straddle an image
// boundary (every tile but the last is full; the slab needs KH - 1 extra rows per image it can touch) and tiles cut per
// image (slab = R + KH - 1 rows, the last tile of an image may be short).  With N = 64 (two CTAs per SM, 113 KB each)
// the smaller slab buys a third weight stage, which matters more than the few percent of idle GEMM rows: the weight
// stage of k-block kb+1 can only be requested when k-block kb-1 retires, and its L2 latency (~1000 cycles) exceeds one
// k-block of MMAs (448 cycles) -- with two stages every k-block waited for its weights.
// n-tile width: 128 output channels per CTA (one CTA per SM) when N allows it, else 64 (two CTAs per SM).
// ISS_B200_F16_BN=64 forces 64 everywhere (experiment: the slab fill / epilogue of one CTA then overlaps the other's main loop).
int iss_f16_bn_for(int N)
{
    static const int forced = [] { const char *e = getenv("ISS_B200_F16_BN"); return e ? atoi(e) : 0; }();
    if (forced == 64) return 64;
    return N % 128 == 0 ? 128 : 64;
}

static int slab_plan(const ConvArgs &a, int *R_out, int *rows_out, int *tpi_out)
{
    const int R = TBM / a.OW;
    const int cross = (R - 1) / a.OH + 1;
    int rows = R + (a.KH - 1) * (1 + cross);
    int tpi = 0;
    const int tiles_img = (a.OH + R - 1) / R;
    const double eff_straddle = (double)(R * a.OW) / TBM, eff_img = (double)(a.OH * a.OW) / ((double)tiles_img * TBM);
    static const bool tpi_off = [] { const char *e = getenv("ISS_B200_F16_TPI"); return e && e[0] == '0'; }();   // A/B experiments
    if (!tpi_off && iss_f16_bn_for(a.N) == 64 && R <= a.OH && eff_img >= 0.92 * eff_straddle) { tpi = tiles_img; rows = R + a.KH - 1; }
    int slab_bytes = rows * a.W * a.C * 4;
    if (slab_bytes < 32768) slab_bytes = 32768;                  // doubles as the 4 x 4 KB epilogue transpose buffers
    *R_out = R; *rows_out = rows; *tpi_out = tpi;
    return slab_bytes;
}

// Does this engine's slab kernel cover the layer?  (Also decides whether the layer in front may hand it
// split-half words: cnn.cu asks before choosing the producer's output format.)
bool iss_conv_f16_slab_covers(const ConvArgs &a)
{
    if (!a.wt_f16 || a.SH != 1 || a.SW != 1 || a.PT != 0 || a.PL != 0 || a.KH * a.KW <= 1) return false;
    if (a.OH != a.H - a.KH + 1 || a.OW != a.W - a.KW + 1 || a.OW > TBM || a.Kp != a.K) return false;
    if (a.N % 64 != 0 || a.C % HBK != 0 || a.K % HBK != 0) return false;
    int R, rows, tpi;
    const int slab_bytes = slab_plan(a, &R, &rows, &tpi);
    if (iss_f16_bn_for(a.N) == 128) return TcHCfg<128, 2, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX;
    return TcHCfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_CTA_MAX;
}

// Returns 1 when the layer is not covered (caller continues with the TF32 engine).
int iss_launch_conv_tc_f16(ConvArgs &a, cudaStream_t st)
{
    if (!iss_conv_f16_slab_covers(a)) return 1;
    int R, rows, tpi;
    const int slab_bytes = slab_plan(a, &R, &rows, &tpi);
    a.slab_R = R;
    a.slab_tpi = tpi;
    a.slab_rows = rows;
    a.in_elems = a.M / ((int64_t)a.OH * a.OW) * a.H * a.W * a.C;
    const int BN = iss_f16_bn_for(a.N);                              // same n-tiling as iss_prepare_f16_weights
    F16Args h{reinterpret_cast<const unsigned char *>(a.wt_f16), a.wt_f16_inv_scale};
    if (BN == 128) {
        if (TcHCfg<128, 4, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3h<128, 4, 4>(a, h, slab_bytes, st);
        if (TcHCfg<128, 3, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3h<128, 3, 4>(a, h, slab_bytes, st);
        if (TcHCfg<128, 2, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3h<128, 2, 4>(a, h, slab_bytes, st);
        return 1;
    }
    if (TcHCfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3h<64, 3, 2>(a, h, slab_bytes, st);
    if (TcHCfg<64, 2, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3h<64, 2, 2>(a, h, slab_bytes, st);
    if (TcHCfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3h<64, 3, 2>(a, h, slab_bytes, st);
    return 1;
}

// Host-side preparation of a layer's weights for this engine: W[K][N] (Keras / blob layout) -> device image
// [n-tile][k-block of 64][hi | lo][BN rows x 128 B, SWIZZLE_128B] of fp16 halves, pre-scaled by a power of two
// (undone in the epilogue with *inv_scale).  Owned by the layer (freed with cudaFree by its destructor).
int iss_prepare_f16_weights(const float *h_w, int K, int N, void **d_out, float *inv_scale)
{
    *d_out = nullptr; *inv_scale = 1.f;
    // K = 32 (a 1x1 convolution / Dense layer over 32 channels) is zero-padded to one 64-wide k-block: only the direct
    // kernel (conv_gemm_tc_f16d.cu) reads such an image
    // (likewise N = 32 is padded with zero rows to one 64-wide n-tile)
    const int Kp = K == 32 ? HBK : K;
    const int Np = N == 32 ? 64 : N;
    if (Kp % HBK != 0 || Np % 64 != 0) return ISS_OK;            // not a shape this engine takes
    std::vector<float> wt((size_t)Np * Kp, 0.f);                 // transposed [Np][Kp]
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) wt[(size_t)n * Kp + k] = h_w[(size_t)k * N + n];
    std::vector<__half> img;
    const float scale = iss_f16_build_image(wt.data(), Np, Kp, Kp, iss_f16_bn_for(Np), img);
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, img.size() * sizeof(__half));
    if (e != cudaSuccess) { iss_set_error("cudaMalloc f16 weights: %s", cudaGetErrorString(e)); return ISS_ERR_NOMEM; }
    e = cudaMemcpy(d, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); iss_set_error("cudaMemcpy f16 weights: %s", cudaGetErrorString(e)); return ISS_ERR_CUDA; }
    *d_out = d; *inv_scale = 1.f / scale;
    return ISS_OK;
}
