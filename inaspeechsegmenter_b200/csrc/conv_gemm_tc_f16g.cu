// conv_gemm_tc_f16g.cu -- engine 3, gather variant: the fp16-split tcgen05 pipeline of conv_gemm_tc_f16.cu for the
// convolutions its slab kernel does not take -- 1x1 convolutions (plain GEMMs over pixels), padded and strided
// KHxKW convolutions -- i.e. every ResNet101 layer of stages 2-4 (K5, 90 % of the x-vector FLOPs).
//
//   * A operand: cp.async im2col gathers of 16-byte chunks (zero-fill for padding / tail rows) into a warp-private
//     DA-deep ring; a k-block is 64 input channels of one filter tap = two 128-byte row segments; each lane then reads
//     ITS row (lane = GEMM row = TMEM lane), turns the 64 values into fp16 hi / lo operand columns -- PRMT only when
//     the input tensor already holds split-half words, conversions when it holds fp32 -- and writes them with
//     tcgen05.st into a ring in tensor memory next to the accumulators.
//   * B operand, MMA pattern (Ah.[Bh | Bl] then Al.Bh, kind::f16, fp32 accumulation in TMEM), epilogue (bias / BN affine /
//     residual / ReLU, fp32 or split-half output): as in conv_gemm_tc_f16.cu.  The residual may itself be a
//     split-half tensor (ConvArgs::residual_packed).
// Accuracy: the same 22-significant-bit split as the slab kernel (fp32-class; ResNet101 embedding error vs the real
// resnet.py measured in tests/test_vbx.py).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

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

__device__ __forceinline__ uint4 lds128u(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

template <int BN, int DA, int SB, int ST>
struct TcGCfg {
    static constexpr int THREADS = 160;
    static constexpr int A_SLOT = 8192;                                  // 2 k-halves x 32 rows x 128 B per warp
    static constexpr int A_RING = 4 * DA * A_SLOT;
    static constexpr int B_TILE = BN * 128;                              // BN rows x 64 halves
    static constexpr int B_STAGE = 2 * B_TILE;                           // hi | lo
    static constexpr int SMEM = A_RING + SB * B_STAGE + 1024 + 256;
    static constexpr uint32_t ACC_COLS = 2 * BN;                         // D_main | D_lo
    static constexpr uint32_t A_COLS = 64;                               // 32 packed columns hi + 32 lo per stage
    static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(ACC_COLS + ST * A_COLS);
};

struct F16GArgs {
    const unsigned char *wt;    // tiled fp16 image [n-tile][k-block][hi|lo][BN rows x 128 B, SWIZZLE_128B]
    float inv_scale;
};

template <int BN, int DA, int SB, int ST, bool PACKED>
__global__ void __launch_bounds__(160, (TcGCfg<BN, DA, SB, ST>::TMEM_COLS <= 256 ? 2 : 1))
conv_gemm_tc2h_kernel(const ConvArgs a, const F16GArgs h)
{
    using Cfg = TcGCfg<BN, DA, SB, ST>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *b_ring = smem;                                        // 1024-aligned operand tiles first
    unsigned char *a_ring = smem + SB * Cfg::B_STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_ring + Cfg::A_RING);
    uint64_t *fullA = bars, *emptyA = bars + ST, *emptyB = bars + 2 * ST, *fullB = bars + 2 * ST + SB, *accum = bars + 2 * ST + 2 * SB;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * ST + 2 * SB + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TBM;
    const int n0 = blockIdx.y * BN;
    const int nkb = a.K / HBK;

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
        // ============================ A producers ============================
        const int quad = warp & 3;                           // TMEM lane quadrant this warp may access
        const int sub = lane >> 3, chunk = lane & 7;
        uint32_t row_base[8];                                // element offset of (img, ih0, iw0) + chunk*4 (wraps for padded rows)
        int row_ih0[8], row_iw0[8];
        uint32_t row_dst[8];                                 // swizzled byte offset inside one k-half of a ring slot
        uint32_t ok_mask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + quad * 32 + 4 * i + sub;
            const bool okr = m < a.M;
            ok_mask |= (okr ? 1u : 0u) << i;
            const int64_t mm = okr ? m : 0;
            const int ohw = a.OH * a.OW;
            const int64_t img = mm / ohw;
            const int rem = (int)(mm - img * ohw);
            const int oh = rem / a.OW, ow = rem - oh * a.OW;
            row_ih0[i] = oh * a.SH - a.PT; row_iw0[i] = ow * a.SW - a.PL;
            row_base[i] = (uint32_t)(img * ((int64_t)a.H * a.W * a.C) + ((int64_t)row_ih0[i] * a.W + row_iw0[i]) * a.C + chunk * 4);
            const int rl = 4 * i + sub;
            row_dst[i] = (uint32_t)(rl * 128 + ((chunk ^ (rl & 7)) << 4));
        }
        const bool padded = (a.PT | a.PL) != 0 || (a.OH - 1) * a.SH + a.KH > a.H || (a.OW - 1) * a.SW + a.KW > a.W;
        unsigned char *my_ring = a_ring + quad * DA * Cfg::A_SLOT;
        const uint32_t ring_u32 = smem_u32(my_ring);
        // incremental tap state of the next k-block to gather (a k-block = 64 channels of one tap: C % 64 == 0)
        int is_c0 = 0, is_ss = 0, is_rr = 0, is_kb = 0, is_n = 0;
        uint32_t is_off = 0;                                 // (rr*W + ss)*C + c0
        const uint32_t wrap_step = (uint32_t)((a.W - a.KW) * a.C);
        auto issue_a = [&]() {
            if (is_kb < nkb) {
                const uint32_t slot = ring_u32 + (uint32_t)(is_n % DA) * Cfg::A_SLOT;
                uint32_t okm = ok_mask;
                if (padded) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ih = row_ih0[i] + is_rr, iw = row_iw0[i] + is_ss;
                        if (!(ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)) okm &= ~(1u << i);
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool ok = (okm >> i) & 1u;
                        const float *src = a.in + (ok ? (uint32_t)(row_base[i] + is_off + 32 * g) : 0u);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot + (uint32_t)(g * 4096) + row_dst[i]), "l"(src), "r"(ok ? 16 : 0) : "memory");
                    }
                ++is_n; ++is_kb;
                is_c0 += HBK; is_off += HBK;
                if (is_c0 == a.C) {
                    is_c0 = 0;
                    if (++is_ss == a.KW) { is_ss = 0; ++is_rr; is_off += wrap_step; }
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int p = 0; p < DA; ++p) issue_a();
        const uint32_t lane_addr = ((uint32_t)(quad * 32)) << 16;
        uint32_t lds_off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lds_off[j] = (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4));
        uint32_t hi[32], lo[32];                              // 64 k-elements, two halves per register
        auto load_split = [&](int n) {
            cp_async_wait<DA - 1>();
            __syncwarp();
            const uint32_t slot = ring_u32 + (uint32_t)(n % DA) * Cfg::A_SLOT;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 q = lds128u(slot + (uint32_t)(g * 4096) + lds_off[j]);
                    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if constexpr (PACKED) {                   // word = lo << 16 | hi; operand column = k even (low) | k odd (high)
                        hi[g * 16 + i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x5410);
                        lo[g * 16 + i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x7632);
                    } else {
                        const float f0 = __uint_as_float(v[2 * i]), f1 = __uint_as_float(v[2 * i + 1]);
                        const __half2 hh = __floats2half2_rn(f0, f1);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
                        hi[g * 16 + i] = *reinterpret_cast<const uint32_t *>(&hh);
                        lo[g * 16 + i] = *reinterpret_cast<const uint32_t *>(&ll);
                    }
                }
            }
            __syncwarp();                                    // slot fully read before it is refilled
            issue_a();
        };
        load_split(0);
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % ST;
            if (lane == 0) mbar_wait(&emptyA[st], ((kb / ST) & 1) ^ 1, 1);
            __syncwarp();
            tc_fence_after();
            const uint32_t ta = tmem_base + lane_addr + Cfg::ACC_COLS + st * Cfg::A_COLS;
            tmem_st32(ta, hi);
            tmem_st32(ta + 32, lo);
            if (kb + 1 < nkb) load_split(kb + 1);             // overlaps the TMEM store latency
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA[st]);
        }
        cp_async_wait<0>();

        // ============================ epilogue ============================
        if (lane == 0) mbar_wait(accum, 0, 5);
        __syncwarp();
        tc_fence_after();
        unsigned char *stage_buf = my_ring;                  // the A ring is idle now: reuse slot 0 as transpose buffer
        const bool has_bias = a.flags & ISS_F_BIAS, pre = a.flags & ISS_F_AFFINE_PRE, post = a.flags & ISS_F_AFFINE_POST;
        const bool relu = a.flags & ISS_F_RELU, resid = a.flags & ISS_F_RESIDUAL;
        const float inv_s = h.inv_scale;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t acc[32];
            {
                uint32_t corr[32];
                tmem_ld32(tmem_base + lane_addr + c, acc);
                tmem_ld32(tmem_base + lane_addr + BN + c, corr);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint((__uint_as_float(acc[j]) + __uint_as_float(corr[j])) * inv_s);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                    make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            __syncwarp();
            const int nb = n0 + c + chunk * 4;
            float eb[4], es1[4], et1[4], es2[4], et2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                eb[q] = has_bias ? __ldg(a.bias + nb + q) : 0.f;
                es1[q] = pre ? __ldg(a.pre_scale + nb + q) : 1.f;  et1[q] = pre ? __ldg(a.pre_shift + nb + q) : 0.f;
                es2[q] = post ? __ldg(a.post_scale + nb + q) : 1.f; et2[q] = post ? __ldg(a.post_shift + nb + q) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rl = 4 * i + sub;
                const int64_t m = m0 + quad * 32 + rl;
                const uint4 q4 = *reinterpret_cast<const uint4 *>(stage_buf + rl * 128 + ((chunk ^ (rl & 7)) << 4));
                if (m < a.M) {
                    float y[4] = {__uint_as_float(q4.x), __uint_as_float(q4.y), __uint_as_float(q4.z), __uint_as_float(q4.w)};
                    float rv[4] = {0.f, 0.f, 0.f, 0.f};
                    if (resid) {
                        const uint4 rw = __ldg(reinterpret_cast<const uint4 *>(a.residual + m * a.N + nb));
                        if (a.residual_packed) { rv[0] = iss_unpack_split(rw.x); rv[1] = iss_unpack_split(rw.y); rv[2] = iss_unpack_split(rw.z); rv[3] = iss_unpack_split(rw.w); }
                        else { rv[0] = __uint_as_float(rw.x); rv[1] = __uint_as_float(rw.y); rv[2] = __uint_as_float(rw.z); rv[3] = __uint_as_float(rw.w); }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float t = y[q] + eb[q];
                        if (pre) t = fmaf(t, es1[q], et1[q]);
                        if (resid) t += rv[q];
                        if (relu) t = fmaxf(t, 0.f);
                        if (post) t = fmaf(t, es2[q], et2[q]);
                        y[q] = t;
                    }
                    if (a.out_packed)
                        *reinterpret_cast<uint4 *>(a.out + m * a.N + nb) = make_uint4(iss_pack_split(y[0]), iss_pack_split(y[1]), iss_pack_split(y[2]), iss_pack_split(y[3]));
                    else
                        *reinterpret_cast<float4 *>(a.out + m * a.N + nb) = make_float4(y[0], y[1], y[2], y[3]);
                }
            }
        }
        tc_fence_before();
    } else {
        // ============================ B loader + MMA issuer (warp 4, warp-uniform control flow) ============================
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

template <int BN, int DA, int SB, int ST>
int launch_tc2h(const ConvArgs &a, const F16GArgs &h, cudaStream_t st)
{
    using Cfg = TcGCfg<BN, DA, SB, ST>;
    const int64_t gm = (a.M + TBM - 1) / TBM;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv_tc_f16g: M too large");
    dim3 grid((unsigned)gm, (unsigned)(a.N / BN));
    if (a.in_packed) {
        auto kern = conv_gemm_tc2h_kernel<BN, DA, SB, ST, true>;
        ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), Cfg::SMEM));
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a, h);
    } else {
        auto kern = conv_gemm_tc2h_kernel<BN, DA, SB, ST, false>;
        ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), Cfg::SMEM));
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a, h);
    }
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

}  // namespace

// Does the gather variant of engine 3 cover this layer?  (Any stride / padding / filter size; needs 64-channel k-blocks,
// n-tiles of 64 and a prepared fp16 weight image.)
bool iss_conv_f16_gather_covers(const ConvArgs &a)
{
    if (!a.wt_f16 || a.Kp != a.K) return false;
    if (a.N % 64 != 0 || a.C % HBK != 0 || a.K % HBK != 0) return false;
    return (int64_t)a.H * a.W * a.C * (a.M / ((int64_t)a.OH * a.OW) + 1) < (1ll << 32);       // 32-bit element offsets in the gathers
}

// Returns 1 when the layer is not covered (caller continues with the TF32 engine).
int iss_launch_conv_tc_f16g(const ConvArgs &a, cudaStream_t st)
{
    if (!iss_conv_f16_gather_covers(a)) return 1;
    F16GArgs h{reinterpret_cast<const unsigned char *>(a.wt_f16), a.wt_f16_inv_scale};
    // n-tile width must match iss_prepare_f16_weights (128 when N allows it, else 64)
    if (iss_f16_bn_for(a.N) == 128) return launch_tc2h<128, 3, 3, 4>(a, h, st);   // 193 KB smem, 512 TMEM cols, 1 CTA/SM
    return launch_tc2h<64, 2, 2, 2>(a, h, st);                                      //  98 KB smem, 256 TMEM cols, 2 CTAs/SM
}
