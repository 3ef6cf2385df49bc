// conv_gemm_tc.cu -- implicit-GEMM convolution / dense layer on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM) with fp32-class accuracy ("3xTF32"), for every layer
// shape: padded / strided / 1x1 convolutions (ResNet101) and Dense layers.  The un-padded stride-1
// KHxKW convolutions of the segmenter CNNs take the faster fp16-split slab kernel of
// conv_gemm_tc_f16.cu (engine 3, the default) and only fall back to this kernel when it does not
// cover them.
//
// Why 3xTF32: the reference evaluates its CNNs in fp32 (TF-CPU) and the parity bar is
// 1e-4 on the per-frame softmax; a single TF32 pass (10-bit mantissa) misses it.  Every
// fp32 operand x is split exactly into hi = x with the 13 low mantissa bits cleared (a
// valid TF32 number) and lo = x - hi; the product is accumulated as
//      A.B ~= Ah.Bh + Ah.Bl + Al.Bh          (dropped Al.Bl ~ 2^-22 relative)
// Tile: 128 output positions (UMMA_M = 128, cta_group::1) x BN output channels, K in blocks of 32 floats
// (= one 128-byte swizzle row); 4 producer/epilogue warps + 1 TMEM-allocator / B-loader / MMA-issuer warp.
// A operand: cp.async im2col gathers -> warp-private ring -> registers -> hi/lo split -> TMEM (tcgen05.st);
// weights: one bulk copy per stage of host-pre-swizzled tiles; two accumulators (main + correction).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "conv_gemm.cuh"

#include "tc_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// B loader + MMA issuer, run by ALL 32 lanes of warp 4 with warp-uniform control flow; only the
// tcgen05 / bulk-copy instructions themselves sit under elect.sync.  This matters more than anything
// else in the kernel: with the loop inside `if (lane == 0)` ptxas cannot prove the tcgen05 operands
// (TMEM addresses, smem descriptors) warp-uniform and wraps EVERY tcgen05.mma in an
// ELECT / 4 x R2UR.BROADCAST / BRA.U.ANY "waterfall" -- ~50 issue cycles per MMA, ~330 instructions per
// k-block on one thread, i.e. 600-1100 cycles per k-block against a tensor-time floor of 392 (measured with
// what-if builds in round 1, profiles/r01_tc_whatif.txt).  With uniform control flow and the TMEM base
// passed through REDUX (a uniform-register producer) the 8 MMAs of a k-block are 8 back-to-back UTCHMMA.
// The weights are static, so the host stores them already tiled and swizzled exactly as the smem operand
// image ([n-tile][k-block][hi|lo][BN x 128 B, SWIZZLE_128B]); one stage is a single 1-D bulk copy.
template <int BN, int SB, int ST, int B_STAGE, uint32_t ACC_COLS>
__device__ __forceinline__ void tc_issuer_warp(const ConvArgs &a, uint32_t tmem_base_any, unsigned char *b_ring, uint64_t *fullA,
                                               uint64_t *emptyA, uint64_t *fullB, uint64_t *emptyB, uint64_t *accum, int nkb)
{
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, tmem_base_any);
    const unsigned char *wt = reinterpret_cast<const unsigned char *>(a.wt_tiled) + (size_t)blockIdx.y * nkb * B_STAGE;
    auto issue_b = [&](int kb) {
        if (kb < nkb) {
            const int sl = kb % SB;
            if (elect_one()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fullB[sl])), "r"((uint32_t)B_STAGE) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(b_ring + sl * B_STAGE)), "l"(wt + (size_t)kb * B_STAGE),
                               "r"((uint32_t)B_STAGE), "r"(smem_u32(&fullB[sl])) : "memory");
            }
            __syncwarp();
        }
    };
    for (int p = 0; p < SB - 1; ++p) issue_b(p);
    const uint32_t d_main = tmem_base, d_lo = tmem_base + BN;
    for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % ST, sl = kb % SB;
        mbar_wait(&fullB[sl], (kb / SB) & 1, 2);
        mbar_wait(&fullA[st], (kb / ST) & 1, 3);
        tc_fence_after();
        const uint64_t dbh = make_sw128_desc(smem_u32(b_ring + sl * B_STAGE));
        const uint32_t ta = tmem_base + ACC_COLS + st * 2 * TBK;
        if (elect_one()) {
            // the B stage is [hi rows | lo rows] and the accumulators are [D_main | D_lo] in adjacent TMEM
            // columns, hence Ah.Bh and Ah.Bl are ONE MMA of width 2*BN (D_main += Ah.Bh, D_lo += Ah.Bl);
            // Al.Bh follows into D_lo.
#pragma unroll
            for (int kk = 0; kk < TBK / 8; ++kk) {
                const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                umma_tf32_ts(d_main, ta + kk * 8, dbh + 2 * kk, idesc2, first);          // Ah.[Bh | Bl]
                umma_tf32_ts(d_lo, ta + TBK + kk * 8, dbh + 2 * kk, idesc, 1u);          // Al.Bh
            }
            umma_commit(&emptyA[st]);
            umma_commit(&emptyB[sl]);
            if (kb == nkb - 1) umma_commit(accum);
        }
        __syncwarp();
        // refill the slot of k-block kb-1 with kb+SB-1: its MMAs retire before those just issued start,
        // so this wait is short and the tensor pipe is never drained
        if (kb + SB - 1 < nkb) {
            if (kb >= 1) mbar_wait(&emptyB[(kb - 1) % SB], ((kb - 1) / SB) & 1, 4);   // slot (kb-1)%SB; unused so far when kb == 0
            issue_b(kb + SB - 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Asynchronous TS pipeline (the production path).  Differences from the kernel above:
//   * A rows are fetched with cp.async (16-byte, L2 -> smem, zero-fill for padding / tail rows)
//     into a warp-private DA-deep ring, so DA k-blocks of gathers are in flight per warp with no
//     register staging; the ring doubles as the transpose buffer (lane = row on the way out).
//   * the weight tiles are fetched by the MMA warp itself (32 lanes of cp.async into an SB-deep
//     ring), which removes them from the producers' critical path;
//   * operand A goes registers -> TMEM (tcgen05.st) into an ST-deep ring next to the two
//     accumulators, so shared memory only carries B for the MMAs.
template <int BN, int DA, int SB, int ST>
struct Tc2Cfg {
    static constexpr int A_SLOT = 4096;                                  // 32 rows x 128 B per warp
    static constexpr int A_RING = 4 * DA * A_SLOT;
    static constexpr int THREADS = 160;
    static constexpr int B_TILE = BN * TBK * 4;
    static constexpr int B_STAGE = 2 * B_TILE;                           // hi | lo
    static constexpr int SMEM = A_RING + SB * B_STAGE + 1024 + 256;
    static constexpr uint32_t ACC_COLS = 2 * BN;
    static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(ACC_COLS + ST * 2 * TBK);
};

template <int BN, int DA, int SB, int ST>
__global__ void __launch_bounds__(160, (Tc2Cfg<BN, DA, SB, ST>::TMEM_COLS <= 256 ? 2 : 1))
conv_gemm_tc2_kernel(const ConvArgs a)
{
    using Cfg = Tc2Cfg<BN, DA, SB, ST>;
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
    const int nkb = a.Kp / TBK;

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
        // Issue-slot budget matters here (the kernel is instruction-issue bound before it is tensor
        // bound): per-row state is 32-bit and precomputed, the filter tap advances incrementally
        // (no integer division in the loop) and un-padded convolutions skip all bounds checks.
        const int sub = lane >> 3, chunk = lane & 7;
        uint32_t row_base[8];                                // element offset of (img, ih0, iw0) + chunk*4 (wraps for padded rows)
        int row_ih0[8], row_iw0[8];
        uint32_t row_dst[8];                                 // swizzled byte offset inside a ring slot
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
        // incremental tap state of the next k-block to gather
        int is_c0 = 0, is_ss = 0, is_rr = 0, is_kb = 0, is_n = 0;
        uint32_t is_off = 0;                                 // (rr*W + ss)*C + c0
        const uint32_t wrap_step = (uint32_t)((a.W - a.KW) * a.C + TBK);
        auto advance = [&]() {
            ++is_kb;
            is_c0 += TBK; is_off += TBK;
            if (is_c0 == a.C) {
                is_c0 = 0;
                if (++is_ss == a.KW) { is_ss = 0; ++is_rr; is_off += wrap_step - TBK; }
            }
        };
        auto issue_a = [&]() {
            if (is_kb < nkb) {
                const uint32_t slot = ring_u32 + (uint32_t)(is_n % DA) * Cfg::A_SLOT;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool ok = (ok_mask >> i) & 1u;
                    if (padded) {
                        const int ih = row_ih0[i] + is_rr, iw = row_iw0[i] + is_ss;
                        ok = ok && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
                    }
                    const float *src = a.in + (ok ? (uint32_t)(row_base[i] + is_off) : 0u);
                    // .cg (L2 only); L1-allocating gathers (.ca) measured no different
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot + row_dst[i]), "l"(src), "r"(ok ? 16 : 0) : "memory");
                }
                ++is_n;
                advance();
            }
            cp_async_commit();
        };
#pragma unroll
        for (int p = 0; p < DA; ++p) issue_a();
        const uint32_t lane_addr = ((uint32_t)(quad * 32)) << 16;
        uint32_t lds_off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lds_off[j] = (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4));
        // Software pipeline: the TMEM stores of k-block kb are issued, then the ring slot of kb+1 is
        // read and split while they drain; only then tcgen05.wait::st + arrive.  The per-warp critical
        // path per k-block is max(store latency, load+split) instead of their sum.
        uint32_t hi[32], lo[32];
        auto load_split = [&](int n) {                        // n-th k-block of this set
            cp_async_wait<DA - 1>();
            __syncwarp();
            const unsigned char *slot = my_ring + (n % DA) * Cfg::A_SLOT;
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 q = *reinterpret_cast<const uint4 *>(slot + lds_off[j]);
                v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
            __syncwarp();                                    // slot fully read before it is refilled
            issue_a();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                hi[j] = v[j] & 0xFFFFE000u;
                lo[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(hi[j]));
            }
        };
        load_split(0);
        int nloc = 0;
        for (int kb = 0; kb < nkb; ++kb, ++nloc) {
            const int st = kb % ST;
            if (lane == 0) mbar_wait(&emptyA[st], ((kb / ST) & 1) ^ 1, 1);   // one poller / one arrival per warp:
            __syncwarp();                                                     // 128 threads hammering the mbarriers cost more than the MMAs
            tc_fence_after();
            const uint32_t ta = tmem_base + lane_addr + Cfg::ACC_COLS + st * 2 * TBK;
            tmem_st32(ta, hi);
            tmem_st32(ta + TBK, lo);
            if (kb + 1 < nkb) load_split(nloc + 1);       // overlaps the TMEM store latency
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
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t acc[32];
            {
                uint32_t corr[32];
                tmem_ld32(tmem_base + lane_addr + c, acc);
                tmem_ld32(tmem_base + lane_addr + BN + c, corr);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(corr[j]));
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
                    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (resid) rs = __ldg(reinterpret_cast<const float4 *>(a.residual + m * a.N + nb));
                    const float rv[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float t = y[q] + eb[q];
                        if (pre) t = fmaf(t, es1[q], et1[q]);
                        if (resid) t += rv[q];
                        if (relu) t = fmaxf(t, 0.f);
                        if (post) t = fmaf(t, es2[q], et2[q]);
                        y[q] = t;
                    }
                    *reinterpret_cast<float4 *>(a.out + m * a.N + nb) = make_float4(y[0], y[1], y[2], y[3]);
                }
            }
        }
        tc_fence_before();
    } else {
        // ============================ B loader + MMA issuer (warp 4) ============================
        tc_issuer_warp<BN, SB, ST, Cfg::B_STAGE, Cfg::ACC_COLS>(a, tmem_base, b_ring, fullA, emptyA, fullB, emptyB, accum, nkb);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
}


constexpr int SMEM_CTA_MAX = 232448;       // 227 KB opt-in limit per CTA
constexpr int SMEM_HALF_SM = 115712;       // two CTAs per SM: 2 * (x + 1 KB reserved) <= 228 KB

template <int BN, int DA, int SB, int ST>
int launch_tc2(const ConvArgs &a, cudaStream_t st)
{
    using Cfg = Tc2Cfg<BN, DA, SB, ST>;
    auto kern = conv_gemm_tc2_kernel<BN, DA, SB, ST>;
    ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), Cfg::SMEM));
    const int64_t gm = (a.M + TBM - 1) / TBM;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv_tc: M too large");
    dim3 grid((unsigned)gm, (unsigned)(a.N / BN));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

int g_gemm_mode = -1;      // -1 = read ISS_B200_GEMM on first use

}  // namespace

extern "C" int iss_set_gemm_mode(int mode)
{
    ISS_REQUIRE(mode == ISS_GEMM_FP32 || mode == ISS_GEMM_TC_TS || mode == ISS_GEMM_TC_F16, ISS_ERR_INVALID, "iss_set_gemm_mode: %d (0 = fp32 cores, 2 = 3xTF32, 3 = fp16 split)", mode);
    g_gemm_mode = mode;
    return ISS_OK;
}

extern "C" int iss_get_gemm_mode(void)
{
    if (g_gemm_mode < 0) {
        const char *e = getenv("ISS_B200_GEMM");
        g_gemm_mode = ISS_GEMM_DEFAULT;
        if (e && !strcmp(e, "fp32")) g_gemm_mode = ISS_GEMM_FP32;
        else if (e && !strcmp(e, "tc_ts")) g_gemm_mode = ISS_GEMM_TC_TS;
        else if (e && !strcmp(e, "tc_f16")) g_gemm_mode = ISS_GEMM_TC_F16;
    }
    return g_gemm_mode;
}

int iss_launch_conv_tc_f16(ConvArgs &a, cudaStream_t st);              // conv_gemm_tc_f16.cu; 1 = layer not covered
int iss_launch_conv_tc_f16g(const ConvArgs &a, cudaStream_t st);       // conv_gemm_tc_f16g.cu; 1 = layer not covered
int iss_launch_conv_tc_f16d(ConvArgs &a, cudaStream_t st);              // conv_gemm_tc_f16d.cu; 1 = layer not covered

bool iss_conv_tc_eligible(const ConvArgs &a)
{
    return a.wt_hi && a.wt_lo && a.wt_tiled && a.Kp > 0 && a.C % 32 == 0 && a.K % 32 == 0 && a.N % 32 == 0 && a.N >= 32;
}

int iss_launch_conv_tc(const ConvArgs &a_in, int mode, cudaStream_t st)
{
    ConvArgs a = a_in;
    if (mode == ISS_GEMM_TC_F16) {                                      // fp16-split kernels where they apply: direct, slab, then gather
        int rc = iss_launch_conv_tc_f16d(a, st);
        if (rc != 1) return rc;
        rc = iss_launch_conv_tc_f16(a, st);
        if (rc != 1) return rc;
        static const bool gather_off = [] { const char *e = getenv("ISS_B200_F16_GATHER"); return e && e[0] == '0'; }();   // A/B experiments
        if (!gather_off || a.in_packed || a.out_packed || a.residual_packed) {
            rc = iss_launch_conv_tc_f16g(a, st);
            if (rc != 1) return rc;
        }
    }
    ISS_REQUIRE(!a.in_packed && !a.out_packed && !a.residual_packed, ISS_ERR_UNSUPPORTED, "conv_tc: split-half tensors need the fp16-split engine");
    // two accumulators per tile (main + correction) => BN <= 128 (2 x 128 + A ring <= 512 TMEM columns)
    if (a.N % 128 == 0) return launch_tc2<128, 4, 4, 4>(a, st);          // 193 KB smem, 512 TMEM cols (2x128 acc + 4 A stages), 1 CTA/SM
    if (a.N % 64 == 0) return launch_tc2<64, 3, 3, 2>(a, st);            //  97 KB smem, 256 TMEM cols, 2 CTAs/SM
    return launch_tc2<32, 3, 4, 3>(a, st);                                //  81 KB smem, 256 TMEM cols
}

static int tc_block_n(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32); }

// Host-side preparation of a layer's weights for the tensor-core path:
// W[K][N] (Keras / our blob layout) ->
//   (1) transposed, zero-padded, split row-major [2][N][Kp]           (SS kernel), followed by
//   (2) the same values tiled and pre-swizzled as smem operand images
//       [N/BN][Kp/32][hi|lo][BN rows x 128 B, 16-byte chunk j of row n stored at chunk j ^ (n & 7)]
//       (TS kernel: one bulk copy per stage).
int iss_prepare_tc_weights(const float *h_w, int K, int N, float **d_out, int *Kp_out)
{
    const int Kp = (K + TBK - 1) / TBK * TBK;
    const size_t plane = (size_t)N * Kp;
    std::vector<float> buf(4 * plane, 0.f);
    float *hi = buf.data(), *lo = buf.data() + plane, *tiled = buf.data() + 2 * plane;
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            const float w = h_w[(size_t)k * N + n];
            uint32_t u;
            memcpy(&u, &w, 4);
            u &= 0xFFFFE000u;
            float h;
            memcpy(&h, &u, 4);
            hi[(size_t)n * Kp + k] = h;
            lo[(size_t)n * Kp + k] = w - h;
        }
    const int BN = tc_block_n(N), nkb = Kp / TBK;
    for (int nt = 0; nt < N / BN; ++nt)
        for (int kb = 0; kb < nkb; ++kb)
            for (int part = 0; part < 2; ++part) {
                float *dst = tiled + (((size_t)nt * nkb + kb) * 2 + part) * (size_t)BN * TBK;
                const float *src = part ? lo : hi;
                for (int n = 0; n < BN; ++n)
                    for (int k = 0; k < TBK; ++k) {
                        const int chunk = (k >> 2) ^ (n & 7);
                        dst[n * TBK + chunk * 4 + (k & 3)] = src[(size_t)(nt * BN + n) * Kp + kb * TBK + k];
                    }
            }
    float *d = nullptr;
    cudaError_t e = cudaMalloc(&d, buf.size() * sizeof(float));
    if (e != cudaSuccess) { iss_set_error("cudaMalloc tc weights: %s", cudaGetErrorString(e)); return ISS_ERR_NOMEM; }
    e = cudaMemcpy(d, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); iss_set_error("cudaMemcpy tc weights: %s", cudaGetErrorString(e)); return ISS_ERR_CUDA; }
    *d_out = d;
    *Kp_out = Kp;
    return ISS_OK;
}
