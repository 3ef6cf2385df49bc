// conv_gemm_tc.cu -- implicit-GEMM convolution / dense layer on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM) with fp32-class accuracy ("3xTF32").
//
// Why 3xTF32: the reference evaluates its CNNs in fp32 (TF-CPU) and the parity bar is
// 1e-4 on the per-frame softmax; a single TF32 pass (10-bit mantissa) misses it.  Every
// fp32 operand x is split exactly into hi = x with the 13 low mantissa bits cleared (a
// valid TF32 number) and lo = x - hi; the product is accumulated as
//      A.B ~= Ah.Bh + Ah.Bl + Al.Bh          (dropped Al.Bl ~ 2^-22 relative)
// i.e. three kind::tf32 MMAs per K-step into the same fp32 TMEM accumulator.
//
// Two kernels live here:
//   conv_gemm_tc2_kernel  (engine 2, the default): A operand registers -> TMEM (tcgen05.st), weights as
//       one bulk copy per stage of host-pre-swizzled tiles, two accumulators (main + correction),
//       asynchronous cp.async gather ring.  See the block comment above that kernel and DESIGN.md 4.1.
//   conv_gemm_tc_kernel   (engine 1, kept as the all-shared-memory comparison point): both operands in
//       the canonical K-major SWIZZLE_128B smem layout, producers split hi/lo into two A tiles.
// Tile: 128 output positions (UMMA_M = 128, cta_group::1) x BN output channels, K in blocks of 32 floats
// (= one 128-byte swizzle row); 4 producer/epilogue warps + 1 TMEM-allocator / MMA-issuer warp.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "conv_gemm.cuh"

#include "tc_common.cuh"

namespace {

template <int BN, int STAGES, bool A_TMEM>
struct TcCfg {
    static constexpr int A_TILE = A_TMEM ? 0 : TBM * TBK * 4;           // bytes per hi (or lo) A tile in smem
    static constexpr int B_TILE = BN * TBK * 4;
    static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/ + 4 * 4096 /*per-warp transpose*/;
    static constexpr uint32_t ACC_COLS = 2 * BN;                        // D_main | D_lo (correction terms)
    static constexpr uint32_t TMEM_USED = ACC_COLS + (A_TMEM ? STAGES * 2 * TBK : 0);
    static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(TMEM_USED);
};

template <int BN, int STAGES, bool A_TMEM>
__global__ void __launch_bounds__(160, (BN <= 64 ? 2 : 1))
conv_gemm_tc_kernel(const ConvArgs a)
{
    using Cfg = TcCfg<BN, STAGES, A_TMEM>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * Cfg::STAGE);
    uint64_t *full = bars, *empty = bars + STAGES, *accum = bars + 2 * STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TBM;
    const int n0 = blockIdx.y * BN;
    const int nkb = a.Kp / TBK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
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

    if (warp < 4) {
        // ============================ producers ============================
        // Coalesced gather: for load i (0..7) lane l fetches 16-byte chunk (l & 7) of tile row
        // 32*warp + 4*i + (l >> 3), so one warp instruction reads 4 full 128-byte segments
        // (4 L1 wavefronts instead of 32).  Rows are then regrouped per lane through a
        // warp-private swizzled 4 KB staging buffer (TS mode) or written straight to the
        // swizzled operand tile (SS mode).
        const int sub = lane >> 3, chunk = lane & 7;
        int64_t row_off[8];                                  // element offset of the (img, ih0, iw0) origin
        int row_ih0[8], row_iw0[8];
        bool row_ok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + warp * 32 + 4 * i + sub;
            row_ok[i] = m < a.M;
            const int64_t mm = row_ok[i] ? m : 0;
            const int ohw = a.OH * a.OW;
            const int64_t img = mm / ohw;
            const int rem = (int)(mm - img * ohw);
            const int oh = rem / a.OW, ow = rem - oh * a.OW;
            row_ih0[i] = oh * a.SH - a.PT; row_iw0[i] = ow * a.SW - a.PL;
            row_off[i] = img * ((int64_t)a.H * a.W * a.C) + ((int64_t)row_ih0[i] * a.W + row_iw0[i]) * a.C;
        }
        unsigned char *stage_buf = smem + STAGES * Cfg::STAGE + 256 + warp * 4096;   // warp-private transpose buffer
        const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t use = kb / STAGES;
            const int k0 = kb * TBK;
            const int tap = k0 / a.C, c0 = k0 - tap * a.C;
            const int rr = tap / a.KW, ss = tap - rr * a.KW;
            const int64_t tap_off = ((int64_t)rr * a.W + ss) * a.C + c0 + chunk * 4;
            uint4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ih = row_ih0[i] + rr, iw = row_iw0[i] + ss;
                const bool ok = row_ok[i] && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
                x[i] = make_uint4(0u, 0u, 0u, 0u);
                if (ok) x[i] = __ldg(reinterpret_cast<const uint4 *>(a.in + row_off[i] + tap_off));
            }
            mbar_wait(&empty[s], (use & 1) ^ 1);             // fresh barrier: passes immediately
            unsigned char *st = smem + s * Cfg::STAGE;
            if (A_TMEM) {
                tc_fence_after();
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = 4 * i + sub;
                    *reinterpret_cast<uint4 *>(stage_buf + rl * 128 + ((chunk ^ (rl & 7)) << 4)) = x[i];
                }
                __syncwarp();
                uint32_t v[32], hi[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4));
                    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) hi[j] = v[j] & 0xFFFFE000u;
                const uint32_t ta = tmem_base + lane_addr + Cfg::ACC_COLS + s * 2 * TBK;
                tmem_st32(ta, hi);
#pragma unroll
                for (int j = 0; j < 32; ++j) hi[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(hi[j]));
                tmem_st32(ta + TBK, hi);
            } else {
                unsigned char *ah = st, *al = st + Cfg::A_TILE;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = warp * 32 + 4 * i + sub;
                    uint4 h, l;
                    h.x = x[i].x & 0xFFFFE000u; h.y = x[i].y & 0xFFFFE000u; h.z = x[i].z & 0xFFFFE000u; h.w = x[i].w & 0xFFFFE000u;
                    l.x = __float_as_uint(__uint_as_float(x[i].x) - __uint_as_float(h.x));
                    l.y = __float_as_uint(__uint_as_float(x[i].y) - __uint_as_float(h.y));
                    l.z = __float_as_uint(__uint_as_float(x[i].z) - __uint_as_float(h.z));
                    l.w = __float_as_uint(__uint_as_float(x[i].w) - __uint_as_float(h.w));
                    const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4 *>(ah + off) = h;
                    *reinterpret_cast<uint4 *>(al + off) = l;
                }
            }
            // ---- B: pre-split transposed weights [N][Kp] -> [BN][32] swizzled tiles ----
            {
                unsigned char *bh = st + 2 * Cfg::A_TILE, *bl = bh + Cfg::B_TILE;
#pragma unroll
                for (int i = 0; i < (BN * 8) / 128; ++i) {
                    const int idx = tid + i * 128;
                    const int n = idx >> 3, j = idx & 7;
                    const int64_t g = (int64_t)(n0 + n) * a.Kp + k0 + 4 * j;
                    const uint4 h = __ldg(reinterpret_cast<const uint4 *>(a.wt_hi + g));
                    const uint4 l = __ldg(reinterpret_cast<const uint4 *>(a.wt_lo + g));
                    const int off = n * 128 + ((j ^ (n & 7)) << 4);
                    *reinterpret_cast<uint4 *>(bh + off) = h;
                    *reinterpret_cast<uint4 *>(bl + off) = l;
                }
            }
            fence_proxy_async();                              // generic-proxy smem writes -> async proxy (UMMA)
            if (A_TMEM) { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); tc_fence_before(); }
            mbar_arrive(&full[s]);
        }

        // ============================ epilogue ============================
        // TMEM quadrant -> registers (lane = row) -> swizzled staging -> coalesced rows:
        // lane l then owns columns 4*(l&7)..+3 of rows 4*i + (l>>3), so bias / BN vectors are
        // per-lane constants and every store instruction writes four full 128-byte segments.
        mbar_wait(accum, 0);
        tc_fence_after();
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
                const int64_t m = m0 + warp * 32 + rl;
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
        // ============================ MMA issuer ============================
        // instruction descriptor: D = f32, A = B = tf32, K-major both, N = BN, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t use = kb / STAGES;
            mbar_wait(&full[s], use & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_u32(smem + s * Cfg::STAGE);
                const uint64_t dbh = make_sw128_desc(st + 2 * Cfg::A_TILE);
                const uint64_t dbl = make_sw128_desc(st + 2 * Cfg::A_TILE + Cfg::B_TILE);
#pragma unroll
                for (int kk = 0; kk < TBK / 8; ++kk) {
                    const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                    // The tensor core's accumulate-add truncates, so every MMA into an accumulator
                    // costs ~half an ulp of bias relative to that accumulator's magnitude: the two
                    // small correction products go to their own accumulator (D_lo) and are added
                    // to D_main with a correctly rounded fp32 add in the epilogue.
                    const uint32_t d_main = tmem_base, d_lo = tmem_base + BN;
                    if (A_TMEM) {
                        const uint32_t ta = tmem_base + Cfg::ACC_COLS + s * 2 * TBK + kk * 8;
                        umma_tf32_ts(d_main, ta, dbh + 2 * kk, idesc, first);                 // Ah.Bh
                        umma_tf32_ts(d_lo, ta, dbl + 2 * kk, idesc, first);                   // Ah.Bl
                        umma_tf32_ts(d_lo, ta + TBK, dbh + 2 * kk, idesc, 1u);                // Al.Bh
                    } else {
                        const uint64_t dah = make_sw128_desc(st), dal = make_sw128_desc(st + Cfg::A_TILE);
                        umma_tf32_ss(d_main, dah + 2 * kk, dbh + 2 * kk, idesc, first);
                        umma_tf32_ss(d_lo, dah + 2 * kk, dbl + 2 * kk, idesc, first);
                        umma_tf32_ss(d_lo, dal + 2 * kk, dbh + 2 * kk, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);                        // stage reusable once these MMAs have read it
                if (kb == nkb - 1) umma_commit(accum);         // accumulator complete
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
}



// ---------------------------------------------------------------------------------------------
// B loader + MMA issuer, run by ALL 32 lanes of warp 4 with warp-uniform control flow; only the
// tcgen05 / bulk-copy instructions themselves sit under elect.sync.  This matters more than anything
// else in the kernel: with the loop inside `if (lane == 0)` ptxas cannot prove the tcgen05 operands
// (TMEM addresses, smem descriptors) warp-uniform and wraps EVERY tcgen05.mma in an
// ELECT / 4 x R2UR.BROADCAST / BRA.U.ANY "waterfall" -- ~50 issue cycles per MMA, ~330 instructions per
// k-block on one thread, i.e. 600-1100 cycles per k-block against a tensor-time floor of 392 (measured with
// the what-if switches of ISS_B200_TC_DEBUG: removing all MMAs saved 12 %, removing everything but
// the issue loop skeleton still cost 60 % of the run time).  With uniform control flow and the TMEM base
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
    const bool dbg_nob = a.debug_same_addr & 8, dbg_nomma = a.debug_same_addr & 16, dbg_nocommit_b = a.debug_same_addr & 32;   // timing experiments
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
    for (int p = 0; p < SB - 1 + (dbg_nob ? 1 : 0); ++p) issue_b(p);
    const uint32_t d_main = tmem_base, d_lo = tmem_base + BN;
    const bool prof = a.prof != nullptr;
    long long w_b = 0, w_a = 0, w_e = 0;
    const long long t_i0 = prof ? clock64() : 0;
    for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % ST, sl = kb % SB;
        const long long c0 = prof ? clock64() : 0;
        if (!dbg_nob || kb < SB) mbar_wait(&fullB[sl], (kb / SB) & 1, 2);
        const long long c1 = prof ? clock64() : 0;
        mbar_wait(&fullA[st], (kb / ST) & 1, 3);
        if (prof) { w_b += c1 - c0; w_a += clock64() - c1; }
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
                if (dbg_nomma && first) continue;
                umma_tf32_ts(d_main, ta + kk * 8, dbh + 2 * kk, idesc2, first);          // Ah.[Bh | Bl]
                umma_tf32_ts(d_lo, ta + TBK + kk * 8, dbh + 2 * kk, idesc, 1u);          // Al.Bh
            }
            umma_commit(&emptyA[st]);
            if (!dbg_nocommit_b) umma_commit(&emptyB[sl]);
            if (kb == nkb - 1) umma_commit(accum);
        }
        __syncwarp();
        // refill the slot of k-block kb-1 with kb+SB-1: its MMAs retire before those just issued start,
        // so this wait is short and the tensor pipe is never drained
        if (kb + SB - 1 < nkb && !dbg_nob) {
            const long long c2 = prof ? clock64() : 0;
            if (kb >= 1) mbar_wait(&emptyB[(kb - 1) % SB], ((kb - 1) / SB) & 1, 4);   // slot (kb-1)%SB; unused so far when kb == 0
            if (prof) w_e += clock64() - c2;
            issue_b(kb + SB - 1);
        }
    }
    if (prof && (threadIdx.x & 31) == 0) {
        atomicAdd(a.prof + 0, (unsigned long long)(clock64() - t_i0)); atomicAdd(a.prof + 1, (unsigned long long)w_b);
        atomicAdd(a.prof + 2, (unsigned long long)w_a); atomicAdd(a.prof + 3, (unsigned long long)w_e);
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
template <int BN, int DA, int SB, int ST, int NSETS = 1>
struct Tc2Cfg {
    static constexpr int A_SLOT = 4096;                                  // 32 rows x 128 B per warp
    static constexpr int A_RING = 4 * NSETS * DA * A_SLOT;
    static constexpr int THREADS = 32 * (4 * NSETS + 1);
    static constexpr int B_TILE = BN * TBK * 4;
    static constexpr int B_STAGE = 2 * B_TILE;                           // hi | lo
    static constexpr int SMEM = A_RING + SB * B_STAGE + 1024 + 256;
    static constexpr uint32_t ACC_COLS = 2 * BN;
    static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(ACC_COLS + ST * 2 * TBK);
};

// NSETS = 2: one CTA per SM with TWO producer warp-sets (warps 0-3 and 5-8; a warp's TMEM lane
// quadrant is warp_id % 4) that alternate k-blocks, and a TMEM A ring as deep as 512 columns allow.
// The micro-benchmark (tools/umma_microbench.cu, profiles/r01_umma_microbench.txt) shows one issuing
// thread saturates the tensor pipe with this MMA pattern, so the limit is how decoupled producers
// and issuer are, not the number of issuers.
template <int BN, int DA, int SB, int ST, int NSETS>
__global__ void __launch_bounds__(32 * (4 * NSETS + 1), (Tc2Cfg<BN, DA, SB, ST, NSETS>::TMEM_COLS <= 256 ? 2 : 1))
conv_gemm_tc2_kernel(const ConvArgs a)
{
    using Cfg = Tc2Cfg<BN, DA, SB, ST, NSETS>;
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
        const int pset = (warp < 4) ? 0 : 1;                 // producer set: handles k-blocks kb % NSETS == pset
        const int pwarp = pset * 4 + quad;                   // index of this warp's private ring
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
        unsigned char *my_ring = a_ring + pwarp * DA * Cfg::A_SLOT;
        const uint32_t ring_u32 = smem_u32(my_ring);
        // incremental tap state of the NEXT k-block this set issues (k-blocks pset, pset+NSETS, ...)
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
        for (int q = 0; q < pset; ++q) advance();
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
                    const float *src = a.in + (ok ? ((a.debug_same_addr & 1) ? (uint32_t)(chunk * 4) : (uint32_t)(row_base[i] + is_off)) : 0u);
                    // .cg (L2 only) by default; L1-allocating gathers (.ca, experiment bit 2) measured no different
                    if (a.debug_same_addr & 2)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(slot + row_dst[i]), "l"(src), "r"(ok ? 16 : 0) : "memory");
                    else
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot + row_dst[i]), "l"(src), "r"(ok ? 16 : 0) : "memory");
                }
                ++is_n;
#pragma unroll
                for (int q = 0; q < NSETS; ++q) advance();
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
        if (pset < nkb) load_split(0);
        int nloc = 0;
        for (int kb = pset; kb < nkb; kb += NSETS, ++nloc) {
            const int st = kb % ST;
            if (lane == 0) mbar_wait(&emptyA[st], ((kb / ST) & 1) ^ 1, 1);   // one poller / one arrival per warp:
            __syncwarp();                                                     // 128 threads hammering the mbarriers cost more than the MMAs
            tc_fence_after();
            const uint32_t ta = tmem_base + lane_addr + Cfg::ACC_COLS + st * 2 * TBK;
            tmem_st32(ta, hi);
            tmem_st32(ta + TBK, lo);
            if (kb + NSETS < nkb) load_split(nloc + 1);       // overlaps the TMEM store latency
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
        for (int c = pset * 32; c < BN; c += 32 * NSETS) {     // the sets share the columns
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


// ---------------------------------------------------------------------------------------------
// Slab variant of the TS pipeline for un-padded stride-1 convolutions with KH*KW > 1 (the CNN's
// 5x4 / 3x3 layers).  The im2col gathers of conv_gemm_tc2_kernel re-read every input element
// KH*KW times from L2 (ncu: 17.9 GB of L2->SM traffic per launch of the 64->64 5x4 layer, 9 TB/s,
// i.e. ~3/4 of the measured L2 throughput cap) -- here the input rows a tile needs are brought into
// shared memory ONCE (coalesced cp.async, pixel-swizzled so that the lane = GEMM-row reads are
// bank-conflict free) and every filter tap is served from that slab.
//   tile = R = 128 / OW consecutive output rows of the global row sequence q = img * OH + oh
//   slab = input rows g(q0) .. g(q1) + KH - 1 with g(q) = q + (q / OH) * (KH - 1): one contiguous range
//          of the NHWC tensor even when the tile straddles images
//   GEMM row r <-> (q0 + r / OW, r % OW); output offset = q0 * OW + r (contiguous)
// Everything downstream of the A fetch (hi/lo split, tcgen05.st ring, B bulk stages, MMA issue,
// epilogue) is the tc2 design.
template <int BN, int SB, int ST, int NSETS = 1>
struct Tc3Cfg {
    static constexpr int THREADS = 32 * (4 * NSETS + 1);
    static constexpr int PRODUCERS = 128 * NSETS;
    static constexpr int B_TILE = BN * TBK * 4;
    static constexpr int B_STAGE = 2 * B_TILE;
    static constexpr uint32_t ACC_COLS = 2 * BN;
    static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(ACC_COLS + ST * 2 * TBK);
    static constexpr int FIXED = SB * B_STAGE + 1024 + 256;              // + slab bytes
};

// NSETS = 2: two producer warp-sets (warps 0-3 and 5-8, TMEM lane quadrant = warp % 4) alternate k-blocks.
// V2 (ISS_B200_TC3_V2=1, prepared for round 2, not yet run on hardware): swizzle key (x + row * OW) & 7 instead of
// p & 7 (no bank conflicts at image-row wraps: 0.42 -> 0.06 extra wavefronts per wavefront, tests/test_slab_indexing.py)
// and explicit ld.shared for the slab reads (the generic-pointer form compiles to LD.E.128 + 64-bit address math).
// POOLIN (ISS_B200_FUSE_POOL=1, implies V2; prepared, not yet run on hardware): the slab is filled with the 2x2/2
// max-pooling of the un-pooled input tensor, so the pooling layer in front of this convolution is never launched
// and its output never written (same comparison order and NaN propagation as maxpool_nhwc_kernel).
template <int BN, int SB, int ST, int NSETS, bool V2 = false, bool POOLIN = false>
__global__ void __launch_bounds__(32 * (4 * NSETS + 1), (Tc3Cfg<BN, SB, ST, NSETS>::TMEM_COLS <= 256 ? 2 : 1))
conv_gemm_tc3_kernel(const ConvArgs a)
{
    using Cfg = Tc3Cfg<BN, SB, ST, NSETS>;
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
    const int nkb = a.Kp / TBK;
    const int R = a.slab_R, KH1 = a.KH - 1;
    const int64_t Q = a.M / a.OW;                                // output rows in the whole batch
    const int64_t q0 = (int64_t)blockIdx.x * R;
    const int64_t mbase = q0 * a.OW;
    const int nq = (int)((Q - q0) < (int64_t)R ? (Q - q0) : (int64_t)R);
    const int valid = nq * a.OW;                                 // GEMM rows of this tile that exist
    const int64_t img0 = q0 / a.OH;

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&fullA[s], 4); mbar_init(&emptyA[s], 1); }   // 4 = the warps of ONE producer set
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
        // ============================ slab fill (128 threads) ============================
        const int quad = warp & 3;                               // TMEM lane quadrant this warp may access
        const int pset = (warp < 4) ? 0 : 1;                     // producer set: k-blocks kb % NSETS == pset
        const int ptid = pset * 128 + quad * 32 + lane;          // 0 .. PRODUCERS-1
        const bool prof = a.prof != nullptr && ptid == 0;
        long long t_start = prof ? clock64() : 0, t_wait = 0, t_fill = 0, t_loop = 0, t_acc = 0;
        const uint32_t slab_u32 = smem_u32(slab);
        const uint32_t pix_bytes = (uint32_t)a.C * 4;
        {
            const int cpp = a.C >> 2;                            // 16-byte chunks per pixel
            const int64_t g0 = q0 + img0 * KH1;                  // first input row (global index) of the slab
            const int64_t img1 = (q0 + nq - 1) / a.OH;
            const int rows = nq + KH1 * (int)(img1 - img0 + 1);
            const int64_t first = g0 * a.W * a.C;                // element offset of the slab in `in`
            const float *src0 = a.in + first;
            const int64_t avail = (a.in_elems - first) >> 2;     // chunks that exist past `first`
            (void)src0; (void)avail;
            const int total = rows * a.W * cpp;
            int p = ptid / cpp, j = ptid - p * cpp;              // chunk ptid, then += PRODUCERS per iteration
            const int dp = Cfg::PRODUCERS / cpp, dj = Cfg::PRODUCERS - dp * cpp;
            int prow = 0, px = 0;                                // V2: slab row / column of pixel p
            if constexpr (V2) { prow = p / a.W; px = p - prow * a.W; }
            if constexpr (POOLIN) {
                static_assert(!POOLIN || V2, "POOLIN builds on the V2 row/column tracking");
                const int64_t Gtot = a.M / ((int64_t)a.OH * a.OW) * a.H;          // pooled input rows in the batch
                int64_t img = (g0 + prow) / a.H;
                int prr = (int)((g0 + prow) - img * a.H);                          // pooled row inside its image
                const size_t row_stride = (size_t)a.inW * a.C;
                for (int q = ptid; q < total; q += Cfg::PRODUCERS) {
                    const int key = px + prow * a.OW;
                    const uint32_t dst = slab_u32 + (uint32_t)p * pix_bytes + (uint32_t)(((j & ~7) | ((j ^ key) & 7)) << 4);
                    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g0 + prow < Gtot) {
                        const float *s0 = a.in + ((size_t)(img * a.inH + 2 * prr) * a.inW + 2 * px) * a.C + j * 4;
                        const float4 v00 = __ldg(reinterpret_cast<const float4 *>(s0));
                        const float4 v01 = __ldg(reinterpret_cast<const float4 *>(s0 + a.C));
                        const float4 v10 = __ldg(reinterpret_cast<const float4 *>(s0 + row_stride));
                        const float4 v11 = __ldg(reinterpret_cast<const float4 *>(s0 + row_stride + a.C));
                        auto mx = [](float best, float v) { return (v > best || v != v) ? v : best; };   // as maxpool_nhwc_kernel
                        m.x = mx(mx(mx(mx(-INFINITY, v00.x), v01.x), v10.x), v11.x);
                        m.y = mx(mx(mx(mx(-INFINITY, v00.y), v01.y), v10.y), v11.y);
                        m.z = mx(mx(mx(mx(-INFINITY, v00.z), v01.z), v10.z), v11.z);
                        m.w = mx(mx(mx(mx(-INFINITY, v00.w), v01.w), v10.w), v11.w);
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(m.x), "f"(m.y), "f"(m.z), "f"(m.w) : "memory");
                    p += dp; j += dj; px += dp;
                    if (j >= cpp) { j -= cpp; ++p; ++px; }
                    while (px >= a.W) {
                        px -= a.W; ++prow;
                        if (++prr == a.H) { prr = 0; ++img; }
                    }
                }
            } else {
                for (int q = ptid; q < total; q += Cfg::PRODUCERS) {
                    const int key = V2 ? (px + prow * a.OW) : p;
                    const uint32_t dst = slab_u32 + (uint32_t)p * pix_bytes + (uint32_t)(((j & ~7) | ((j ^ key) & 7)) << 4);
                    const bool ok = q < avail;
                    cp_async16_u32(dst, src0 + (ok ? (size_t)q * 4 : 0), ok ? 16 : 0);
                    p += dp; j += dj;
                    if constexpr (V2) px += dp;
                    if (j >= cpp) { j -= cpp; ++p; if constexpr (V2) ++px; }
                    if constexpr (V2) { while (px >= a.W) { px -= a.W; ++prow; } }
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::PRODUCERS) : "memory");
        }
        if (prof) t_fill = clock64() - t_start;
        const long long t_l0 = prof ? clock64() : 0;
        // ============================ A producers ============================
        const int r = quad * 32 + lane;                          // GEMM row = TMEM lane
        int pix0 = 0, key0 = 0;
        if (r < valid) {
            const int dq = r / a.OW, ow = r - dq * a.OW;
            const int srow = dq + KH1 * (int)((q0 + dq) / a.OH - img0);
            pix0 = srow * a.W + ow;
            if constexpr (V2) key0 = ow + srow * a.OW;
        }
        int is_c0 = 0, is_ss = 0, is_poff = 0;                   // tap state of the next k-block to load
        int is_key = 0;                                          // V2: swizzle key offset of that tap
        const uint32_t lane_addr = ((uint32_t)(quad * 32)) << 16;
        uint32_t hi[32], lo[32];
        auto advance = [&]() {
            is_c0 += TBK;
            if (is_c0 == a.C) {
                is_c0 = 0; ++is_poff;
                if constexpr (V2) ++is_key;
                if (++is_ss == a.KW) { is_ss = 0; is_poff += a.W - a.KW; if constexpr (V2) is_key += a.OW - a.KW; }
            }
        };
        for (int q = 0; q < pset; ++q) advance();
        auto load_split = [&]() {
            const int p = pix0 + is_poff;
            const unsigned char *base = slab + (size_t)p * pix_bytes + is_c0 * 4;
            const uint32_t x = (uint32_t)((V2 ? (key0 + is_key) : p) & 7) << 4;
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                if constexpr (V2) {
                    const uint32_t addr = slab_u32 + (uint32_t)p * pix_bytes + (uint32_t)is_c0 * 4 + (((uint32_t)j << 4) ^ x);
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
                } else {
                    q = *reinterpret_cast<const uint4 *>(base + (((uint32_t)j << 4) ^ x));
                }
                v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
#pragma unroll
            for (int q = 0; q < NSETS; ++q) advance();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                hi[j] = v[j] & 0xFFFFE000u;
                lo[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(hi[j]));
            }
        };
        if (pset < nkb) load_split();
        for (int kb = pset; kb < nkb; kb += NSETS) {
            const int st = kb % ST;
            const long long tw = prof ? clock64() : 0;
            if (lane == 0) mbar_wait(&emptyA[st], ((kb / ST) & 1) ^ 1, 1);   // one poller / one arrival per warp
            __syncwarp();
            if (prof) t_wait += clock64() - tw;
            tc_fence_after();
            const uint32_t ta = tmem_base + lane_addr + Cfg::ACC_COLS + st * 2 * TBK;
            if (!(a.debug_same_addr & 4)) {                       // (timing experiment 4: producers only hand-shake)
                tmem_st32(ta, hi);
                tmem_st32(ta + TBK, lo);
                if (kb + NSETS < nkb) load_split();              // overlaps the TMEM store latency
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA[st]);
        }

        // ============================ epilogue ============================
        if (prof) t_loop = clock64() - t_l0;
        const long long ta0 = prof ? clock64() : 0;
        if (lane == 0) mbar_wait(accum, 0, 5);                   // all MMAs done => every warp is done with the slab
        __syncwarp();
        if (prof) t_acc = clock64() - ta0;
        tc_fence_after();
        unsigned char *stage_buf = slab + (pset * 4 + quad) * 4096;           // the slab is idle now: per-warp transpose buffer
        const int sub = lane >> 3, chunk = lane & 7;
        const bool has_bias = a.flags & ISS_F_BIAS, pre = a.flags & ISS_F_AFFINE_PRE, post = a.flags & ISS_F_AFFINE_POST;
        const bool relu = a.flags & ISS_F_RELU, resid = a.flags & ISS_F_RESIDUAL;
#pragma unroll 1
        for (int c = pset * 32; c < BN; c += 32 * NSETS) {       // the sets share the columns
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
                const int rr = quad * 32 + rl;
                const int64_t m = mbase + rr;
                const uint4 q4 = *reinterpret_cast<const uint4 *>(stage_buf + rl * 128 + ((chunk ^ (rl & 7)) << 4));
                if (rr < valid) {
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
        if (prof) {
            atomicAdd(a.prof + 4, (unsigned long long)t_loop); atomicAdd(a.prof + 5, (unsigned long long)t_wait);
            atomicAdd(a.prof + 7, (unsigned long long)t_fill); atomicAdd(a.prof + 9, (unsigned long long)t_acc);
            atomicAdd(a.prof + 8, (unsigned long long)(clock64() - ta0 - t_acc)); atomicAdd(a.prof + 10, 1ull);
            atomicAdd(a.prof + 11, (unsigned long long)(clock64() - t_start));
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

template <int BN, int SB, int ST, int NSETS = 1, bool V2 = false, bool POOLIN = false>
int launch_tc3(const ConvArgs &a, int slab_bytes, cudaStream_t st)
{
    using Cfg = Tc3Cfg<BN, SB, ST, NSETS>;
    auto kern = conv_gemm_tc3_kernel<BN, SB, ST, NSETS, V2, POOLIN>;
    static bool configured = false;
    if (!configured) {
        ISS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CTA_MAX));
        configured = true;
    }
    const int64_t Q = a.M / a.OW;
    const int64_t gm = (Q + a.slab_R - 1) / a.slab_R;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv_tc: M too large");
    dim3 grid((unsigned)gm, (unsigned)(a.N / BN));
    kern<<<grid, Cfg::THREADS, Cfg::FIXED + slab_bytes, st>>>(a);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

unsigned long long *g_prof = nullptr;
void prof_dump()
{
    unsigned long long h[12];
    if (!g_prof || cudaMemcpy(h, g_prof, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess || !h[10]) return;
    const double n = (double)h[10];
    fprintf(stderr, "libiss_b200 tc3 prof (cycles per CTA, %.0f CTAs): issuer loop %.0f = wait fullB %.0f + wait fullA %.0f + wait emptyB %.0f + issue; "
            "producer: fill %.0f, loop %.0f (wait emptyA %.0f), accum wait %.0f, epilogue %.0f, total %.0f\n",
            n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[7] / n, h[4] / n, h[5] / n, h[9] / n, h[8] / n, h[11] / n);
}

// Slab kernel dispatch; returns 1 if the layer is not eligible (caller falls back to tc2).
int try_launch_slab(ConvArgs &a, cudaStream_t st)
{
    static const int want_prof = [] { const char *e = getenv("ISS_B200_TC_PROF"); return e ? atoi(e) : 0; }();
    if (want_prof && !g_prof) {
        if (cudaMalloc(&g_prof, 12 * sizeof(unsigned long long)) == cudaSuccess) { cudaMemset(g_prof, 0, 12 * sizeof(unsigned long long)); atexit(prof_dump); }
    }
    a.prof = (want_prof && a.N == want_prof) ? g_prof : nullptr;      // ISS_B200_TC_PROF=<N of the layers to instrument>
    if (a.SH != 1 || a.SW != 1 || a.PT != 0 || a.PL != 0 || a.KH * a.KW <= 1) return 1;
    if (a.OH != a.H - a.KH + 1 || a.OW != a.W - a.KW + 1 || a.OW > TBM || a.Kp != a.K || a.N % 64 != 0) return 1;
    const int R = TBM / a.OW;
    const int cross = (R - 1) / a.OH + 1;                        // image boundaries a tile can contain
    const int rows = R + (a.KH - 1) * (1 + cross);
    int slab_bytes = rows * a.W * a.C * 4;
    if (slab_bytes < 32768) slab_bytes = 32768;                  // doubles as the 8 x 4 KB epilogue transpose buffers
    a.slab_R = R;
    a.slab_rows = rows;
    a.in_elems = a.M / ((int64_t)a.OH * a.OW) * a.H * a.W * a.C;
    static const int cfg = [] { const char *e = getenv("ISS_B200_TC3_CFG"); return e ? atoi(e) : 0; }();   // experiments
    if (cfg == 1 && a.N % 128 != 0 && Tc3Cfg<64, 4, 6>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<64, 4, 6, 1>(a, slab_bytes, st);
    if (cfg == 2 && a.N % 128 != 0 && Tc3Cfg<64, 4, 6>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<64, 4, 6, 2>(a, slab_bytes, st);
    if (cfg == 2 && a.N % 128 == 0 && Tc3Cfg<128, 3, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 3, 4, 2>(a, slab_bytes, st);
    if (a.pool_in) {                                             // fused input pooling: POOLIN instantiations only
        if (a.N % 128 == 0) {
            if (Tc3Cfg<128, 4, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 4, 4, 1, true, true>(a, slab_bytes, st);
            if (Tc3Cfg<128, 3, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 3, 4, 1, true, true>(a, slab_bytes, st);
        } else {
            if (Tc3Cfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3<64, 3, 2, 1, true, true>(a, slab_bytes, st);
            if (Tc3Cfg<64, 2, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3<64, 2, 2, 1, true, true>(a, slab_bytes, st);
        }
        return 1;
    }
    static const int v2 = [] { const char *e = getenv("ISS_B200_TC3_V2"); return (e && e[0] == '1') ? 1 : 0; }();      // prepared, not yet validated
    if (v2) {
        if (a.N % 128 == 0) {
            if (Tc3Cfg<128, 4, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 4, 4, 1, true>(a, slab_bytes, st);
            if (Tc3Cfg<128, 3, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 3, 4, 1, true>(a, slab_bytes, st);
        } else {
            if (Tc3Cfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3<64, 3, 2, 1, true>(a, slab_bytes, st);
            if (Tc3Cfg<64, 2, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3<64, 2, 2, 1, true>(a, slab_bytes, st);
        }
    }
    if (a.N % 128 == 0) {
        if (Tc3Cfg<128, 4, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 4, 4>(a, slab_bytes, st);
        if (Tc3Cfg<128, 3, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 3, 4>(a, slab_bytes, st);
        if (Tc3Cfg<128, 2, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<128, 2, 4>(a, slab_bytes, st);
        return 1;
    }
    if (Tc3Cfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3<64, 3, 2>(a, slab_bytes, st);
    if (Tc3Cfg<64, 2, 2>::FIXED + slab_bytes <= SMEM_HALF_SM) return launch_tc3<64, 2, 2>(a, slab_bytes, st);
    if (Tc3Cfg<64, 3, 2>::FIXED + slab_bytes <= SMEM_CTA_MAX) return launch_tc3<64, 3, 2>(a, slab_bytes, st);
    return 1;
}

template <int BN, int DA, int SB, int ST, int NSETS = 1>
int launch_tc2(const ConvArgs &a, cudaStream_t st)
{
    using Cfg = Tc2Cfg<BN, DA, SB, ST, NSETS>;
    auto kern = conv_gemm_tc2_kernel<BN, DA, SB, ST, NSETS>;
    static bool configured = false;
    if (!configured) {
        ISS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        configured = true;
    }
    const int64_t gm = (a.M + TBM - 1) / TBM;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv_tc: M too large");
    dim3 grid((unsigned)gm, (unsigned)(a.N / BN));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

template <int BN, int STAGES, bool A_TMEM>
int launch_tc(const ConvArgs &a, cudaStream_t st)
{
    using Cfg = TcCfg<BN, STAGES, A_TMEM>;
    auto kern = conv_gemm_tc_kernel<BN, STAGES, A_TMEM>;
    static bool configured = false;
    if (!configured) {
        ISS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        configured = true;
    }
    const int64_t gm = (a.M + TBM - 1) / TBM;
    ISS_REQUIRE(gm < (1ll << 31), ISS_ERR_INVALID, "conv_tc: M too large");
    dim3 grid((unsigned)gm, (unsigned)(a.N / BN));
    kern<<<grid, 160, Cfg::SMEM, st>>>(a);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

int g_gemm_mode = -1;      // -1 = read ISS_B200_GEMM on first use

}  // namespace

extern "C" int iss_set_gemm_mode(int mode)
{
    ISS_REQUIRE(mode >= 0 && mode <= 3, ISS_ERR_INVALID, "iss_set_gemm_mode: %d", mode);
    g_gemm_mode = mode;
    return ISS_OK;
}

extern "C" int iss_get_gemm_mode(void)
{
    if (g_gemm_mode < 0) {
        const char *e = getenv("ISS_B200_GEMM");
        g_gemm_mode = ISS_GEMM_DEFAULT;
        if (e && !strcmp(e, "fp32")) g_gemm_mode = ISS_GEMM_FP32;
        else if (e && !strcmp(e, "tc_ss")) g_gemm_mode = ISS_GEMM_TC_SS;
        else if (e && !strcmp(e, "tc_ts")) g_gemm_mode = ISS_GEMM_TC_TS;
        else if (e && !strcmp(e, "tc_f16")) g_gemm_mode = ISS_GEMM_TC_F16;
    }
    return g_gemm_mode;
}

int iss_launch_conv_tc_f16(ConvArgs &a, cudaStream_t st);              // conv_gemm_tc_f16.cu; 1 = layer not covered

bool iss_conv_tc_eligible(const ConvArgs &a)
{
    return a.wt_hi && a.wt_lo && a.wt_tiled && a.Kp > 0 && a.C % 32 == 0 && a.K % 32 == 0 && a.N % 32 == 0 && a.N >= 32;
}

// Mirrors the checks of try_launch_slab + the POOLIN dispatch above (kept next to them on purpose).
bool iss_conv_poolin_supported(const ConvArgs &a, int mode)
{
    if (mode != ISS_GEMM_TC_TS || !iss_conv_tc_eligible(a)) return false;
    if (a.SH != 1 || a.SW != 1 || a.PT != 0 || a.PL != 0 || a.KH * a.KW <= 1) return false;
    if (a.OH != a.H - a.KH + 1 || a.OW != a.W - a.KW + 1 || a.OW > TBM || a.Kp != a.K || a.N % 64 != 0) return false;
    const int R = TBM / a.OW;
    const int cross = (R - 1) / a.OH + 1;
    int slab_bytes = (R + (a.KH - 1) * (1 + cross)) * a.W * a.C * 4;
    if (slab_bytes < 32768) slab_bytes = 32768;
    if (a.N % 128 == 0) return Tc3Cfg<128, 3, 4>::FIXED + slab_bytes <= SMEM_CTA_MAX;
    return Tc3Cfg<64, 2, 2>::FIXED + slab_bytes <= SMEM_HALF_SM;
}

int iss_launch_conv_tc(const ConvArgs &a_in, int mode, cudaStream_t st)
{
    static const int dbg = [] { const char *e = getenv("ISS_B200_TC_DEBUG"); return e ? atoi(e) : 0; }();   // timing experiments: 1 = same-address gathers, 2 = L1-allocating gathers
    ConvArgs a = a_in;
    a.debug_same_addr = dbg;
    if (mode == ISS_GEMM_TC_F16) {                                      // experimental engine: slab convolutions only
        const int rc = iss_launch_conv_tc_f16(a, st);
        if (rc != 1) return rc;
    }
    const bool ts = (mode == ISS_GEMM_TC_TS || mode == ISS_GEMM_TC_F16);
    ISS_REQUIRE(!a.pool_in || mode == ISS_GEMM_TC_TS, ISS_ERR_UNSUPPORTED, "conv_tc: fused input pooling needs engine 2");
    // BN: the widest of {256,128,64,32} dividing N
    // two accumulators per tile (main + correction) => BN <= 128 (2 x 128 + A ring <= 512 TMEM columns)
    if (ts) {
        static const int two_sets = [] { const char *e = getenv("ISS_B200_TC_SETS"); return (e && e[0] == '2') ? 1 : 0; }();
        if (two_sets && !a.pool_in) {
            if (a.N % 128 == 0) return launch_tc2<128, 2, 4, 4, 2>(a, st);   // 193 KB smem, 512 TMEM cols
            if (a.N % 64 == 0) return launch_tc2<64, 2, 4, 6, 2>(a, st);     // 129 KB smem (=> ~96 KB L1), 512 TMEM cols
        }
        static const int slab = [] { const char *e = getenv("ISS_B200_TC_SLAB"); return (e && e[0] == '0') ? 0 : 1; }();
        if (slab || a.pool_in) {                                         // un-padded stride-1 KHxKW convs: input slab in smem
            const int rc = try_launch_slab(a, st);
            if (rc != 1) return rc;
        }
        ISS_REQUIRE(!a.pool_in, ISS_ERR_UNSUPPORTED, "conv_tc: fused input pooling needs the slab kernel");
        if (a.N % 128 == 0) return launch_tc2<128, 4, 4, 4>(a, st);      // 193 KB smem, 512 TMEM cols (2x128 acc + 4 A stages), 1 CTA/SM
        if (a.N % 64 == 0) return launch_tc2<64, 3, 3, 2>(a, st);        //  97 KB smem, 256 TMEM cols, 2 CTAs/SM
        return launch_tc2<32, 3, 4, 3>(a, st);                            //  81 KB smem, 256 TMEM cols
    }
    if (a.N % 128 == 0) return launch_tc<128, 2, false>(a, st);
    if (a.N % 64 == 0) return launch_tc<64, 2, false>(a, st);
    return launch_tc<32, 3, false>(a, st);
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
