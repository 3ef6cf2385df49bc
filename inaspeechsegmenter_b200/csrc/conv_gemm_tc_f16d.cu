// conv_gemm_tc_f16d.cu -- engine 3, DIRECT variant: the fp16-split convolution with BOTH tensor-core operands read
// from shared memory (tcgen05.mma SS form); there are no A-operand producer warps at all.
//
// Idea.  Lay the NHWC input of a whole batch out as ONE tall image of n_img * H rows of W pixels and give every pixel
// p = (img * H + ih) * W + iw a 128-byte shared-memory row per 64-channel plane (fp16 hi plane, fp16 lo plane,
// SWIZZLE_128B K-major).  Number the outputs the same way -- "slot" s = (img * H + oh) * W + ow -- and the A tile of
// filter tap (kh, kw) for the 128 slots s0 .. s0 + 127 is simply the 128 pixel rows starting at s0 + kh * W + kw:
// a contiguous, row-shifted window of the slab, which a UMMA shared-memory descriptor can address directly (the
// 128-byte swizzle is a function of the absolute shared-memory address, so a descriptor may start at any row:
// tools/umma_desc_offset_test.cu, profiles/r02_umma_desc_offset_test.txt: base-offset field 0 for every start row).  The price: slots with ow >= OW (KW - 1 per row) and oh >= OH (KH - 1 rows per
// image) are computed and thrown away -- (OW / W) * (OH / H) = 77 % useful rows for the 5x4 layer of the segmenter
// CNNs -- in exchange for deleting the producers' 32 KB of LDS + PRMT + tcgen05.st per k-block, which bounded the
// TMEM-operand kernel (conv_gemm_tc_f16.cu) at ~36 % of the tensor pipe.
//
// One persistent CTA per SM, 512 threads, warp-specialised:
//   warp 0      MMA issuer (warp-uniform loop, tcgen05 under elect.sync)
//   warp 1      weight loader: one cp.async.bulk per 16 KB stage [Bh | Bl], 3-deep ring
//   warps 2-5   epilogue (one TMEM lane quadrant each; 2-9 for the residual layers, one per quadrant and sub-tile):
//               accumulators -> bias/BN (+ residual) / ReLU -> split-half words or fp32 -> HBM
//   the rest    slab fill: de-interleaves split-half words from HBM (IN_PACKED), or takes the 2x2 / stride-2 maximum of
//               the un-pooled tensor on the way (IN_POOL: the MaxPooling2D layer in front never runs), or EVALUATES the
//               one-channel first convolution from the shared two-float map Yh + Yl (IN_FIRST, FirstFuse in conv_gemm.cuh)
// A tile is DT = 2 sub-tiles of 128 slots that share the slab and every weight stage (half the weight traffic of a
// 128-row tile); slab and accumulators are double-buffered, so fill(i+1), MMA(i) and epilogue(i-1) overlap.
// Per k-block (= 64 input channels of one filter tap) and sub-tile: Ah.[Bh | Bl] (N = 128) and Al.Bh (N = 64), as in
// the other fp16-split kernels; TMEM: 2 buffers x 2 sub-tiles x 128 columns = 512.
// 1x1 convolutions / Dense layers are the degenerate case (one tap, every slot useful): the kernel then is a persistent
// GEMM whose A tile (256 rows x C) stays in shared memory for all NT n-tile passes -- used for ResNet101's "expand"
// convolutions (K = 32 .. 128, N = 128 .. 512, + residual + ReLU), which are memory-bound and ran at ~1/6 of HBM speed as
// one-tile CTAs of the gather kernel (2 k-blocks of MMAs per CTA, un-overlapped 128 x 128 epilogue).  C = 32 is zero-padded
// to one 64-channel k-block; the input may be fp32 (IN_F32: split in the fill); the residual may be fp32 or words.
// Wider layers: C = 64 * CB input channels are CB planes pairs per slab buffer (k-block = (tap, channel block)); N = 64 * NT
// output channels are NT passes over the SAME slab (pass = (tile, n-tile), accumulators alternate per pass).  When two
// slab buffers do not fit (C = 128: 139 KB each) the kernel runs with one: the fill of the next tile then waits for the
// last pass of the current one (NBUF = 1; ~10 % bubble for the 3x3 128 -> 128 layer).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

int iss_launch_conv_tc_f16d(ConvArgs &a, cudaStream_t st);
bool iss_conv_f16_direct_covers(const ConvArgs &a);

namespace {

constexpr int HBK = 64;
// (DT, the number of 128-slot sub-tiles per tile, is a template parameter of the kernel: 2, or 1 for the 1x1 layers whose
//  two slab buffers only fit with 128-row tiles)
constexpr int DSB = 3;                                  // weight stages
constexpr int DBN = 64;                                 // output channels per pass (n-tile)
constexpr int D_NMAX = 512;                             // output channels of a layer
constexpr int D_WARPS = 16;                             // MMA issuer, weight loader, n_epi epilogue warps (4 or 8), the rest fill the slab
constexpr int D_THREADS = 32 * D_WARPS;
constexpr int D_B_STAGE = 2 * DBN * 128;                // [hi rows | lo rows]
constexpr int D_TAB = 8;                                // images a slab may touch (IN_FIRST table)
constexpr int D_SMEM_MAX = 232448;

constexpr int DIN_PACKED = 1, DIN_FIRST = 2, DIN_POOL = 3, DIN_F32 = 4, DIN_PAD = 5;
constexpr int D_RES_BOX = 128 * 128;                    // one residual / output box: 128 rows x 32 channels (128 B, SWIZZLE_128B)
constexpr int D_RES_MAX = 4;                            // boxes in the ring at most

struct DirectArgs {
    const unsigned char *wt;    // tiled fp16 image [k-block][hi | lo][64 rows x 128 B, SWIZZLE_128B]
    float inv_scale;
    int npix;                   // slab rows (pixels), multiple of 8
    int n_tiles;
    int n_img;
    int64_t total_pix;          // n_img * H * W
    int cb;                     // 64-channel blocks of the input (ceil(C / 64)), a power of two
    int c_real;                 // input channels in memory (32 is zero-padded to one block)
    int nt;                     // 64-channel n-tiles (N / 64)
    int bn_img;                 // n-tile width of the weight image (64 or 128, iss_f16_bn_for)
    int n_epi;                  // epilogue warps: 4 (one per TMEM lane quadrant, both sub-tiles) or 8 (one per quadrant and sub-tile)
    int tma;                    // residual 1x1 layers: bit 0 = residual boxes by tensor-map TMA (cp.async.bulk.tensor), bit 1 = outputs by TMA store
    int rs;                     // boxes in the residual ring (2 .. D_RES_MAX)
};
// tensor maps of the residual and the output tensor ([M rows][N words], box 32 words x 128 rows, SWIZZLE_128B); zero when unused
struct alignas(64) DirectMaps { CUtensorMap res, out; };

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
// four epilogue constants (channels n .. n+3) as one LDS.128 broadcast.  The table is written once before the CTA-wide
// barrier and never again, so the load is plain (not volatile) asm: the compiler may schedule it freely.
// (Indexing the table through its generic pointer compiled to one 4-byte generic LD per constant -- 64-128 per 32
// columns and warp, each a shared-memory wavefront competing with the tensor core's operand reads.)
__device__ __forceinline__ void lds_f4(uint32_t addr, float (&v)[4])
{
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

struct DSmem {                                           // everything behind the 1024-aligned operand buffers
    long long tab_row[2][D_TAB];                         // IN_FIRST: Y row of input row 0 of the images a slab touches (-1: none)
    uint64_t slab_full[2], slab_empty[2], acc_full[2], acc_empty[2], b_full[DSB], b_empty[DSB];
    uint64_t res_full[D_RES_MAX], res_empty[D_RES_MAX];  // residual boxes (TMA mode)
    uint32_t tmem_slot;
};

__device__ __forceinline__ void lds128(uint32_t addr, uint32_t &x, uint32_t &y, uint32_t &z, uint32_t &w)
{
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
}
// one box of a [rows][words] tensor -> shared memory, completion on an mbarrier (UTMALDG)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// one box shared memory -> tensor (UTMASTG), bulk async-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
template <int MODE, int NBUF, int DT>
__global__ void __launch_bounds__(D_THREADS, 1)
conv_gemm_tc4h_kernel(const ConvArgs a, const FirstFuse ff, const DirectArgs d, const __grid_constant__ DirectMaps maps)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const uint32_t plane = (uint32_t)d.npix * 128u;       // bytes of one fp16 plane of the slab (multiple of 1024)
    unsigned char *b_ring = smem;
    unsigned char *slab = smem + DSB * D_B_STAGE;         // [NBUF buffers][channel block][hi plane | lo plane]
    const uint32_t slab_buf = 2u * (uint32_t)d.cb * plane;
    unsigned char *res_ring = slab + (size_t)NBUF * slab_buf;          // [d.rs boxes] (TMA mode only), 1024-aligned
    DSmem *sm = reinterpret_cast<DSmem *>(res_ring + (d.tma ? (size_t)d.rs * D_RES_BOX : 0));
    // k1 | k0 | es2 | et2, N floats each (y = acc * k1 + k0, ReLU, y * es2 + et2); 16-byte aligned for the epilogue's LDS.128
    float *cst = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(sm + 1) + 15) & ~(uintptr_t)15);
    // TMA mode: the last warp feeds the residual ring instead of filling the slab
    const int n_fill_warps = D_WARPS - 2 - d.n_epi - (d.tma ? 1 : 0), first_fill = 2 + d.n_epi, nfill = 32 * n_fill_warps;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = a.KH * a.KW * d.cb;                   // k-block = 64 channels of one filter tap
    const uint32_t HW = (uint32_t)(a.H * a.W);

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm->slab_full[b], n_fill_warps); mbar_init(&sm->slab_empty[b], 1);
            mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], d.n_epi);
        }
        for (int s = 0; s < DSB; ++s) { mbar_init(&sm->b_full[s], 1); mbar_init(&sm->b_empty[s], 1); }
        // a box is released by the 4 warps that read it (outputs by TMA store: once their own rows have left)
        for (int s = 0; s < D_RES_MAX; ++s) { mbar_init(&sm->res_full[s], 1); mbar_init(&sm->res_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        for (int n = tid; n < a.N; n += D_THREADS) {      // epilogue constants, once per CTA
            const bool has_bias = a.flags & ISS_F_BIAS, pre = a.flags & ISS_F_AFFINE_PRE, post = a.flags & ISS_F_AFFINE_POST;
            const float eb = has_bias ? __ldg(a.bias + n) : 0.f;
            const float s1 = pre ? __ldg(a.pre_scale + n) : 1.f, t1 = pre ? __ldg(a.pre_shift + n) : 0.f;
            cst[n] = d.inv_scale * s1; cst[a.N + n] = fmaf(eb, s1, t1);
            cst[2 * a.N + n] = post ? __ldg(a.post_scale + n) : 1.f; cst[3 * a.N + n] = post ? __ldg(a.post_shift + n) : 0.f;
        }
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
        uint32_t p = 0;                                   // running pass count (accumulator buffer = p & 1)
        int i = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x, ++i) {
            const int sbuf = NBUF == 2 ? (i & 1) : 0;
            const uint32_t suse = (uint32_t)(NBUF == 2 ? (i >> 1) : i) & 1u;
            const uint32_t slab_b = slab_u32 + (uint32_t)sbuf * slab_buf;
            for (int nt = 0; nt < d.nt; ++nt, ++p) {
                const uint32_t abuf = p & 1u, ause = (p >> 1) & 1u;
                if (nt == 0) mbar_wait(&sm->slab_full[sbuf], suse, 1);
                mbar_wait(&sm->acc_empty[abuf], ause ^ 1u, 2);
                tc_fence_after();
                int kh = 0, kw = 0, cb = 0;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const uint32_t sl = g % DSB;
                    mbar_wait(&sm->b_full[sl], (g / DSB) & 1u, 3);
                    tc_fence_after();
                    const uint64_t db = make_sw128_desc(b_u32 + sl * D_B_STAGE);
                    const uint32_t tap = slab_b + (uint32_t)cb * 2u * plane + (uint32_t)(kh * a.W + kw) * 128u;
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < DT; ++t) {
                            const uint32_t arow = tap + (uint32_t)t * (128u * 128u);
                            const uint64_t dah = make_sw128_desc(arow), dal = make_sw128_desc(arow + plane);   // base-offset field stays 0
                            const uint32_t dm = tb + abuf * 256u + (uint32_t)t * 128u;
                            // (K-step-major order -- both sub-tiles back to back on the same weight columns -- measured 0.5 % slower: r02k)
#pragma unroll
                            for (int kk = 0; kk < HBK / 16; ++kk) {
                                umma_f16_ss(dm, dah + 2 * kk, db + 2 * kk, idesc2, (kb > 0 || kk > 0) ? 1u : 0u);     // Ah.[Bh | Bl]
                                umma_f16_ss(dm + DBN, dal + 2 * kk, db + 2 * kk, idesc, 1u);                          // Al.Bh
                            }
                        }
                        umma_commit(&sm->b_empty[sl]);
                        if (kb == nkb - 1) {
                            if (nt == d.nt - 1) umma_commit(&sm->slab_empty[sbuf]);
                            umma_commit(&sm->acc_full[abuf]);
                        }
                    }
                    __syncwarp();
                    if (++cb == d.cb) { cb = 0; if (++kw == a.KW) { kw = 0; ++kh; } }
                }
            }
        }
        tc_fence_before();
    } else if (warp == 1) {
        // ============================ weight loader ============================
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x) {
            for (int nt = 0; nt < d.nt; ++nt) {
                // rows nt * 64 .. + 63 of the layer's weight image [n-tile of bn_img][k-block][hi rows | lo rows][128 B]
                const int row0 = nt * DBN, it = row0 / d.bn_img, sub = row0 - it * d.bn_img;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const uint32_t sl = g % DSB;
                    mbar_wait(&sm->b_empty[sl], ((g / DSB) & 1u) ^ 1u, 4);
                    if (elect_one()) {
                        const unsigned char *hi_src = d.wt + (((size_t)it * nkb + kb) * 2 * d.bn_img + sub) * 128;
                        const uint32_t dst = smem_u32(b_ring + sl * D_B_STAGE), bar = smem_u32(&sm->b_full[sl]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)D_B_STAGE) : "memory");
                        if (d.bn_img == DBN) {
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst), "l"(hi_src), "r"((uint32_t)D_B_STAGE), "r"(bar) : "memory");
                        } else {
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst), "l"(hi_src), "r"((uint32_t)(D_B_STAGE / 2)), "r"(bar) : "memory");
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst + D_B_STAGE / 2), "l"(hi_src + (size_t)d.bn_img * 128), "r"((uint32_t)(D_B_STAGE / 2)), "r"(bar) : "memory");
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < first_fill) {
        // ============================ epilogue ============================
        const int quad = warp & 3;                        // the TMEM lanes this warp may read
        const uint32_t lane_addr = ((uint32_t)(quad * 32)) << 16;
        const bool relu = a.flags & ISS_F_RELU, post = a.flags & ISS_F_AFFINE_POST, resid = a.flags & ISS_F_RESIDUAL;
        // 8 epilogue warps: warps 2-5 take sub-tile 0, warps 6-9 sub-tile 1 (twice the loads / stores in flight: the residual
        // layers are bound by memory latency x bytes in flight, not by instructions)
        // DT = 2: one group of four warps per sub-tile; DT = 1: the two groups split the 32-column chunks of the one sub-tile
        const int egrp = d.n_epi == 8 ? ((warp - 2) >> 2) : 0;
        const bool split_t = DT == 2 && d.n_epi == 8, split_c = DT == 1 && d.n_epi == 8;
        const int t_first = split_t ? egrp : 0, t_last = split_t ? t_first + 1 : DT;
        const int c_first = split_c ? 32 * egrp : 0, c_step = split_c ? 64 : 32;
        const uint32_t cst_u32 = smem_u32(cst), cst_arr = 4u * (uint32_t)a.N;         // byte address of the table, bytes per array
        // 8 channels (n0 ..) of this lane's row: bias/BN affine (+ residual words rw) / ReLU / second affine -> 8 output words
        auto finish8 = [&](const uint32_t *acc, const u32x8 &rw, int n0) -> u32x8 {
            float y[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float k1[4], k0[4], s2[4], t2[4];
                const uint32_t cj = cst_u32 + 4u * (uint32_t)(n0 + 4 * h);
                lds_f4(cj, k1); lds_f4(cj + cst_arr, k0);
                if (post) { lds_f4(cj + 2u * cst_arr, s2); lds_f4(cj + 3u * cst_arr, t2); }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v = fmaf(__uint_as_float(acc[4 * h + q]), k1[q], k0[q]);
                    if (resid) v += a.residual_packed ? iss_unpack_split(rw.v[4 * h + q]) : __uint_as_float(rw.v[4 * h + q]);
                    if (relu) v = fmaxf(v, 0.f);
                    if (post) v = fmaf(v, s2[q], t2[q]);
                    y[4 * h + q] = v;
                }
            }
            u32x8 w;
            if (a.out_packed) {
#pragma unroll
                for (int q = 0; q < 4; ++q) iss_pack_split2(y[2 * q], y[2 * q + 1], w.v[2 * q], w.v[2 * q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) w.v[q] = __float_as_uint(y[q]);
            }
            return w;
        };
        // TMA mode (residual 1x1 layers, 8 epilogue warps): the residual words of (pass, 32-channel half h, sub-tile t) arrive as
        // box DT * chunk + t of a ring fed by the last warp; with bit 1 the outputs are written back in place and leave by TMA store
        const bool tma = d.tma != 0, tma_st = (d.tma & 2) != 0;
        const uint32_t res_u32 = smem_u32(res_ring);
        uint32_t chunk = 0;                                   // running (tile, pass, half) count of this CTA
        uint32_t p = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x) {
            for (int nt = 0; nt < d.nt; ++nt, ++p) {
                const uint32_t abuf = p & 1u, ause = (p >> 1) & 1u;
                const int nb = nt * DBN;
                const int ncol = a.N - nb < DBN ? a.N - nb : DBN;    // (N = 32: only the first 32 columns exist)
                // this lane's row of sub-tile t: slot -> (image, oh, ow), validity, offset of its channels nb ..
                uint32_t slot; bool valid; int64_t orow;
                auto geometry = [&](int t) {
                    slot = (uint32_t)tile * (DT * 128u) + (uint32_t)t * 128u + (uint32_t)(quad * 32 + lane);
                    const uint32_t img = slot / HW, rem = slot - img * HW;
                    const uint32_t oh = rem / (uint32_t)a.W, ow = rem - oh * (uint32_t)a.W;
                    valid = img < (uint32_t)d.n_img && oh < (uint32_t)a.OH && ow < (uint32_t)a.OW;
                    orow = (((int64_t)img * a.OH + oh) * a.OW + ow) * a.N + nb;
                };
                if (lane == 0) mbar_wait(&sm->acc_full[abuf], ause, 5);
                __syncwarp();
                tc_fence_after();
#pragma unroll 1
                for (int t = t_first; t < t_last; ++t) {
                    geometry(t);
                    if (resid && valid && !tma) {                   // the residual of the pass behind this one -> L2 while this one is computed
                        const float *nx = nt + 1 < d.nt ? a.residual + orow + DBN
                                                        : a.residual + orow - nb + (int64_t)gridDim.x * (DT * 128) * a.N;     // (1x1 layers: row = slot)
                        if (nt + 1 < d.nt || (a.KH * a.KW == 1 && tile + (int)gridDim.x < d.n_tiles && (int64_t)slot + (int64_t)gridDim.x * (DT * 128) < a.M)) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
                        }
                    }
                    float *dst = a.out + orow;
                    const uint4 *res = resid ? reinterpret_cast<const uint4 *>(a.residual + orow) : nullptr;
                    const int64_t row0 = (int64_t)tile * (DT * 128) + (int64_t)t * 128;          // first row of the sub-tile (1x1: row = slot)
#pragma unroll 1
                    for (int c = c_first; c < ncol; c += c_step) {
                        uint32_t acc[32];
                        {
                            uint32_t corr[32];
                            const uint32_t col = tmem_base + lane_addr + abuf * 256u + (uint32_t)t * 128u + (uint32_t)c;
                            tmem_ld32(col, acc);
                            tmem_ld32(col + DBN, corr);
#pragma unroll
                            for (int q = 0; q < 32; ++q) acc[q] = __float_as_uint(__uint_as_float(acc[q]) + __uint_as_float(corr[q]));
                        }
                        if (tma) {
                            const uint32_t ck = chunk + (uint32_t)(c >> 5);
                            {
                                // each group of four warps owns one half of the ring: its k-th box sits in slot k % half of that half
                                // (one consumer per slot: a waiter can never be two phases away from the phase it waits for)
                                const uint32_t half = (uint32_t)d.rs >> 1, k = DT == 2 ? ck : ck >> 1;
                                const uint32_t sl = (uint32_t)egrp * half + k % half, use = k / half;
                                mbar_wait(&sm->res_full[sl], use & 1u, 8);      // every reading thread observes the completion itself
                                // this lane's row of the box: 128 bytes, 16-byte chunk q at q ^ (row & 7) (SWIZZLE_128B)
                                const uint32_t r = (uint32_t)(quad * 32 + lane);
                                const uint32_t rowa = res_u32 + sl * (uint32_t)D_RES_BOX + r * 128u, x7 = r & 7u;
                                u32x8 rw[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    lds128(rowa + (((uint32_t)(2 * j) ^ x7) << 4), rw[j].v[0], rw[j].v[1], rw[j].v[2], rw[j].v[3]);
                                    lds128(rowa + (((uint32_t)(2 * j + 1) ^ x7) << 4), rw[j].v[4], rw[j].v[5], rw[j].v[6], rw[j].v[7]);
                                }
                                if (!tma_st) {
                                    if (valid) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) stg256(dst + c + 8 * j, finish8(acc + 8 * j, rw[j], nb + c + 8 * j));
                                    }
                                    // The box is released BEHIND the stores, i.e. after the residual words have been USED.  An arrive
                                    // issued right behind the LDS (the first form of this mode) gave wrong residuals in most windows
                                    // of a large batch (gpurun_out r02t / r02u / r02v): the loads are still queued -- the tensor
                                    // core's operand reads keep the shared memory busy -- when the barrier flips, and the producer's
                                    // refill (an L2 hit thanks to the prefetch cursor; fewer mismatches without it) lands first.  A
                                    // variant that XOR-ed every loaded word into the barrier address "& 0" changed nothing because
                                    // ptxas folds the AND and with it the dependency (checked in the SASS).  This order: 0 mismatches
                                    // at 37 / 261 / 700 windows, both tilings, every run.  An early release with a dependency the
                                    // assembler cannot fold is the obvious next experiment; until it has run on hardware the mode
                                    // stays opt-in and the plain LDG epilogue is the default.
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive(&sm->res_empty[sl]);          // 4 warps read a box
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const u32x8 w = finish8(acc + 8 * j, rw[j], nb + c + 8 * j);
                                        sts128(rowa + (((uint32_t)(2 * j) ^ x7) << 4), w.v[0], w.v[1], w.v[2], w.v[3]);
                                        sts128(rowa + (((uint32_t)(2 * j + 1) ^ x7) << 4), w.v[4], w.v[5], w.v[6], w.v[7]);
                                    }
                                    // every warp stores its own 32 rows of the box (a 32 x 32-word sub-box whose 4 KB start on a
                                    // 1024-byte boundary, so the swizzle pattern is the box's own): no barrier across the four warps
                                    fence_proxy_async();             // generic-proxy stores -> visible to the TMA store
                                    __syncwarp();
                                    if (lane == 0) {
                                        const int64_t rq = row0 + quad * 32;
                                        if (rq < a.M) {              // rows behind the last one are clipped by the TMA unit
                                            tma_store_2d(&maps.out, res_u32 + sl * (uint32_t)D_RES_BOX + (uint32_t)quad * 4096u, nb + c, (int)rq);
                                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the rows have been read: the box may be refilled
                                        }
                                        mbar_arrive(&sm->res_empty[sl]);
                                    }
                                }
                            }
                        } else if (valid) {
                            // (requesting the residual words one chunk ahead of use -- the first chunk's before the accumulators are
                            //  awaited -- measured 0.4 % SLOWER on ResNet101, r02n: the residual layers are not bound by this latency)
                            u32x8 rw[4] = {};                        // 32 channels of residual: four 32-byte loads
                            if (resid) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) rw[j] = ldg256(res + (c >> 2) + 2 * j);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) stg256(dst + c + 8 * j, finish8(acc + 8 * j, rw[j], nb + c + 8 * j));
                        }
                    }
                }
                chunk += (uint32_t)(ncol >> 5);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm->acc_empty[abuf]);
            }
        }
        if (tma_st) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every TMA store of this thread has completed
    } else if (d.tma && warp == D_WARPS - 1) {
        // ============================ residual boxes (TMA mode) ============================
        // order = the consumers': per (tile, pass, 32-channel half) one box per sub-tile.  The ring is short (2-4 boxes: what the
        // slab leaves of the shared memory), so a second cursor runs D_RES_AHEAD boxes in front and pulls them into L2
        // (cp.async.bulk.prefetch.tensor): the ring then turns over at L2 latency instead of HBM latency.
        struct Cur { int tile, nt, c, t; };
        auto box_of = [&](const Cur &k, int &c0, int &r0) {
            // (a sub-tile wholly behind the last row still gets a box -- rows 0.. of the tensor, never used -- so that box
            //  indices, ring slots and barrier phases stay in step on both sides)
            int64_t row0 = (int64_t)k.tile * (DT * 128) + (int64_t)k.t * 128;
            if (row0 >= a.M) row0 = 0;
            c0 = k.nt * DBN + k.c; r0 = (int)row0;
        };
        auto advance = [&](Cur &k) {                                  // false behind the last box of this CTA
            const int ncol = a.N - k.nt * DBN < DBN ? a.N - k.nt * DBN : DBN;
            if (++k.t < DT) return true;
            k.t = 0;
            if ((k.c += 32) < ncol) return true;
            k.c = 0;
            if (++k.nt < d.nt) return true;
            k.nt = 0;
            return (k.tile += (int)gridDim.x) < d.n_tiles;
        };
        constexpr int D_RES_AHEAD = 8;
        const uint32_t res_u32 = smem_u32(res_ring);
        Cur cur = {(int)blockIdx.x, 0, 0, 0}, pf = cur;
        bool more = cur.tile < d.n_tiles, pf_more = more;
        auto prefetch = [&]() {
            if (!pf_more) return;
            int c0, r0;
            box_of(pf, c0, r0);
            if (elect_one())
                asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                             ::"l"(reinterpret_cast<uint64_t>(&maps.res)), "r"(c0), "r"(r0) : "memory");
            __syncwarp();
            pf_more = advance(pf);
        };
        for (int i = 0; i < D_RES_AHEAD; ++i) prefetch();
        uint32_t ck = 0;                                              // running (tile, pass, half) count, as in the epilogue
        while (more) {
            // the consumers' slot rule: group = sub-tile (DT = 2) or 32-channel half (DT = 1), k = that group's box count
            const uint32_t half = (uint32_t)d.rs >> 1, grp = DT == 2 ? (uint32_t)cur.t : (uint32_t)(cur.c >> 5), k = DT == 2 ? ck : ck >> 1;
            const uint32_t sl = grp * half + k % half, use = k / half;
            int c0, r0;
            box_of(cur, c0, r0);
            mbar_wait(&sm->res_empty[sl], (use & 1u) ^ 1u, 7);
            if (elect_one()) {
                const uint32_t bar = smem_u32(&sm->res_full[sl]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)D_RES_BOX) : "memory");
                tma_load_2d(res_u32 + sl * (uint32_t)D_RES_BOX, &maps.res, c0, r0, bar);
            }
            __syncwarp();
            prefetch();
            if (cur.t == DT - 1) ++ck;                                // the last sub-tile of a (tile, pass, half) chunk
            more = advance(cur);
        }
    } else {
        // ============================ slab fill ============================
        const int ftid = tid - first_fill * 32;
        const int j = ftid & 7;                           // 16-byte chunk = channels 8j .. 8j+7
        // fused first layer: the second affine (behind the ReLU) is per channel; alpha / beta are per (patch, channel)
        float fs2[8], ft2[8];
        if constexpr (MODE == DIN_FIRST) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = j * 8 + e;
                fs2[e] = (ff.flags & ISS_F_AFFINE_POST) ? ff.post_scale[c] : 1.f; ft2[e] = (ff.flags & ISS_F_AFFINE_POST) ? ff.post_shift[c] : 0.f;
            }
        }
        int i = 0;
        for (int tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x, ++i) {
            const int buf = NBUF == 2 ? (i & 1) : 0;
            const uint32_t use = (uint32_t)(NBUF == 2 ? (i >> 1) : i) & 1u;
            const int tbuf = i & 1;                                     // the image table alternates even with one slab buffer
            if (lane == 0) mbar_wait(&sm->slab_empty[buf], use ^ 1u, 6);
            __syncwarp();
            const uint32_t s0 = (uint32_t)tile * (DT * 128u);          // global pixel of slab row 0
            const uint32_t hi_base = slab_u32 + (uint32_t)buf * slab_buf;
            if constexpr (MODE == DIN_FIRST) {
                const uint32_t img0 = s0 / HW;
                if (ftid < D_TAB) {
                    const int64_t img = (int64_t)img0 + ftid;
                    sm->tab_row[tbuf][ftid] = img < ff.n_img ? (long long)ff.row0[img] - ff.y_f0 : -1ll;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(nfill) : "memory");
                const int f_flags = ff.flags;
                const int64_t row_len = (int64_t)a.W * HBK;            // floats per Y row
                // this thread's pixels: slab rows (ftid >> 3) + 32 k; (t, ih, x) = image relative to img0, input row, column
                uint32_t t, ih, x;
                {
                    const uint32_t gp = s0 + (uint32_t)(ftid >> 3);
                    const uint32_t img = gp / HW, rem = gp - img * HW;
                    t = img - img0; ih = rem / (uint32_t)a.W; x = rem - ih * (uint32_t)a.W;
                }
                int cur_t = -1;
                float al[8], bh[8], bl[8];                              // alpha, beta_hi, beta_lo of image cur_t, channels 8j..8j+7
                constexpr int FB = 2;                                   // pixels whose loads are in flight together
                const int PSTEP = nfill >> 3;
                for (int pl0 = ftid >> 3; pl0 < d.npix; pl0 += FB * PSTEP) {
                    u32x8 yh[FB], yl[FB];
                    int ti[FB];
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        ti[f] = -1;                                     // -1: nothing to write, D_TAB: zero fill
                        if (pl0 + f * PSTEP < d.npix) {
                            ti[f] = D_TAB;
                            if (t < (uint32_t)D_TAB && sm->tab_row[tbuf][t] >= 0) {
                                ti[f] = (int)t;
                                const int64_t o = (sm->tab_row[tbuf][t] + ih) * row_len + x * HBK + j * 8;
                                yh[f] = ldg256(ff.Yh + o); yl[f] = ldg256(ff.Yl + o);
                            }
                            x += PSTEP;
                            while (x >= (uint32_t)a.W) { x -= (uint32_t)a.W; if (++ih == (uint32_t)a.H) { ih = 0; ++t; } }
                        }
                    }
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        if (ti[f] < 0) continue;
                        const int pl = pl0 + f * PSTEP;
                        uint32_t hw[4] = {0u, 0u, 0u, 0u}, lw[4] = {0u, 0u, 0u, 0u};
                        if (ti[f] < D_TAB) {
                            if (ti[f] != cur_t) {
                                cur_t = ti[f];
                                const float4 *cp = reinterpret_cast<const float4 *>(ff.coef + ((int64_t)img0 + cur_t) * (3 * HBK) + j * 8);
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const float4 va = __ldg(cp + q), vh = __ldg(cp + HBK / 4 + q), vl = __ldg(cp + 2 * (HBK / 4) + q);
                                    al[4 * q] = va.x; al[4 * q + 1] = va.y; al[4 * q + 2] = va.z; al[4 * q + 3] = va.w;
                                    bh[4 * q] = vh.x; bh[4 * q + 1] = vh.y; bh[4 * q + 2] = vh.z; bh[4 * q + 3] = vh.w;
                                    bl[4 * q] = vl.x; bl[4 * q + 1] = vl.y; bl[4 * q + 2] = vl.z; bl[4 * q + 3] = vl.w;
                                }
                            }
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float u = fmaf(__uint_as_float(yh[f].v[e]), al[e], bh[e]) + fmaf(__uint_as_float(yl[f].v[e]), al[e], bl[e]);
                                if (f_flags & ISS_F_RELU) u = fmaxf(u, 0.f);
                                if (f_flags & ISS_F_AFFINE_POST) u = fmaf(u, fs2[e], ft2[e]);
                                v[e] = u;
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) {               // hi = fp16(v), lo = fp16(v - hi): two channels per word
                                const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                                const float2 hf = __half22float2(hh);
                                const __half2 ll = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                                hw[e] = *reinterpret_cast<const uint32_t *>(&hh);
                                lw[e] = *reinterpret_cast<const uint32_t *>(&ll);
                            }
                        }
                        const uint32_t dst = hi_base + (uint32_t)pl * 128u + (uint32_t)((j ^ (pl & 7)) << 4);
                        sts128(dst, hw[0], hw[1], hw[2], hw[3]);
                        sts128(dst + plane, lw[0], lw[1], lw[2], lw[3]);
                    }
                }
            } else if constexpr (MODE == DIN_POOL) {
                // task = (pooled pixel pl, chunk jj of 8 channels): the 2x2 window of the un-pooled NHWC tensor [img][PH][PW][C],
                // the WORD of the largest value is kept (as maxpool_nhwc_packed_kernel does: NaN wins, first maximum wins)
                const uint4 *src = reinterpret_cast<const uint4 *>(a.in);
                const int cshift = 3 + (d.cb == 1 ? 0 : d.cb == 2 ? 1 : 2), cpp = 1 << cshift;
                const int total = d.npix << cshift;
                constexpr int FB = 2;                                   // tasks whose 8 loads each are in flight together (HBM latency)
                for (int idx0 = ftid; idx0 < total; idx0 += FB * nfill) {
                    u32x8 u[FB][4];                                     // the 2x2 window: 8 channels (32 bytes) of each pixel
                    bool live[FB];
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        const int idx = idx0 + f * nfill;
                        const int pl = idx >> cshift, jj = idx & (cpp - 1);
                        const uint32_t gp = s0 + (uint32_t)pl;
                        live[f] = idx < total && (int64_t)gp < d.total_pix;
                        if (live[f]) {
                            const uint32_t img = gp / HW, rem = gp - img * HW;
                            const uint32_t ih = rem / (uint32_t)a.W, x = rem - ih * (uint32_t)a.W;
                            const uint4 *p00 = src + (((int64_t)img * a.pool_h + 2 * ih) * a.pool_w + 2 * x) * (2 * cpp) + jj * 2;
                            const uint4 *p10 = p00 + (int64_t)a.pool_w * (2 * cpp);
                            u[f][0] = ldg256(p00); u[f][1] = ldg256(p00 + 2 * cpp); u[f][2] = ldg256(p10); u[f][3] = ldg256(p10 + 2 * cpp);
                        }
                    }
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        const int idx = idx0 + f * nfill;
                        if (idx >= total) continue;
                        const int pl = idx >> cshift, jj = idx & (cpp - 1);
                        uint32_t bw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                        if (live[f]) {
                            float bv[8];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {               // window order (0,0) (0,1) (1,0) (1,1), as the pooling kernel
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float fv = iss_unpack_split(u[f][q].v[e]);
                                    if (q == 0 || fv > bv[e] || fv != fv) { bv[e] = fv; bw[e] = u[f][q].v[e]; }
                                }
                            }
                        }
                        const uint32_t dst = hi_base + (uint32_t)(jj >> 3) * 2u * plane + (uint32_t)pl * 128u + (uint32_t)(((jj & 7) ^ (pl & 7)) << 4);
                        sts128(dst, __byte_perm(bw[0], bw[1], 0x5410), __byte_perm(bw[2], bw[3], 0x5410), __byte_perm(bw[4], bw[5], 0x5410), __byte_perm(bw[6], bw[7], 0x5410));
                        sts128(dst + plane, __byte_perm(bw[0], bw[1], 0x7632), __byte_perm(bw[2], bw[3], 0x7632), __byte_perm(bw[4], bw[5], 0x7632), __byte_perm(bw[6], bw[7], 0x7632));
                    }
                }
            } else {
                // task = (pixel pl, chunk jj of 8 channels): two 16-byte loads (split-half words, or fp32 values split here)
                // -> one 16-byte chunk of the hi plane and one of the lo plane of channel block jj >> 3; chunks past the
                // real channel count (C = 32 padded to 64) are zero
                // DIN_PAD ('same' 3x3, padding 1): the tall image is the PADDED one -- a.H = H + 1 rows of a.W = W + 1 pixels per
                // image, row 0 and column 0 zero; the zero column of the next row / zero row of the next image serve as this
                // row's right / this image's bottom padding, so the tap windows stay plain row shifts -- and pixel (r, cc) of it
                // is pixel (r - 1, cc - 1) of the un-padded NHWC input
                const uint4 *src = reinterpret_cast<const uint4 *>(a.in);
                const int cshift = 3 + (d.cb == 1 ? 0 : d.cb == 2 ? 1 : 2), cpp = 1 << cshift;      // chunks per pixel
                const int creal = d.c_real >> 3, q4 = d.c_real >> 2;                                // real chunks / uint4 per pixel
                const int total = d.npix << cshift;
                const int64_t Hr = a.H - 1, Wr = a.W - 1;               // (DIN_PAD) the un-padded input
                constexpr int FB = 4;                                   // tasks whose loads are in flight together
                for (int idx0 = ftid; idx0 < total; idx0 += FB * nfill) {
                    u32x8 u[FB];
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        const int idx = idx0 + f * nfill;
#pragma unroll
                        for (int e = 0; e < 8; ++e) u[f].v[e] = 0u;
                        if (idx < total) {
                            const int pl = idx >> cshift, jj = idx & (cpp - 1);
                            int64_t gp = (int64_t)s0 + pl;
                            bool live = gp < d.total_pix && jj < creal;
                            if constexpr (MODE == DIN_PAD) {
                                if (live) {
                                    const uint32_t img = (uint32_t)gp / HW, rem = (uint32_t)gp - img * HW;
                                    const uint32_t r = rem / (uint32_t)a.W, cc = rem - r * (uint32_t)a.W;
                                    live = r >= 1u && cc >= 1u;
                                    gp = ((int64_t)img * Hr + (int64_t)r - 1) * Wr + (int64_t)cc - 1;
                                }
                            }
                            if (live) u[f] = ldg256(src + gp * q4 + jj * 2);
                        }
                    }
#pragma unroll
                    for (int f = 0; f < FB; ++f) {
                        const int idx = idx0 + f * nfill;
                        if (idx >= total) continue;
                        const int pl = idx >> cshift, jj = idx & (cpp - 1);
                        const uint32_t dst = hi_base + (uint32_t)(jj >> 3) * 2u * plane + (uint32_t)pl * 128u + (uint32_t)(((jj & 7) ^ (pl & 7)) << 4);
                        if constexpr (MODE == DIN_F32) {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {               // hi = fp16(v), lo = fp16(v - hi): two channels per word
                                const float v0 = __uint_as_float(u[f].v[2 * e]), v1 = __uint_as_float(u[f].v[2 * e + 1]);
                                const __half2 hh = __floats2half2_rn(v0, v1);
                                const float2 hf = __half22float2(hh);
                                const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                                hw[e] = *reinterpret_cast<const uint32_t *>(&hh);
                                lw[e] = *reinterpret_cast<const uint32_t *>(&ll);
                            }
                            sts128(dst, hw[0], hw[1], hw[2], hw[3]);
                            sts128(dst + plane, lw[0], lw[1], lw[2], lw[3]);
                        } else {
                            sts128(dst, __byte_perm(u[f].v[0], u[f].v[1], 0x5410), __byte_perm(u[f].v[2], u[f].v[3], 0x5410), __byte_perm(u[f].v[4], u[f].v[5], 0x5410), __byte_perm(u[f].v[6], u[f].v[7], 0x5410));
                            sts128(dst + plane, __byte_perm(u[f].v[0], u[f].v[1], 0x7632), __byte_perm(u[f].v[2], u[f].v[3], 0x7632), __byte_perm(u[f].v[4], u[f].v[5], 0x7632), __byte_perm(u[f].v[6], u[f].v[7], 0x7632));
                        }
                    }
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

int direct_npix(const ConvArgs &a, int dt)
{
    const int npix = dt * 128 + (a.KH - 1) * a.W + a.KW - 1;
    return (npix + 7) & ~7;
}

int direct_cb(const ConvArgs &a) { return (a.C + HBK - 1) / HBK; }

size_t direct_smem(const ConvArgs &a, int nbuf, int dt)
{
    return (size_t)DSB * D_B_STAGE + (size_t)nbuf * 2 * direct_cb(a) * direct_npix(a, dt) * 128 + sizeof(DSmem) + 1024 + 16 + (size_t)4 * a.N * sizeof(float);
}

// Tiling of a layer: {slab buffers, 128-slot sub-tiles per tile}; nbuf = 0: the layer does not fit.
//  * two slab buffers and 256-slot tiles when they fit (fill of tile i+1 under the MMAs of tile i, weight stages shared by two
//    sub-tiles);
//  * 1x1 layers whose two buffers only fit with 128-row tiles take those: they are memory-bound, and with ONE buffer the fill
//    of the next A tile (128 KB for C = 128) cannot start before the last pass of the current one -- measured on the
//    128 -> 512 expand layers of ResNet101: fill ~13 us + 8 passes ~14 us per tile, nothing overlapped;
//  * one buffer is accepted only when a tile has >= 4 n-tile passes to amortise the exposed fill (on the 3x3 128 -> 128
//    layer, 2 passes, one buffer measured slower than the TMEM-operand slab kernel: 286 vs 208 us; 1x1 layers with C = 128
//    and N = 32 / 64 measured 273 / 292 us with one buffer against 230 / 262 us on the gather kernels -- gpurun_out r02j).
struct DirectPlan { int nbuf, dt; };
DirectPlan direct_plan(const ConvArgs &a)
{
    const char *e = getenv("ISS_B200_DIRECT_DT1");                       // A/B runs: 0 = never use 128-row tiles
    const bool dt1 = !(e && e[0] == '0');
    if (direct_smem(a, 2, 2) <= (size_t)D_SMEM_MAX) return {2, 2};
    if (dt1 && a.KH * a.KW == 1 && !a.first && a.pool_h == 0 && a.N / DBN >= 4 && direct_smem(a, 2, 1) <= (size_t)D_SMEM_MAX) return {2, 1};
    if (direct_smem(a, 1, 2) <= (size_t)D_SMEM_MAX && a.N / DBN >= 4) return {1, 2};
    return {0, 0};
}

// 'same' 3x3 / stride-1 convolution with padding 1 (ResNet101's 3x3 layers, Keras padding='same'): runs as the un-padded
// convolution of the padded tall image (H + 1 rows of W + 1 pixels per image, see the DIN_PAD fill) -- 92 % useful slots
// on the 16 x 36 maps of ResNet101's stage 3, where the gather kernel spends its time on im2col copies.
// ISS_B200_DIRECT_PAD=1 turns it on (validated on hardware, parity identical; with one slab buffer and two passes the exposed
// fill makes it 0.6 % slower than the gather kernel on ResNet101, so it is not the default).
bool direct_is_same3x3(const ConvArgs &a)
{
    const char *e = getenv("ISS_B200_DIRECT_PAD");                       // opt-in: measured 0.6 % SLOWER than the gather kernel on ResNet101 (r02s)
    if (!(e && e[0] == '1')) return false;
    return a.KH == 3 && a.KW == 3 && a.SH == 1 && a.SW == 1 && a.PT == 1 && a.PL == 1 && a.OH == a.H && a.OW == a.W && a.in_packed && !a.first &&
           a.pool_h == 0 && !(a.flags & ISS_F_RESIDUAL);
}
ConvArgs direct_padded_view(const ConvArgs &a)
{
    ConvArgs b = a;
    b.H = a.H + 1; b.W = a.W + 1; b.PT = 0; b.PL = 0;                    // OH / OW stay the real output dims
    return b;
}
// (one slab buffer, two n-tile passes: accepted for the padded 3x3 layers -- on ResNet101's 128 -> 128 layers the exposed
//  fill costs less than the gather kernel's producers)
DirectPlan direct_plan_padded(const ConvArgs &b)
{
    if (direct_smem(b, 2, 2) <= (size_t)D_SMEM_MAX) return {2, 2};
    if (direct_smem(b, 1, 2) <= (size_t)D_SMEM_MAX && b.N / DBN >= 2) return {1, 2};
    return {0, 0};
}

// ---- tensor maps for the TMA mode of the residual layers -------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point: fetched through the runtime so that the library does not link libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled()
{
    static const EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// [rows][n] 32-bit words, box = 32 words (128 bytes, SWIZZLE_128B) x 128 rows; elements behind the last row read as zero and
// are not written
bool make_rows_map(CUtensorMap *m, const void *base, int64_t rows, int n, int box_rows = 128)
{
    const EncodeTiledFn enc = encode_tiled();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || n % 32 != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE, int NBUF, int DT>
int launch_tc4h(const ConvArgs &a, const FirstFuse &ff, DirectArgs d, unsigned grid, cudaStream_t st)
{
    auto kern = conv_gemm_tc4h_kernel<MODE, NBUF, DT>;
    ISS_CUDA_OK(iss_optin_smem(reinterpret_cast<const void *>(kern), D_SMEM_MAX));
    DirectMaps maps;
    memset(&maps, 0, sizeof(maps));
    size_t smem = direct_smem(a, NBUF, DT);
    if (d.tma) {
        // the ring takes what is left of the shared memory: at least 2 boxes, else the mode is off for this layer
        const int fit = (int)(((size_t)D_SMEM_MAX - smem) / D_RES_BOX);
        d.rs = (fit < D_RES_MAX ? fit : D_RES_MAX) & ~1;               // one half per group of four epilogue warps
        if (d.rs < 2 || !make_rows_map(&maps.res, a.residual, a.M, a.N) || ((d.tma & 2) && !make_rows_map(&maps.out, a.out, a.M, a.N, 32))) { d.tma = 0; d.rs = 0; }
        else smem += (size_t)d.rs * D_RES_BOX;
    }
    kern<<<grid, D_THREADS, smem, st>>>(a, ff, d, maps);
    ISS_CUDA_OK(cudaGetLastError());
    iss_count_launch();
    return ISS_OK;
}

}  // namespace

// Does the direct kernel take this layer?  Un-padded stride-1 KHxKW convolution (1x1 / Dense included) with C in {64, 128,
// 256} input channels (1x1 also C = 32, zero-padded to one k-block) and N = 64 .. 512 (multiple of 64) output channels;
// input split-half words, fp32, or (C = 64) the fused first layer; optional residual.
// ISS_B200_F16_DIRECT=0 turns it off (A/B runs against the TMEM-operand kernels).
bool iss_conv_f16_direct_covers(const ConvArgs &a)
{
    const char *env = getenv("ISS_B200_F16_DIRECT");                     // read per call: the GPU tests toggle it
    if (env && env[0] == '0') return false;
    if (a.wt_f16 && direct_is_same3x3(a)) {
        const ConvArgs b = direct_padded_view(a);
        const bool c_ok = a.C == 64 || a.C == 128 || a.C == 256;
        if (!c_ok || a.N % DBN != 0 || a.N > D_NMAX || a.K != 9 * a.C || a.Kp != a.K || (a.flags & ISS_F_SIGMOID)) return false;
        const int64_t n_img = a.M / ((int64_t)a.OH * a.OW);
        if (n_img * b.H * b.W >= (1ll << 31) - 4096) return false;
        return direct_plan_padded(b).nbuf > 0;
    }
    if (!a.wt_f16 || a.SH != 1 || a.SW != 1 || a.PT != 0 || a.PL != 0) return false;
    if (a.OH != a.H - a.KH + 1 || a.OW != a.W - a.KW + 1 || a.Kp != a.K) return false;
    const bool c_ok = a.C == 64 || a.C == 128 || a.C == 256 || (a.C == 32 && a.KH * a.KW == 1);
    const bool n_ok = (a.N % DBN == 0 && a.N <= D_NMAX) || (a.N == 32 && a.KH * a.KW == 1);     // N = 32: one n-tile, upper half masked
    if (!n_ok || !c_ok || a.K != a.KH * a.KW * a.C) return false;
    if (a.flags & ISS_F_SIGMOID) return false;
    if (a.first && (a.C != HBK || a.pool_h > 0 || (a.flags & ISS_F_RESIDUAL))) return false;
    if (a.pool_h > 0 && (!a.in_packed || a.pool_h / 2 != a.H || a.pool_w / 2 != a.W)) return false;   // fused 2x2 / stride-2 'valid' pooling only
    const int64_t n_img = a.M / ((int64_t)a.OH * a.OW);
    if (n_img * a.H * a.W >= (1ll << 31) - 4096) return false;
    // the IN_FIRST table covers D_TAB images per slab
    if (a.first && (direct_npix(a, 2) - 1) / (a.H * a.W) + 2 > D_TAB) return false;
    return direct_plan(a).nbuf > 0;
}

// Returns 1 when the layer is not covered (caller continues with the other fp16-split kernels).
int iss_launch_conv_tc_f16d(ConvArgs &a_in, cudaStream_t st)
{
    if (!iss_conv_f16_direct_covers(a_in)) return 1;
    const bool pad = direct_is_same3x3(a_in);
    ConvArgs a = pad ? direct_padded_view(a_in) : a_in;                  // (padded: the kernel sees the un-padded convolution of the padded image)
    const DirectPlan plan = pad ? direct_plan_padded(a) : direct_plan(a);
    DirectArgs d = {};
    d.wt = reinterpret_cast<const unsigned char *>(a.wt_f16);
    d.inv_scale = a.wt_f16_inv_scale;
    d.npix = direct_npix(a, plan.dt);
    d.n_img = (int)(a.M / ((int64_t)a.OH * a.OW));
    d.total_pix = (int64_t)d.n_img * a.H * a.W;
    d.cb = direct_cb(a);
    d.c_real = a.C;
    d.nt = (a.N + DBN - 1) / DBN;
    d.bn_img = iss_f16_bn_for(a.N == 32 ? 64 : a.N);                     // tiling of the weight image (iss_prepare_f16_weights)
    const char *nepi_env = getenv("ISS_B200_NEPI");                      // A/B runs: 4 or 8
    d.n_epi = nepi_env ? (atoi(nepi_env) == 8 ? 8 : 4) : ((a.flags & ISS_F_RESIDUAL) ? 8 : 4);
    // TMA mode of the residual 1x1 layers (ISS_B200_TMA_EPI: 0 off, 1 residual boxes by TMA, 3 + outputs by TMA store)
    {
        // default 0: the mode is validated bit-identical (tests/test_vbx.py) and 1.7 % faster (107.9 vs 106.1 TFLOP/s on ResNet101, r02v),
        // but its first form had a write-after-read race on the ring (see the epilogue) that the single-window checks of the time
        // did not see; it stays opt-in until it has more hardware mileage
        const char *te = getenv("ISS_B200_TMA_EPI");
        const int want = te ? atoi(te) & 3 : 0;
        const bool ok = (a.flags & ISS_F_RESIDUAL) && a.KH * a.KW == 1 && a.N % 32 == 0 && d.n_epi == 8 && !a.first && a.pool_h == 0;
        d.tma = ok ? ((want & 1) ? want : 0) : 0;
    }
    const int64_t total_slots = ((int64_t)(d.n_img - 1) * a.H + a.OH - 1) * a.W + a.OW;
    d.n_tiles = (int)((total_slots + plan.dt * 128 - 1) / (plan.dt * 128));
    int dev = 0, sms = 0;
    ISS_CUDA_OK(cudaGetDevice(&dev));
    ISS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(d.n_tiles < sms ? d.n_tiles : sms);
    const bool two = plan.nbuf == 2;
    FirstFuse ff = {};
    if (a.first) {
        ff = *a.first;
        return two ? launch_tc4h<DIN_FIRST, 2, 2>(a, ff, d, grid, st) : launch_tc4h<DIN_FIRST, 1, 2>(a, ff, d, grid, st);
    }
    if (a.pool_h > 0) return two ? launch_tc4h<DIN_POOL, 2, 2>(a, ff, d, grid, st) : launch_tc4h<DIN_POOL, 1, 2>(a, ff, d, grid, st);
    if (pad) return two ? launch_tc4h<DIN_PAD, 2, 2>(a, ff, d, grid, st) : launch_tc4h<DIN_PAD, 1, 2>(a, ff, d, grid, st);
    if (plan.dt == 1)                                                    // (128-row tiles: 1x1 layers only, always two buffers)
        return a.in_packed ? launch_tc4h<DIN_PACKED, 2, 1>(a, ff, d, grid, st) : launch_tc4h<DIN_F32, 2, 1>(a, ff, d, grid, st);
    if (!a.in_packed) return two ? launch_tc4h<DIN_F32, 2, 2>(a, ff, d, grid, st) : launch_tc4h<DIN_F32, 1, 2>(a, ff, d, grid, st);
    return two ? launch_tc4h<DIN_PACKED, 2, 2>(a, ff, d, grid, st) : launch_tc4h<DIN_PACKED, 1, 2>(a, ff, d, grid, st);
}
