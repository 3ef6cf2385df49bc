// umma_contention_bench.cu -- does the tcgen05.mma stream of the convolution kernel slow down when the
// producer warps run next to it?  (round-2 question of DESIGN.md 4.1: in situ the 8 MMAs of a k-block take
// ~750-835 cycles against 392 stand-alone.)  Warp 4 issues the kernel's MMA pattern (TS N=128 + TS N=64,
// warp-uniform elect-based loop, all operands resident); warps 0-3 run one kind of background load until
// the issuer is done:
//   0 none | 1 tcgen05.st (2 x 32 columns) + wait::st | 2 LDS.128 x 8 (lane = row, swizzled, conflict-free)
//   3 cp.async 16 B x 8 from global + wait | 4 one 16 KB cp.async.bulk at a time (thread 0) | 5 = 1 + 2
//   6 = 1 + 2 + mbarrier arrive/try_wait ping (the hand-shake instructions, no dependency on the MMAs)
// Output: cycles per k-step (floor 96) at 1 and 2 CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_contention_bench umma_contention_bench.cu && ./umma_contention_bench
// STATUS: written at the end of round 1 (GPU budget spent): compiles, not yet run.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t a)
{
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

constexpr int SM_B = 0, SM_SLAB = 16 * 1024, SM_DST = 48 * 1024, SM_MISC = 80 * 1024;      // 16 KB B tile, 32 KB slab, 32 KB copy target

template <int LOAD>
__global__ void __launch_bounds__(160) bench(int iters, long long *out, const float *gsrc, float *sink)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + SM_MISC);            // [0] MMA done, [1] bulk copy, [2..5] per-warp ping
    uint32_t *slot = (uint32_t *)(bar + 8);
    volatile int *stop = (volatile int *)(slot + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < SM_MISC / 4; i += 160) ((float *)smem)[i] = 0.0f;
    if (tid == 0) {
        *stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar[2 + w])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_any = *slot;
    // TMEM columns: D_main|D_lo 0..191, A (hi and lo share) 192..223, background store target 224..255
    if (warp < 4) {
        const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
        for (int c = 192; c < 256; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm_any + lane_addr + c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        uint32_t z[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) z[c] = 0;
        const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
        float acc = 0.f;
        uint32_t ping = 0;
        long long n = 0;
        for (;;) {
            if (__shfl_sync(0xffffffffu, *stop, 0)) break;      // warp-uniform exit
            ++n;
            if (LOAD == 1 || LOAD == 5 || LOAD == 6) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                             "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                             "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                             ::"r"(tm_any + lane_addr + 224),
                               "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]),
                               "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]),
                               "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]), "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]),
                               "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]), "r"(z[30]), "r"(z[31]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                             "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                             "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                             ::"r"(tm_any + lane_addr + 224),
                               "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]),
                               "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]),
                               "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]), "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]),
                               "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]), "r"(z[30]), "r"(z[31]) : "memory");
            }
            if (LOAD == 2 || LOAD == 5 || LOAD == 6) {
                const unsigned char *base = smem + SM_SLAB + warp * 8192 + (n & 1) * 4096;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = *reinterpret_cast<const float4 *>(base + lane * 128 + ((j ^ (lane & 7)) << 4));
                    acc += q.x + q.y + q.z + q.w;
                }
            }
            if (LOAD == 1 || LOAD == 5 || LOAD == 6) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (LOAD == 3) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + SM_DST + warp * 8192 + (j * 32 + lane) * 16)),
                                 "l"(gsrc + ((size_t)blockIdx.x * 65536 + (size_t)((n * 8 + j) & 1023) * 512 + warp * 128 + lane * 4)) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            if (LOAD == 4 && tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(16384u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + SM_DST)), "l"(gsrc + (size_t)blockIdx.x * 65536 + (size_t)(n & 7) * 4096), "r"(16384u), "r"(smem_u32(&bar[1])) : "memory");
                while (!try_wait(&bar[1], (uint32_t)((n - 1) & 1))) { }
            }
            if (LOAD == 6) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[2 + warp])) : "memory");
                while (!try_wait(&bar[2 + warp], ping & 1)) { }
                ++ping;
            }
        }
        if (acc == 123.456f) sink[0] = acc;
        if (tid == 0) out[gridDim.x + blockIdx.x] = n;
    } else {
        const uint32_t tm = __reduce_or_sync(0xffffffffu, tm_any);
        const uint64_t db = sw128_desc(smem_u32(smem + SM_B));
        const uint32_t ta = tm + 192;
        constexpr uint32_t i64 = (1u << 4) | (2u << 7) | (2u << 10) | (8u << 17) | (8u << 24);
        constexpr uint32_t i128 = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    mma_ts(tm, ta + kk * 8, db + 2 * kk, i128, 1);
                    mma_ts(tm + 64, ta + kk * 8, db + 2 * kk, i64, 1);
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        __syncwarp();
        while (!try_wait(&bar[0], 0)) { }
        const long long t1 = clock64();
        *stop = 1;
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_any), "r"(256) : "memory");
}

template <int LOAD>
void run(int ctas, int iters, long long *d, const float *gsrc, float *sink, const char *name)
{
    cudaFuncSetAttribute(bench<LOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    bench<LOAD><<<ctas, 160, 84 * 1024, 0>>>(iters, d, gsrc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("load %d: %s\n", LOAD, cudaGetErrorString(e)); return; }
    static long long h[8192];
    cudaMemcpy(h, d, 2 * ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0, n = 0;
    for (int i = 0; i < ctas; ++i) { s += (double)h[i]; n += (double)h[ctas + i]; }
    printf("%d CTA/SM  %-46s %7.1f cycles per k-step (floor 96)   background iterations per k-block of 4 k-steps: %.2f\n",
           ctas / 148, name, s / ctas / iters, n / ctas / (iters / 4.0));
}

int main()
{
    long long *d;
    float *gsrc, *sink;
    cudaMalloc(&d, 8192 * sizeof(long long));
    cudaMalloc(&gsrc, (size_t)296 * 65536 * sizeof(float) + (1 << 20));
    cudaMemset(gsrc, 0, (size_t)296 * 65536 * sizeof(float) + (1 << 20));
    cudaMalloc(&sink, 16);
    const int iters = 8192;
    for (int ctas = 148; ctas <= 296; ctas += 148) {
        run<0>(ctas, iters, d, gsrc, sink, "no background load");
        run<1>(ctas, iters, d, gsrc, sink, "tcgen05.st 2x32 columns + wait::st");
        run<2>(ctas, iters, d, gsrc, sink, "8 x LDS.128 per lane");
        run<3>(ctas, iters, d, gsrc, sink, "8 x cp.async 16 B per lane + wait");
        run<4>(ctas, iters, d, gsrc, sink, "16 KB cp.async.bulk, one at a time");
        run<5>(ctas, iters, d, gsrc, sink, "tcgen05.st + LDS (producer work)");
        run<6>(ctas, iters, d, gsrc, sink, "tcgen05.st + LDS + mbarrier ping");
    }
    return 0;
}
