// umma_microbench.cu -- cycles per tcgen05.mma (kind::tf32, M = 128, K = 8) issued back-to-back by one
// thread with all operands resident, for the shapes / accumulator patterns conv_gemm_tc.cu uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench umma_microbench.cu && ./umma_microbench
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

// variant: 0 TS N=64 same D | 1 TS N=128 same D | 2 TS N=256 same D | 3 TS N=64 alternating 2 D | 4 SS N=64 | 5 SS N=256
//          6 TS N=128 then N=64 (the conv kernel's pattern, both touch D_lo) | 7 same but the N=64 MMA goes to a third D
// The variant is a template parameter and the 4 k-steps are unrolled with constant offsets, so the loop
// body is just the MMA issues (what a tuned kernel's issue loop looks like).
template <int V, bool UNIFORM>
__global__ void __launch_bounds__(160) bench(int iters, long long *out)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + 96 * 1024);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 160) ((float *)smem)[i] = 0.0f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
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
    const uint32_t tm = *slot;
    if (warp < 4) {
        uint32_t z = 0;
        for (int c = 192; c < 256; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + c), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (UNIFORM ? (warp == 4) : (tid == 128)) {
        const bool leader = (tid == 128);
        const uint64_t da = sw128_desc(smem_u32(smem)), db = sw128_desc(smem_u32(smem + 16 * 1024));
        const uint32_t ta = tm + 192;
        constexpr uint32_t i64 = (1u << 4) | (2u << 7) | (2u << 10) | (8u << 17) | (8u << 24);
        constexpr uint32_t i128 = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);
        constexpr uint32_t i256 = (1u << 4) | (2u << 7) | (2u << 10) | (32u << 17) | (8u << 24);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (leader) {
                    if (V == 0) mma_ts(tm, ta + kk * 8, db + 2 * kk, i64, 1);
                    if (V == 1) mma_ts(tm, ta + kk * 8, db + 2 * kk, i128, 1);
                    if (V == 2) mma_ts(tm, ta + kk * 8, db + 2 * kk, i256, 1);
                    if (V == 3) mma_ts(tm + 64 * (kk & 1), ta + kk * 8, db + 2 * kk, i64, 1);
                    if (V == 4) mma_ss(tm, da + 2 * kk, db + 2 * kk, i64, 1);
                    if (V == 5) mma_ss(tm, da + 2 * kk, db + 2 * kk, i256, 1);
                    if (V == 6) { mma_ts(tm, ta + kk * 8, db + 2 * kk, i128, 1); mma_ts(tm + 64, ta + 32 + kk * 8, db + 2 * kk, i64, 1); }
                    if (V == 7) { mma_ts(tm, ta + kk * 8, db + 2 * kk, i128, 1); mma_ts(tm + 128, ta + 32 + kk * 8, db + 2 * kk, i64, 1); }
                }
            }
        }
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
        const long long t1 = clock64();
        if (leader) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

template <int V, bool U>
double run(int ctas, int iters, long long *d)
{
    cudaFuncSetAttribute(bench<V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    bench<V, U><<<ctas, 160, 98 * 1024, 0>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", V, cudaGetErrorString(e)); return -1; }
    static long long h[4096];
    cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < ctas; ++i) s += (double)h[i];
    return s / ctas / iters;
}

int main()
{
    const char *names[] = {"TS N=64 same D", "TS N=128 same D", "TS N=256 same D", "TS N=64 alternating D0/D1", "SS N=64 same D", "SS N=256 same D",
                           "TS N=128 + TS N=64 (conv pattern, shared D_lo)", "TS N=128 + TS N=64 (third accumulator)"};
    const double floor_[] = {32, 64, 128, 32, 32, 128, 96, 96};
    long long *d;
    cudaMalloc(&d, 4096 * sizeof(long long));
    const int iters = 8192;
    for (int ctas = 148; ctas <= 296; ctas += 148) {
        double r[2][8];
        r[0][0] = run<0, false>(ctas, iters, d); r[1][0] = run<0, true>(ctas, iters, d);
        r[0][1] = run<1, false>(ctas, iters, d); r[1][1] = run<1, true>(ctas, iters, d);
        r[0][2] = run<2, false>(ctas, iters, d); r[1][2] = run<2, true>(ctas, iters, d);
        r[0][3] = run<3, false>(ctas, iters, d); r[1][3] = run<3, true>(ctas, iters, d);
        r[0][4] = run<4, false>(ctas, iters, d); r[1][4] = run<4, true>(ctas, iters, d);
        r[0][5] = run<5, false>(ctas, iters, d); r[1][5] = run<5, true>(ctas, iters, d);
        r[0][6] = run<6, false>(ctas, iters, d); r[1][6] = run<6, true>(ctas, iters, d);
        r[0][7] = run<7, false>(ctas, iters, d); r[1][7] = run<7, true>(ctas, iters, d);
        for (int v = 0; v < 8; ++v)
            printf("%d CTA/SM  %-50s single-lane loop %6.1f | warp-uniform loop %6.1f cycles per k-step (tensor-time floor %.0f)\n",
                   ctas / 148, names[v], r[0][v], r[1][v], floor_[v]);
    }
    return 0;
}
