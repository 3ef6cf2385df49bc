// umma_desc_offset_test.cu -- can a K-major SWIZZLE_128B shared-memory operand of tcgen05.mma START at any 128-byte
// row of a 1024-byte-aligned buffer?  (Needed to feed a convolution's A operand straight from an input slab: the tile of
// filter tap (kh, kw) is the same pixel array shifted by kh * W + kw rows.)
//
// The buffer holds rows of 64 halves (128 B); 16-byte chunk j of row p is stored at chunk j ^ (p & 7) (the layout TMA's
// SWIZZLE_128B produces for a 1024-byte-aligned buffer, i.e. the XOR is a function of the ABSOLUTE row index).  For every
// shift s = 0..16 one MMA (M = 128, N = 64, K = 64: four K = 16 steps) is issued with the A descriptor's start address at
// row s, once with the descriptor's base-offset field = 0 and once with base-offset = s & 7, and D is compared with a CPU
// reference.  B is a plain 1024-byte-aligned tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_desc_offset_test umma_desc_offset_test.cu && ./umma_desc_offset_test
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t a, uint32_t base_off)
{
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}

constexpr int ROWS = 160;      // A buffer rows (128 + max shift + slack)

__global__ void __launch_bounds__(128) test_kernel(const __half *gA, const __half *gB, float *gD, int shift, int use_base_off)
{
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = sm;                       // ROWS x 128 B
    unsigned char *sB = sm + 24 * 1024;           // 64 x 128 B
    uint64_t *bar = (uint64_t *)(sm + 40 * 1024);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // swizzled fill: chunk j of row p at chunk j ^ (p & 7)
    for (int i = tid; i < ROWS * 8; i += 128) {
        const int p = i >> 3, j = i & 7;
        *(uint4 *)(sA + p * 128 + ((j ^ (p & 7)) << 4)) = *(const uint4 *)(gA + p * 64 + j * 8);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int n = i >> 3, j = i & 7;
        *(uint4 *)(sB + n * 128 + ((j ^ (n & 7)) << 4)) = *(const uint4 *)(gB + n * 64 + j * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (tid == 0) {
        // instruction descriptor: D = F32, A = B = F16, K-major both, N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(sA) + (uint32_t)shift * 128u;
        const uint64_t da = sw128_desc(a0, use_base_off ? (uint32_t)shift : 0u), db = sw128_desc(smem_u32(sB), 0u);
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = kk > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait (bounded)
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 2000000000ll) { if (tid == 0) printf("timeout\n"); break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    for (int c = 0; c < 64; c += 32) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) gD[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

int main()
{
    std::vector<__half> hA(ROWS * 64), hB(64 * 64);
    std::vector<float> fA(ROWS * 64), fB(64 * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (float)(rand() % 17 - 8) / 8.f; hA[i] = __float2half(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (float)(rand() % 13 - 6) / 4.f; hB[i] = __float2half(v); fB[i] = v; }
    __half *dA, *dB; float *dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    std::vector<float> hD(128 * 64);
    for (int use = 0; use < 2; ++use)
        for (int s = 0; s <= 17; ++s) {
            cudaMemset(dD, 0, 128 * 64 * 4);
            test_kernel<<<1, 128, 44 * 1024>>>(dA, dB, dD, s, use);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift %d base_off %d: %s\n", s, use, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0;
            int bad_rows = 0;
            for (int r = 0; r < 128; ++r) {
                double rowbad = 0;
                for (int n = 0; n < 64; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 64; ++k) ref += (double)fA[(s + r) * 64 + k] * fB[n * 64 + k];
                    rowbad = fmax(rowbad, fabs(ref - hD[r * 64 + n]));
                }
                worst = fmax(worst, rowbad);
                if (rowbad > 1e-3) ++bad_rows;
            }
            printf("shift %2d  base_offset field %s : max |D - ref| = %.3e  rows wrong: %d %s\n", s, use ? "= shift & 7" : "= 0        ", worst, bad_rows,
                   worst < 1e-3 ? "OK" : "MISMATCH");
        }
    return 0;
}
