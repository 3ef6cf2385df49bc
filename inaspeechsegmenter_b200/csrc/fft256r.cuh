// fft256r.cuh -- warp-level 256-point complex FFT held in REGISTERS: 8 points per lane, ONE pass through shared
// memory (round 1's radix-4 Stockham made four; with float64 data that was the kernel's bound: 33 % of its
// shared-memory wavefronts were bank conflicts and the fp64 pipe sat at 22 %).
//
//   256 = 8 x 32:   X[k1 + 8 k2] = sum_n2 W256^(n2 k1) [ sum_n1 x[32 n1 + n2] W8^(n1 k1) ] W32^(n2 k2)
//   step 1  lane n2 holds x[32 n1 + n2], n1 = 0..7: 8-point DFT in registers, times W256^(n2 k1)
//   step 2  transpose through shared memory: lane L = (k1 = L >> 2, r = L & 3) receives A[k1][4 m + r], m = 0..7
//           (rows of 36 complex: every quarter-warp reads / writes 128 contiguous-or-64-byte-staggered bytes: no conflicts)
//   step 3  32 = 8 x 4 in the same way: 8-point DFT over m in registers, times W32^(r k2a), then the 4-point DFT over r
//           across the four lanes of a group with two rounds of shuffles
//   result  lane L holds X[k1 + 8 (k2a + 8 k2b)] in out[k2a], with k2b = bitrev2(L & 3)
#pragma once

// forward 8-point DFT (e^{-2 pi i n k / 8}) in place, natural order in and out
template <typename T>
__device__ __forceinline__ void dft8(T (&re)[8], T (&im)[8])
{
    const T h = (T)0.70710678118654752440;
    T ar[4], ai[4], br[4], bi[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        ar[n] = re[n] + re[n + 4]; ai[n] = im[n] + im[n + 4];
        br[n] = re[n] - re[n + 4]; bi[n] = im[n] - im[n + 4];
    }
    // b[n] *= W8^n :  W8^1 = (1 - i) / sqrt2,  W8^2 = -i,  W8^3 = (-1 - i) / sqrt2
    { const T x = br[1], y = bi[1]; br[1] = (x + y) * h; bi[1] = (y - x) * h; }
    { const T x = br[2], y = bi[2]; br[2] = y; bi[2] = -x; }
    { const T x = br[3], y = bi[3]; br[3] = (y - x) * h; bi[3] = -(x + y) * h; }
    // two 4-point DFTs: even outputs from a, odd outputs from b
    auto dft4 = [](const T (&cr)[4], const T (&ci)[4], T (&outr)[4], T (&outi)[4]) {
        const T s0r = cr[0] + cr[2], s0i = ci[0] + ci[2];
        const T s1r = cr[0] - cr[2], s1i = ci[0] - ci[2];
        const T s2r = cr[1] + cr[3], s2i = ci[1] + ci[3];
        const T s3r = ci[1] - ci[3], s3i = -(cr[1] - cr[3]);          // (c1 - c3) * (-i)
        outr[0] = s0r + s2r; outi[0] = s0i + s2i;
        outr[1] = s1r + s3r; outi[1] = s1i + s3i;
        outr[2] = s0r - s2r; outi[2] = s0i - s2i;
        outr[3] = s1r - s3r; outi[3] = s1i - s3i;
    };
    T er[4], ei[4], orr[4], oi[4];
    dft4(ar, ai, er, ei);
    dft4(br, bi, orr, oi);
#pragma unroll
    for (int k = 0; k < 4; ++k) { re[2 * k] = er[k]; im[2 * k] = ei[k]; re[2 * k + 1] = orr[k]; im[2 * k + 1] = oi[k]; }
}

template <typename T> struct Cplx { T x, y; };      // 16-byte (double) / 8-byte (float) shared-memory element

constexpr int FFT_ROW = 36;                         // complex elements per transpose row (32 + 4: rows 64 bytes apart mod 128)
constexpr int FFT_BUF = 8 * FFT_ROW;                // 288 complex per warp

// twA: [7][32] W256^(lane * k1), k1 = 1..7;  twB: [7][4] W32^(r * k2a), k2a = 1..7 (both (cos, -sin) pairs)
// in: zr/zi[n1] = z[32 n1 + lane].  out: zr/zi[k2a] = Z[k1 + 8 (k2a + 8 k2b)], k1 = lane >> 2, k2b = bitrev2(lane & 3).
template <typename T>
__device__ __forceinline__ void warp_fft256_reg(T (&zr)[8], T (&zi)[8], Cplx<T> *buf, const Cplx<T> *twA, const Cplx<T> *twB, int lane)
{
    dft8<T>(zr, zi);
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) {
        const Cplx<T> w = twA[(k1 - 1) * 32 + lane];
        const T x = zr[k1], y = zi[k1];
        zr[k1] = x * w.x - y * w.y;
        zi[k1] = x * w.y + y * w.x;
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) { Cplx<T> v; v.x = zr[k1]; v.y = zi[k1]; buf[k1 * FFT_ROW + lane] = v; }
    __syncwarp();
    const int g = lane >> 2, r = lane & 3;
#pragma unroll
    for (int m = 0; m < 8; ++m) { const Cplx<T> v = buf[g * FFT_ROW + 4 * m + r]; zr[m] = v.x; zi[m] = v.y; }
    __syncwarp();
    dft8<T>(zr, zi);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const Cplx<T> w = twB[(k - 1) * 4 + r];
        const T x = zr[k], y = zi[k];
        zr[k] = x * w.x - y * w.y;
        zi[k] = x * w.y + y * w.x;
    }
    // 4-point DFT across lanes r = 0..3 (decimation in frequency): partner r ^ 2, then r ^ 1
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const T pr = __shfl_xor_sync(0xffffffffu, zr[k], 2), pi = __shfl_xor_sync(0xffffffffu, zi[k], 2);
        if (r < 2) { zr[k] += pr; zi[k] += pi; }
        else {
            const T dr = pr - zr[k], di = pi - zi[k];              // z[r-2] - z[r]
            if (r == 3) { zr[k] = di; zi[k] = -dr; }               // * W4^1 = -i
            else { zr[k] = dr; zi[k] = di; }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const T pr = __shfl_xor_sync(0xffffffffu, zr[k], 1), pi = __shfl_xor_sync(0xffffffffu, zi[k], 1);
        if ((r & 1) == 0) { zr[k] += pr; zi[k] += pi; }
        else { zr[k] = pr - zr[k]; zi[k] = pi - zi[k]; }
    }
}
