// fft256.cuh -- warp-level 256-point complex FFT (radix-4 Stockham, 4 passes through
// a warp-private shared buffer stored SoA and skewed by i + (i >> 5) against bank
// conflicts).  A 512-point real FFT is computed as this FFT of z[n] = x[2n] + i x[2n+1]
// followed by the even/odd split X[k] = E[k] + W512^k O[k].
#pragma once

// skew by one element per 128 bytes: every 32 floats / every 16 doubles
__device__ __forceinline__ int skew(int i) { return i + (i >> 5); }
template <typename T>
__device__ __forceinline__ int skewT(int i) { return sizeof(T) == 8 ? i + (i >> 4) : i + (i >> 5); }

// One radix-4 Stockham pass set over a warp-private 256-point complex buffer.
template <typename T>
__device__ __forceinline__ void warp_fft256(T *re, T *im, const T *tw, int lane)
{
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int Ns = 1 << (2 * s);
        T yr[2][4], yi[2][4];
        int j0[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = lane + 32 * b;
            const int k = j & (Ns - 1);
            const int step = k * (64 / Ns);
            T vr[4], vi[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int idx = skewT<T>(j + 64 * t);
                const T xr = re[idx], xi = im[idx];
                if (t == 0 || s == 0) { vr[t] = xr; vi[t] = xi; }
                else {
                    const T c = tw[2 * (t * step)], sn = tw[2 * (t * step) + 1];   // (cos, -sin)
                    vr[t] = xr * c - xi * sn;
                    vi[t] = xr * sn + xi * c;
                }
            }
            const T a0r = vr[0] + vr[2], a0i = vi[0] + vi[2];
            const T a1r = vr[0] - vr[2], a1i = vi[0] - vi[2];
            const T a2r = vr[1] + vr[3], a2i = vi[1] + vi[3];
            const T a3r = vi[1] - vi[3], a3i = -(vr[1] - vr[3]);     // (v1 - v3) * (-i)
            yr[b][0] = a0r + a2r; yi[b][0] = a0i + a2i;
            yr[b][1] = a1r + a3r; yi[b][1] = a1i + a3i;
            yr[b][2] = a0r - a2r; yi[b][2] = a0i - a2i;
            yr[b][3] = a1r - a3r; yi[b][3] = a1i - a3i;
            j0[b] = ((j - k) << 2) + k;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int idx = skewT<T>(j0[b] + t * Ns);
                re[idx] = yr[b][t];
                im[idx] = yi[b][t];
            }
        __syncwarp();
    }
}

