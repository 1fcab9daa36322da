// Synthetic-phantom generator on the device (benchmark utility; see phantom.py for the recipe and the
// NumPy twin).  Not part of the reference path: it exists so that batches of 512x512x256 scans can be
// created in HBM without crossing PCIe.
#include "common.cuh"

#include <math.h>

// Ellipsoid inside test in float32 with separately rounded operations, in the NumPy twin's order.
__global__ void __launch_bounds__(256) k_paint(uint16_t* __restrict__ vol, int nx, int ny, int x0, int y0, int z0, int bx,
                                               int by, int bz, float cx, float cy, float cz, float ax, float ay, float az,
                                               uint16_t value) {
    const long long n = (long long)bx * by * bz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int lx = int(i % bx), ly = int((i / bx) % by), lz = int(i / ((long long)bx * by));
        const int x = x0 + lx, y = y0 + ly, z = z0 + lz;
        float tx = __fdiv_rn(__fsub_rn(float(x), cx), ax), ty = __fdiv_rn(__fsub_rn(float(y), cy), ay),
              tz = __fdiv_rn(__fsub_rn(float(z), cz), az);
        float q = __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz));
        if (q <= 1.0f) vol[((size_t)z * ny + y) * nx + x] = value;
    }
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint16_t rician(uint16_t a, uint32_t w1, uint32_t w2, float sigma) {
    const float u1 = __fmul_rn(__fadd_rn(float(w1 >> 8), 0.5f), 5.9604644775390625e-08f);   // 2^-24
    const float u2 = __fmul_rn(__fadd_rn(float(w2 >> 8), 0.5f), 5.9604644775390625e-08f);
    const float rad = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(u1))), sigma);
    const float ang = __fmul_rn(3.14159274101257324f, __fmul_rn(2.0f, u2));
    const float n1 = __fmul_rn(rad, cosf(ang)), n2 = __fmul_rn(rad, sinf(ang));
    const float s = __fadd_rn(float(a), n1);
    float v = rintf(sqrtf(__fadd_rn(__fmul_rn(s, s), __fmul_rn(n2, n2))));
    v = fminf(fmaxf(v, 0.0f), 65535.0f);
    return uint16_t(v);
}

// One Philox call per voxel pair (see add_rician_noise in phantom.py).
__global__ void __launch_bounds__(256) k_noise(uint16_t* __restrict__ vol, size_t n, float sigma, uint32_t k0, uint32_t k1,
                                               uint32_t scan_index) {
    const size_t n_pairs = (n + 1) / 2;
    for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pairs; p += size_t(gridDim.x) * blockDim.x) {
        uint32_t r[4];
        philox4x32_10(uint32_t(p & 0xFFFFFFFFull), uint32_t(p >> 32), scan_index, 0x50484E54u, k0, k1, r);
        const size_t i = 2 * p;
        if (i + 1 < n) {
            const uint32_t in = *reinterpret_cast<const uint32_t*>(vol + i);
            const uint16_t a = rician(uint16_t(in & 0xFFFFu), r[0], r[1], sigma);
            const uint16_t b = rician(uint16_t(in >> 16), r[2], r[3], sigma);
            *reinterpret_cast<uint32_t*>(vol + i) = uint32_t(a) | (uint32_t(b) << 16);
        } else {
            vol[i] = rician(vol[i], r[0], r[1], sigma);
        }
    }
}

cudaError_t launch_phantom(uint16_t* d_volume, int nx, int ny, int nz, const float* h_ell, int n_ell, float sigma,
                           unsigned long long seed, unsigned int scan_index, cudaStream_t s) {
    const size_t n = size_t(nx) * ny * nz;
    cudaError_t e = cudaMemsetAsync(d_volume, 0, n * sizeof(uint16_t), s);
    if (e != cudaSuccess) return e;
    for (int k = 0; k < n_ell; ++k) {
        const float* p = h_ell + 7 * k;
        const float cx = p[0], cy = p[1], cz = p[2], ax = p[3], ay = p[4], az = p[5];
        int x0 = int(floorf(cx - ax)), x1 = int(ceilf(cx + ax));
        int y0 = int(floorf(cy - ay)), y1 = int(ceilf(cy + ay));
        int z0 = int(floorf(cz - az)), z1 = int(ceilf(cz + az));
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (z0 < 0) z0 = 0;
        if (x1 > nx - 1) x1 = nx - 1;
        if (y1 > ny - 1) y1 = ny - 1;
        if (z1 > nz - 1) z1 = nz - 1;
        if (x0 > x1 || y0 > y1 || z0 > z1) continue;
        const int bx = x1 - x0 + 1, by = y1 - y0 + 1, bz = z1 - z0 + 1;
        long long nb = (long long)bx * by * bz;
        long long blocks = (nb + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        k_paint<<<unsigned(blocks), 256, 0, s>>>(d_volume, nx, ny, x0, y0, z0, bx, by, bz, cx, cy, cz, ax, ay, az,
                                                 uint16_t(p[6]));
    }
    if (sigma > 0.0f) {
        if (reinterpret_cast<uintptr_t>(d_volume) & 3u) return cudaErrorMisalignedAddress;
        size_t blocks = ((n + 1) / 2 + 255) / 256;
        if (blocks > 148 * 64) blocks = 148 * 64;
        if (blocks == 0) blocks = 1;
        k_noise<<<unsigned(blocks), 256, 0, s>>>(d_volume, n, sigma, uint32_t(seed & 0xFFFFFFFFull), uint32_t(seed >> 32),
                                                 scan_index);
    }
    return cudaGetLastError();
}
