// Micro-benchmark: achievable HBM read / write bandwidth on this box for the access patterns the
// threshold (read 128-bit, tiny output) and materialise (write-only) kernels use.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int U, bool NC>
__global__ void __launch_bounds__(256) k_read(const uint4* __restrict__ p, size_t n, uint32_t* out) {
    const size_t warp = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (size_t(gridDim.x) * blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    uint32_t acc = 0;
    for (size_t b = warp * U * 32; b < n; b += nw * U * 32) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { size_t i = b + u * 32 + lane; v[u] = i < n ? (NC ? ld_nc(p + i) : p[i]) : make_uint4(0, 0, 0, 0); }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_write(uint4* __restrict__ p, size_t n) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        p[i] = make_uint4(1, 2, 3, 4);
}

template <typename F>
float time_ms(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); f();
    float best = 1e9;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    const size_t big = size_t(1) << 30;            // 1 GiB arena, 8 distinct 128 MiB windows to defeat L2
    uint4* buf; cudaMalloc(&buf, big + (size_t(512) << 20)); cudaMemset(buf, 1, big);
    uint32_t* out; cudaMalloc(&out, 4);
    uint4* wbuf = buf + big / 16;
    const size_t n128 = (size_t(128) << 20) / 16;
    int win = 0;
    auto next = [&]() { win = (win + 1) & 7; return buf + size_t(win) * n128; };
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        float a = time_ms([&] { k_read<4, true><<<grid, 256>>>(next(), n128, out); });
        float b = time_ms([&] { k_read<8, true><<<grid, 256>>>(next(), n128, out); });
        float c = time_ms([&] { k_read<4, false><<<grid, 256>>>(next(), n128, out); });
        float d = time_ms([&] { k_read<2, true><<<grid, 256>>>(next(), n128, out); });
        printf("read 128MiB grid %5d: U4nc %.1f us %.0f GB/s | U8nc %.1f us %.0f GB/s | U4 %.1f us %.0f GB/s | U2nc %.1f us %.0f GB/s\n", grid,
               a * 1e3, 134.2 / a, b * 1e3, 134.2 / b, c * 1e3, 134.2 / c, d * 1e3, 134.2 / d);
    }
    float r1 = time_ms([&] { k_read<4, true><<<148 * 16, 256>>>(buf, big / 16, out); });
    printf("read 1GiB: %.1f us %.0f GB/s\n", r1 * 1e3, 1073.7 / r1);
    for (int grid : {148 * 8, 148 * 16, 148 * 64}) {
        float w = time_ms([&] { k_write<<<grid, 256>>>(wbuf, (size_t(320) << 20) / 16); });
        printf("write 320MiB grid %5d: %.1f us %.0f GB/s\n", grid, w * 1e3, 335.5 / w);
    }
    float ms = time_ms([&] { cudaMemsetAsync(wbuf, 0, size_t(320) << 20); });
    printf("cudaMemset 320MiB: %.1f us %.0f GB/s\n", ms * 1e3, 335.5 / ms);
    float cp = time_ms([&] { cudaMemcpyAsync(wbuf, buf, size_t(256) << 20, cudaMemcpyDeviceToDevice); });
    printf("D2D copy 256MiB: %.1f us %.0f GB/s (read+write)\n", cp * 1e3, 2 * 268.4 / cp);
    return 0;
}
