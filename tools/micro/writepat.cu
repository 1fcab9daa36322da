// Micro-benchmark: the materialise kernel's store pattern (u32 labels as 16 B/lane + u8 mask) in variants.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

// A: per 4 words: lane -> uint4 label store + u32 mask store (current kernel)
__global__ void __launch_bounds__(256) k_a(uint32_t* __restrict__ lab, uint8_t* __restrict__ msk, uint32_t n_words) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w0 = warp * 32; w0 < n_words; w0 += nw * 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t wi = w0 + k * 4 + (lane >> 3);
            const uint32_t v = wi * 32 + (lane & 7) * 4;
            *reinterpret_cast<uint4*>(lab + v) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint32_t*>(msk + v) = 0;
        }
    }
}
// B: labels as in A; mask as 2 x (16 B per lane)
__global__ void __launch_bounds__(256) k_b(uint32_t* __restrict__ lab, uint8_t* __restrict__ msk, uint32_t n_words) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w0 = warp * 32; w0 < n_words; w0 += nw * 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t wi = w0 + k * 4 + (lane >> 3);
            const uint32_t v = wi * 32 + (lane & 7) * 4;
            *reinterpret_cast<uint4*>(lab + v) = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
            *reinterpret_cast<uint4*>(msk + size_t(w0) * 32 + k * 512 + lane * 16) = make_uint4(0, 0, 0, 0);
    }
}
// C: labels only
__global__ void __launch_bounds__(256) k_c(uint32_t* __restrict__ lab, uint32_t n_words) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w0 = warp * 32; w0 < n_words; w0 += nw * 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t wi = w0 + k * 4 + (lane >> 3);
            *reinterpret_cast<uint4*>(lab + wi * 32 + (lane & 7) * 4) = make_uint4(0, 0, 0, 0);
        }
    }
}
// D: like B but streaming stores (st.global.cs)
__device__ __forceinline__ void st_cs(uint4* p, uint4 v) { asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)); }
__global__ void __launch_bounds__(256) k_d(uint32_t* __restrict__ lab, uint8_t* __restrict__ msk, uint32_t n_words) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w0 = warp * 32; w0 < n_words; w0 += nw * 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t wi = w0 + k * 4 + (lane >> 3);
            st_cs(reinterpret_cast<uint4*>(lab + wi * 32 + (lane & 7) * 4), make_uint4(0, 0, 0, 0));
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) st_cs(reinterpret_cast<uint4*>(msk + size_t(w0) * 32 + k * 512 + lane * 16), make_uint4(0, 0, 0, 0));
    }
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
    const uint32_t n_words = 16u * 512 * 256;   // C2
    const size_t nvox = size_t(n_words) * 32;
    uint32_t* lab; uint8_t* msk;
    cudaMalloc(&lab, nvox * 4 * 2); cudaMalloc(&msk, nvox * 2);
    int flip = 0;
    for (int per_sm : {4, 8, 16, 32}) {
        const int grid = 148 * per_sm;
        auto L = [&] { flip ^= 1; return lab + (flip ? nvox : 0); };
        auto M = [&] { return msk + (flip ? nvox : 0); };
        float a = time_ms([&] { uint32_t* l = L(); k_a<<<grid, 256>>>(l, M(), n_words); });
        float b = time_ms([&] { uint32_t* l = L(); k_b<<<grid, 256>>>(l, M(), n_words); });
        float c = time_ms([&] { k_c<<<grid, 256>>>(L(), n_words); });
        float d = time_ms([&] { uint32_t* l = L(); k_d<<<grid, 256>>>(l, M(), n_words); });
        printf("ctas/sm %2d: A(cur) %.1f us | B(mask 16B) %.1f us | C(labels only) %.1f us | D(B + st.cs) %.1f us\n", per_sm, a * 1e3, b * 1e3, c * 1e3, d * 1e3);
    }
    return 0;
}
