// Dependent-access latencies on the GPU at hand, as the labelling kernels see them: pointer chase through an
// L2-resident table with plain / volatile / atomic accesses, one thread and many threads; __syncthreads, __threadfence,
// cluster barrier round trips.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/latency tools/micro/latency.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__global__ void chase(const uint32_t* __restrict__ next, int hops, int mode, uint32_t* out, long long* cyc, unsigned long long* ns) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 97u % 65536u;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long c0 = clock64();
    if (mode == 0) for (int i = 0; i < hops; ++i) x = next[x];
    else if (mode == 1) for (int i = 0; i < hops; ++i) x = *reinterpret_cast<const volatile uint32_t*>(next + x);
    else if (mode == 2) for (int i = 0; i < hops; ++i) x = __ldcg(next + x);
    else for (int i = 0; i < hops; ++i) x = atomicMin(const_cast<uint32_t*>(next) + x, 0xFFFFFFFFu);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { *cyc = c1 - c0; *ns = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void syncs(int n, int mode, long long* cyc) {
    __shared__ int s;
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (mode == 0) __syncthreads();
        else if (mode == 1) { __threadfence(); __syncthreads(); }
        else if (mode == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
        else { __threadfence(); asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { *cyc = c1 - c0; s = 0; }
}

__global__ void empty_k(int* p) { if (p && threadIdx.x == 12345) *p = 1; }

int main() {
    const int N = 65536;
    std::vector<uint32_t> h(N);
    for (int i = 0; i < N; ++i) h[i] = (uint32_t)((i * 40503u + 12345u) % N);
    uint32_t *d, *out; long long* cyc; unsigned long long* ns;
    cudaMalloc(&d, N * 4); cudaMalloc(&out, 1 << 22); cudaMallocManaged(&cyc, 8); cudaMallocManaged(&ns, 8);
    cudaMemcpy(d, h.data(), N * 4, cudaMemcpyHostToDevice);
    const char* names[4] = {"plain ld (L1 may hit)", "ld.volatile", "ld.cg", "atomicMin (return)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int cfg = 0; cfg < 3; ++cfg) {
            int blocks = cfg == 0 ? 1 : (cfg == 1 ? 148 : 148 * 4), threads = cfg == 0 ? 32 : 256;
            chase<<<blocks, threads>>>(d, 200, mode, out, cyc, ns);
            cudaDeviceSynchronize();
            chase<<<blocks, threads>>>(d, 200, mode, out, cyc, ns);
            cudaDeviceSynchronize();
            printf("%-24s %4d x %3d threads: %7.1f cycles/hop  %7.1f ns/hop\n", names[mode], blocks, threads, *cyc / 200.0, *ns / 200.0);
            if (mode == 3) cudaMemcpy(d, h.data(), N * 4, cudaMemcpyHostToDevice);
        }
    const char* sn[4] = {"__syncthreads", "__threadfence + __syncthreads", "cluster barrier", "__threadfence + cluster barrier"};
    for (int mode = 0; mode < 4; ++mode)
        for (int threads = 256; threads <= 1024; threads *= 4) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(mode >= 2 ? 16 : 1); cfg.blockDim = dim3(threads);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = mode >= 2 ? 16 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaFuncSetAttribute(syncs, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchKernelEx(&cfg, syncs, 100, mode, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            printf("%-32s %4d threads%s: %7.1f cycles each (%s)\n", sn[mode], threads, mode >= 2 ? " x 16 CTAs" : "", *cyc / 100.0, cudaGetErrorString(e));
        }
    // launch gaps: 20 dependent empty kernels in a stream / in a graph
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0, s);
        for (int i = 0; i < 20; ++i) empty_k<<<148, 256, 0, s>>>(nullptr);
        cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("20 empty kernels in a stream: %.2f us each\n", ms * 1e3 / 20);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < 20; ++i) empty_k<<<148, 256, 0, s>>>(nullptr);
    cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("20 empty kernels in a graph: %.2f us each (incl. graph launch)\n", ms * 1e3 / 20);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock attr %d kHz\n", clk);
    return 0;
}
