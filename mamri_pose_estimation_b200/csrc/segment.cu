// Stage 1: intensity threshold -> bit-packed mask, and binary closing with ITK's ball on the
// bit-packed mask.  Replaces sitk.BinaryThreshold + sitk.BinaryMorphologicalClosing at
// Mamri/Mamri.py:1308.  HBM-bound: the voxel volume is read exactly once with 128-bit loads; all
// morphology runs on 1 bit/voxel data (L2-resident) with shared-memory halo tiles.
#include "common.cuh"

#include <limits>
#include <math.h>

// ------------------------------------------------------------------------------------------------
// threshold + pack
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool in_range(T v, T lo, T hi) { return v >= lo && v <= hi; }  // NaN -> false

__device__ __forceinline__ uint4 ld_stream_128(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Fast path: nx % 32 == 0 and 16-byte aligned base, so the packed mask is one flat bit array.
// A warp load instruction covers 512 contiguous bytes; each lane turns its 16 bytes into E bits and
// 32/E neighbouring lanes are merged into one 32-voxel word by shuffles.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) k_threshold_pack_flat(const uint4* __restrict__ vol, size_t n_vec, T lo, T hi,
                                                             uint32_t* __restrict__ bits) {
    constexpr int E = 16 / sizeof(T);   // voxels per 128-bit load
    constexpr int G = 32 / E;           // lanes per output word
    const unsigned lane = lane_id();
    const size_t warp = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
    for (size_t base = warp * (UNROLL * 32); base < n_vec; base += n_warps * (UNROLL * 32)) {
        uint4 v[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            size_t i = base + u * 32 + lane;
            ok[u] = i < n_vec;
            v[u] = ok[u] ? ld_stream_128(vol + i) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            union { uint4 q; T e[E]; } cvt;
            cvt.q = v[u];
            uint32_t b = 0;
#pragma unroll
            for (int k = 0; k < E; ++k) b |= (in_range(cvt.e[k], lo, hi) ? 1u : 0u) << k;
            if (!ok[u]) b = 0;
#pragma unroll
            for (int s = 1; s < G; s <<= 1) b |= __shfl_down_sync(0xFFFFFFFFu, b, s) << (E * s);
            size_t i = base + u * 32 + lane;
            if (ok[u] && (lane % G) == 0) bits[i / G] = b;
        }
    }
}

// General path (ragged nx or unaligned base): one warp per output word, one voxel per lane.
template <typename T>
__global__ void __launch_bounds__(256) k_threshold_pack_rows(const T* __restrict__ vol, int nx, int W, size_t n_words,
                                                             T lo, T hi, uint32_t* __restrict__ bits) {
    const unsigned lane = lane_id();
    const size_t warp = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
    for (size_t wi = warp; wi < n_words; wi += n_warps) {
        size_t row = wi / W;
        int x = int(wi - row * W) * 32 + int(lane);
        bool p = x < nx && in_range(vol[row * size_t(nx) + x], lo, hi);
        uint32_t b = __ballot_sync(0xFFFFFFFFu, p);
        if (lane == 0) bits[wi] = b;
    }
}

// static_cast<InputPixelType>(double) as itk::BinaryThresholdImageFilter applies to its bounds;
// out-of-range bounds are clamped (documented deviation for int16).
template <typename T>
static T cast_bound(double v) {
    if (v != v) return T(0);
    double t = v < 0 ? -floor(-v) : floor(v);
    double lo = double(std::numeric_limits<T>::lowest()), hi = double(std::numeric_limits<T>::max());
    if (t < lo) t = lo;
    if (t > hi) t = hi;
    return T(t);
}
template <>
float cast_bound<float>(double v) { return float(v); }

template <typename T>
static cudaError_t threshold_pack_t(const void* d_vol, int nx, int ny, int nz, double lo, double hi, uint32_t* d_bits,
                                    cudaStream_t s) {
    const size_t rows = size_t(ny) * nz;
    const int W = (nx + 31) / 32;
    const size_t n_words = rows * W;
    const T tlo = cast_bound<T>(lo), thi = cast_bound<T>(hi);
    const bool flat = (nx % 32 == 0) && ((reinterpret_cast<uintptr_t>(d_vol) & 15u) == 0);
    if (flat) {
        constexpr int E = 16 / sizeof(T);
        constexpr int UNROLL = 4;
        const size_t n_vec = rows * size_t(nx) / E;
        size_t warps = (n_vec + UNROLL * 32 - 1) / (UNROLL * 32);
        size_t blocks = (warps + 7) / 8;
        const size_t cap = 148 * 8 * 4;           // a few waves of 8 resident CTAs per SM
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        k_threshold_pack_flat<T, UNROLL><<<unsigned(blocks), 256, 0, s>>>(static_cast<const uint4*>(d_vol), n_vec, tlo,
                                                                         thi, d_bits);
    } else {
        size_t blocks = (n_words + 7) / 8;
        const size_t cap = 148 * 8 * 8;
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        k_threshold_pack_rows<T><<<unsigned(blocks), 256, 0, s>>>(static_cast<const T*>(d_vol), nx, W, n_words, tlo, thi,
                                                                  d_bits);
    }
    return cudaGetLastError();
}

cudaError_t launch_threshold_pack(const void* d_vol, int dtype, int nx, int ny, int nz, double lo, double hi,
                                  uint32_t* d_bits, cudaStream_t s) {
    switch (dtype) {
        case MAMRI_U8:  return threshold_pack_t<uint8_t>(d_vol, nx, ny, nz, lo, hi, d_bits, s);
        case MAMRI_I16: return threshold_pack_t<int16_t>(d_vol, nx, ny, nz, lo, hi, d_bits, s);
        case MAMRI_U16: return threshold_pack_t<uint16_t>(d_vol, nx, ny, nz, lo, hi, d_bits, s);
        case MAMRI_I32: return threshold_pack_t<int32_t>(d_vol, nx, ny, nz, lo, hi, d_bits, s);
        case MAMRI_F32: return threshold_pack_t<float>(d_vol, nx, ny, nz, lo, hi, d_bits, s);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// closing = dilation on the r-grown domain, then erosion back on the image domain
// ------------------------------------------------------------------------------------------------
// ITK's ball of radius R (FlatStructuringElement::Ball, radiusIsParametric = false) is
// {d : dx^2+dy^2+dz^2 <= R^2+R}.  On bit-packed rows it is, for every (dy,dz) with
// dy^2+dz^2 <= R^2+R, an x-interval of half-width h = floor(sqrt(R^2+R-dy^2-dz^2)).  So with
// S_h(row) = the row dilated (eroded) along x by h, the result word is the OR (AND) of S_h over
// those rows.  A CTA stages a (TZ+2R) x (TY+2R) x (TXW+2) halo tile of source words in shared
// memory, derives S_1..S_R once per source word, then every output word is <= (2R+1)^2 LDS + ORs.
constexpr int TY = 8, TZ = 8, TXW_MAX = 32;

__host__ __device__ constexpr int isqrt_c(int v) { int r = 0; while ((r + 1) * (r + 1) <= v) ++r; return r; }

template <int R, bool ERODE>
__global__ void __launch_bounds__(256) k_morph_tile(BitVol src, BitVol dst, int ox, int oy, int oz, uint32_t tail_mask) {
    // src coordinate = dst coordinate + (ox, oy, oz)  (words, rows, slices)
    extern __shared__ uint32_t sm[];
    const int txw = min(dst.w - int(blockIdx.x) * TXW_MAX, TXW_MAX);   // output words along x in this tile
    constexpr int HY = TY + 2 * R, HZ = TZ + 2 * R;
    const int rw = txw + 2;                                           // raw tile row length (1-word halo each side)
    uint32_t* raw = sm;                                               // [HZ][HY][rw]
    uint32_t* sh = sm + HZ * HY * rw;                                 // [R][HZ][HY][txw]
    const int x0 = int(blockIdx.x) * TXW_MAX, y0 = int(blockIdx.y) * TY, z0 = int(blockIdx.z) * TZ;

    for (int i = threadIdx.x; i < HZ * HY * rw; i += blockDim.x) {
        int tx = i % rw, t = i / rw, ty = t % HY, tz = t / HY;
        int sx = x0 + tx - 1 + ox, sy = y0 + ty - R + oy, sz = z0 + tz - R + oz;
        uint32_t v = 0;
        if (sx >= 0 && sx < src.w && sy >= 0 && sy < src.h && sz >= 0 && sz < src.d)
            v = src.p[(size_t(sz) * src.h + sy) * src.w + sx];
        raw[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HZ * HY * txw; i += blockDim.x) {
        int tx = i % txw, row = i / txw;
        const uint32_t l = raw[row * rw + tx], c = raw[row * rw + tx + 1], r = raw[row * rw + tx + 2];
        uint32_t acc = c;
#pragma unroll
        for (int k = 1; k <= R; ++k) {
            uint32_t a = (c << k) | (l >> (32 - k)), b = (c >> k) | (r << (32 - k));
            acc = ERODE ? (acc & a & b) : (acc | a | b);
            sh[(size_t(k - 1) * HZ * HY + row) * txw + tx] = acc;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TZ * TY * txw; i += blockDim.x) {
        int tx = i % txw, t = i / txw, ty = t % TY, tz = t / TY;
        int x = x0 + tx, y = y0 + ty, z = z0 + tz;
        if (y >= dst.h || z >= dst.d) continue;
        uint32_t acc = ERODE ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int dz = -R; dz <= R; ++dz) {
#pragma unroll
            for (int dy = -R; dy <= R; ++dy) {
                constexpr int R2 = R * R + R;
                const int rem = R2 - dy * dy - dz * dz;
                if (rem < 0) continue;
                const int h = isqrt_c(rem);
                const int row = (tz + R + dz) * HY + (ty + R + dy);
                uint32_t v = (h == 0) ? raw[row * rw + tx + 1] : sh[(size_t(h - 1) * HZ * HY + row) * txw + tx];
                acc = ERODE ? (acc & v) : (acc | v);
            }
        }
        if (x == dst.w - 1) acc &= tail_mask;
        dst.p[(size_t(z) * dst.h + y) * dst.w + x] = acc;
    }
}

template <int R>
static cudaError_t closing_r(mamri_ctx* c, int nx, int ny, int nz, cudaStream_t s) {
    const int W = (nx + 31) / 32;
    BitVol raw{c->d_raw, W, ny, nz};
    BitVol dil{c->d_dil, W + 2, ny + 2 * R, nz + 2 * R};
    BitVol out{c->d_closed, W, ny, nz};
    constexpr int HY = TY + 2 * R, HZ = TZ + 2 * R;
    auto smem_for = [&](int w) {
        int txw = w < TXW_MAX ? w : TXW_MAX;
        return size_t(HZ * HY) * (txw + 2 + R * txw) * sizeof(uint32_t);
    };
    static bool attr_set = false;
    if (!attr_set) {
        size_t mx = smem_for(TXW_MAX);
        cudaError_t e = cudaFuncSetAttribute(k_morph_tile<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(mx));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_morph_tile<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(mx));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // dilation: padded output coordinate (xw, y, z) reads raw (xw-1, y-R, z-R)
    dim3 gd((dil.w + TXW_MAX - 1) / TXW_MAX, (dil.h + TY - 1) / TY, (dil.d + TZ - 1) / TZ);
    k_morph_tile<R, false><<<gd, 256, smem_for(dil.w), s>>>(raw, dil, -1, -R, -R, 0xFFFFFFFFu);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // erosion: image output coordinate (xw, y, z) reads the padded dilation at (xw+1, y+R, z+R)
    const uint32_t tail = (nx % 32) ? (0xFFFFFFFFu >> (32 - nx % 32)) : 0xFFFFFFFFu;
    dim3 ge((out.w + TXW_MAX - 1) / TXW_MAX, (out.h + TY - 1) / TY, (out.d + TZ - 1) / TZ);
    k_morph_tile<R, true><<<ge, 256, smem_for(out.w), s>>>(dil, out, 1, R, R, tail);
    return cudaGetLastError();
}

cudaError_t launch_closing(mamri_ctx* c, int nx, int ny, int nz, int radius, cudaStream_t s) {
    switch (radius) {
        case 1: return closing_r<1>(c, nx, ny, nz, s);
        case 2: return closing_r<2>(c, nx, ny, nz, s);
        case 3: return closing_r<3>(c, nx, ny, nz, s);
        default: return cudaErrorInvalidValue;
    }
}
