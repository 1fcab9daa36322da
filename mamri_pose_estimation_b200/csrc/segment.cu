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
// dy^2+dz^2 <= R^2+R, an x-interval of half-width h = floor(sqrt(R^2+R-dy^2-dz^2)).  With
// S_h(row) = the row dilated (eroded) along x by h (funnel shifts over three neighbouring words),
//   P_a(y, z') = OP_{dy} S_{h(dy,a)}(y+dy, z')          (the in-slice part for |dz| = a)
//   out(y, z)  = OP_{dz} P_|dz|(y, z+dz)                 (OP = OR for dilation, AND for erosion).
// A thread owns one output word column (xw, y) over a chunk of ZC slices and slides a register
// window of P values along z: every step loads (2R+1) rows x 3 words of ONE new slice (coalesced,
// L1-shared with the neighbouring threads), so a source word is loaded ~(2R+1)*3 times per output
// instead of (2R+1)^2*3 -- no shared memory, no barriers.  The bit volumes are L2-resident (1 bit/voxel).
__host__ __device__ constexpr int isqrt_c(int v) { int r = 0; while ((r + 1) * (r + 1) <= v) ++r; return r; }

template <int R, bool ERODE>
__device__ __forceinline__ void slice_patterns(const BitVol& src, int sx, int sy, int sz, uint32_t (&P)[R + 1]) {
    constexpr int R2 = R * R + R;
#pragma unroll
    for (int a = 0; a <= R; ++a) P[a] = ERODE ? 0xFFFFFFFFu : 0u;
    const bool z_ok = sz >= 0 && sz < src.d;
#pragma unroll
    for (int dy = -R; dy <= R; ++dy) {
        const int y = sy + dy;
        uint32_t l = 0, c = 0, r = 0;
        if (z_ok && y >= 0 && y < src.h) {
            const uint32_t* row = src.p + (size_t(sz) * src.h + y) * src.w;
            if (sx >= 0 && sx < src.w) c = row[sx];
            if (sx - 1 >= 0 && sx - 1 < src.w) l = row[sx - 1];
            if (sx + 1 >= 0 && sx + 1 < src.w) r = row[sx + 1];
        }
        uint32_t S[R + 1];
        S[0] = c;
#pragma unroll
        for (int k = 1; k <= R; ++k) {
            const uint32_t a = (c << k) | (l >> (32 - k)), b = (c >> k) | (r << (32 - k));
            S[k] = ERODE ? (S[k - 1] & a & b) : (S[k - 1] | a | b);
        }
#pragma unroll
        for (int a = 0; a <= R; ++a) {
            const int rem = R2 - dy * dy - a * a;
            if (rem >= 0) P[a] = ERODE ? (P[a] & S[isqrt_c(rem)]) : (P[a] | S[isqrt_c(rem)]);
        }
    }
}

template <int R, bool ERODE>
__global__ void __launch_bounds__(256) k_morph_sweep(BitVol src, BitVol dst, int ox, int oy, int oz, uint32_t tail_mask,
                                                     int zc, int n_chunks) {
    // src coordinate = dst coordinate + (ox, oy, oz)  (words, rows, slices)
    const size_t per_chunk = size_t(dst.w) * dst.h;
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= per_chunk * n_chunks) return;
    const int chunk = int(t / per_chunk);
    const size_t in_slice = t - size_t(chunk) * per_chunk;
    const int y = int(in_slice / dst.w), x = int(in_slice - size_t(y) * dst.w);
    const int z0 = chunk * zc, z1 = min(z0 + zc, dst.d);
    const int sx = x + ox, sy = y + oy;
    uint32_t win[2 * R + 1][R + 1];                 // win[j] = patterns of source slice (z + oz) - R + j
#pragma unroll
    for (int j = 0; j < 2 * R; ++j) slice_patterns<R, ERODE>(src, sx, sy, z0 + oz - R + j, win[j]);
    const uint32_t tm = (x == dst.w - 1) ? tail_mask : 0xFFFFFFFFu;
    for (int z = z0; z < z1; ++z) {
        slice_patterns<R, ERODE>(src, sx, sy, z + oz + R, win[2 * R]);
        uint32_t acc = ERODE ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int dz = -R; dz <= R; ++dz) {
            const uint32_t v = win[dz + R][dz < 0 ? -dz : dz];
            acc = ERODE ? (acc & v) : (acc | v);
        }
        dst.p[(size_t(z) * dst.h + y) * dst.w + x] = acc & tm;
#pragma unroll
        for (int j = 0; j < 2 * R; ++j)
#pragma unroll
            for (int a = 0; a <= R; ++a) win[j][a] = win[j + 1][a];
    }
}

template <int R, bool ERODE>
static cudaError_t morph_launch(BitVol src, BitVol dst, int ox, int oy, int oz, uint32_t tail, cudaStream_t s) {
    // chunk the z sweep so that the grid has a few hundred thousand threads (148 SMs x 2048)
    const size_t per_chunk = size_t(dst.w) * dst.h;
    int zc = 16;
    while (zc > 4 && per_chunk * ((dst.d + zc - 1) / zc) < size_t(148) * 2048) zc >>= 1;
    const int n_chunks = (dst.d + zc - 1) / zc;
    const size_t threads = per_chunk * n_chunks;
    k_morph_sweep<R, ERODE><<<unsigned((threads + 255) / 256), 256, 0, s>>>(src, dst, ox, oy, oz, tail, zc, n_chunks);
    return cudaGetLastError();
}

template <int R>
static cudaError_t closing_r(mamri_ctx* c, int nx, int ny, int nz, cudaStream_t s) {
    const int W = (nx + 31) / 32;
    BitVol raw{c->d_raw, W, ny, nz};
    BitVol dil{c->d_dil, W + 2, ny + 2 * R, nz + 2 * R};
    BitVol out{c->d_closed, W, ny, nz};
    // dilation: padded output coordinate (xw, y, z) reads raw (xw-1, y-R, z-R); reads outside raw are 0
    cudaError_t e = morph_launch<R, false>(raw, dil, -1, -R, -R, 0xFFFFFFFFu, s);
    if (e != cudaSuccess) return e;
    // erosion: image output coordinate (xw, y, z) reads the padded dilation at (xw+1, y+R, z+R); every
    // read of an image-domain output stays inside the padded volume
    const uint32_t tail = (nx % 32) ? (0xFFFFFFFFu >> (32 - nx % 32)) : 0xFFFFFFFFu;
    return morph_launch<R, true>(dil, out, 1, R, R, tail, s);
}

cudaError_t launch_closing(mamri_ctx* c, int nx, int ny, int nz, int radius, cudaStream_t s) {
    switch (radius) {
        case 1: return closing_r<1>(c, nx, ny, nz, s);
        case 2: return closing_r<2>(c, nx, ny, nz, s);
        case 3: return closing_r<3>(c, nx, ny, nz, s);
        default: return cudaErrorInvalidValue;
    }
}
