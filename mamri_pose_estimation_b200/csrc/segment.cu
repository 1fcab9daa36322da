// Stage 1: intensity threshold -> bit-packed mask, and binary closing with ITK's ball on the
// bit-packed mask.  Replaces sitk.BinaryThreshold + sitk.BinaryMorphologicalClosing at
// Mamri/Mamri.py:1308.  HBM-bound: the voxel volume is read exactly once with 128-bit loads; all
// morphology runs on 1 bit/voxel data that stays L2-resident.
#include "common.cuh"

#include <limits>
#include <math.h>
#include <stdlib.h>
#include <type_traits>

// ------------------------------------------------------------------------------------------------
// threshold + pack
// ------------------------------------------------------------------------------------------------
// Where the packed words go: either the plain [nz][ny][W] mask, or the interior of the zero-apron
// padded volume the closing reads (row/slice strides and a start offset).
struct BitDst {
    uint32_t* p;
    uint32_t W, ny;
    uint32_t row_stride, slice_stride, off;
    bool linear;
    __device__ __forceinline__ uint32_t index(uint32_t i) const {
        if (linear) return i;
        const uint32_t row = i / W, xw = i - row * W;
        const uint32_t z = row / ny, y = row - z * ny;
        return off + z * slice_stride + y * row_stride + xw;
    }
};

template <typename T>
__device__ __forceinline__ bool in_range(T v, T lo, T hi) { return v >= lo && v <= hi; }  // NaN -> false

__device__ __forceinline__ uint4 ld_stream_128(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// Same, with an L2 evict-first policy: the voxels are read exactly once, so their lines should leave L2
// before the bit-packed intermediates of the scans in flight do.
__device__ __forceinline__ uint4 ld_stream_128(const uint4* p, uint64_t policy) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy));
    return r;
}

// Fast path: nx % 32 == 0 and 16-byte aligned base, so every row is a whole number of 128-bit vectors
// and of mask words.  blockIdx.y = slice, warps stride over the rows of the slice two at a time (no
// index divisions anywhere).  A warp load instruction covers 512 contiguous bytes; each lane turns its
// 16 bytes into E bits and 32/E neighbouring lanes are merged into one 32-voxel word by shuffles.
// Generic: one compare pair per voxel.
template <typename T>
struct RangeTest {
    T lo, hi;
    __device__ __forceinline__ uint32_t pack(uint4 q) const {
        constexpr int E = 16 / sizeof(T);
        union { uint4 q; T e[E]; } cvt;
        cvt.q = q;
        uint32_t b = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) b |= (in_range(cvt.e[k], lo, hi) ? 1u : 0u) << k;
        return b;
    }
};

// 16-bit voxels (the MR case): two voxels per 32-bit register, compared without unpacking.  For
// 16-bit lanes a, l:  d = (a | 0x8000) - (l & 0x7FFF) never borrows across lanes and its top bit is
// (a & 0x7FFF) >= (l & 0x7FFF); the lanes' own top bits settle the rest:  a >= l  <=>
// msb(l) ? msb(a & d) : msb(a | d).  a <= h is h >= a.  Signed voxels are biased by 0x8000 first
// (flips the order of the top bit only).  ~2.6 instructions per voxel when the upper bound is the
// type's maximum (the reference's 65535), ~4.6 otherwise, against ~10 for compare-and-select.
template <bool LO_MSB, bool HI_MSB, bool CHECK_HI, bool SIGNED>
struct RangeTest16 {
    uint32_t lo_c;      // (lo & 0x7FFF) in both lanes
    uint32_t hi_c;      // (hi | 0x8000) in both lanes
    __device__ __forceinline__ uint32_t lanes(uint32_t w) const {
        const uint32_t H = 0x80008000u;
        if (SIGNED) w ^= H;
        const uint32_t d = (w | H) - lo_c;
        uint32_t r = LO_MSB ? (w & d) : (w | d);
        if (CHECK_HI) {
            const uint32_t d2 = hi_c - (w & ~H);
            r &= HI_MSB ? (~w | d2) : (~w & d2);
        }
        return r & H;                       // bit 15 / bit 31 = low / high voxel in range
    }
    __device__ __forceinline__ uint32_t pack(uint4 q) const {
        const uint32_t s = (lanes(q.x) >> 15) | (lanes(q.y) >> 13) | (lanes(q.z) >> 11) | (lanes(q.w) >> 9);
        return (s | (s >> 15)) & 0xFFu;     // low voxels sit on even bits, high voxels 16 above: interleave
    }
};

// Occupancy cells of the raw mask, in image coordinates (rows x slices; a cell spans all of x).  A cell holds the
// tag of the last scan that saw foreground in it (occ_tag: 1..255 from the launch generation), so the cells never
// need clearing: a stale tag that happens to equal the current one only makes the closing read a tile of air.
constexpr uint32_t OCC_CY = 8, OCC_CZ = 4;
__device__ __forceinline__ uint8_t occ_tag(uint32_t gen) { return uint8_t(gen % 255u + 1u); }

// WHOLE: ny is even and a row is a whole number of warp trips (vec_per_row % 64 == 0), so nothing in the loop needs a
// bounds predicate -- the common case (512- and 1024-voxel rows), and the kernel is close to issue-bound.
template <int VOXEL_BYTES, typename Test, bool WHOLE>
__global__ void __launch_bounds__(256) k_threshold_pack_vec(const DynArgs* __restrict__ dyn, uint32_t vec_per_row,
                                                            uint32_t ny, Test test, uint32_t* __restrict__ dst,
                                                            uint32_t row_stride, uint32_t slice_stride, uint32_t off,
                                                            int evict_first, uint8_t* __restrict__ occ, uint32_t occ_ncy) {
    pdl_wait();
    ktrace(KT_THRESHOLD);
    uint64_t policy = 0;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const uint4* __restrict__ vol = static_cast<const uint4*>(dyn->vol);
    const uint8_t tag = occ_tag(dyn->gen);
    constexpr int E = 16 / VOXEL_BYTES; // voxels per 128-bit load
    constexpr int G = 32 / E;           // lanes per output word
    constexpr int U = 2;                // vectors per lane per row in flight (x 2 rows)
    const unsigned lane = lane_id();
    const uint32_t z = blockIdx.y;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint4* slice_src = vol + size_t(z) * ny * vec_per_row;
    uint32_t* slice_dst = dst + off + z * slice_stride;
    for (uint32_t y0 = warp * 2; y0 < ny; y0 += n_warps * 2) {
        uint32_t seen = 0;                                        // any foreground in this pair of rows
        for (uint32_t u0 = 0; u0 < vec_per_row; u0 += 32 * U) {
            uint4 v[2][U];
            bool ok[2][U];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t i = u0 + u * 32 + lane;
                    ok[r][u] = WHOLE || ((y0 + r < ny) && i < vec_per_row);
                    const uint4* src = slice_src + size_t(y0 + r) * vec_per_row + i;
                    v[r][u] = !ok[r][u] ? make_uint4(0, 0, 0, 0) : evict_first ? ld_stream_128(src, policy) : ld_stream_128(src);
                }
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t b = ok[r][u] ? test.pack(v[r][u]) : 0u;
#pragma unroll
                    for (int s = 1; s < G; s <<= 1) b |= __shfl_down_sync(0xFFFFFFFFu, b, s) << (E * s);
                    if (ok[r][u] && (lane % G) == 0) slice_dst[(y0 + r) * row_stride + (u0 + u * 32 + lane) / G] = b;
                    seen |= b;
                }
        }
        // occupancy cell (OCC_CY rows x OCC_CZ slices) of the raw mask: lets the closing skip tiles of air unread
        if (occ && __any_sync(0xFFFFFFFFu, seen != 0u) && lane == 0) occ[(z / OCC_CZ) * occ_ncy + y0 / OCC_CY] = tag;
    }
}


// 16-bit voxels, rows that are a whole number of 512-voxel warp trips, 32-byte aligned base: 256-bit loads
// (ld.global.nc.v8.u32, sm_100).  A lane turns its 32 bytes = 16 voxels into 16 mask bits, two lanes make a word (one
// shuffle), ROWS rows are in flight per warp.  About half the instructions per voxel of the 128-bit kernel above
// (which sits between the issue and the memory limit): the compare is the same SWAR test, but the lane merge, the
// address arithmetic and the loop overhead are paid once per 16 voxels instead of once per 8.
__device__ __forceinline__ void ld_stream_256(const void* p, uint64_t policy, bool hint, uint4& a, uint4& b) {
    if (hint)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p), "l"(policy));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

template <typename Test, int ROWS>
__global__ void __launch_bounds__(256) k_threshold_pack_v8(const DynArgs* __restrict__ dyn, uint32_t vec_per_row /* 32-byte vectors */,
                                                           uint32_t ny, Test test, uint32_t* __restrict__ dst, uint32_t row_stride,
                                                           uint32_t slice_stride, uint32_t off, int evict_first,
                                                           uint8_t* __restrict__ occ, uint32_t occ_ncy) {
    pdl_wait();
    ktrace(KT_THRESHOLD);
    uint64_t policy = 0;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const char* __restrict__ vol = static_cast<const char*>(dyn->vol);
    const uint8_t tag = occ_tag(dyn->gen);
    const unsigned lane = lane_id();
    const uint32_t z = blockIdx.y;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const char* slice_src = vol + size_t(z) * ny * vec_per_row * 32u;
    uint32_t* slice_dst = dst + off + z * slice_stride;
    for (uint32_t y0 = warp * ROWS; y0 < ny; y0 += n_warps * ROWS) {           // ny % ROWS == 0
        uint32_t seen = 0;
        for (uint32_t u0 = 0; u0 < vec_per_row; u0 += 32) {                    // vec_per_row % 32 == 0
            uint4 a[ROWS], b[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                ld_stream_256(slice_src + (size_t(y0 + r) * vec_per_row + u0 + lane) * 32u, policy, evict_first != 0, a[r], b[r]);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const uint32_t h = test.pack(a[r]) | (test.pack(b[r]) << 8);   // 16 voxels of this lane
                const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, h, 1);
                const uint32_t w = (lane & 1u) ? 0u : (h | (other << 16));
                if (!(lane & 1u)) slice_dst[(y0 + r) * row_stride + ((u0 + lane) >> 1)] = w;
                seen |= w;
            }
        }
        if (occ && __any_sync(0xFFFFFFFFu, seen != 0u) && lane == 0) occ[(z / OCC_CZ) * occ_ncy + y0 / OCC_CY] = tag;
    }
    ktrace_last(KT_THR_LAST);
}

// General path (ragged nx or unaligned base): one warp per output word, one voxel per lane.
template <typename T>
__global__ void __launch_bounds__(256) k_threshold_pack_rows(const DynArgs* __restrict__ dyn, uint32_t nx, uint32_t n_words,
                                                             T lo, T hi, BitDst dst, uint8_t* __restrict__ occ, uint32_t occ_ncy) {
    pdl_wait();
    ktrace(KT_THRESHOLD);
    const T* __restrict__ vol = static_cast<const T*>(dyn->vol);
    const uint8_t tag = occ_tag(dyn->gen);
    const unsigned lane = lane_id();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t wi = warp; wi < n_words; wi += n_warps) {
        const uint32_t row = wi / dst.W;
        const uint32_t x = (wi - row * dst.W) * 32 + lane;
        const bool p = x < nx && in_range(vol[size_t(row) * nx + x], lo, hi);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, p);
        if (lane == 0) {
            dst.p[dst.index(wi)] = b;
            if (occ && b) { const uint32_t z = row / dst.ny, y = row - z * dst.ny; occ[(z / OCC_CZ) * occ_ncy + y / OCC_CY] = tag; }
        }
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
template <>
double cast_bound<double>(double v) { return v; }

template <typename T>
static cudaError_t threshold_pack_t(mamri_ctx* c, int vol_aligned16, int nx, int ny, int nz, double lo, double hi,
                                    BitDst dst, cudaStream_t s) {
    const DynArgs* dyn = c->d_dyn;
    const uint32_t rows = uint32_t(ny) * nz;
    const uint32_t n_words = rows * dst.W;
    const T tlo = cast_bound<T>(lo), thi = cast_bound<T>(hi);
    const bool flat = (nx % 32 == 0) && vol_aligned16;
    uint8_t* occ = dst.linear ? nullptr : c->d_occ_raw;          // only the closing reads the cells
    const uint32_t occ_ncy = (uint32_t(ny) + OCC_CY - 1) / OCC_CY;
    if (flat) {
        constexpr int E = 16 / sizeof(T);
        // 8 warps per CTA, each warp takes two rows per trip; ~12 CTAs per SM in total (two waves of the 6 resident
        // ones): fewer, longer-lived warps amortise the per-warp set-up of a kernel that runs close to its issue
        // limit -- alone it is ~2 us slower than with 32 per SM, in a batch the machine gets ~1 % more done
        uint32_t gx = (uint32_t(ny) + 15) / 16;
        static const int per_sm = [] { const char* e = getenv("MAMRI_THR_CTAS_PER_SM"); return e ? atoi(e) : 12; }();
        const uint32_t want = (uint32_t(148 * per_sm) + uint32_t(nz) - 1) / uint32_t(nz);
        if (gx > want) gx = want;
        if (gx == 0) gx = 1;
        const dim3 grid(gx, uint32_t(nz));
        static const int evict_first = [] { const char* e = getenv("MAMRI_STREAM_HINTS"); return e ? atoi(e) : 1; }();
        const DynArgs* src = dyn;
        if constexpr (sizeof(T) == 2) {
            constexpr bool SG = std::is_signed<T>::value;
            const uint32_t bias = SG ? 0x8000u : 0u;
            uint32_t l = (uint32_t(uint16_t(tlo)) ^ bias), h = (uint32_t(uint16_t(thi)) ^ bias);
            if (l > h) { l = 0xFFFFu; h = 0u; }                 // empty range: a >= 0xFFFF && a <= 0
            const uint32_t lo_c = (l & 0x7FFFu) * 0x10001u, hi_c = (h | 0x8000u) * 0x10001u;
            const bool lm = (l & 0x8000u) != 0, hm = (h & 0x8000u) != 0, ch = h != 0xFFFFu;
            const uint32_t vpr = uint32_t(nx) / E;
            const bool full = (ny % 2 == 0) && (vpr % 64u == 0);
            static const int use_v8 = [] { const char* e = getenv("MAMRI_THR_V8"); return e ? atoi(e) : 1; }();
            const bool wide = use_v8 && vol_aligned16 >= 2 && (nx % 512 == 0) && (ny % 4 == 0);
            if (wide) {
                // ROWS = 4 rows per warp trip; the grid keeps the same number of warps per SM as the 128-bit kernel
                uint32_t gx8 = (uint32_t(ny) + 31) / 32;
                if (gx8 > want) gx8 = want;
                if (gx8 == 0) gx8 = 1;
                const dim3 grid8(gx8, uint32_t(nz));
                const uint32_t vpr8 = uint32_t(nx) / 16u;
#define MAMRI_T8(LM, HM, CH) LK(k_threshold_pack_v8<RangeTest16<LM, HM, CH, SG>, 4>, grid8, 256, s, true, src, vpr8, uint32_t(ny), RangeTest16<LM, HM, CH, SG>{lo_c, hi_c}, dst.p, dst.row_stride, dst.slice_stride, dst.off, evict_first, occ, occ_ncy)
                if (!ch) { if (lm) MAMRI_T8(true, true, false); else MAMRI_T8(false, true, false); }
                else if (lm) { if (hm) MAMRI_T8(true, true, true); else MAMRI_T8(true, false, true); }
                else { if (hm) MAMRI_T8(false, true, true); else MAMRI_T8(false, false, true); }
#undef MAMRI_T8
                prof_mark(c, s, "threshold_pack");
                return cudaGetLastError();
            }
#define MAMRI_T16(LM, HM, CH)                                                                                         \
    do { if (full) LK(k_threshold_pack_vec<2, RangeTest16<LM, HM, CH, SG>, true>, grid, 256, s, true, src, vpr, uint32_t(ny), RangeTest16<LM, HM, CH, SG>{lo_c, hi_c}, dst.p, dst.row_stride, dst.slice_stride, dst.off, evict_first, occ, occ_ncy); \
         else LK(k_threshold_pack_vec<2, RangeTest16<LM, HM, CH, SG>, false>, grid, 256, s, true, src, vpr, uint32_t(ny), RangeTest16<LM, HM, CH, SG>{lo_c, hi_c}, dst.p, dst.row_stride, dst.slice_stride, dst.off, evict_first, occ, occ_ncy); } while (0)
            if (!ch) { if (lm) MAMRI_T16(true, true, false); else MAMRI_T16(false, true, false); }
            else if (lm) { if (hm) MAMRI_T16(true, true, true); else MAMRI_T16(true, false, true); }
            else { if (hm) MAMRI_T16(false, true, true); else MAMRI_T16(false, false, true); }
#undef MAMRI_T16
        } else {
            RangeTest<T> t{tlo, thi};
            LK(k_threshold_pack_vec<int(sizeof(T)), RangeTest<T>, false>, grid, 256, s, true, src, uint32_t(nx) / E, uint32_t(ny), t, dst.p, dst.row_stride, dst.slice_stride, dst.off, evict_first, occ, occ_ncy);
        }
    } else {
        uint32_t blocks = (n_words + 7) / 8;
        const uint32_t cap = 148 * 8 * 8;
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        LK(k_threshold_pack_rows<T>, blocks, 256, s, true, dyn, uint32_t(nx), n_words, tlo, thi, dst, occ, occ_ncy);
    }
    prof_mark(c, s, "threshold_pack");
    return cudaGetLastError();
}

// Geometry of the padded bit volumes of one closing: 1 pad word left and >= 1 right of a row (row stride a
// multiple of 16 bytes, so whole rows can be moved with bulk copies), 2R pad rows and
// slices on each side (R for the dilation's apron + R for the reach of the ball from there).
struct PadGeom {
    uint32_t W, Wp, Hp, Dp, slice, words;
    __host__ PadGeom(int nx, int ny, int nz, int R) {
        W = uint32_t(nx + 31) / 32;
        Wp = (W + 2 + 3) & ~3u; Hp = uint32_t(ny + 4 * R); Dp = uint32_t(nz + 4 * R);
        slice = Wp * Hp;
        words = slice * Dp;
    }
};

// The apron of the padded raw mask must be zero; it is never written, so it is cleared only when the
// geometry changes (outside the captured graph).
cudaError_t prepare_raw_apron(mamri_ctx* c, int nx, int ny, int nz, int radius, cudaStream_t s) {
    if (radius == 0) return cudaSuccess;
    if (c->raw_nx == nx && c->raw_ny == ny && c->raw_nz == nz && c->raw_r == radius) return cudaSuccess;
    const PadGeom g(nx, ny, nz, radius);
    cudaError_t e = cudaMemsetAsync(c->d_raw, 0, size_t(g.words) * 4, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_open, 0, size_t(g.words) * 4, s);   // same rule for the opening's scratch
    if (e != cudaSuccess) return e;
    c->raw_nx = nx; c->raw_ny = ny; c->raw_nz = nz; c->raw_r = radius;
    return cudaSuccess;
}

cudaError_t launch_threshold_pack(mamri_ctx* c, int vol_aligned16, int dtype, int nx, int ny, int nz, double lo,
                                  double hi, int radius, cudaStream_t s) {
    BitDst dst;
    dst.W = uint32_t(nx + 31) / 32;
    dst.ny = uint32_t(ny);
    if (radius == 0) {                       // no closing: straight into the mask the labelling reads
        dst.p = c->d_closed;
        dst.row_stride = dst.W; dst.slice_stride = dst.W * dst.ny; dst.off = 0; dst.linear = true;
    } else {
        const PadGeom g(nx, ny, nz, radius);
        dst.p = c->d_raw;
        dst.row_stride = g.Wp; dst.slice_stride = g.slice;
        dst.off = uint32_t(2 * radius) * g.slice + uint32_t(2 * radius) * g.Wp + 1;
        dst.linear = false;
    }
    switch (dtype) {
        case MAMRI_U8:  return threshold_pack_t<uint8_t>(c, vol_aligned16, nx, ny, nz, lo, hi, dst, s);
        case MAMRI_I16: return threshold_pack_t<int16_t>(c, vol_aligned16, nx, ny, nz, lo, hi, dst, s);
        case MAMRI_U16: return threshold_pack_t<uint16_t>(c, vol_aligned16, nx, ny, nz, lo, hi, dst, s);
        case MAMRI_I32: return threshold_pack_t<int32_t>(c, vol_aligned16, nx, ny, nz, lo, hi, dst, s);
        case MAMRI_F32: return threshold_pack_t<float>(c, vol_aligned16, nx, ny, nz, lo, hi, dst, s);
        case MAMRI_F64: return threshold_pack_t<double>(c, vol_aligned16, nx, ny, nz, lo, hi, dst, s);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// closing = dilation on the r-grown domain, then erosion back on the image domain
// ------------------------------------------------------------------------------------------------
// ITK's ball of radius R (FlatStructuringElement::Ball, radiusIsParametric = false) is
// {d : dx^2+dy^2+dz^2 <= R^2+R}.  On bit-packed rows it is, for every (dy,dz) with
// dy^2+dz^2 <= R^2+R, an x-interval of half-width h(dy,dz) = floor(sqrt(R^2+R-dy^2-dz^2)).  With
// S_h(row) = the row dilated (eroded) along x by h (funnel shifts over three neighbouring words):
//   pass A   P_a(y', z) = OP_{dz} S_{h(a,dz)}(y', z+dz)      the x-z part of the ball at |dy| = a
//   pass B   out(y, z)  = OP_{dy} P_|dy|(y+dy, z)              (OP = OR for dilation, AND for erosion)
// Pass A: a thread owns one word column (xw, y') over a chunk of slices and slides a register window
// of S values along z, so each source word is loaded once (+ chunk halo) and shifted once; all loads
// and stores are coalesced along the slice.  Values of a with identical h(a, .) share one plane
// (for R = 2: a = 0 and 1).  Pass B is 2R+1 coalesced loads per output word.  All volumes share one
// padded layout with a zero apron, so neither pass has bounds checks, shared memory or barriers.
__host__ __device__ constexpr int isqrt_c(int v) { int r = 0; while ((r + 1) * (r + 1) <= v) ++r; return r; }

template <int R>
struct Ball {
    static constexpr int R2 = R * R + R;
    __host__ __device__ static constexpr int h(int a, int d) {
        return (R2 - a * a - d * d) < 0 ? -1 : isqrt_c(R2 - a * a - d * d);
    }
    __host__ __device__ static constexpr bool same(int a, int b) {
        for (int d = -R; d <= R; ++d)
            if (h(a, d) != h(b, d)) return false;
        return true;
    }
    __host__ __device__ static constexpr int canon(int a) {
        for (int b = 0; b < a; ++b)
            if (same(a, b)) return b;
        return a;
    }
    __host__ __device__ static constexpr int plane(int a) {      // index of a's plane among the distinct ones
        int n = 0;
        for (int b = 0; b < canon(a); ++b) n += (canon(b) == b) ? 1 : 0;
        return n;
    }
    __host__ __device__ static constexpr int n_planes() {
        int n = 0;
        for (int b = 0; b <= R; ++b) n += (canon(b) == b) ? 1 : 0;
        return n;
    }
};

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

template <int R, bool ERODE>
__device__ __forceinline__ void load_shifted(const uint32_t* __restrict__ p, bool has_l, bool has_r, uint32_t (&S)[R + 1]) {
    const uint32_t c = p[0];
    const uint32_t l = has_l ? p[-1] : 0u, r = has_r ? p[1] : 0u;
    S[0] = c;
#pragma unroll
    for (int k = 1; k <= R; ++k) {
        const uint32_t a = (c << k) | (l >> (32 - k)), b = (c >> k) | (r << (32 - k));
        S[k] = ERODE ? (S[k - 1] & a & b) : (S[k - 1] | a | b);
    }
}

template <int R, bool ERODE>
__global__ void __launch_bounds__(256) k_morph_planes(const uint32_t* __restrict__ src, uint32_t* __restrict__ planes,
                                                      uint32_t Wp, uint32_t slice, uint32_t words, uint32_t y_lo,
                                                      uint32_t y_cnt, uint32_t z_lo, uint32_t z_hi, uint32_t zc,
                                                      uint32_t n_chunks) {
    pdl_wait();
    const uint32_t per_chunk = Wp * y_cnt;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per_chunk * n_chunks) return;
    const uint32_t chunk = t / per_chunk, rem = t - chunk * per_chunk;
    const uint32_t yy = rem / Wp, xw = rem - yy * Wp;
    const uint32_t row_off = (y_lo + yy) * Wp + xw;
    const uint32_t z0 = z_lo + chunk * zc, z1 = min(z0 + zc, z_hi);
    const bool has_l = xw > 0, has_r = xw + 1 < Wp;
    uint32_t S[2 * R + 1][R + 1];                   // S[j] = shifted rows of slice z - R + j
#pragma unroll
    for (int j = 0; j < 2 * R; ++j) load_shifted<R, ERODE>(src + (z0 - R + j) * slice + row_off, has_l, has_r, S[j]);
    for (uint32_t z = z0; z < z1; ++z) {
        load_shifted<R, ERODE>(src + (z + R) * slice + row_off, has_l, has_r, S[2 * R]);
        static_for<0, R + 1>([&](auto ia) {
            constexpr int a = decltype(ia)::value;
            if constexpr (Ball<R>::canon(a) == a) {
                uint32_t acc = ERODE ? 0xFFFFFFFFu : 0u;
                static_for<0, 2 * R + 1>([&](auto id) {
                    constexpr int dz = decltype(id)::value - R;
                    constexpr int hh = Ball<R>::h(a, dz);
                    if constexpr (hh >= 0) acc = ERODE ? (acc & S[dz + R][hh]) : (acc | S[dz + R][hh]);
                });
                constexpr uint32_t pl = uint32_t(Ball<R>::plane(a));
                planes[pl * words + z * slice + row_off] = acc;
            }
        });
#pragma unroll
        for (int j = 0; j < 2 * R; ++j)
#pragma unroll
            for (int k = 0; k <= R; ++k) S[j][k] = S[j + 1][k];
    }
}

// Pass B.  TO_IMAGE = false: writes the padded dilation.  TO_IMAGE = true: writes the plain [nz][ny][W]
// closed mask (erosion), masking the bits beyond nx in the last word of a row.
template <int R, bool ERODE, bool TO_IMAGE>
__global__ void __launch_bounds__(256) k_morph_combine(const uint32_t* __restrict__ planes, uint32_t* __restrict__ dst,
                                                       uint32_t Wp, uint32_t slice, uint32_t words, uint32_t x_lo,
                                                       uint32_t x_cnt, uint32_t y_lo, uint32_t y_cnt, uint32_t z_lo,
                                                       uint32_t z_cnt, uint32_t tail_mask) {
    pdl_wait();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= x_cnt * y_cnt * z_cnt) return;
    const uint32_t row = t / x_cnt, xx = t - row * x_cnt;
    const uint32_t zz = row / y_cnt, yy = row - zz * y_cnt;
    const uint32_t idx = (z_lo + zz) * slice + (y_lo + yy) * Wp + (x_lo + xx);
    uint32_t acc = ERODE ? 0xFFFFFFFFu : 0u;
    static_for<0, 2 * R + 1>([&](auto id) {
        constexpr int dy = decltype(id)::value - R;
        constexpr uint32_t pl = uint32_t(Ball<R>::plane(dy < 0 ? -dy : dy));
        const uint32_t v = planes[pl * words + idx + uint32_t(dy * int(Wp))];
        acc = ERODE ? (acc & v) : (acc | v);
    });
    if (TO_IMAGE) {
        if (xx == x_cnt - 1) acc &= tail_mask;
        dst[t] = acc;                               // (zz * ny + yy) * W + xx == t
    } else {
        dst[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// One-kernel-per-operation variant: no intermediate planes in memory.
// ------------------------------------------------------------------------------------------------
// A thread owns one word column xw for a strip of SY output rows and walks a chunk of slices.  Every
// step it reads the SY + 2R source rows of the incoming slice z' (three words each for the x shifts,
// all hits in L1/L2), folds them over y into one value per output row and per |dz| class
//   Q_c(y, z') = OP_{dy} S_{h(dy, c)}(y + dy, z')
// and scatters those into a ring of 2R+1 pending output slices, acc[z' - dz] OP= Q_|dz|; the slice
// z' - R is then complete and is stored.  The ring lives in registers (SY x (2R+1) words); source
// words are read (SY + 2R) / SY x (zc + 2R) / zc times, the result is written once.
template <int R>
struct BallZ {          // classes of dz with identical y-profiles h(., dz)
    __host__ __device__ static constexpr bool same(int a, int b) {
        for (int d = -R; d <= R; ++d)
            if (Ball<R>::h(d < 0 ? -d : d, a) != Ball<R>::h(d < 0 ? -d : d, b)) return false;
        return true;
    }
    __host__ __device__ static constexpr int canon(int a) {
        for (int b = 0; b < a; ++b)
            if (same(a, b)) return b;
        return a;
    }
    __host__ __device__ static constexpr int cls(int a) {
        int n = 0;
        for (int b = 0; b < canon(a); ++b) n += (canon(b) == b) ? 1 : 0;
        return n;
    }
    __host__ __device__ static constexpr int n_cls() {
        int n = 0;
        for (int b = 0; b <= R; ++b) n += (canon(b) == b) ? 1 : 0;
        return n;
    }
    __host__ __device__ static constexpr int rep(int c) {        // a |dz| of class c
        for (int b = 0; b <= R; ++b)
            if (cls(b) == c) return b;
        return 0;
    }
};

template <int R>
static cudaError_t closing_r(mamri_ctx* c, int nx, int ny, int nz, cudaStream_t s) {
    const PadGeom g(nx, ny, nz, R);
    auto chunks = [&](uint32_t per_chunk, uint32_t depth, uint32_t& zc, uint32_t& n_chunks) {
        zc = 16;                                    // keep >= one full wave of threads (148 SMs x 2048)
        while (zc > 4 && size_t(per_chunk) * ((depth + zc - 1) / zc) < size_t(148) * 2048) zc >>= 1;
        n_chunks = (depth + zc - 1) / zc;
    };
    uint32_t zc, nch;
    // ---- dilation: P_a on every padded row, slices [R, nz+3R); D on rows [R, ny+3R), same slices
    const uint32_t dz_lo = R, dz_hi = uint32_t(nz) + 3 * R;
    chunks(g.Wp * g.Hp, dz_hi - dz_lo, zc, nch);
    uint32_t threads = g.Wp * g.Hp * nch;
    LK(k_morph_planes<R, false>, (threads + 255) / 256, 256, s, false, c->d_raw, c->d_planes, g.Wp, g.slice, g.words, 0, g.Hp, dz_lo, dz_hi, zc, nch);
    prof_mark(c, s, "morph_planes_dilate");
    const uint32_t dy_cnt = uint32_t(ny) + 2 * R;
    threads = g.Wp * dy_cnt * (dz_hi - dz_lo);
    LK(k_morph_combine<R, false, false>, (threads + 255) / 256, 256, s, false, c->d_planes, c->d_dil, g.Wp, g.slice, g.words, 0, g.Wp, R, dy_cnt, dz_lo, dz_hi - dz_lo, 0xFFFFFFFFu);
    prof_mark(c, s, "morph_combine_dilate");
    // ---- erosion: P_a on rows [R, ny+3R), image slices [2R, nz+2R); E on the image domain
    const uint32_t ez_lo = 2 * R, ez_hi = uint32_t(nz) + 2 * R;
    chunks(g.Wp * dy_cnt, ez_hi - ez_lo, zc, nch);
    threads = g.Wp * dy_cnt * nch;
    LK(k_morph_planes<R, true>, (threads + 255) / 256, 256, s, false, c->d_dil, c->d_planes, g.Wp, g.slice, g.words, R, dy_cnt, ez_lo, ez_hi, zc, nch);
    prof_mark(c, s, "morph_planes_erode");
    const uint32_t tail = (nx % 32) ? (0xFFFFFFFFu >> (32 - nx % 32)) : 0xFFFFFFFFu;
    threads = g.W * uint32_t(ny) * uint32_t(nz);
    LK(k_morph_combine<R, true, true>, (threads + 255) / 256, 256, s, false, c->d_planes, c->d_closed, g.Wp, g.slice, g.words, 1, g.W, 2 * R, uint32_t(ny), 2 * R, uint32_t(nz), tail);
    prof_mark(c, s, "morph_combine_erode");
    return cudaGetLastError();
}

// One CTA produces a tile of TY rows x TZ slices x all word columns.  Thread 0 fetches the source tile
// with its halo of R rows / slices -- the rows of one slice are contiguous in the padded layout, so it is
// one bulk copy (cp.async.bulk, completion on an mbarrier) per slice, no address arithmetic in the other
// threads and no registers staging the data.  Then each thread owns one word column for a strip of SY
// output rows and walks the slices of the tile in shared memory as described above.
struct TileArgs {
    uint32_t Wp, slice;                 // padded row / slice strides (words); Wp % 4 == 0
    uint32_t x_lo, x_cnt;               // output word columns
    uint32_t y_lo, y_cnt;               // output rows (padded coordinates)
    uint32_t z_lo, z_hi;                // output slices (padded coordinates)
    uint32_t TY, TZ;                    // tile: output rows / slices per CTA (TY % SY == 0)
    uint32_t out_w, out_ny;             // TO_IMAGE: row length / rows per slice of the plain mask
    uint32_t tail_mask;
    // Occupancy of the source: a grid of cells (cy rows x cz slices) over rows [oy, oy + dom_y) and slices
    // [oz, oz + dom_z) of the padded volume; everything outside that domain is the zero apron.  A tile whose
    // source cells are all clear writes zeros without reading its source.
    const uint8_t* occ_in;
    uint32_t occ_stride, cy, cz, oy, oz, dom_y, dom_z;
    uint8_t* occ_out;                   // one flag per tile of this pass: any output word non-zero
    uint32_t occ_tagged;                // occ_in holds scan tags (raw-mask cells) instead of 0/1 flags
    uint32_t invert_out;                // padded output only: store the complement (inside the row's tail mask)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template <int R, bool ERODE, bool TO_IMAGE, int SY>
__global__ void __launch_bounds__(1024) k_morph_tile(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                     const DynArgs* __restrict__ dyn, TileArgs a) {
    extern __shared__ __align__(128) uint32_t tile[];              // [TZ+2R][TY+2R][Wp], then the mbarrier
    constexpr int RING = 2 * R + 1, NC = BallZ<R>::n_cls();
    constexpr uint32_t NEUTRAL = ERODE ? 0xFFFFFFFFu : 0u;
    const uint32_t rows_alloc = a.TY + 2 * R;
    const uint32_t y0t = a.y_lo + blockIdx.x * a.TY, z0 = a.z_lo + blockIdx.y * a.TZ;
    const uint32_t y_end = a.y_lo + a.y_cnt, z1 = min(z0 + a.TZ, a.z_hi);
    const uint32_t ys0 = y0t - R, zs0 = z0 - R;                    // first source row / slice of the tile
    const uint32_t rows_src = min(rows_alloc, y_end + R - ys0);    // rows past y_end - 1 + R feed no output
    const uint32_t n_slices = z1 + R - zs0;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + (a.TZ + 2 * R) * rows_alloc * a.Wp);
    const uint32_t bar_s = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                                    // the source is the previous kernel's output
    ktrace(ERODE ? KT_ERODE : KT_CLOSE);
    if (a.occ_in) {
        const uint8_t want = a.occ_tagged ? occ_tag(dyn->gen) : uint8_t(1);
        // source rows [ys0, ys0 + rows_src) x slices [zs0, zs0 + n_slices), clipped to the occupancy domain
        const int ya = max(int(ys0) - int(a.oy), 0), yb = min(int(ys0 + rows_src) - int(a.oy), int(a.dom_y));
        const int za = max(int(zs0) - int(a.oz), 0), zb = min(int(zs0 + n_slices) - int(a.oz), int(a.dom_z));
        int hit = 0;
        if (ya < yb && za < zb) {
            const int c0 = ya / int(a.cy), nyc = (yb - 1) / int(a.cy) - c0 + 1;
            const int d0 = za / int(a.cz), nzc = (zb - 1) / int(a.cz) - d0 + 1;
            for (int i = threadIdx.x; i < nyc * nzc; i += blockDim.x)
                hit |= a.occ_in[(d0 + i / nyc) * a.occ_stride + c0 + i % nyc] == want;
        }
        if (!__syncthreads_or(hit)) {                              // air: the output of either pass is zero
            if (a.occ_out && threadIdx.x == 0) a.occ_out[blockIdx.y * gridDim.x + blockIdx.x] = 0;
            const uint32_t n_strips = a.TY / SY;
            if (threadIdx.x >= a.x_cnt * n_strips) return;
            const uint32_t strip = threadIdx.x / a.x_cnt, xx = threadIdx.x - strip * a.x_cnt;
            const uint32_t xw = a.x_lo + xx, y0 = y0t + strip * SY;
            for (uint32_t zo = z0; zo < z1; ++zo)
#pragma unroll
                for (int y = 0; y < SY; ++y)
                    if (y0 + y < y_end) {
                        if (TO_IMAGE) dst[((zo - a.z_lo) * a.out_ny + (y0 + y - a.y_lo)) * a.out_w + xx] = 0u;
                        else dst[zo * a.slice + (y0 + y) * a.Wp + xw] = 0u;
                    }
            return;
        }
    }
    if (a.occ_out) {
        if (threadIdx.x == 0) a.occ_out[blockIdx.y * gridDim.x + blockIdx.x] = 0;
        __syncthreads();                                           // ... before any thread of the CTA raises it
    }
    if (threadIdx.x == 0) {
        const uint32_t bytes = rows_src * a.Wp * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes * n_slices) : "memory");
        for (uint32_t k = 0; k < n_slices; ++k) {
            const uint32_t* g = src + size_t(zs0 + k) * a.slice + ys0 * a.Wp;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(tile + k * rows_alloc * a.Wp)), "l"(g), "r"(bytes), "r"(bar_s) : "memory");
        }
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_s) : "memory");
    }
    const uint32_t n_strips = a.TY / SY;
    if (threadIdx.x >= a.x_cnt * n_strips) return;
    const uint32_t strip = threadIdx.x / a.x_cnt, xx = threadIdx.x - strip * a.x_cnt;
    const uint32_t xw = a.x_lo + xx;
    const uint32_t y0 = y0t + strip * SY;                          // first output row of the strip
    if (y0 >= y_end) return;
    const bool has_l = xw > 0, has_r = xw + 1 < a.Wp;
    const uint32_t* col = tile + (strip * SY) * a.Wp + xw;         // local row strip*SY = source row y0 - R

    uint32_t acc[SY][RING];
    uint32_t out_any = 0;
#pragma unroll
    for (int y = 0; y < SY; ++y)
#pragma unroll
        for (int j = 0; j < RING; ++j) acc[y][j] = NEUTRAL;

    for (uint32_t k = 0; k < n_slices; ++k) {                      // incoming source slice zs0 + k
        uint32_t Q[NC][SY];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int y = 0; y < SY; ++y) Q[c][y] = NEUTRAL;
        const uint32_t* p = col + k * rows_alloc * a.Wp;
        static_for<0, SY + 2 * R>([&](auto ir) {
            constexpr int r = decltype(ir)::value;                 // source row y0 - R + r
            uint32_t S[R + 1];
            load_shifted<R, ERODE>(p + r * a.Wp, has_l, has_r, S); // rows past rows_src: unloaded, feed discarded outputs only
            static_for<0, SY>([&](auto iy) {
                constexpr int y = decltype(iy)::value;
                constexpr int dy = r - R - y;
                if constexpr (dy >= -R && dy <= R) {
                    static_for<0, NC>([&](auto ic) {
                        constexpr int c = decltype(ic)::value;
                        constexpr int hh = Ball<R>::h(dy < 0 ? -dy : dy, BallZ<R>::rep(c));
                        if constexpr (hh >= 0) Q[c][y] = ERODE ? (Q[c][y] & S[hh]) : (Q[c][y] | S[hh]);
                    });
                }
            });
        });
        // acc[y][j] is the output slice zs0 + k - R + j; it sees the incoming slice at dz = R - j
#pragma unroll
        for (int y = 0; y < SY; ++y)
            static_for<0, RING>([&](auto ij) {
                constexpr int j = decltype(ij)::value;
                constexpr int adz = (R - j) < 0 ? (j - R) : (R - j);
                constexpr int c = BallZ<R>::cls(adz);
                acc[y][j] = ERODE ? (acc[y][j] & Q[c][y]) : (acc[y][j] | Q[c][y]);
            });
        if (k >= 2 * R) {                                           // output slice zs0 + k - R = z0 + (k - 2R) is complete
            const uint32_t zo = z0 + k - 2 * R;
#pragma unroll
            for (int y = 0; y < SY; ++y) {
                if (y0 + y < y_end) {
                    uint32_t v = acc[y][0];
                    if (TO_IMAGE) {
                        if (xx == a.x_cnt - 1) v &= a.tail_mask;
                        dst[((zo - a.z_lo) * a.out_ny + (y0 + y - a.y_lo)) * a.out_w + xx] = v;
                    } else {
                        if (a.invert_out) v = ~v;
                        if (xx == a.x_cnt - 1) v &= a.tail_mask;
                        dst[zo * a.slice + (y0 + y) * a.Wp + xw] = v;
                        out_any |= v;
                    }
                }
            }
        }
#pragma unroll
        for (int y = 0; y < SY; ++y) {
#pragma unroll
            for (int j = 0; j + 1 < RING; ++j) acc[y][j] = acc[y][j + 1];
            acc[y][RING - 1] = NEUTRAL;
        }
    }
    if (!TO_IMAGE && a.occ_out && out_any) a.occ_out[blockIdx.y * gridDim.x + blockIdx.x] = 1;
}

// Tile shape for a padded row of Wp words: the largest of a few candidates whose tile + halo fits the
// shared-memory budget (two CTAs per SM) with at most 1024 threads; false = row too wide, use the planes kernels.
template <int R, int SY>
static bool tile_plan(uint32_t Wp, uint32_t& TY, uint32_t& TZ, uint32_t& smem) {
    static const uint32_t cand[][2] = {{32, 16}, {32, 8}, {16, 8}, {16, 4}, {8, 4}};
    for (auto& c : cand) {
        const uint32_t bytes = Wp * (c[0] + 2 * R) * (c[1] + 2 * R) * 4u + 16u;
        if (bytes <= 100u * 1024u && Wp * (c[0] / SY) <= 1024u && c[0] % SY == 0) {
            TY = c[0]; TZ = c[1]; smem = bytes;
            return true;
        }
    }
    return false;
}

template <int R, int SY>
static cudaError_t closing_tile_r(mamri_ctx* c, int nx, int ny, int nz, uint32_t TY, uint32_t TZ, uint32_t smem,
                                  cudaStream_t s) {
    const PadGeom g(nx, ny, nz, R);
    TileArgs a;
    a.Wp = g.Wp; a.slice = g.slice; a.TY = TY; a.TZ = TZ;
    // dilation on the r-grown domain: rows [R, ny+3R), slices [R, nz+3R), every padded word column
    a.x_lo = 0; a.x_cnt = g.Wp; a.y_lo = R; a.y_cnt = uint32_t(ny) + 2 * R; a.z_lo = R; a.z_hi = uint32_t(nz) + 3 * R;
    a.out_w = 0; a.out_ny = 0; a.tail_mask = 0xFFFFFFFFu;
    uint32_t threads = ((a.x_cnt * (TY / SY) + 31) / 32) * 32;
    dim3 grid((a.y_cnt + TY - 1) / TY, (a.z_hi - a.z_lo + TZ - 1) / TZ);
    static const int use_occ = [] { const char* e = getenv("MAMRI_TILE_OCC"); return e ? atoi(e) : 1; }();
    const uint32_t ncy = (uint32_t(ny) + OCC_CY - 1) / OCC_CY, ncz = (uint32_t(nz) + OCC_CZ - 1) / OCC_CZ;
    const bool occ = use_occ && size_t(ncy) * ncz <= c->occ_cap && size_t(grid.x) * grid.y <= c->occ_cap;
    // dilation: source = raw mask, occupancy cells in image coordinates (image row y is padded row y + 2R)
    a.occ_in = occ ? c->d_occ_raw : nullptr;
    a.occ_stride = ncy; a.cy = OCC_CY; a.cz = OCC_CZ; a.oy = 2 * R; a.oz = 2 * R; a.dom_y = uint32_t(ny); a.dom_z = uint32_t(nz);
    a.occ_out = occ ? c->d_occ_dil : nullptr;
    a.occ_tagged = 1; a.invert_out = 0;
    const dim3 dil_grid = grid;
    LKS(k_morph_tile<R, false, false, SY>, grid, threads, smem, s, false, c->d_raw, c->d_dil, c->d_dyn, a);
    prof_mark(c, s, "dilate");
    // erosion: source = dilated mask, occupancy = the dilation's per-tile flags (its tiles start at padded row / slice R)
    a.occ_in = occ ? c->d_occ_dil : nullptr;
    a.occ_stride = dil_grid.x; a.cy = TY; a.cz = TZ; a.oy = R; a.oz = R; a.dom_y = uint32_t(ny) + 2 * R; a.dom_z = uint32_t(nz) + 2 * R;
    a.occ_out = nullptr;
    a.occ_tagged = 0;
    // erosion back on the image domain, straight into the plain [nz][ny][W] mask
    a.x_lo = 1; a.x_cnt = g.W; a.y_lo = 2 * R; a.y_cnt = uint32_t(ny); a.z_lo = 2 * R; a.z_hi = uint32_t(nz) + 2 * R;
    a.out_w = g.W; a.out_ny = uint32_t(ny);
    a.tail_mask = (nx % 32) ? (0xFFFFFFFFu >> (32 - nx % 32)) : 0xFFFFFFFFu;
    threads = ((a.x_cnt * (TY / SY) + 31) / 32) * 32;
    grid = dim3((a.y_cnt + TY - 1) / TY, (a.z_hi - a.z_lo + TZ - 1) / TZ);
    LKS(k_morph_tile<R, true, true, SY>, grid, threads, smem, s, false, c->d_dil, c->d_closed, c->d_dyn, a);
    prof_mark(c, s, "erode");
    return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// Fused closing: dilation and erosion of one tile in ONE kernel, the dilated tile never leaves shared memory
// ------------------------------------------------------------------------------------------------
// A CTA produces TY rows x TZ slices x all word columns of the closed mask.  It fetches the raw tile with a halo
// of 2R rows / slices (one bulk copy per slice, completion on an mbarrier), dilates it into a second shared-memory
// tile that carries a halo of R (the r-grown domain of the safe-border rule falls out of the zero apron of the
// padded layout), and erodes that tile straight into the plain [nz][ny][W] mask.  Both passes use the same
// strip walk as k_morph_tile: a unit of work is one word column x a strip of SY rows x a chunk of slices, walked
// with a register ring of 2R+1 partial results.  The source is read (TY+4R)(TZ+4R)/(TY TZ) times from L2 instead of
// the dilated volume making a round trip through memory, and one launch replaces two.
struct FusedArgs {
    uint32_t Wp, slice;                 // padded row / slice strides (words); Wp % 4 == 0
    uint32_t W, ny, nz;                 // image: words per row, rows, slices
    uint32_t pad;                       // apron rows / slices on each side of the padded volume (>= 2R)
    uint32_t TY, TZ;                    // output tile; (TY + 2R) % SYD == 0 and TY % SYE == 0
    uint32_t zs_d, zs_e;                // slice chunks per strip in the dilation / erosion (more, shorter walks)
    uint32_t tail_mask;
    const uint8_t* occ;                 // occupancy cells of the raw mask (scan tags), NULL = read every tile
    uint32_t occ_stride;
    uint32_t tiles_y, n_tiles;          // tiles along y / in all; CTAs draw them from DevScalars::ticket_close
};

// The strip walk: `col` points at (first source row of the strip, word column) of the first source slice of a
// shared-memory tile; source slices [k0, k1) are folded, output slice m (centred on source slice m + R) is
// complete after source slice m + 2R and handed to store(m, y, value) for the SY rows of the strip.
template <int R, bool ERODE, int SY, typename StoreF>
__device__ __forceinline__ void ball_walk(const uint32_t* col, uint32_t row_stride, uint32_t slice_stride, uint32_t k0,
                                          uint32_t k1, bool has_l, bool has_r, StoreF&& store) {
    constexpr int RING = 2 * R + 1, NC = BallZ<R>::n_cls();
    constexpr uint32_t NEUTRAL = ERODE ? 0xFFFFFFFFu : 0u;
    uint32_t acc[SY][RING];
#pragma unroll
    for (int y = 0; y < SY; ++y)
#pragma unroll
        for (int j = 0; j < RING; ++j) acc[y][j] = NEUTRAL;
    for (uint32_t k = k0; k < k1; ++k) {
        uint32_t Q[NC][SY];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int y = 0; y < SY; ++y) Q[c][y] = NEUTRAL;
        const uint32_t* p = col + k * slice_stride;
        static_for<0, SY + 2 * R>([&](auto ir) {
            constexpr int r = decltype(ir)::value;
            uint32_t S[R + 1];
            load_shifted<R, ERODE>(p + r * row_stride, has_l, has_r, S);
            static_for<0, SY>([&](auto iy) {
                constexpr int y = decltype(iy)::value;
                constexpr int dy = r - R - y;
                if constexpr (dy >= -R && dy <= R) {
                    static_for<0, NC>([&](auto ic) {
                        constexpr int c = decltype(ic)::value;
                        constexpr int hh = Ball<R>::h(dy < 0 ? -dy : dy, BallZ<R>::rep(c));
                        if constexpr (hh >= 0) Q[c][y] = ERODE ? (Q[c][y] & S[hh]) : (Q[c][y] | S[hh]);
                    });
                }
            });
        });
#pragma unroll
        for (int y = 0; y < SY; ++y)
            static_for<0, RING>([&](auto ij) {
                constexpr int j = decltype(ij)::value;
                constexpr int adz = (R - j) < 0 ? (j - R) : (R - j);
                constexpr int c = BallZ<R>::cls(adz);
                acc[y][j] = ERODE ? (acc[y][j] & Q[c][y]) : (acc[y][j] | Q[c][y]);
            });
        if (k - k0 >= 2 * R) {
#pragma unroll
            for (int y = 0; y < SY; ++y) store(k - 2 * R, y, acc[y][0]);
        }
#pragma unroll
        for (int y = 0; y < SY; ++y) {
#pragma unroll
            for (int j = 0; j + 1 < RING; ++j) acc[y][j] = acc[y][j + 1];
            acc[y][RING - 1] = NEUTRAL;
        }
    }
}

// (Strides and tile shape are run-time values on purpose: variants with the padded row length and the tile shape as
// template constants -- every shared-memory address base + immediate, a fifth fewer instructions -- measured 13 us
// SLOWER in the dilation pass on the B200, so they were dropped.)
template <int R, int SYD, int SYE>
__global__ void __launch_bounds__(512) k_close_fused(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                     const DynArgs* __restrict__ dyn, DevScalars* sc, FusedArgs a) {
    extern __shared__ __align__(128) uint32_t tile[];
    __shared__ uint32_t s_ticket;
    const uint32_t SY_ = a.TY + 4 * R, SZ_ = a.TZ + 4 * R;          // source rows / slices held
    const uint32_t DY = a.TY + 2 * R, DZ = a.TZ + 2 * R;            // dilated rows / slices held
    uint32_t* s_src = tile;                                         // [SZ_][SY_][Wp]
    uint32_t* s_dil = tile + SZ_ * SY_ * a.Wp;                      // [DZ][DY][Wp]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_dil + DZ * DY * a.Wp);
    const uint32_t bar_s = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    ktrace(KT_CLOSE);
    const uint8_t want = a.occ ? occ_tag(dyn->gen) : uint8_t(0);
    uint32_t phase = 0;                                             // parity of the mbarrier phase the next fetch completes
    // Persistent CTAs draw tiles from a ticket counter: most tiles of an MR volume are air and cost next to nothing, so
    // the tiles that do hold something end up spread over all the CTAs instead of queueing behind each other in the
    // launch order of a static grid (measured: the last CTA of a static grid finished 7 us after the first).
    while (true) {
        __syncthreads();                                            // everyone is done with the previous tile's shared memory
        if (threadIdx.x == 0) s_ticket = atomicAdd(&sc->ticket_close, 1u);
        __syncthreads();
        const uint32_t t = s_ticket;
        if (t >= a.n_tiles) break;
        const uint32_t ty = t % a.tiles_y, tz = t / a.tiles_y;
        const uint32_t Y0 = ty * a.TY, Z0 = tz * a.TZ;              // first output row / slice (image coordinates)
        const uint32_t py0 = Y0 + a.pad - 2 * R, pz0 = Z0 + a.pad - 2 * R;   // first source row / slice (padded coordinates)
        const uint32_t rows_src = min(SY_, a.ny + 2 * a.pad - py0); // rows / slices past the padded volume feed no output
        const uint32_t n_sl = min(SZ_, a.nz + 2 * a.pad - pz0);
        const uint32_t out_rows = min(a.TY, a.ny - Y0), out_sl = min(a.TZ, a.nz - Z0);
        if (a.occ) {
            // source rows / slices of the tile in image coordinates, clipped to the image (the apron is air)
            const int ya = max(int(Y0) - 2 * R, 0), yb = min(int(Y0 + a.TY) + 2 * R, int(a.ny));
            const int za = max(int(Z0) - 2 * R, 0), zb = min(int(Z0 + a.TZ) + 2 * R, int(a.nz));
            const int c0 = ya / int(OCC_CY), nyc = (yb - 1) / int(OCC_CY) - c0 + 1;
            const int d0 = za / int(OCC_CZ), nzc = (zb - 1) / int(OCC_CZ) - d0 + 1;
            int hit = 0;
            for (int i = threadIdx.x; i < nyc * nzc; i += blockDim.x)
                hit |= a.occ[(d0 + i / nyc) * a.occ_stride + c0 + i % nyc] == want;
            if (!__syncthreads_or(hit)) {                           // air: the closing of nothing is nothing
                const uint32_t per_slice = out_rows * a.W;
                for (uint32_t i = threadIdx.x; i < per_slice * out_sl; i += blockDim.x) {
                    const uint32_t zo = i / per_slice;
                    dst[(size_t(Z0 + zo) * a.ny + Y0) * a.W + (i - zo * per_slice)] = 0u;
                }
                continue;
            }
        }
        if (threadIdx.x == 0) {
            const uint32_t bytes = rows_src * a.Wp * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes * n_sl) : "memory");
            for (uint32_t k = 0; k < n_sl; ++k) {
                const uint32_t* g = src + size_t(pz0 + k) * a.slice + py0 * a.Wp;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(s_src + k * SY_ * a.Wp)), "l"(g), "r"(bytes), "r"(bar_s) : "memory");
            }
        }
        {
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
            phase ^= 1u;
        }
        ktrace(KT_CLOSE_LD);
        // ---- dilation: dilated row j (image row Y0 - R + j) folds source rows j .. j + 2R; same along the slices.
        // Slices / rows past the data that was fetched hold stale shared memory; they only reach outputs that are dropped.
        {
            const uint32_t n_strips = DY / SYD;
            const uint32_t cz = (DZ + a.zs_d - 1) / a.zs_d;
            const uint32_t n_units = a.Wp * n_strips * a.zs_d;
            for (uint32_t u = threadIdx.x; u < n_units; u += blockDim.x) {
                const uint32_t xw = u % a.Wp, rest = u / a.Wp;
                const uint32_t strip = rest % n_strips, q = rest / n_strips;
                const uint32_t m0 = q * cz, m1 = min(m0 + cz, DZ);
                if (m0 >= m1) continue;
                uint32_t* out = s_dil + (strip * SYD) * a.Wp + xw;
                ball_walk<R, false, SYD>(s_src + (strip * SYD) * a.Wp + xw, a.Wp, SY_ * a.Wp, m0, m1 + 2 * R, xw > 0, xw + 1 < a.Wp,
                                         [&](uint32_t m, int y, uint32_t v) { out[m * DY * a.Wp + uint32_t(y) * a.Wp] = v; });
            }
        }
        __syncthreads();
        ktrace(KT_CLOSE_DIL);
        // ---- erosion: output row y (image row Y0 + y) folds dilated rows y .. y + 2R
        {
            const uint32_t n_strips = a.TY / SYE;
            const uint32_t cz = (a.TZ + a.zs_e - 1) / a.zs_e;
            const uint32_t n_units = a.W * n_strips * a.zs_e;
            for (uint32_t u = threadIdx.x; u < n_units; u += blockDim.x) {
                const uint32_t xx = u % a.W, rest = u / a.W;
                const uint32_t strip = rest % n_strips, q = rest / n_strips;
                const uint32_t m0 = q * cz, m1 = min(min(m0 + cz, a.TZ), out_sl);
                if (m0 >= m1 || strip * SYE >= out_rows) continue;
                const uint32_t tail = (xx == a.W - 1) ? a.tail_mask : 0xFFFFFFFFu;
                uint32_t* out = dst + (size_t(Z0) * a.ny + Y0 + strip * SYE) * a.W + xx;
                const uint32_t rows_left = out_rows - strip * SYE;
                ball_walk<R, true, SYE>(s_dil + (strip * SYE) * a.Wp + xx + 1, a.Wp, DY * a.Wp, m0, m1 + 2 * R, true, true,
                                        [&](uint32_t m, int y, uint32_t v) {
                                            if (uint32_t(y) < rows_left) out[size_t(m) * a.ny * a.W + uint32_t(y) * a.W] = v & tail;
                                        });
            }
        }
        ktrace(KT_CLOSE_ERO);
    }
    ktrace_last(KT_CLOSE_LAST);
}

// Tile of the fused kernel for rows of Wp padded words: TY from {SYD k - 2R} with TY % SYE == 0 near the wanted size,
// TZ halved until source + dilated tile fit the shared-memory budget.  false = does not fit (very wide rows).
template <int R, int SYD, int SYE>
static bool fused_plan(uint32_t Wp, uint32_t want_ty, uint32_t want_tz, uint32_t budget, uint32_t& TY, uint32_t& TZ, uint32_t& smem) {
    for (uint32_t tz = want_tz; tz >= 2; tz >>= 1) {
        for (uint32_t ty = want_ty + 2; ty >= 4; --ty) {
            if ((ty + 2 * R) % SYD != 0 || ty % SYE != 0) continue;
            const uint32_t bytes = Wp * 4u * ((ty + 4 * R) * (tz + 4 * R) + (ty + 2 * R) * (tz + 2 * R)) + 16u;
            if (bytes <= budget) { TY = ty; TZ = tz; smem = bytes; return true; }
        }
    }
    return false;
}

constexpr uint32_t FUSED_SMEM_MAX = 200u * 1024u;

template <int R, int SYD, int SYE>
static cudaError_t closing_fused_r(mamri_ctx* c, int nx, int ny, int nz, int geom_r, bool& done, cudaStream_t s) {
    static const int e_ty = [] { const char* e = getenv("MAMRI_CLOSE_TY"); return e ? atoi(e) : 16; }();
    static const int e_tz = [] { const char* e = getenv("MAMRI_CLOSE_TZ"); return e ? atoi(e) : 16; }();
    static const int e_zsd = [] { const char* e = getenv("MAMRI_CLOSE_ZSD"); return e ? atoi(e) : 1; }();
    static const int e_zse = [] { const char* e = getenv("MAMRI_CLOSE_ZSE"); return e ? atoi(e) : 1; }();
    static const int e_budget = [] { const char* e = getenv("MAMRI_CLOSE_SMEM_KB"); return e ? atoi(e) : 100; }();
    static const int use_occ = [] { const char* e = getenv("MAMRI_TILE_OCC"); return e ? atoi(e) : 1; }();
    const PadGeom g(nx, ny, nz, geom_r);                 // the apron may be wider than this ball needs (opening before it)
    FusedArgs a;
    uint32_t smem = 0;
    done = false;
    uint32_t budget = uint32_t(e_budget) * 1024u;
    if (budget > FUSED_SMEM_MAX) budget = FUSED_SMEM_MAX;
    if (!fused_plan<R, SYD, SYE>(g.Wp, uint32_t(e_ty), uint32_t(e_tz), budget, a.TY, a.TZ, smem) &&
        !fused_plan<R, SYD, SYE>(g.Wp, uint32_t(e_ty), uint32_t(e_tz), FUSED_SMEM_MAX, a.TY, a.TZ, smem))
        return cudaSuccess;                                         // too wide: the caller takes the two-pass kernels
    a.Wp = g.Wp; a.slice = g.slice; a.W = g.W; a.ny = uint32_t(ny); a.nz = uint32_t(nz); a.pad = 2 * uint32_t(geom_r);
    a.zs_d = uint32_t(e_zsd < 1 ? 1 : e_zsd); a.zs_e = uint32_t(e_zse < 1 ? 1 : e_zse);
    a.tail_mask = (nx % 32) ? (0xFFFFFFFFu >> (32 - nx % 32)) : 0xFFFFFFFFu;
    const uint32_t ncy = (uint32_t(ny) + OCC_CY - 1) / OCC_CY, ncz = (uint32_t(nz) + OCC_CZ - 1) / OCC_CZ;
    a.occ = (use_occ && size_t(ncy) * ncz <= c->occ_cap) ? c->d_occ_raw : nullptr;
    a.occ_stride = ncy;
    const uint32_t units_d = a.Wp * ((a.TY + 2 * R) / SYD) * a.zs_d, units_e = a.W * (a.TY / SYE) * a.zs_e;
    uint32_t threads = ((units_d > units_e ? units_d : units_e) + 31) / 32 * 32;
    if (threads > 512) threads = 512;
    a.tiles_y = (uint32_t(ny) + a.TY - 1) / a.TY;
    a.n_tiles = a.tiles_y * ((uint32_t(nz) + a.TZ - 1) / a.TZ);
    // persistent CTAs: as many as fit the machine at this tile size (shared memory is what limits them), at most one per tile
    static const int e_ctas = [] { const char* e = getenv("MAMRI_CLOSE_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    uint32_t per_sm = (227u * 1024u) / (smem + 1024u);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    if (e_ctas > 0) per_sm = uint32_t(e_ctas);
    uint32_t grid = 148u * per_sm;
    if (grid > a.n_tiles) grid = a.n_tiles;
    LKS(k_close_fused<R, SYD, SYE>, grid, threads, smem, s, false, c->d_raw, c->d_closed, c->d_dyn, c->d_scalars, a);
    prof_mark(c, s, "close_fused");
    done = true;
    return cudaGetLastError();
}

template <int R>
static cudaError_t closing_fused_dispatch(mamri_ctx* c, int nx, int ny, int nz, int geom_r, bool& done, cudaStream_t s) {
    // strip heights of the two passes: 4 rows amortise the shifted loads best, 2 rows give more, shorter walks
    static const int syd = [] { const char* e = getenv("MAMRI_CLOSE_SYD"); return e ? atoi(e) : 4; }();
    static const int sye = [] { const char* e = getenv("MAMRI_CLOSE_SYE"); return e ? atoi(e) : 2; }();
    if (syd == 4 && sye == 4 && R % 2 == 0) return closing_fused_r<R, 4, 4>(c, nx, ny, nz, geom_r, done, s);
    if (syd == 4) return closing_fused_r<R, 4, 2>(c, nx, ny, nz, geom_r, done, s);
    return closing_fused_r<R, 2, 2>(c, nx, ny, nz, geom_r, done, s);
}

// Opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute: set for every tile kernel
// whenever a context is created on a device (mamri_create, under its device guard, outside any stream capture).
template <int R>
static cudaError_t tile_attrs_r() {
    cudaError_t e = cudaFuncSetAttribute(k_morph_tile<R, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 16);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_morph_tile<R, true, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 16);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_close_fused<R, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(FUSED_SMEM_MAX));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_close_fused<R, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(FUSED_SMEM_MAX));
    if constexpr (R % 2 == 0)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_close_fused<R, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(FUSED_SMEM_MAX));
    return e;
}

cudaError_t segment_init_device() {
    cudaError_t e = tile_attrs_r<1>();
    if (e == cudaSuccess) e = tile_attrs_r<2>();
    if (e == cudaSuccess) e = tile_attrs_r<3>();
    return e;
}

template <int R, int SY>
static cudaError_t closing_dispatch(mamri_ctx* c, int nx, int ny, int nz, int geom_r, cudaStream_t s) {
    static const int use_planes = [] { const char* e = getenv("MAMRI_MORPH_PLANES"); return e ? atoi(e) : 0; }();
    static const int use_fused = [] { const char* e = getenv("MAMRI_CLOSE_FUSED"); return e ? atoi(e) : 1; }();
    // MAMRI_CLOSE_FUSED: 1 (default) = the fused kernel for volumes of up to 2^27 voxels -- a clinical scan, mostly air,
    // where the launch it saves and the occupancy skip count -- and the two-pass kernels above that (a noisy high-
    // resolution volume has foreground in every tile: the fused tile's halo of 2R is then redundant work, measured
    // 159 us against 116 us on config C4); 2 = always fused, 0 = never.
    const bool small = (unsigned long long)nx * ny * nz <= (1ull << 27);
    if (((use_fused == 2 || (use_fused == 1 && small)) && !use_planes) || geom_r != R) {
        bool done = false;
        const cudaError_t e = closing_fused_dispatch<R>(c, nx, ny, nz, geom_r, done, s);
        if (e != cudaSuccess || done) return e;
        if (geom_r != R) return cudaErrorInvalidValue;     // the two-pass kernels assume an apron of exactly 2R
    }
    uint32_t TY, TZ, smem;
    if (!use_planes && tile_plan<R, SY>(PadGeom(nx, ny, nz, R).Wp, TY, TZ, smem)) {
        if (const char* e = getenv("MAMRI_TILE_TZ")) {              // experiments only
            TZ = uint32_t(atoi(e));
            smem = PadGeom(nx, ny, nz, R).Wp * (TY + 2 * R) * (TZ + 2 * R) * 4u + 16u;
        }
        return closing_tile_r<R, SY>(c, nx, ny, nz, TY, TZ, smem, s);
    }
    return closing_r<R>(c, nx, ny, nz, s);
}

// Radius the padded layout is built for: the closing needs an apron of 2 Rc, an opening before it one of Ro.
int morph_geom_radius(int open_radius, int close_radius) {
    const int ro = (open_radius + 1) / 2;
    return close_radius > ro ? close_radius : ro;
}

// Copies the interior of the padded raw mask into the plain [nz][ny][W] mask (opening without a closing after it).
__global__ void __launch_bounds__(256) k_unpad(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t Wp, uint32_t slice,
                                               uint32_t W, uint32_t ny, uint32_t n_words, uint32_t pad) {
    pdl_wait();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) {
        const uint32_t row = i / W, xw = i - row * W;
        const uint32_t z = row / ny, y = row - z * ny;
        dst[i] = src[(z + pad) * slice + (y + pad) * Wp + 1u + xw];
    }
}

// `geom_r` = radius the padded layout was built for (morph_geom_radius); the mask ends up in c->d_closed.
cudaError_t launch_closing(mamri_ctx* c, int nx, int ny, int nz, int radius, int geom_r, cudaStream_t s) {
    switch (radius) {
        case 0: {
            if (geom_r == 0) return cudaSuccess;                     // threshold wrote the plain mask itself
            const PadGeom g(nx, ny, nz, geom_r);
            const uint32_t n_words = g.W * uint32_t(ny) * uint32_t(nz);
            uint32_t blocks = (n_words + 255) / 256;
            if (blocks > 148 * 8) blocks = 148 * 8;
            LK(k_unpad, blocks, 256, s, false, c->d_raw, c->d_closed, g.Wp, g.slice, g.W, uint32_t(ny), n_words, 2u * uint32_t(geom_r));
            prof_mark(c, s, "unpad");
            return cudaGetLastError();
        }
        case 1: return closing_dispatch<1, 4>(c, nx, ny, nz, geom_r, s);
        case 2: return closing_dispatch<2, 4>(c, nx, ny, nz, geom_r, s);
        case 3: return closing_dispatch<3, 4>(c, nx, ny, nz, geom_r, s);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// opening (north_star "open/close"; the reference only closes): erosion with the outside of the image counted as
// foreground, then dilation -- itk::BinaryMorphologicalOpeningImageFilter
// ------------------------------------------------------------------------------------------------
// With X' the complement of X inside the image (and 0 outside), erode(X; outside = 1) = complement of dilate(X'), so
// the opening is two DILATIONS on the image domain of the padded layout (zero apron = the outside of both):
//   raw <- ~raw (image only);  open <- ~dilate(raw) (image only) = the erosion;  raw <- dilate(open).
// `d_open` is a buffer whose apron is zero and is never written.  The occupancy cells of the raw mask stay valid for
// the closing that follows (an opening only removes voxels).
__global__ void __launch_bounds__(256) k_invert_image(uint32_t* __restrict__ buf, uint32_t Wp, uint32_t slice, uint32_t W, uint32_t ny,
                                                      uint32_t n_words, uint32_t pad, uint32_t tail_mask) {
    pdl_wait();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) {
        const uint32_t row = i / W, xw = i - row * W;
        const uint32_t z = row / ny, y = row - z * ny;
        uint32_t* p = buf + (z + pad) * slice + (y + pad) * Wp + 1u + xw;
        *p = ~*p & (xw == W - 1 ? tail_mask : 0xFFFFFFFFu);
    }
}

template <int R>
static cudaError_t opening_r(mamri_ctx* c, int nx, int ny, int nz, int geom_r, cudaStream_t s) {
    const PadGeom g(nx, ny, nz, geom_r);
    const uint32_t pad = 2u * uint32_t(geom_r);
    uint32_t TY, TZ, smem;
    if (!tile_plan<R, 4>(g.Wp, TY, TZ, smem)) return cudaErrorInvalidValue;     // rows of more than ~8000 voxels
    const uint32_t tail = (nx % 32) ? (0xFFFFFFFFu >> (32 - nx % 32)) : 0xFFFFFFFFu;
    const uint32_t n_words = g.W * uint32_t(ny) * uint32_t(nz);
    uint32_t blocks = (n_words + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    LK(k_invert_image, blocks, 256, s, false, c->d_raw, g.Wp, g.slice, g.W, uint32_t(ny), n_words, pad, tail);
    prof_mark(c, s, "open_invert");
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.Wp = g.Wp; a.slice = g.slice; a.TY = TY; a.TZ = TZ;
    a.x_lo = 1; a.x_cnt = g.W; a.y_lo = pad; a.y_cnt = uint32_t(ny); a.z_lo = pad; a.z_hi = pad + uint32_t(nz);
    a.tail_mask = tail;
    const uint32_t threads = ((a.x_cnt * (TY / 4) + 31) / 32) * 32;
    const dim3 grid((a.y_cnt + TY - 1) / TY, (a.z_hi - a.z_lo + TZ - 1) / TZ);
    a.invert_out = 1;
    LKS(k_morph_tile<R, false, false, 4>, grid, threads, smem, s, false, c->d_raw, c->d_open, c->d_dyn, a);
    prof_mark(c, s, "open_erode");
    a.invert_out = 0;
    LKS(k_morph_tile<R, false, false, 4>, grid, threads, smem, s, false, c->d_open, c->d_raw, c->d_dyn, a);
    prof_mark(c, s, "open_dilate");
    return cudaGetLastError();
}

cudaError_t launch_opening(mamri_ctx* c, int nx, int ny, int nz, int radius, int geom_r, cudaStream_t s) {
    switch (radius) {
        case 0: return cudaSuccess;
        case 1: return opening_r<1>(c, nx, ny, nz, geom_r, s);
        case 2: return opening_r<2>(c, nx, ny, nz, geom_r, s);
        case 3: return opening_r<3>(c, nx, ny, nz, geom_r, s);
        default: return cudaErrorInvalidValue;
    }
}

KTRACE_TU(segment)
