// Skin-surface candidates of the body segment: the step between the body labelmap of
// volume_threshold_segmentation (Mamri/Mamri.py:1322-1323) and the candidate loop of
// findAndSetEntryPoint (:1008-1023).  The reference gets its candidates from Slicer's closed-surface
// representation of "AutoBodySegmentation" (:1338-1339, _get_body_polydata :994) plus
// vtkPolyDataNormals (:997-1003); neither is reproducible outside Slicer, so this stage defines the
// candidate set on the voxel grid instead (validated geometrically, see tests):
//
//   candidate  = body voxel with at least one of its 6 face neighbours outside the body (outside the
//                volume counts as outside), in ascending linear index -> deterministic point ids;
//   point      = physical centre of the voxel, LPS -> RAS (float32, as vtkPoints stores them);
//   normal     = outward unit normal from the first moment of the body occupancy inside the ITK ball of
//                radius 2 around the voxel (81 voxels, d.d <= 6): g = sum d * body(p + d) points into the
//                body, n_index = -g, n_physical = Direction * (n_index / spacing), LPS -> RAS, normalised.
//
// Everything works on a 1 bit/voxel copy of the body: either packed from the caller's uint8 labelmap
// (one streaming read, 1 B/voxel) or -- d_body_mask == NULL -- taken straight from the last scan's closed
// mask and run labels (no per-voxel read at all).  One single-pass kernel then finds the surface bits of a
// tile of words, gets its output offset by decoupled look-back and emits points + normals, the surface
// voxels of every 32-word round spread evenly over the warp's lanes.
#include "common.cuh"

constexpr int SF_THREADS = 256;
constexpr int SF_TILE = 8192;          // words per tile; mamri_create sizes d_scan_runs for tiles of 8192 words

struct SurfArgs {
    int nx, ny, nz, W;
    double sp[3], org[3], dir[9];
};

// ---- 1 bit/voxel body ---------------------------------------------------------------------------------
// nx % 32 == 0 and 16-byte aligned: 16 voxels per thread (one 128-bit load), two lanes make a word.
__global__ void __launch_bounds__(256) k_body_bits_vec(const uint4* __restrict__ body, uint32_t* __restrict__ bits,
                                                       uint32_t n_words) {
    const uint32_t n_half = n_words * 2u;
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < ((n_half + 31u) & ~31u); h += gridDim.x * blockDim.x) {
        uint32_t half = 0;
        if (h < n_half) {
            uint4 q;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(body + h));
            const uint32_t v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = __vcmpne4(v[k], 0u) & 0x01010101u;        // bytes -> 0/1
                half |= ((t * 0x01020408u) >> 24 & 0xFu) << (4 * k);         // gather the four low bits
            }
        }
        const uint32_t other = __shfl_down_sync(FULL, half, 1);
        if (!(threadIdx.x & 1u) && h < n_half) bits[h >> 1] = half | (other << 16);
    }
}

// any row length / alignment: one warp per word, lane = voxel.
__global__ void __launch_bounds__(256) k_body_bits_any(const uint8_t* __restrict__ body, uint32_t* __restrict__ bits, int nx,
                                                       int W, uint32_t n_words) {
    const unsigned lane = lane_id();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w0 = warp * 4u; w0 < n_words; w0 += n_warps * 4u) {
        uint8_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                       // four loads in flight per lane
            const uint32_t wi = w0 + k;
            const uint32_t row = wi / uint32_t(W);
            const int x = int(wi - row * uint32_t(W)) * 32 + int(lane);
            v[k] = (wi < n_words && x < nx) ? body[size_t(row) * nx + x] : uint8_t(0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t m = __ballot_sync(FULL, v[k] != 0);
            if (lane == 0 && w0 + k < n_words) bits[w0 + k] = m;
        }
    }
}

// body of the last scan: runs of the closed mask whose label is the body label.
__global__ void __launch_bounds__(256) k_body_bits_runs(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ word_base,
                                                        const uint32_t* __restrict__ run_label, int W, uint32_t n_words,
                                                        uint32_t body, uint32_t* __restrict__ bits) {
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
        const uint32_t m = mask[wi];
        uint32_t out = 0;
        if (m && body) {
            const uint32_t prev = (wi % uint32_t(W)) ? mask[wi - 1] : 0u;
            const uint32_t starts = run_starts(m, prev), base = word_base[wi];
            uint32_t rem = m;
            while (rem) {
                const int b = __ffs(rem) - 1;
                const uint32_t t = m >> b;
                const int len = (t == (0xFFFFFFFFu >> b)) ? 32 - b : __ffs(~t) - 1;
                const uint32_t seg = (len == 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << b;
                if (run_label[run_id_in_word(base, starts, b)] == body) out |= seg;
                rem &= ~seg;
            }
        }
        bits[wi] = out;
    }
}

// ---- surface + normals -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t surface_word(const uint32_t* __restrict__ bits, uint32_t wi, const SurfArgs& a) {
    const uint32_t b = bits[wi];
    if (!b) return 0u;
    const uint32_t W = uint32_t(a.W), row = wi / W, xw = wi - row * W;
    const uint32_t y = row % uint32_t(a.ny), z = row / uint32_t(a.ny);
    const uint32_t slice = W * uint32_t(a.ny);
    const uint32_t xl = (b << 1) | (xw > 0 ? bits[wi - 1] >> 31 : 0u);
    const uint32_t xr = (b >> 1) | (xw + 1 < W ? bits[wi + 1] << 31 : 0u);      // bits beyond nx are never set
    const uint32_t yu = y > 0 ? bits[wi - W] : 0u, yd = y + 1 < uint32_t(a.ny) ? bits[wi + W] : 0u;
    const uint32_t zu = z > 0 ? bits[wi - slice] : 0u, zd = z + 1 < uint32_t(a.nz) ? bits[wi + slice] : 0u;
    return b & ~(xl & xr & yu & yd & zu & zd);
}

// First moment of the body inside the radius-2 ITK ball around (x,y,z): 21 rows, 5 or 3 voxels each.
__device__ __forceinline__ void ball_moment(const uint32_t* __restrict__ bits, int x, int y, int z, const SurfArgs& a,
                                            int& gx, int& gy, int& gz) {
    gx = gy = gz = 0;
    const int xw = x >> 5, bit = x & 31;
#pragma unroll
    for (int dz = -2; dz <= 2; ++dz) {
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            const int rem = 6 - dz * dz - dy * dy;          // dx*dx <= rem
            if (rem < 0) continue;
            const uint32_t keep = rem >= 4 ? 0x1Fu : 0x0Eu; // |dx| <= 2 or |dx| <= 1 (bit 2 = dx 0)
            const int yy = y + dy, zz = z + dz;
            if (yy < 0 || yy >= a.ny || zz < 0 || zz >= a.nz) continue;
            const uint32_t* r = bits + (size_t(zz) * a.ny + yy) * a.W;
            const uint32_t lo = __ldg(r + xw);
            uint32_t t;                                     // bits x-2 .. x+2 of the row -> t[0..4]
            if (bit >= 2) {
                const uint32_t hi = (bit > 29 && xw + 1 < a.W) ? __ldg(r + xw + 1) : 0u;
                t = __funnelshift_r(lo, hi, bit - 2);
            } else {
                const uint32_t pv = xw > 0 ? __ldg(r + xw - 1) : 0u;
                t = __funnelshift_r(pv, lo, 30 + bit);
            }
            t &= keep;
            const int c = __popc(t);
            gx += int((t >> 3) & 1u) - int((t >> 1) & 1u) + 2 * (int((t >> 4) & 1u) - int(t & 1u));
            gy += dy * c;
            gz += dz * c;
        }
    }
}

__device__ __forceinline__ void emit_candidate(const uint32_t* __restrict__ bits, const SurfArgs& a, int x, int y, int z,
                                               float* __restrict__ pts, float* __restrict__ nrm, size_t pos) {
    // physical centre, in the oracle's operation order with separately rounded operations (no FMA contraction)
    const double ix = __dmul_rn(a.sp[0], double(x)), iy = __dmul_rn(a.sp[1], double(y)), iz = __dmul_rn(a.sp[2], double(z));
    double lps[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        lps[k] = __dadd_rn(a.org[k], __dadd_rn(__dadd_rn(__dmul_rn(a.dir[3 * k], ix), __dmul_rn(a.dir[3 * k + 1], iy)),
                                               __dmul_rn(a.dir[3 * k + 2], iz)));
    pts[3 * pos] = float(-lps[0]); pts[3 * pos + 1] = float(-lps[1]); pts[3 * pos + 2] = float(lps[2]);
    int gx, gy, gz;
    ball_moment(bits, x, y, z, a, gx, gy, gz);
    const double qx = __ddiv_rn(double(-gx), a.sp[0]), qy = __ddiv_rn(double(-gy), a.sp[1]), qz = __ddiv_rn(double(-gz), a.sp[2]);
    double v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        v[k] = __dadd_rn(__dadd_rn(__dmul_rn(a.dir[3 * k], qx), __dmul_rn(a.dir[3 * k + 1], qy)), __dmul_rn(a.dir[3 * k + 2], qz));
    const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2])));
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (len > 0.0) { n0 = float(-__ddiv_rn(v[0], len)); n1 = float(-__ddiv_rn(v[1], len)); n2 = float(__ddiv_rn(v[2], len)); }
    nrm[3 * pos] = n0; nrm[3 * pos + 1] = n1; nrm[3 * pos + 2] = n2;
}

__global__ void __launch_bounds__(SF_THREADS) k_surface(const uint32_t* __restrict__ bits, SurfArgs a, uint32_t n_words,
                                                        volatile unsigned long long* state, uint32_t gen,
                                                        float* __restrict__ pts, float* __restrict__ nrm,
                                                        unsigned long long capacity, SurfScalars* sc) {
    __shared__ uint32_t s_surf[SF_TILE];
    __shared__ uint32_t s_warp[SF_THREADS / 32];
    __shared__ uint32_t s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(&sc->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, n_tiles = gridDim.x;
    const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t warp_w0 = tile * SF_TILE + wid * (SF_TILE / (SF_THREADS / 32));
    constexpr int ROUNDS = SF_TILE / SF_THREADS;              // 32 rounds of 32 words per warp
    uint32_t cnt = 0, body = 0;
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t wi = warp_w0 + r * 32 + lane;
        uint32_t s = 0;
        if (wi < n_words) { s = surface_word(bits, wi, a); body += __popc(bits[wi]); }
        s_surf[wid * (ROUNDS * 32) + r * 32 + lane] = s;
        cnt += __popc(s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(FULL, cnt, o); body += __shfl_xor_sync(FULL, body, o); }
    if (lane == 0) { s_warp[wid] = cnt; if (body) atomicAdd(&sc->n_body, (unsigned long long)body); }
    __syncthreads();
    uint32_t total = 0, before_warp = 0;
#pragma unroll
    for (int w = 0; w < SF_THREADS / 32; ++w) { if (w < int(wid)) before_warp += s_warp[w]; total += s_warp[w]; }
    if (threadIdx.x < 32) {
        const uint32_t before = scan_lookback(state, tile, total, gen);
        if (threadIdx.x == 0) {
            s_prefix = before;
            if (tile == n_tiles - 1) sc->n_points = (unsigned long long)before + total;
        }
    }
    __syncthreads();
    if (cnt == 0 || !pts) return;                             // (count query: nothing to emit)
    size_t base = size_t(s_prefix) + before_warp;
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t sw = s_surf[wid * (ROUNDS * 32) + r * 32 + lane];
        if (!__any_sync(FULL, sw != 0u)) continue;
        const uint32_t c = __popc(sw);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= unsigned(o)) incl += t; }
        const uint32_t excl = incl - c, n_r = __shfl_sync(FULL, incl, 31);
        for (uint32_t k0 = 0; k0 < n_r; k0 += 32) {           // item k of the round -> (word j, its (k - excl_j)-th set bit)
            const uint32_t k = k0 + lane;
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int cand = j + step;
                const uint32_t e = __shfl_sync(FULL, excl, cand & 31);
                if (cand < 32 && e <= k) j = cand;
            }
            const uint32_t wj = __shfl_sync(FULL, sw, j), ej = __shfl_sync(FULL, excl, j);
            if (k < n_r && base + k < capacity) {
                const int bit = int(__fns(wj, 0u, int(k - ej) + 1));
                const uint32_t wi = warp_w0 + r * 32 + uint32_t(j);
                const uint32_t row = wi / uint32_t(a.W), xw = wi - row * uint32_t(a.W);
                emit_candidate(bits, a, int(xw) * 32 + bit, int(row % uint32_t(a.ny)), int(row / uint32_t(a.ny)), pts, nrm, base + k);
            }
        }
        base += n_r;
    }
}

cudaError_t launch_body_surface(mamri_ctx* c, const mamri_volume_desc* desc, const uint8_t* d_body_mask, uint32_t body_label,
                                float* d_points, float* d_normals, unsigned long long capacity, cudaStream_t s) {
    const int nx = desc->nx, ny = desc->ny, nz = desc->nz, W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * uint32_t(ny) * uint32_t(nz);
    uint32_t* bits = c->d_planes;                             // free between scans: >= cap_words words
    cudaError_t e = cudaMemsetAsync(c->d_surf, 0, sizeof(SurfScalars), s);
    if (e != cudaSuccess) return e;
    if (d_body_mask) {
        if (nx % 32 == 0 && (reinterpret_cast<uintptr_t>(d_body_mask) & 15u) == 0) {
            uint32_t blocks = (2u * n_words + 255u) / 256u;
            if (blocks > 148u * 16u) blocks = 148u * 16u;
            k_body_bits_vec<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(d_body_mask), bits, n_words);
        } else {
            uint32_t blocks = (n_words + 31u) / 32u;
            if (blocks > 148u * 16u) blocks = 148u * 16u;
            k_body_bits_any<<<blocks, 256, 0, s>>>(d_body_mask, bits, nx, W, n_words);
        }
    } else {
        uint32_t blocks = (n_words + 255u) / 256u;
        if (blocks > 148u * 16u) blocks = 148u * 16u;
        k_body_bits_runs<<<blocks, 256, 0, s>>>(c->d_closed, c->d_word_base, c->d_run_label, W, n_words, body_label, bits);
    }
    SurfArgs a;
    a.nx = nx; a.ny = ny; a.nz = nz; a.W = W;
    for (int i = 0; i < 3; ++i) { a.sp[i] = desc->spacing[i]; a.org[i] = desc->origin[i]; }
    for (int i = 0; i < 9; ++i) a.dir[i] = desc->direction[i];
    uint32_t gen = ++c->gen;                                  // new generation of the look-back states
    if ((gen & 0x3FFFFFFFu) == 0) gen = ++c->gen;
    const uint32_t tiles = (n_words + SF_TILE - 1) / SF_TILE;
    k_surface<<<tiles, SF_THREADS, 0, s>>>(bits, a, n_words, c->d_scan_runs, gen, d_points, d_normals, capacity, c->d_surf);
    return cudaGetLastError();
}
