// Stage 2: 3-D connected-component labelling of the bit-packed closed mask, 6- or 26-connected,
// with ITK's consecutive numbering.  Replaces sitk.ConnectedComponent at Mamri/Mamri.py:1309
// (itk::ConnectedComponentImageFilter: provisional labels on x-runs, smaller label wins a union,
// final labels consecutive in increasing root order = rank of each component's minimum linear index).
//
// Run-based: the union-find nodes are the x-runs of the mask, numbered in raster order by an
// exclusive scan of per-word run-start counts.  Run ids are monotone in the linear index of the
// run's first voxel, so hooking the larger root under the smaller with atomicMin makes the root of
// every component its first run in raster order, independent of scheduling -> deterministic labels.
// The parent array is 4 B per RUN (not per voxel) and stays L2-resident; the only per-voxel traffic
// of the stage is the final label write (materialise.cu side of this file).
#include "common.cuh"

constexpr int SCAN_THREADS = 1024;

// ------------------------------------------------------------------------------------------------
// block-wide primitives (blockDim.x == SCAN_THREADS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= unsigned(o)) v += t;
    }
    return v;
}

// Exclusive prefix of `v` over the block; `total` = block sum.  `ws` is 33 words of shared memory.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* ws, uint32_t& total) {
    const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = ws[lane];
        uint32_t ti = warp_incl_scan(t);
        ws[lane] = ti - t;
        if (lane == 31) ws[32] = ti;
    }
    __syncthreads();
    uint32_t r = ws[wid] + inc - v;
    total = ws[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* ws) {
    const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) ws[wid] = v;
    __syncthreads();
    uint32_t t = 0;
    if (wid == 0) {
        t = ws[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    }
    __syncthreads();
    return t;   // valid in warp 0
}

__device__ __forceinline__ void chunk_of(uint32_t n, uint32_t& begin, uint32_t& end) {
    uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
    begin = uint32_t(blockIdx.x) * chunk;
    end = begin + chunk < n ? begin + chunk : n;
    if (begin > n) begin = n;
}

// ------------------------------------------------------------------------------------------------
// run numbering: exclusive scan of run starts per word (3 small launches)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_runs_count(const uint32_t* __restrict__ mask, int W, uint32_t n_words,
                                                             uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t ws[33];
    uint32_t begin, end;
    chunk_of(n_words, begin, end);
    uint32_t sum = 0;
    for (uint32_t i = begin + threadIdx.x; i < end; i += SCAN_THREADS) {
        uint32_t m = mask[i];
        if (m) {
            uint32_t prev = (i % W) ? mask[i - 1] : 0u;
            sum += __popc(run_starts(m, prev));
        }
    }
    uint32_t t = block_sum(sum, ws);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

// One CTA: exclusive scan of the per-CTA partials in place; total -> *out_total (capacity-checked).
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partials(uint32_t* __restrict__ block_sums, int n_blocks,
                                                                unsigned int* out_total, unsigned int limit,
                                                                DevScalars* sc) {
    __shared__ uint32_t ws[33];
    uint32_t v = int(threadIdx.x) < n_blocks ? block_sums[threadIdx.x] : 0u;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, ws, total);
    if (int(threadIdx.x) < n_blocks) block_sums[threadIdx.x] = ex;
    if (threadIdx.x == 0) {
        if (total > limit) { sc->status = MAMRI_ERR_CAPACITY; total = 0; }
        *out_total = total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_runs_assign(const uint32_t* __restrict__ mask, int W, uint32_t n_words,
                                                              const uint32_t* __restrict__ block_sums,
                                                              uint32_t* __restrict__ word_base, const DevScalars* sc) {
    __shared__ uint32_t ws[33];
    if (sc->status != MAMRI_OK) return;
    uint32_t begin, end;
    chunk_of(n_words, begin, end);
    uint32_t running = block_sums[blockIdx.x];
    for (uint32_t i0 = begin; i0 < end; i0 += SCAN_THREADS) {
        uint32_t i = i0 + threadIdx.x;
        uint32_t starts = 0;
        if (i < end) {
            uint32_t m = mask[i];
            if (m) starts = run_starts(m, (i % W) ? mask[i - 1] : 0u);
        }
        uint32_t cnt = __popc(starts), total;
        uint32_t base = running + block_excl_scan(cnt, ws, total);
        if (i < end) {
            word_base[i] = base;
        }
        running += total;
    }
}

// ------------------------------------------------------------------------------------------------
// union-find over runs
// ------------------------------------------------------------------------------------------------
// `parent` may live in shared memory (per-slice phase) or in global memory (cross-slice phase);
// volatile generic loads keep every hop coherent (L2 for global) while other threads hook roots.
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
    volatile uint32_t* vp = parent;
    uint32_t p = vp[x];
    while (p != x) {
        uint32_t gp = vp[p];
        if (gp != p) vp[x] = gp;       // path halving; any smaller same-set node is a valid parent
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }       // hook the larger root under the smaller
        uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;                               // a was still a root: done
        a = old;                                            // someone re-parented a meanwhile: keep merging
    }
}

// Joins the runs of word (xw,y,z) with the runs of one earlier neighbour row.  DIAG adds the
// x+-1 contacts (26-connectivity).  Every maximal piece of overlapping bits lies in exactly one run
// on either side, so one union per piece start suffices; pieces continuing from the previous word
// are skipped (their start was handled there).  Node ids are run ids minus `id_off`.
template <bool DIAG>
__device__ __forceinline__ void join_row(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ word_base,
                                         uint32_t* parent, uint32_t id_off, int W, int xw, uint32_t m, uint32_t m_prev,
                                         uint32_t base_m, uint32_t starts_m, uint32_t ni) {
    const uint32_t up = mask[ni];
    const uint32_t up_prev = xw > 0 ? mask[ni - 1] : 0u;
    const uint32_t up_next = (DIAG && xw + 1 < W) ? mask[ni + 1] : 0u;
    if (!DIAG && !up) return;
    if (DIAG && !(up | (up_prev >> 31) | (up_next & 1u))) return;
    const uint32_t starts_up = run_starts(up, up_prev);
    const uint32_t base_up = word_base[ni] - id_off;
    base_m -= id_off;
    // direct (same x) contacts
    uint32_t ov = m & up;
    uint32_t ps = ov & ~((ov << 1) | ((m_prev & up_prev) >> 31));
    while (ps) {
        int b = __ffs(ps) - 1;
        ps &= ps - 1;
        uf_union(parent, run_id_in_word(base_m, starts_m, b), run_id_in_word(base_up, starts_up, b));
    }
    if (DIAG) {
        // bit b of m touching up(b-1) where up(b) is clear: the run of `up` ending at b-1
        // (skipped when m(b-1) is set too: the direct contact at b-1 already joined the same two runs)
        uint32_t a = m & ~up & ((up << 1) | (up_prev >> 31)) & ~((m << 1) | (m_prev >> 31));
        while (a) {
            int b = __ffs(a) - 1;
            a &= a - 1;
            uint32_t rid_up = b > 0 ? run_id_in_word(base_up, starts_up, b - 1) : base_up - 1u;
            uf_union(parent, run_id_in_word(base_m, starts_m, b), rid_up);
        }
        // bit b of m touching up(b+1) where up(b) is clear: the run of `up` starting at b+1
        uint32_t c = m & ~up & ((up >> 1) | (up_next << 31)) & ~(m >> 1);
        while (c) {
            int b = __ffs(c) - 1;
            c &= c - 1;
            uint32_t rid_up = b < 31 ? run_id_in_word(base_up, starts_up, b + 1) : word_base[ni + 1] - id_off;
            uf_union(parent, run_id_in_word(base_m, starts_m, b), rid_up);
        }
    }
}

// Phase 1 -- block-local: one CTA per z-slice.  The runs of a slice are contiguous in id space, so the
// CTA keeps their parents in shared memory (local index = run id - first run of the slice), joins
// every row with the row above it at shared-memory latency (simultaneous hooking builds chains as
// long as the object is tall; in shared memory a hop costs ~30 cycles instead of an L2 round trip),
// flattens, and publishes parent[run] = slice-local root as a global id.  Slices with more runs than
// fit fall back to the same code on the global array.
constexpr int SLICE_THREADS = 512;
constexpr uint32_t SLICE_SMEM_RUNS = 12000;      // 48 KB static shared memory

template <bool CONN26>
__global__ void __launch_bounds__(SLICE_THREADS) k_union_slices(const uint32_t* __restrict__ mask,
                                                               const uint32_t* __restrict__ word_base, uint32_t* parent,
                                                               int W, int ny, int nz, const DevScalars* sc) {
    __shared__ uint32_t sp[SLICE_SMEM_RUNS];
    if (sc->status != MAMRI_OK) return;
    const int z = blockIdx.x;
    const uint32_t slice_words = uint32_t(W) * ny;
    const uint32_t w0 = uint32_t(z) * slice_words;
    const uint32_t r0 = word_base[w0];
    const uint32_t r1 = (z + 1 < nz) ? word_base[w0 + slice_words] : sc->n_runs;
    const uint32_t n = r1 - r0;
    if (n == 0) return;
    uint32_t* P = (n <= SLICE_SMEM_RUNS) ? sp : parent + r0;
    for (uint32_t i = threadIdx.x; i < n; i += SLICE_THREADS) P[i] = i;
    __syncthreads();
    for (uint32_t i = uint32_t(W) + threadIdx.x; i < slice_words; i += SLICE_THREADS) {     // rows y >= 1
        const uint32_t wi = w0 + i;
        const uint32_t m = mask[wi];
        if (!m) continue;
        const int xw = int(i % W);
        const uint32_t m_prev = xw > 0 ? mask[wi - 1] : 0u;
        join_row<CONN26>(mask, word_base, P, r0, W, xw, m, m_prev, word_base[wi], run_starts(m, m_prev), wi - W);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += SLICE_THREADS) {
        uint32_t x = i, p = P[x];
        while (p != x) { x = p; p = P[x]; }
        P[i] = x;                                 // roots stay fixed points: concurrent walkers remain correct
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += SLICE_THREADS) parent[r0 + i] = r0 + P[i];
}

// Phase 2 -- global boundary merge between slices, with atomicMin on the roots.  Joining all slice
// boundaries at once would again chain the roots nz deep (every hop an L2 round trip), so the
// boundaries are merged in two rounds: first those inside blocks of `radix` slices (chains <= radix),
// then the boundaries between blocks (again <= radix of them per chain).
template <bool CONN26>
__global__ void __launch_bounds__(256) k_union_z(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ word_base,
                                                 uint32_t* parent, int W, int ny, int nz, uint32_t n_words, int radix,
                                                 int between_blocks, const DevScalars* sc) {
    if (sc->status != MAMRI_OK) return;
    const uint32_t slice = uint32_t(W) * ny;
    for (uint32_t wi = slice + uint32_t(blockIdx.x) * blockDim.x + threadIdx.x; wi < n_words;
         wi += uint32_t(gridDim.x) * blockDim.x) {
        const int z = int(wi / slice);
        if (((z % radix) == 0) != (between_blocks != 0)) continue;
        const uint32_t m = mask[wi];
        if (!m) continue;
        const uint32_t row = wi / W;
        const int xw = int(wi - row * W);
        const int y = int(row % ny);
        const uint32_t m_prev = xw > 0 ? mask[wi - 1] : 0u;
        const uint32_t starts_m = run_starts(m, m_prev);
        const uint32_t base_m = word_base[wi];
        join_row<CONN26>(mask, word_base, parent, 0u, W, xw, m, m_prev, base_m, starts_m, wi - slice);
        if (CONN26) {
            if (y > 0) join_row<true>(mask, word_base, parent, 0u, W, xw, m, m_prev, base_m, starts_m, wi - slice - W);
            if (y + 1 < ny) join_row<true>(mask, word_base, parent, 0u, W, xw, m, m_prev, base_m, starts_m, wi - slice + W);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// flatten, rank the roots (ITK-consecutive labels), propagate to every run
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_flatten_count(uint32_t* parent, uint32_t* __restrict__ block_sums,
                                                                const DevScalars* sc) {
    __shared__ uint32_t ws[33];
    uint32_t begin, end;
    chunk_of(sc->n_runs, begin, end);
    uint32_t roots = 0;
    for (uint32_t r = begin + threadIdx.x; r < end; r += SCAN_THREADS) {
        uint32_t x = uint32_t(r), p = parent[x];
        while (p != x) { x = p; p = parent[x]; }
        parent[r] = x;
        roots += (x == uint32_t(r));
    }
    uint32_t t = block_sum(roots, ws);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_rank_roots(const uint32_t* __restrict__ parent,
                                                             const uint32_t* __restrict__ block_sums,
                                                             uint32_t* __restrict__ run_label,
                                                             uint32_t* __restrict__ label_count, const DevScalars* sc) {
    __shared__ uint32_t ws[33];
    uint32_t begin, end;
    chunk_of(sc->n_runs, begin, end);
    uint32_t running = block_sums[blockIdx.x];
    for (uint32_t i0 = begin; i0 < end; i0 += SCAN_THREADS) {
        uint32_t r = i0 + threadIdx.x;
        uint32_t is_root = (r < end && parent[r] == uint32_t(r)) ? 1u : 0u, total;
        uint32_t rank = running + block_excl_scan(is_root, ws, total);
        if (is_root) {
            run_label[r] = rank + 1u;
            label_count[rank] = 0u;
        }
        running += total;
    }
}

__global__ void __launch_bounds__(256) k_propagate_labels(const uint32_t* __restrict__ parent, uint32_t* run_label,
                                                          const DevScalars* sc) {
    const uint32_t n = sc->n_runs;
    for (uint32_t r = uint32_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += uint32_t(gridDim.x) * blockDim.x) {
        uint32_t p = parent[r];
        if (p != uint32_t(r)) run_label[r] = run_label[p];   // roots were written by k_rank_roots
    }
}

cudaError_t launch_ccl(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int connectivity, cudaStream_t s) {
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    const int G = MAMRI_SCAN_CTAS;
    uint32_t* bs_runs = c->d_block_sums;
    uint32_t* bs_roots = c->d_block_sums + 1024;
    k_runs_count<<<G, SCAN_THREADS, 0, s>>>(d_mask, W, n_words, bs_runs);
    k_scan_partials<<<1, SCAN_THREADS, 0, s>>>(bs_runs, G, &c->d_scalars->n_runs, c->max_runs, c->d_scalars);
    k_runs_assign<<<G, SCAN_THREADS, 0, s>>>(d_mask, W, n_words, bs_runs, c->d_word_base, c->d_scalars);
    int radix = 1;
    while (radix * radix < nz) radix <<= 1;
    uint32_t ub = (n_words + 255) / 256;
    if (ub > 148 * 16) ub = 148 * 16;
    if (ub == 0) ub = 1;
    if (connectivity == 26) {
        k_union_slices<true><<<nz, SLICE_THREADS, 0, s>>>(d_mask, c->d_word_base, c->d_parent, W, ny, nz, c->d_scalars);
        if (nz > 1) k_union_z<true><<<unsigned(ub), 256, 0, s>>>(d_mask, c->d_word_base, c->d_parent, W, ny, nz, n_words, radix, 0, c->d_scalars);
        if (nz > radix) k_union_z<true><<<unsigned(ub), 256, 0, s>>>(d_mask, c->d_word_base, c->d_parent, W, ny, nz, n_words, radix, 1, c->d_scalars);
    } else {
        k_union_slices<false><<<nz, SLICE_THREADS, 0, s>>>(d_mask, c->d_word_base, c->d_parent, W, ny, nz, c->d_scalars);
        if (nz > 1) k_union_z<false><<<unsigned(ub), 256, 0, s>>>(d_mask, c->d_word_base, c->d_parent, W, ny, nz, n_words, radix, 0, c->d_scalars);
        if (nz > radix) k_union_z<false><<<unsigned(ub), 256, 0, s>>>(d_mask, c->d_word_base, c->d_parent, W, ny, nz, n_words, radix, 1, c->d_scalars);
    }
    k_flatten_count<<<G, SCAN_THREADS, 0, s>>>(c->d_parent, bs_roots, c->d_scalars);
    k_scan_partials<<<1, SCAN_THREADS, 0, s>>>(bs_roots, G, &c->d_scalars->n_labels, 0xFFFFFFFFu, c->d_scalars);
    k_rank_roots<<<G, SCAN_THREADS, 0, s>>>(c->d_parent, bs_roots, c->d_run_label, c->d_label_count, c->d_scalars);
    k_propagate_labels<<<148 * 4, 256, 0, s>>>(c->d_parent, c->d_run_label, c->d_scalars);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// materialise: closed mask (u8), label volume (u32), body mask (u8) -- the per-voxel outputs
// ------------------------------------------------------------------------------------------------
// Aligned path (nx % 32 == 0): a warp writes 4 words = 128 voxels per iteration; lane l owns word
// l/8, voxels 4*(l%8)..+3 -> one 16-byte label store and one 4-byte mask store per lane, i.e.
// 512 B / 128 B contiguous per warp instruction.
template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_materialise(const uint32_t* __restrict__ mask,
                                                     const uint32_t* __restrict__ word_base,
                                                     const uint32_t* __restrict__ run_label, int nx, int W,
                                                     uint32_t n_words, uint8_t* __restrict__ mask_out,
                                                     uint32_t* __restrict__ labels_out, uint8_t* __restrict__ body_out,
                                                     const DevScalars* sc) {
    const bool ok = sc->status == MAMRI_OK;
    const uint32_t body = body_out ? (0xFFFFFFFFu - uint32_t(sc->body_packed & 0xFFFFFFFFull)) : 0u;
    const bool has_body = body_out && (sc->body_packed >> 32) != 0;
    const unsigned lane = lane_id();
    const uint32_t warp = (uint32_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (uint32_t(gridDim.x) * blockDim.x) >> 5;
    const bool need_labels = (labels_out != nullptr) || (body_out != nullptr);
    if (ALIGNED) {
        // 32 mask words per trip: one coalesced load, then 8 rounds in which lane l serves word
        // 4k + l/8 (fetched by shuffle), voxels 4*(l%8)..+3.  The next trip's words are loaded first.
        uint32_t w0 = warp * 32;
        uint32_t m_next = (w0 + lane < n_words) ? mask[w0 + lane] : 0u;
        for (; w0 < n_words; w0 += n_warps * 32) {
            const uint32_t m_l = m_next;
            const uint32_t wn = w0 + n_warps * 32 + lane;
            m_next = (wn < n_words) ? mask[wn] : 0u;
            const bool resolve = need_labels && ok && __any_sync(0xFFFFFFFFu, m_l != 0u);
            uint32_t starts_l = 0, base_l = 0;
            if (resolve) {
                uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, m_l, 1);
                if (lane == 0) prev = w0 > 0 ? mask[w0 - 1] : 0u;
                if (m_l) {
                    if ((w0 + lane) % W == 0) prev = 0u;
                    starts_l = run_starts(m_l, prev);
                    base_l = word_base[w0 + lane];
                }
            }
            const int sub = int(lane & 7u) * 4;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int src = k * 4 + int(lane >> 3);
                const uint32_t m = __shfl_sync(0xFFFFFFFFu, m_l, src);
                uint32_t starts = 0, base = 0;
                if (resolve) {
                    starts = __shfl_sync(0xFFFFFFFFu, starts_l, src);
                    base = __shfl_sync(0xFFFFFFFFu, base_l, src);
                }
                const uint32_t wi = w0 + src;
                if (wi >= n_words) continue;
                const uint32_t nib = (m >> sub) & 0xFu;
                uint32_t lab[4] = {0u, 0u, 0u, 0u};
                if (nib && resolve) {
                    uint32_t last_rid = MAMRI_NONE, last_lab = 0u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (nib & (1u << q)) {
                            uint32_t rid = run_id_in_word(base, starts, sub + q);
                            if (rid != last_rid) { last_rid = rid; last_lab = run_label[rid]; }
                            lab[q] = last_lab;
                        }
                    }
                }
                const uint32_t v = wi * 32 + sub;
                if (labels_out) *reinterpret_cast<uint4*>(labels_out + v) = make_uint4(lab[0], lab[1], lab[2], lab[3]);
                if (mask_out) {
                    uint32_t mb = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                    *reinterpret_cast<uint32_t*>(mask_out + v) = mb;
                }
                if (body_out) {
                    uint32_t bb = 0;
                    if (has_body) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) bb |= (lab[q] == body ? 1u : 0u) << (8 * q);
                    }
                    *reinterpret_cast<uint32_t*>(body_out + v) = bb;
                }
            }
        }
    } else {
        for (uint32_t wi = warp; wi < n_words; wi += n_warps) {
            const uint32_t row = wi / W;
            const int xw = int(wi - row * W);
            const int x = xw * 32 + int(lane);
            if (x >= nx) continue;
            const uint32_t m = mask[wi];
            const bool on = (m >> lane) & 1u;
            uint32_t lab = 0u;
            if (on && need_labels && ok) {
                const uint32_t prev = xw > 0 ? mask[wi - 1] : 0u;
                lab = run_label[run_id_in_word(word_base[wi], run_starts(m, prev), int(lane))];
            }
            const uint32_t v = row * uint32_t(nx) + x;
            if (labels_out) labels_out[v] = lab;
            if (mask_out) mask_out[v] = on ? 1 : 0;
            if (body_out) body_out[v] = (has_body && lab == body) ? 1 : 0;
        }
    }
}

cudaError_t launch_materialise(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, uint8_t* d_mask_out,
                               uint32_t* d_labels_out, uint8_t* d_body_out, cudaStream_t s) {
    if (!d_mask_out && !d_labels_out && !d_body_out) return cudaSuccess;
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    const bool aligned = (nx % 32 == 0) && ((reinterpret_cast<uintptr_t>(d_labels_out) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(d_mask_out) & 3u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(d_body_out) & 3u) == 0);
    if (aligned) {
        uint32_t blocks = (n_words / 32 + 7) / 8;
        if (blocks > 148 * 8 * 2) blocks = 148 * 8 * 2;
        if (blocks == 0) blocks = 1;
        k_materialise<true><<<unsigned(blocks), 256, 0, s>>>(d_mask, c->d_word_base, c->d_run_label, nx, W, n_words,
                                                             d_mask_out, d_labels_out, d_body_out, c->d_scalars);
    } else {
        uint32_t blocks = (n_words + 7) / 8;
        if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
        if (blocks == 0) blocks = 1;
        k_materialise<false><<<unsigned(blocks), 256, 0, s>>>(d_mask, c->d_word_base, c->d_run_label, nx, W, n_words,
                                                              d_mask_out, d_labels_out, d_body_out, c->d_scalars);
    }
    return cudaGetLastError();
}
