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

#include <stdlib.h>

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
// run numbering: exclusive scan of run starts per word, and the run table
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_runs_count(const uint32_t* __restrict__ mask, int W, uint32_t n_words,
                                                             uint32_t* __restrict__ block_sums) {
    pdl_wait();
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

// Exclusive prefix of this CTA over the per-CTA partials (every CTA rescans the <= 1024 partials
// itself: cheaper than a separate one-CTA launch); `total` = grand total.
__device__ __forceinline__ uint32_t scan_partials(const uint32_t* __restrict__ block_sums, uint32_t* ws, uint32_t& total) {
    __shared__ uint32_t mine;
    const uint32_t v = threadIdx.x < gridDim.x ? block_sums[threadIdx.x] : 0u;
    const uint32_t ex = block_excl_scan(v, ws, total);
    if (threadIdx.x == blockIdx.x) mine = ex;
    __syncthreads();
    return mine;
}

// word_base[w] = number of runs that start before word w; run table: position (word*32 + bit of the
// first voxel) and length of every run, in raster order.
__global__ void __launch_bounds__(SCAN_THREADS) k_runs_assign(const uint32_t* __restrict__ mask, int W, uint32_t n_words,
                                                              const uint32_t* __restrict__ block_sums,
                                                              uint32_t* __restrict__ word_base,
                                                              uint32_t* __restrict__ run_pos, uint32_t* __restrict__ run_len,
                                                              uint32_t max_runs, DevScalars* sc) {
    pdl_wait();
    __shared__ uint32_t ws[33];
    uint32_t total_runs;
    uint32_t running = scan_partials(block_sums, ws, total_runs);
    if (total_runs > max_runs) {                      // every CTA sees the same total
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc->status = MAMRI_ERR_CAPACITY; sc->n_runs = 0; }
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_runs = total_runs;
    uint32_t begin, end;
    chunk_of(n_words, begin, end);
    for (uint32_t i0 = begin; i0 < end; i0 += SCAN_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t m = 0, starts = 0, xw = 0;
        if (i < end) {
            m = mask[i];
            if (m) {
                xw = i % W;
                starts = run_starts(m, xw ? mask[i - 1] : 0u);
            }
        }
        uint32_t cnt = __popc(starts), total;
        uint32_t base = running + block_excl_scan(cnt, ws, total);
        if (i < end) {
            word_base[i] = base;
            uint32_t st = starts, r = base;
            while (st) {
                const int b = __ffs(st) - 1;
                st &= st - 1;
                const uint32_t t = m >> b;                       // run bits from its start
                uint32_t len;
                if (t != (0xFFFFFFFFu >> b)) {
                    len = __ffs(~t) - 1;                         // ends inside this word
                } else {
                    len = 32 - b;                                // reaches bit 31: follow it through the next words
                    for (uint32_t k = 1; xw + k < uint32_t(W); ++k) {
                        const uint32_t nm = mask[i + k];
                        if (nm == 0xFFFFFFFFu) { len += 32; continue; }
                        len += __ffs(~nm) - 1;
                        break;
                    }
                }
                run_pos[r] = i * 32u + uint32_t(b);
                run_len[r] = len;
                ++r;
            }
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

// Joins run `node` (x range [gx0, gx0+len) of its row) with every run of one earlier neighbour row
// that touches it: same x for face connectivity; DIAG widens the range by one voxel each side
// (26-connectivity).  Every maximal piece of neighbour bits inside the range lies in exactly one
// neighbour run, so one union per piece start suffices (a piece continuing from the previous word
// belongs to the run already joined).  Node ids are run ids minus `id_off`.
template <bool DIAG>
__device__ __forceinline__ void join_run(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ word_base,
                                         uint32_t* parent, uint32_t id_off, int W, uint32_t node, int gx0, int len,
                                         uint32_t nbr_row) {
    int lo = gx0 - (DIAG ? 1 : 0), hi = gx0 + len - 1 + (DIAG ? 1 : 0);
    if (lo < 0) lo = 0;
    if (hi > W * 32 - 1) hi = W * 32 - 1;
    const int w_lo = lo >> 5, w_hi = hi >> 5;
    uint32_t carry = 0;
    for (int w = w_lo; w <= w_hi; ++w) {
        const uint32_t up = mask[nbr_row + w];
        uint32_t seg = 0xFFFFFFFFu;
        if (w == w_lo) seg &= 0xFFFFFFFFu << (lo & 31);
        if (w == w_hi) seg &= 0xFFFFFFFFu >> (31 - (hi & 31));
        const uint32_t t = up & seg;
        uint32_t ps = t & ~((t << 1) | carry);
        if (ps) {
            const uint32_t up_prev = w > 0 ? mask[nbr_row + w - 1] : 0u;
            const uint32_t starts_up = run_starts(up, up_prev);
            const uint32_t base_up = word_base[nbr_row + w] - id_off;
            while (ps) {
                const int bit = __ffs(ps) - 1;
                ps &= ps - 1;
                uf_union(parent, node, run_id_in_word(base_up, starts_up, bit));
            }
        }
        carry = t >> 31;
    }
}

// Phase 1 -- block-local: one CTA per z-slice.  The runs of a slice are contiguous in id space, so the
// CTA keeps their parents in shared memory (local index = run id - first run of the slice), joins
// every run with the runs of the row above at shared-memory latency (simultaneous hooking builds
// chains as long as the object is tall; in shared memory a hop costs ~30 cycles instead of an L2 round
// trip), flattens, and publishes parent[run] = slice-local root as a global id.  Slices with more runs
// than fit fall back to the same code on the global array.
constexpr int SLICE_THREADS = 512;
constexpr uint32_t SLICE_SMEM_RUNS = 12000;      // 48 KB static shared memory

template <bool CONN26>
__global__ void __launch_bounds__(SLICE_THREADS) k_union_slices(const uint32_t* __restrict__ mask,
                                                               const uint32_t* __restrict__ word_base,
                                                               const uint32_t* __restrict__ run_pos,
                                                               const uint32_t* __restrict__ run_len, uint32_t* parent,
                                                               int W, int ny, int nz, const DevScalars* sc) {
    pdl_wait();
    __shared__ uint32_t sp[SLICE_SMEM_RUNS];
    if (sc->status != MAMRI_OK) return;
    const uint32_t z = blockIdx.x;
    const uint32_t slice_words = uint32_t(W) * ny;
    const uint32_t w0 = z * slice_words;
    const uint32_t r0 = word_base[w0];
    const uint32_t r1 = (z + 1 < uint32_t(nz)) ? word_base[w0 + slice_words] : sc->n_runs;
    const uint32_t n = r1 - r0;
    if (n == 0) return;
    uint32_t* P = (n <= SLICE_SMEM_RUNS) ? sp : parent + r0;
    for (uint32_t i = threadIdx.x; i < n; i += SLICE_THREADS) P[i] = i;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += SLICE_THREADS) {
        const uint32_t pos = run_pos[r0 + i];
        const uint32_t wi = pos >> 5, row = wi / W;
        if (row == z * ny) continue;                                        // y == 0: no row above in this slice
        const int gx0 = int((wi - row * W) * 32 + (pos & 31u));
        join_run<CONN26>(mask, word_base, P, r0, W, i, gx0, int(run_len[r0 + i]), (row - 1) * W);
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
                                                 const uint32_t* __restrict__ run_pos, const uint32_t* __restrict__ run_len,
                                                 uint32_t* parent, int W, int ny, int radix, int between_blocks,
                                                 const DevScalars* sc) {
    pdl_wait();
    if (sc->status != MAMRI_OK) return;
    const uint32_t n = sc->n_runs;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t pos = run_pos[r];
        const uint32_t wi = pos >> 5, row = wi / W;
        const uint32_t z = row / ny;
        if (z == 0 || ((z % radix) == 0) != (between_blocks != 0)) continue;
        const uint32_t y = row - z * ny;
        const int gx0 = int((wi - row * W) * 32 + (pos & 31u));
        const int len = int(run_len[r]);
        const uint32_t below = (row - ny) * W;                              // row (y, z-1)
        join_run<CONN26>(mask, word_base, parent, 0u, W, r, gx0, len, below);
        if (CONN26) {
            if (y > 0) join_run<true>(mask, word_base, parent, 0u, W, r, gx0, len, below - W);
            if (y + 1 < uint32_t(ny)) join_run<true>(mask, word_base, parent, 0u, W, r, gx0, len, below + W);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// flatten, rank the roots (ITK-consecutive labels)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_flatten_count(uint32_t* parent, uint32_t* __restrict__ block_sums,
                                                                const DevScalars* sc) {
    pdl_wait();
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

// run_label[root] = 1 + number of roots before it (roots are ordered by minimum linear index, so this
// is ITK's consecutive numbering); non-root runs get their label from their root in the next kernel.
__global__ void __launch_bounds__(SCAN_THREADS) k_rank_roots(const uint32_t* __restrict__ parent,
                                                             const uint32_t* __restrict__ block_sums,
                                                             uint32_t* __restrict__ run_label,
                                                             uint32_t* __restrict__ label_count, DevScalars* sc) {
    pdl_wait();
    __shared__ uint32_t ws[33];
    uint32_t total_roots;
    uint32_t running = scan_partials(block_sums, ws, total_roots);
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_labels = total_roots;
    uint32_t begin, end;
    chunk_of(sc->n_runs, begin, end);
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

cudaError_t launch_ccl(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int connectivity, cudaStream_t s) {
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    const int G = MAMRI_SCAN_CTAS;
    uint32_t* bs_runs = c->d_block_sums;
    uint32_t* bs_roots = c->d_block_sums + 1024;
    LK(k_runs_count, G, SCAN_THREADS, s, false, d_mask, W, n_words, bs_runs);
    prof_mark(c, s, "runs_count");
    LK(k_runs_assign, G, SCAN_THREADS, s, false, d_mask, W, n_words, bs_runs, c->d_word_base, c->d_run_pos, c->d_run_len, c->max_runs, c->d_scalars);
    prof_mark(c, s, "runs_assign");
    int radix = 1;
    while (radix * radix < nz) radix <<= 1;
    const int RG = MAMRI_RUN_CTAS;
    if (connectivity == 26) {
        LK(k_union_slices<true>, nz, SLICE_THREADS, s, false, d_mask, c->d_word_base, c->d_run_pos, c->d_run_len, c->d_parent, W, ny, nz, c->d_scalars);
        prof_mark(c, s, "union_slices");
        if (nz > 1) {
            LK(k_union_z<true>, RG, 256, s, false, d_mask, c->d_word_base, c->d_run_pos, c->d_run_len, c->d_parent, W, ny, radix, 0, c->d_scalars);
            prof_mark(c, s, "union_z_within_blocks");
        }
        if (nz > radix) {
            LK(k_union_z<true>, RG, 256, s, false, d_mask, c->d_word_base, c->d_run_pos, c->d_run_len, c->d_parent, W, ny, radix, 1, c->d_scalars);
            prof_mark(c, s, "union_z_between_blocks");
        }
    } else {
        LK(k_union_slices<false>, nz, SLICE_THREADS, s, false, d_mask, c->d_word_base, c->d_run_pos, c->d_run_len, c->d_parent, W, ny, nz, c->d_scalars);
        prof_mark(c, s, "union_slices");
        if (nz > 1) {
            LK(k_union_z<false>, RG, 256, s, false, d_mask, c->d_word_base, c->d_run_pos, c->d_run_len, c->d_parent, W, ny, radix, 0, c->d_scalars);
            prof_mark(c, s, "union_z_within_blocks");
        }
        if (nz > radix) {
            LK(k_union_z<false>, RG, 256, s, false, d_mask, c->d_word_base, c->d_run_pos, c->d_run_len, c->d_parent, W, ny, radix, 1, c->d_scalars);
            prof_mark(c, s, "union_z_between_blocks");
        }
    }
    LK(k_flatten_count, G, SCAN_THREADS, s, false, c->d_parent, bs_roots, c->d_scalars);
    prof_mark(c, s, "flatten_count");
    LK(k_rank_roots, G, SCAN_THREADS, s, false, c->d_parent, bs_roots, c->d_run_label, c->d_label_count, c->d_scalars);
    prof_mark(c, s, "rank_roots");
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
                                                     uint32_t n_words, const DynArgs* __restrict__ dyn,
                                                     const DevScalars* sc) {
    pdl_wait();
    uint8_t* __restrict__ mask_out = dyn->mask_out;
    uint32_t* __restrict__ labels_out = dyn->labels_out;
    uint8_t* __restrict__ body_out = dyn->body_out;
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

cudaError_t launch_materialise(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int outs_aligned,
                               cudaStream_t s) {
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    const bool aligned = (nx % 32 == 0) && outs_aligned;
    if (aligned) {
        uint32_t blocks = (n_words / 32 + 7) / 8;
        static const int per_sm = [] { const char* e = getenv("MAMRI_MAT_CTAS_PER_SM"); return e ? atoi(e) : 16; }();
        if (blocks > uint32_t(148 * per_sm)) blocks = uint32_t(148 * per_sm);
        if (blocks == 0) blocks = 1;
        LK(k_materialise<true>, blocks, 256, s, true, d_mask, c->d_word_base, c->d_run_label, nx, W, n_words, c->d_dyn, c->d_scalars);
    } else {
        uint32_t blocks = (n_words + 7) / 8;
        if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
        if (blocks == 0) blocks = 1;
        LK(k_materialise<false>, blocks, 256, s, true, d_mask, c->d_word_base, c->d_run_label, nx, W, n_words, c->d_dyn, c->d_scalars);
    }
    prof_mark(c, s, "materialise");
    return cudaGetLastError();
}
