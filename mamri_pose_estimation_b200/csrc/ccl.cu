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

// ------------------------------------------------------------------------------------------------
// warp scan
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

// ------------------------------------------------------------------------------------------------
// run numbering: single-pass exclusive scan of run starts per word, and the run table
// ------------------------------------------------------------------------------------------------
// word_base[w] = number of runs that start before word w; run table: position (word*32 + bit) of the
// first and of the last voxel of every run (run_pos, run_end), in raster order.  One pass over the mask: a CTA takes a tile of
// RS_TILE words (16 consecutive words per thread, 128-bit loads), scans its run-start counts, gets
// the count of all earlier tiles by decoupled look-back and writes its part of both tables.
// Tiles of RS_THREADS x ITEMS words: 16 words per thread on large masks (few, large tiles), 4 on small ones (enough
// CTAs to use the machine).
constexpr int RS_THREADS = 512;

template <int WARPS>
__device__ __forceinline__ uint32_t block_excl_scan_w(uint32_t v, uint32_t* ws /*[WARPS]*/, uint32_t& total) {
    const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t t = ws[w];
        if (w < int(wid)) off += t;
        tot += t;
    }
    total = tot;
    __syncthreads();
    return off + inc - v;
}

// Words of one thread's part of a tile, and the run starts / ends in them.  A run starts where a set bit follows a
// clear one and ends where a clear one follows a set one (row ends cut runs); the j-th start and the j-th end in raster
// order belong to the same run, so lengths need no walk along the run: starts go to run_pos, ends to run_end, both
// indexed by the same scan.  `open` = a run is open across the thread's first word (its end belongs to an earlier start).
template <int RS_ITEMS>
__device__ __forceinline__ uint32_t load_runs(const uint32_t* __restrict__ mask, uint32_t i0, uint32_t n_words, int W,
                                              uint32_t (&m)[RS_ITEMS], uint32_t (&starts)[RS_ITEMS], uint32_t (&ends)[RS_ITEMS],
                                              uint32_t& open, bool want_ends) {
    if (i0 + RS_ITEMS <= n_words) {
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(mask + i0 + k);
            m[k] = q.x; m[k + 1] = q.y; m[k + 2] = q.z; m[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) m[k] = (i0 + k < n_words) ? mask[i0 + k] : 0u;
    }
    const uint32_t xw0 = i0 % uint32_t(W);
    uint32_t prev = (m[0] != 0u && i0 > 0) ? mask[i0 - 1] : 0u;
    const uint32_t after = (want_ends && m[RS_ITEMS - 1] != 0u && i0 + RS_ITEMS < n_words) ? mask[i0 + RS_ITEMS] : 0u;
    uint32_t xw = xw0, cnt = 0;
    open = (xw0 != 0u && (prev >> 31) && (m[0] & 1u)) ? 1u : 0u;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const bool row_first = xw == 0u, row_last = xw + 1u == uint32_t(W);
        uint32_t nxt = after;
        if (k + 1 < RS_ITEMS) nxt = m[(k + 1) % RS_ITEMS];
        starts[k] = run_starts(m[k], row_first ? 0u : prev);
        ends[k] = m[k] & ~((m[k] >> 1) | (row_last ? 0u : (nxt << 31)));
        cnt += __popc(starts[k]);
        prev = m[k];
        if (++xw == uint32_t(W)) xw = 0;
    }
    return cnt;
}

// Two phases over a CTA's tiles (tile = blockIdx.x + k gridDim.x), so that no tile's prefix waits for another tile's
// table writes: (A) every tile's run count is published as soon as its words have been counted; (B) each tile sums
// the counts before it (values that phase A of all CTAs has published, or is about to, without waiting for
// anything), re-reads its words from L2 and writes word_base and the run table.  With one CTA per
// tile, or tiles handed out by a ticket, the late tiles of a big mask wait for the full duration of the earlier ones.
template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS) k_runs_scan(const uint32_t* __restrict__ mask, int W, uint32_t n_words,
                                                          volatile unsigned long long* state, const DynArgs* __restrict__ dyn,
                                                          uint32_t* __restrict__ word_base, uint32_t* __restrict__ run_pos,
                                                          uint32_t* __restrict__ run_end, uint32_t* __restrict__ root_count,
                                                          uint32_t max_runs, uint32_t n_tiles, DevScalars* sc) {
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    pdl_wait();
    ktrace(KT_RUNS);
    __shared__ uint32_t ws[RS_THREADS / 32];
    __shared__ uint32_t red[34];
    const uint32_t gen = dyn->gen;
    // ---- phase A: counts
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t m[RS_ITEMS], starts[RS_ITEMS], ends[RS_ITEMS], open;
        uint32_t cnt = load_runs<RS_ITEMS>(mask, tile * RS_TILE + threadIdx.x * RS_ITEMS, n_words, W, m, starts, ends, open, false);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
        __syncthreads();                                               // ws is free again
        if (lane_id() == 0) ws[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; ++w) total += ws[w];
            state[tile] = scan_pack(gen, total, 1u);
        }
    }
    ktrace(KT_RUNS_LB);
    // ---- phase B: prefixes and tables
    uint32_t carry = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t i0 = tile * RS_TILE + threadIdx.x * RS_ITEMS;
        uint32_t m[RS_ITEMS], starts[RS_ITEMS], ends[RS_ITEMS], open;
        const uint32_t cnt = load_runs<RS_ITEMS>(mask, i0, n_words, W, m, starts, ends, open, true);
        // prefix of the tile = runs before the CTA's previous tile and in it (`carry`, known from the round before) + the
        // published counts of the tiles since then -- at most gridDim.x - 1 values, one per thread (a count of this
        // launch that is not there yet is about to be: phase A waits for nothing), instead of all the earlier counts.
        // Measured on the 2048-tile mask of config C4: phase B is ~102 us either way (7 us per tile and CTA: the loads of
        // the words, the counts, two barriers and the table stores in sequence, 16 warps per SM); staging the table
        // entries in shared memory for whole-sector stores was slower (127 us), so the entries are written straight out.
        uint32_t acc = 0;
        {
            const unsigned long long g30 = gen & 0x3FFFFFFFu;
            const uint32_t first = tile >= gridDim.x ? tile - gridDim.x + 1u : 0u;
            for (uint32_t idx = first + threadIdx.x; idx < tile; idx += blockDim.x) {
                unsigned long long v;
                do { v = state[idx]; } while ((v >> 34) != g30);
                acc += uint32_t(v >> 2);
            }
        }
        const uint32_t inc = warp_incl_scan(cnt);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        __syncthreads();                                               // ws / red are free again
        if (lane_id() == 31) ws[threadIdx.x >> 5] = inc;
        if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        uint32_t ex = inc - cnt, total = 0, before = carry;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) {
            const uint32_t t = ws[w];
            if (w < int(threadIdx.x >> 5)) ex += t;
            total += t;
            before += red[w];
        }
        carry = before + total;                                          // runs up to and including this tile
        if (threadIdx.x == 0 && tile == n_tiles - 1) {  // grand total: every later stage keys off n_runs / status
            if (before + total > max_runs) { sc->status = MAMRI_ERR_CAPACITY; sc->n_runs = 0; }
            else sc->n_runs = before + total;
        }
        uint32_t base = before + ex;
        uint32_t e_idx = base - open;                                    // ends before this thread = starts before - open runs
        uint32_t wb[RS_ITEMS];
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            wb[k] = base;
            const uint32_t bit0 = (i0 + uint32_t(k)) * 32u;
            uint32_t st = starts[k];
            while (st) {
                const int b = __ffs(st) - 1;
                st &= st - 1;
                if (base < max_runs) {
                    run_pos[base] = bit0 + uint32_t(b);
                    root_count[base] = 0u;                               // accumulated on the roots by the labelling
                }
                ++base;
            }
            uint32_t en = ends[k];
            while (en) {
                const int b = __ffs(en) - 1;
                en &= en - 1;
                if (e_idx < max_runs) run_end[e_idx] = bit0 + uint32_t(b);
                ++e_idx;
            }
        }
        if (i0 + RS_ITEMS <= n_words) {
#pragma unroll
            for (int k = 0; k < RS_ITEMS; k += 4)
                *reinterpret_cast<uint4*>(word_base + i0 + k) = make_uint4(wb[k], wb[k + 1], wb[k + 2], wb[k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < RS_ITEMS; ++k)
                if (i0 + k < n_words) word_base[i0 + k] = wb[k];
        }
    }
    ktrace_last(KT_RUNS_LAST);
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

// Two finds at once: the hops of the two chains are independent loads, so they travel together (one round trip per
// pair of hops instead of two).  Same path halving as uf_find.
__device__ __forceinline__ void uf_find2(uint32_t* parent, uint32_t& x, uint32_t& y) {
    volatile uint32_t* vp = parent;
    uint32_t px = vp[x], py = vp[y];
    while (px != x || py != y) {
        const uint32_t gx = px != x ? vp[px] : px;
        const uint32_t gy = py != y ? vp[py] : py;
        if (px != x) { if (gx != px) vp[x] = gx; x = px; px = gx; }
        if (py != y) { if (gy != py) vp[y] = gy; y = py; py = gy; }
    }
}

__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    while (true) {
        uf_find2(parent, a, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }       // hook the larger root under the smaller
        uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;                               // a was still a root: done
        a = old;                                            // someone re-parented a meanwhile: keep merging
    }
}

// Joins run `node` (x range [gx0, gx0+len) of its row) with every run of one earlier neighbour row that touches
// it: same x for face connectivity; DIAG widens the range by one voxel each side (26-connectivity).  The neighbour
// runs come from the run table, not from the mask: word_base gives the ids of the runs that start inside the words
// the range covers, plus the one run that may reach into the range from the left -- two or three dependent loads
// whatever the length of the run (walking the mask costs one load per 32 voxels, and the body's runs are hundreds of
// voxels long).  `nbr_row` = word index of the neighbour row's first word; node ids are run ids minus `id_off`.
template <bool DIAG>
__device__ __forceinline__ void join_run(const uint32_t* __restrict__ word_base, const uint32_t* __restrict__ run_pos,
                                         const uint32_t* __restrict__ run_end, uint32_t* parent, uint32_t id_off, int W,
                                         uint32_t node, int gx0, int len, uint32_t nbr_row) {
    int lo = gx0 - (DIAG ? 1 : 0), hi = gx0 + len - 1 + (DIAG ? 1 : 0);
    if (lo < 0) lo = 0;
    if (hi > W * 32 - 1) hi = W * 32 - 1;
    const uint32_t lo_abs = nbr_row * 32u + uint32_t(lo), hi_abs = nbr_row * 32u + uint32_t(hi);
    const uint32_t j_lo = word_base[nbr_row + uint32_t(lo >> 5)];          // runs that start before the first word of the range
    const uint32_t j_hi = word_base[nbr_row + uint32_t(hi >> 5) + 1u];     // ... before the word after its last (a later row at most)
    // the last run that starts left of the range reaches into it iff it ends at or after lo (a run of an earlier row
    // ends before this row begins, so it fails the test by itself)
    if (j_lo > 0u && run_end[j_lo - 1u] >= lo_abs) uf_union(parent, node, j_lo - 1u - id_off);
    for (uint32_t j = j_lo; j < j_hi; ++j) {
        const uint32_t sj = run_pos[j];
        if (sj > hi_abs) break;                                            // starts right of the range: so do the rest
        if (run_end[j] >= lo_abs) uf_union(parent, node, j - id_off);
    }
}

// The runs of row `nbr_row` that touch [gx0, gx0 + len) (widened by one voxel each side for DIAG) have consecutive
// ids [ja, jb): same rule as join_run, without doing the unions.
template <bool DIAG>
__device__ __forceinline__ void nbr_range(const uint32_t* __restrict__ word_base, const uint32_t* __restrict__ run_pos,
                                          const uint32_t* __restrict__ run_end, int W, int gx0, int len, uint32_t nbr_row,
                                          uint32_t& ja, uint32_t& jb) {
    int lo = gx0 - (DIAG ? 1 : 0), hi = gx0 + len - 1 + (DIAG ? 1 : 0);
    if (lo < 0) lo = 0;
    if (hi > W * 32 - 1) hi = W * 32 - 1;
    const uint32_t lo_abs = nbr_row * 32u + uint32_t(lo), hi_abs = nbr_row * 32u + uint32_t(hi);
    const uint32_t j_lo = word_base[nbr_row + uint32_t(lo >> 5)];
    const uint32_t j_hi = word_base[nbr_row + uint32_t(hi >> 5) + 1u];
    ja = j_lo;
    if (j_lo > 0u && run_end[j_lo - 1u] >= lo_abs) ja = j_lo - 1u;
    else while (ja < j_hi && run_end[ja] < lo_abs) ++ja;              // runs of the first word that end left of the range
    jb = ja > j_lo ? ja : j_lo;
    while (jb < j_hi && run_pos[jb] <= hi_abs) ++jb;
}

// Phase 1 -- block-local: one CTA per z-slice.  The runs of a slice are contiguous in id space, so the
// CTA keeps their parents in shared memory (local index = run id - first run of the slice), joins
// every run with the runs of the row above at shared-memory latency (simultaneous hooking builds
// chains as long as the object is tall; in shared memory a hop costs ~30 cycles instead of an L2 round
// trip), flattens, and publishes parent[run] = slice-local root as a global id.  Slices with more runs
// than fit fall back to the same code on the global array.
constexpr int SLICE_THREADS = 512;
constexpr uint32_t SLICE_SMEM_RUNS = 11776;      // 46 KB of parents + 1.4 KB of flags: under the 48 KB of static shared memory

template <bool CONN26>
__global__ void __launch_bounds__(SLICE_THREADS) k_union_slices(const uint32_t* __restrict__ word_base,
                                                               const uint32_t* __restrict__ run_pos,
                                                               const uint32_t* __restrict__ run_end, uint32_t* parent,
                                                               int W, int ny, int nz, const DevScalars* sc) {
    pdl_wait();
    ktrace(KT_USLICE);
    __shared__ uint32_t sp[SLICE_SMEM_RUNS];
    __shared__ uint32_t sflag[SLICE_SMEM_RUNS / 32];                    // runs that touch more than one run of the row above
    if (sc->status != MAMRI_OK) return;
    const uint32_t z = blockIdx.x;
    const uint32_t slice_words = uint32_t(W) * ny;
    const uint32_t w0 = z * slice_words;
    const uint32_t r0 = word_base[w0];
    const uint32_t r1 = (z + 1 < uint32_t(nz)) ? word_base[w0 + slice_words] : sc->n_runs;
    const uint32_t n = r1 - r0;
    ktrace_both(KT_US_N);
    if (n == 0) return;
    const bool in_smem = n <= SLICE_SMEM_RUNS;
    uint32_t* P = in_smem ? sp : parent + r0;
    if (in_smem) {
        for (uint32_t i = threadIdx.x; i < (n + 31u) / 32u; i += blockDim.x) sflag[i] = 0u;
        __syncthreads();
    }
    // (1) Every run points at the first run of the row above that touches it (a smaller id), or at itself: a forest,
    // built with plain stores -- no atomics, no root walks.  Hooking root under root while the other threads do the
    // same leaves chains as long as the object is tall, and the late warps walk them hop by hop.
    constexpr int KEEP = 2;
    uint32_t ja_[KEEP], jb_[KEEP];
    bool more = false;
    {
        int k = 0;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x, ++k) {
            const uint32_t pos = run_pos[r0 + i];
            const uint32_t wi = pos >> 5, row = wi / W;
            uint32_t ja = 0, jb = 0;
            if (row != z * ny) {                                            // y == 0: no row above in this slice
                const int gx0 = int((wi - row * W) * 32 + (pos & 31u));
                nbr_range<CONN26>(word_base, run_pos, run_end, W, gx0, int(run_end[r0 + i] - pos + 1u), (row - 1) * W, ja, jb);
            }
            P[i] = ja < jb ? ja - r0 : i;
            if (jb - ja > 1u) {
                more = true;
                if (in_smem && k >= KEEP) atomicOr(&sflag[i >> 5], 1u << (i & 31u));
            }
            if (k < KEEP) { ja_[k] = ja; jb_[k] = jb; }
        }
    }
#ifdef MAMRI_KTRACE
    __syncthreads();                                                    // trace build: stamp when the whole CTA is through
#endif
    ktrace_both(KT_US_JOIN);
    // (2) flatten by pointer jumping: log2(height) rounds
    auto flatten = [&]() {
        bool again = true;
        while (again) {
            __syncthreads();
            bool changed = false;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t p = P[i], pp = P[p];   // a concurrent update of P[p] still yields an ancestor
                if (pp != p) { P[i] = pp; changed = true; }
            }
            again = __syncthreads_or(changed);
        }
    };
    flatten();
    // (3) the runs that touch more than one run of the row above join the trees of the others (two branches of an object
    // meeting): real unions, on trees that are flat now; flatten again if there were any
    if (__syncthreads_or(more)) {
        int k = 0;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x, ++k) {
            uint32_t ja, jb;
            if (k < KEEP) { ja = ja_[k]; jb = jb_[k]; }
            else {
                // a thread with more than KEEP runs (noisy slices: a dozen per thread) looks the neighbours up again --
                // but only for the flagged runs, a few per slice: without the flags this second pass over every run cost
                // as much as the first (dependent loads from the run table) on every slice that holds an object
                if (in_smem && !((sflag[i >> 5] >> (i & 31u)) & 1u)) continue;
                const uint32_t pos = run_pos[r0 + i];
                const uint32_t wi = pos >> 5, row = wi / W;
                ja = jb = 0;
                if (row != z * ny)
                    nbr_range<CONN26>(word_base, run_pos, run_end, W, int((wi - row * W) * 32 + (pos & 31u)),
                                      int(run_end[r0 + i] - pos + 1u), (row - 1) * W, ja, jb);
            }
            for (uint32_t j = ja + 1; j < jb; ++j) uf_union(P, i, j - r0);
        }
        flatten();
    }
    ktrace_both(KT_US_FLAT);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) parent[r0 + i] = r0 + P[i];
    ktrace_both(KT_US_END);
}

// Phase 2 -- global boundary merge between slices, with atomicMin on the roots.  Joining all slice
// boundaries at once would again chain the roots nz deep (every hop an L2 round trip), so the
// boundaries are merged in two rounds: first those inside blocks of `radix` slices (chains <= radix),
// then the boundaries between blocks (again <= radix of them per chain).
template <bool CONN26>
__global__ void __launch_bounds__(256) k_union_z(const uint32_t* __restrict__ word_base,
                                                 const uint32_t* __restrict__ run_pos, const uint32_t* __restrict__ run_end,
                                                 uint32_t* parent, int W, int ny, int radix, int between_blocks,
                                                 const DevScalars* sc) {
    pdl_wait();
    ktrace(between_blocks ? KT_UZ2 : KT_UZ1);
    if (sc->status != MAMRI_OK) return;
    const uint32_t n = sc->n_runs;
    const unsigned lane = lane_id();
    // One run per lane, the warp in step: most runs of a slice belong to one object (the body), so most lanes of a
    // warp ask for the SAME union (root of the object in this slice, root of it in the slice below).  Lanes with equal
    // root pairs elect one of them to do it: hundreds of atomicMin on one address, and the re-walks they cause,
    // become a handful.
    for (uint32_t r0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; r0 < n; r0 += gridDim.x * blockDim.x) {
        const uint32_t r = r0 + lane;
        bool active = r < n;
        uint32_t row = 0, y = 0;
        int gx0 = 0, len = 0;
        if (active) {
            const uint32_t pos = run_pos[r], end = run_end[r];       // both in flight together (coalesced)
            const uint32_t wi = pos >> 5;
            row = wi / W;
            const uint32_t z = row / ny;
            y = row - z * ny;
            active = z > 0 && (((z % radix) == 0) == (between_blocks != 0));
            gx0 = int((wi - row * W) * 32 + (pos & 31u));
            len = int(end - pos + 1u);
        }
        if (!__any_sync(FULL, active)) continue;
        const uint32_t below = active ? (row - ny) * W : 0u;            // row (y, z-1)
#pragma unroll
        for (int q = 0; q < (CONN26 ? 3 : 1); ++q) {
            uint32_t ja = 0, jb = 0;
            if (active) {
                if (q == 0) nbr_range<CONN26>(word_base, run_pos, run_end, W, gx0, len, below, ja, jb);
                else if (q == 1) { if (y > 0) nbr_range<true>(word_base, run_pos, run_end, W, gx0, len, below - W, ja, jb); }
                else { if (y + 1 < uint32_t(ny)) nbr_range<true>(word_base, run_pos, run_end, W, gx0, len, below + W, ja, jb); }
            }
            const uint32_t most = __reduce_max_sync(FULL, jb - ja);
            for (uint32_t it = 0; it < most; ++it) {
                unsigned long long key = ~0ull;
                uint32_t a = 0, b = 0;
                if (it < jb - ja) {
                    a = r; b = ja + it;
                    uf_find2(parent, a, b);
                    if (a != b) {
                        if (a < b) { const uint32_t t = a; a = b; b = t; }
                        key = ((unsigned long long)a << 32) | b;
                    }
                }
                const unsigned peers = __match_any_sync(FULL, key);
                if (key != ~0ull && int(lane) == __ffs(peers) - 1) uf_union(parent, a, b);
            }
        }
    }
    ktrace_both(between_blocks ? KT_UZ2_END : KT_UZ1_END);
}

// ------------------------------------------------------------------------------------------------
// flatten, count, rank the roots (ITK-consecutive labels) -- one pass over the run table
// ------------------------------------------------------------------------------------------------
// Per run: parent[r] = root; the run's voxels are added to root_count[root] (lanes of a warp that share
// a root are combined by shuffles, then a per-CTA shared-memory cache, then global atomics -- one huge
// component would otherwise serialise every warp on one address).  Roots are ranked by a single-pass
// scan with decoupled look-back: run_label[root] = 1 + number of roots before it; roots are ordered by
// minimum linear index, so this is ITK's consecutive numbering.
constexpr int FR_THREADS = 256, FR_ITEMS = 4, FR_TILE = FR_THREADS * FR_ITEMS;

__global__ void __launch_bounds__(FR_THREADS) k_flatten_rank(uint32_t* parent, const uint32_t* __restrict__ run_pos,
                                                             const uint32_t* __restrict__ run_end,
                                                             volatile unsigned long long* state,
                                                             const DynArgs* __restrict__ dyn, uint32_t* __restrict__ run_label,
                                                             uint32_t* root_count, DevScalars* sc) {
    pdl_wait();
    ktrace(KT_RANK);
    __shared__ CtaCache<1, uint32_t, 64> cache;
    __shared__ uint32_t ws[9];
    __shared__ uint32_t s_tile, s_prefix;
    if (sc->status != MAMRI_OK) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_labels = 0;
        return;
    }
    const uint32_t n = sc->n_runs;
    const uint32_t n_tiles = (n + FR_TILE - 1) / FR_TILE;
    if (n_tiles == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_labels = 0;
        return;
    }
    cache.init();
    // the grid is sized for the run-table capacity; CTAs draw tiles until the table is exhausted
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&sc->ticket_rank, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t r0 = tile * FR_TILE + (threadIdx.x >> 5) * (32 * FR_ITEMS) + lane_id();
        // The FR_ITEMS root walks of a lane advance together (one dependent load per chain and round, all in flight at
        // once) instead of one after the other.  Plain walks, and only parent[r] is written: a walker that shortened
        // other nodes' paths on its way (path halving) could overwrite a parent[r'] that its owner has already set to the
        // root with a nearer ancestor, and the filter pass relies on parent[] holding roots.
        uint32_t x[FR_ITEMS], p[FR_ITEMS], len[FR_ITEMS];
#pragma unroll
        for (int k = 0; k < FR_ITEMS; ++k) {
            const uint32_t r = r0 + k * 32;                          // warp-contiguous: coalesced, one key per lane
            x[k] = r;
            p[k] = r < n ? parent[r] : r;                            // past the table: a fixed point, nothing to walk
            len[k] = r < n ? run_end[r] - run_pos[r] + 1u : 0u;
        }
        bool more;
        do {
            more = false;
#pragma unroll
            for (int k = 0; k < FR_ITEMS; ++k)
                if (p[k] != x[k]) { x[k] = p[k]; p[k] = parent[x[k]]; more = true; }
        } while (more);
        uint32_t roots = 0, is_root[FR_ITEMS];
#pragma unroll
        for (int k = 0; k < FR_ITEMS; ++k) {
            const uint32_t r = r0 + k * 32;
            is_root[k] = (r < n && x[k] == r) ? 1u : 0u;
            if (r < n && !is_root[k]) parent[r] = x[k];              // roots stay fixed points: concurrent walkers stay correct
            roots += is_root[k];
        }
        // rank: items are ordered (warp, k, lane) inside the tile, so scan per k-row
        uint32_t row_tot[FR_ITEMS], row_ex[FR_ITEMS];
#pragma unroll
        for (int k = 0; k < FR_ITEMS; ++k) {
            const uint32_t inc = warp_incl_scan(is_root[k]);
            row_ex[k] = inc - is_root[k];
            row_tot[k] = __shfl_sync(FULL, inc, 31);
        }
        uint32_t warp_total = 0;
#pragma unroll
        for (int k = 0; k < FR_ITEMS; ++k) { const uint32_t t = row_tot[k]; row_tot[k] = warp_total; warp_total += t; }
        const unsigned wid = threadIdx.x >> 5;
        if (lane_id() == 0) ws[wid] = warp_total;
        __syncthreads();
        uint32_t off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t t = ws[w];
            if (w < int(wid)) off += t;
            total += t;
        }
        // the tile's own total is published BEFORE the voxel counts are combined: the tiles behind this one find it
        // there when they look back, instead of waiting for this tile's aggregation as well
        if (threadIdx.x == 0) state[tile] = scan_pack(dyn->gen, total, tile == 0 ? 2u : 1u);
#pragma unroll
        for (int k = 0; k < FR_ITEMS; ++k) {
            const uint32_t r = r0 + k * 32;
            uint32_t v[1] = {len[k]};
            warp_agg_add(r < n ? x[k] : MAMRI_NONE, v, cache, root_count);
        }
        if (threadIdx.x < 32) {
            const uint32_t before = scan_lookback(state, tile, total, dyn->gen);
            if (threadIdx.x == 0) {
                s_prefix = before;
                if (tile == n_tiles - 1) sc->n_labels = before + total;
            }
        }
        __syncthreads();
        const uint32_t base = s_prefix + off;
#pragma unroll
        for (int k = 0; k < FR_ITEMS; ++k)
            if (is_root[k]) run_label[r0 + k * 32] = base + row_tot[k] + row_ex[k] + 1u;
    }
    cache.flush(root_count);
    ktrace_both(KT_RANK_END);
}


// ------------------------------------------------------------------------------------------------
// small run tables: union-find, ranking and the candidate filter in ONE launch by one thread-block cluster
// ------------------------------------------------------------------------------------------------
// A clinical scan has a few ten thousand x-runs; on such a table the five kernels above (per-slice union, two
// boundary-merge rounds, flatten + rank, filter) are five dependent launches of microseconds of work each.  Here one
// cluster of NC CTAs does all of it, with hardware cluster barriers (barrier.cluster, ~0.2 us) where the kernels
// had launch boundaries.  CTA c owns a chunk of consecutive slices holding about 1/NC of the runs:
//   U1  block-local union-find in shared memory over the chunk: joins with the row above AND with the slice below
//       when it belongs to the chunk; pointer-jumping flatten; parents published as global run ids
//   U2  global boundary merge: the runs of each chunk's first slice join the slice below (atomicMin on the roots;
//       at most NC - 1 boundaries, so root chains stay shorter than NC)
//   F   every run finds its root; voxel counts per root (warp shuffles -> CTA cache -> atomics); roots of the chunk
//       are ranked locally in raster order, the CTA totals are exchanged through DevScalars::cl_roots
//   FIX roots get their ITK-consecutive label = 1 + roots of the earlier chunks + local rank
//   S   the volume filter, marker slots, the body bid and the final label of every run (as k_select)
// The result is bit-identical to the multi-kernel path (run ids are raster-ordered, the smaller id wins every hook).
// Any run count is handled correctly (a chunk that does not fit shared memory works on the global array); the host
// picks this path when the previous scans had few enough runs for it to be the faster one.
constexpr int LC_THREADS = 1024;
constexpr uint32_t LC_SMEM_RUNS = 10240;             // 40 KB of parents per CTA

struct LabelArgs {
    const uint32_t* word_base;
    const uint32_t* run_pos;
    const uint32_t* run_end;
    uint32_t* parent;
    uint32_t* run_label;
    uint32_t* root_count;
    uint32_t* label_count;
    uint32_t* label_slot;
    uint32_t* cand_label;
    unsigned long long* sums;
    DevScalars* sc;
    int W, ny, nz;
    uint32_t max_markers;
    int fence;                          // explicit gpu-scope fence before each cluster barrier (MAMRI_CLUSTER_FENCE, default 1)
    double voxel_volume, min_volume, max_volume;
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// Barrier over the whole cluster; global-memory writes made before it are visible to every CTA after it.
__device__ __forceinline__ void cluster_barrier(int fence) {
    if (fence) __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

template <bool CONN26>
__global__ void __launch_bounds__(LC_THREADS, 1) k_label_cluster(LabelArgs a) {
    __shared__ uint32_t sp[LC_SMEM_RUNS];
    __shared__ CtaCache<1, uint32_t, 64> cache;
    __shared__ uint32_t s_zcnt[2];
    __shared__ uint32_t s_wsum[LC_THREADS / 32];
    __shared__ uint32_t s_carry;
    __shared__ unsigned long long s_bid, s_fg;
    pdl_wait();
    ktrace(KT_LABEL);
    DevScalars* sc = a.sc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t rank = cluster_rank(), NC = cluster_size();
    const uint32_t n = sc->status == MAMRI_OK ? sc->n_runs : 0u;     // uniform over the cluster (written by k_runs_scan)
    if (n == 0) return;                                               // n_labels, n_cand, body: all zero already
    const int W = a.W, ny = a.ny, nz = a.nz;
    const uint32_t slice_words = uint32_t(W) * uint32_t(ny);
    // ---- chunk of slices: [zlo, zhi) with zlo = #{z : first run of slice z < rank n / NC} (the slice bases are sorted)
    const uint32_t t_lo = uint32_t((unsigned long long)rank * n / NC), t_hi = uint32_t((unsigned long long)(rank + 1) * n / NC);
    if (tid < 2) s_zcnt[tid] = 0;
    if (tid == 0) { s_bid = 0ull; s_fg = 0ull; }
    __syncthreads();
    {
        uint32_t c_lo = 0, c_hi = 0;
        for (uint32_t z = tid; z < uint32_t(nz); z += blockDim.x) {
            const uint32_t b = a.word_base[z * slice_words];
            c_lo += b < t_lo;
            c_hi += b < t_hi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { c_lo += __shfl_xor_sync(FULL, c_lo, o); c_hi += __shfl_xor_sync(FULL, c_hi, o); }
        if (lane == 0) { if (c_lo) atomicAdd(&s_zcnt[0], c_lo); if (c_hi) atomicAdd(&s_zcnt[1], c_hi); }
    }
    __syncthreads();
    const uint32_t zlo = s_zcnt[0], zhi = rank + 1 == NC ? uint32_t(nz) : s_zcnt[1];
    const uint32_t r0 = zlo < uint32_t(nz) ? a.word_base[zlo * slice_words] : n;
    const uint32_t r1 = zhi < uint32_t(nz) ? a.word_base[zhi * slice_words] : n;
    const uint32_t nr = r1 - r0;                                      // runs of this chunk (ids r0 .. r1 - 1)

    // ---- U1: block-local union-find over the chunk, in two rounds so that no root chain grows longer than a slice is
    // tall or the chunk is deep: (a) every run joins the row above it, flatten; (b) joins with the slice below, flatten
    uint32_t* P = nr <= LC_SMEM_RUNS ? sp : a.parent + r0;
    for (uint32_t i = tid; i < nr; i += blockDim.x) P[i] = i;
    __syncthreads();
    for (int round = 0; round < 2; ++round) {
        for (uint32_t i = tid; i < nr; i += blockDim.x) {
            const uint32_t pos = a.run_pos[r0 + i];
            const uint32_t wi = pos >> 5, row = wi / uint32_t(W);
            const uint32_t z = row / uint32_t(ny), y = row - z * uint32_t(ny);
            const int gx0 = int((wi - row * uint32_t(W)) * 32 + (pos & 31u));
            const int len = int(a.run_end[r0 + i] - pos + 1u);
            if (round == 0) {
                if (y > 0) join_run<CONN26>(a.word_base, a.run_pos, a.run_end, P, r0, W, i, gx0, len, (row - 1) * uint32_t(W));
            } else if (z > zlo) {
                const uint32_t below = (row - uint32_t(ny)) * uint32_t(W);
                join_run<CONN26>(a.word_base, a.run_pos, a.run_end, P, r0, W, i, gx0, len, below);
                if (CONN26) {
                    if (y > 0) join_run<true>(a.word_base, a.run_pos, a.run_end, P, r0, W, i, gx0, len, below - uint32_t(W));
                    if (y + 1 < uint32_t(ny)) join_run<true>(a.word_base, a.run_pos, a.run_end, P, r0, W, i, gx0, len, below + uint32_t(W));
                }
            }
        }
        bool again = true;
        while (again) {                                               // pointer jumping: log2(depth) rounds
            __syncthreads();
            bool changed = false;
            for (uint32_t i = tid; i < nr; i += blockDim.x) {
                const uint32_t p = ld_vol(P + i), pp = ld_vol(P + p);
                if (pp != p) { P[i] = pp; changed = true; }
            }
            again = __syncthreads_or(changed);
        }
    }
    for (uint32_t i = tid; i < nr; i += blockDim.x) a.parent[r0 + i] = r0 + P[i];   // in place when P is the global array
    cluster_barrier(a.fence);
    ktrace(KT_L_U1);

    // ---- U2: the chunk's first slice joins the slice below it (owned by an earlier chunk)
    if (zlo > 0 && zlo < uint32_t(nz) && nr > 0) {
        const uint32_t e0 = r0;
        const uint32_t e1 = zlo + 1 < uint32_t(nz) ? a.word_base[(zlo + 1) * slice_words] : n;
        for (uint32_t r = e0 + tid; r < e1; r += blockDim.x) {
            const uint32_t pos = a.run_pos[r];
            const uint32_t wi = pos >> 5, row = wi / uint32_t(W);
            const uint32_t y = row - zlo * uint32_t(ny);
            const int gx0 = int((wi - row * uint32_t(W)) * 32 + (pos & 31u));
            const int len = int(a.run_end[r] - pos + 1u);
            const uint32_t below = (row - uint32_t(ny)) * uint32_t(W);
            join_run<CONN26>(a.word_base, a.run_pos, a.run_end, a.parent, 0u, W, r, gx0, len, below);
            if (CONN26) {
                if (y > 0) join_run<true>(a.word_base, a.run_pos, a.run_end, a.parent, 0u, W, r, gx0, len, below - uint32_t(W));
                if (y + 1 < uint32_t(ny)) join_run<true>(a.word_base, a.run_pos, a.run_end, a.parent, 0u, W, r, gx0, len, below + uint32_t(W));
            }
        }
    }
    cluster_barrier(a.fence);
    ktrace(KT_L_U2);

    // ---- F: roots, voxel counts per root, local rank of the chunk's roots (kept in run_label until FIX)
    cache.init();
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < nr; b += blockDim.x) {
        const uint32_t i = b + tid;
        uint32_t key = MAMRI_NONE, v[1] = {0u};
        bool is_root = false;
        if (i < nr) {
            const uint32_t r = r0 + i;
            uint32_t x = r, p = ld_vol(a.parent + x);
            while (p != x) { x = p; p = ld_vol(a.parent + x); }
            a.parent[r] = x;                                           // roots stay fixed points: concurrent walkers stay correct
            is_root = x == r;
            key = x;
            v[0] = a.run_end[r] - a.run_pos[r] + 1u;
        }
        warp_agg_add(key, v, cache, a.root_count);
        const unsigned bal = __ballot_sync(FULL, is_root);
        if (lane == 0) s_wsum[wid] = __popc(bal);
        __syncthreads();
        uint32_t off = 0, tot = 0;
        for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
            const uint32_t t = s_wsum[w];
            if (w < wid) off += t;
            tot += t;
        }
        const uint32_t carry = s_carry;
        if (is_root) a.run_label[r0 + i] = carry + off + __popc(bal & ((1u << lane) - 1u));
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
    }
    cache.flush(a.root_count);
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile unsigned int*>(&sc->cl_roots[rank]) = s_carry;
    cluster_barrier(a.fence);
    ktrace(KT_L_F);

    // ---- FIX: consecutive labels in raster order of the roots
    uint32_t base = 0;
    for (uint32_t c = 0; c < rank; ++c) base += ld_vol(&sc->cl_roots[c]);
    if (rank + 1 == NC && tid == 0) sc->n_labels = base + s_carry;
    for (uint32_t i = tid; i < nr; i += blockDim.x) {
        const uint32_t r = r0 + i;
        if (ld_vol(a.parent + r) == r) a.run_label[r] = ld_vol(a.run_label + r) + base + 1u;
    }
    cluster_barrier(a.fence);
    ktrace(KT_L_FIX);

    // ---- S: volume filter (Mamri.py:1310), body bid (:1320-1322), final label of every run
    {
        unsigned long long packed = 0ull, cnt64 = 0ull;
        for (uint32_t i = tid; i < nr; i += blockDim.x) {
            const uint32_t r = r0 + i;
            const uint32_t root = ld_vol(a.parent + r);
            if (root != r) {
                a.run_label[r] = ld_vol(a.run_label + root);
                const double vol = double(ld_vol(a.root_count + root)) * a.voxel_volume;   // as k_select
                a.label_slot[r] = (vol >= a.min_volume && vol <= a.max_volume) ? MAMRI_SLOT_OF_ROOT : MAMRI_NONE;
            } else {
                const uint32_t cnt = ld_vol(a.root_count + r), label = ld_vol(a.run_label + r);
                a.label_count[label - 1u] = cnt;
                cnt64 += cnt;
                const double vol = double(cnt) * a.voxel_volume;        // GetPhysicalSize
                uint32_t slot = MAMRI_NONE;
                if (vol >= a.min_volume && vol <= a.max_volume) {         // inclusive bounds
                    slot = atomicAdd(&sc->n_cand, 1u);
                    if (slot < a.max_markers) a.cand_label[slot] = label; else slot = MAMRI_NONE;
                } else {
                    const unsigned long long bid = ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - label);
                    packed = bid > packed ? bid : packed;
                }
                a.label_slot[r] = slot;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long p2 = __shfl_xor_sync(FULL, packed, o);
            packed = p2 > packed ? p2 : packed;
            cnt64 += __shfl_xor_sync(FULL, cnt64, o);
        }
        if (lane == 0) {
            if (packed) atomicMax(&s_bid, packed);
            if (cnt64) atomicAdd(&s_fg, cnt64);
        }
        __syncthreads();
        if (tid == 0) {
            if (s_bid) atomicMax(&sc->body_packed, s_bid);
            if (s_fg) atomicAdd(&sc->n_foreground, s_fg);
        }
    }
    cluster_barrier(a.fence);
    ktrace(KT_L_S);
    // ---- the first CTA clamps the candidate count, names the body's label in slot `max_markers` and zeroes the
    // moment sums of the slots in use (as the last CTA of k_select does)
    if (rank == 0) {
        uint32_t nc = *(volatile unsigned int*)&sc->n_cand;
        if (nc > a.max_markers) {
            nc = a.max_markers;
            if (tid == 0) sc->status = MAMRI_ERR_CAPACITY;
        }
        const unsigned long long bp = *(volatile unsigned long long*)&sc->body_packed;
        if (tid == 0 && (bp >> 32) != 0ull) a.cand_label[a.max_markers] = 0xFFFFFFFFu - uint32_t(bp & 0xFFFFFFFFull);
        for (uint32_t i = tid; i < nc * 9u; i += blockDim.x) a.sums[i] = 0ull;
        for (uint32_t i = tid; i < 9u; i += blockDim.x) a.sums[uint32_t(a.max_markers) * 9u + i] = 0ull;
    }
}

// Largest cluster the labelling kernel can be launched with on the current device (16 needs the non-portable opt-in),
// 0 = none.  Also sets the per-device function attributes; called from mamri_create under its device guard.
template <bool C26>
static int label_cluster_probe() {
    int best = 0;
    cudaFuncSetAttribute(k_label_cluster<C26>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int nc = 16; nc >= 2 && best == 0; nc >>= 1) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(nc); cfg.blockDim = dim3(LC_THREADS);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, k_label_cluster<C26>, &cfg) == cudaSuccess && n_clusters >= 1) best = nc;
    }
    cudaGetLastError();
    return best;
}

int ccl_init_device() {
    const int a = label_cluster_probe<false>(), b = label_cluster_probe<true>();
    return a < b ? a : b;
}

// Labelling of the closed mask, the volume filter and the body label: everything `materialise` needs.
// c->label_cluster > 0: one cluster of that many CTAs (small run tables); otherwise the scalable kernels.
cudaError_t launch_label(mamri_ctx* c, const uint32_t* d_mask, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s) {
    const int nx = desc->nx, ny = desc->ny, nz = desc->nz, connectivity = prm->connectivity;
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    // Phase B of the run scan spins on counts that phase A of OTHER CTAs publishes, so the whole grid has to be resident
    // at once: never more CTAs than the occupancy calculator says fit on the device next to each other.
    static const int scan_ctas = [] {
        const char* e = getenv("MAMRI_SCAN_CTAS");
        int want = e ? atoi(e) : 296, per_sm = 0, dev = 0, sms = 148;
        int a = 0, b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_runs_scan<16>, RS_THREADS, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_runs_scan<4>, RS_THREADS, 0) == cudaSuccess)
            per_sm = a < b ? a : b;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int cap = (per_sm > 0 ? per_sm : 1) * sms;
        return want < 1 ? 1 : (want > cap ? cap : want);
    }();
    if (n_words >= 148u * RS_THREADS * 16u) {
        const uint32_t n_tiles = (n_words + RS_THREADS * 16 - 1) / (RS_THREADS * 16);
        LK(k_runs_scan<16>, n_tiles < uint32_t(scan_ctas) ? n_tiles : uint32_t(scan_ctas), RS_THREADS, s, false, d_mask, W, n_words, c->d_scan_runs,
           c->d_dyn, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_root_count, c->max_runs, n_tiles, c->d_scalars);
    } else {
        const uint32_t n_tiles = (n_words + RS_THREADS * 4 - 1) / (RS_THREADS * 4);
        LK(k_runs_scan<4>, n_tiles < uint32_t(scan_ctas) ? n_tiles : uint32_t(scan_ctas), RS_THREADS, s, false, d_mask, W, n_words, c->d_scan_runs,
           c->d_dyn, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_root_count, c->max_runs, n_tiles, c->d_scalars);
    }
    prof_mark(c, s, "runs_scan");
    if (c->label_cluster > 0) {
        static const int fence = [] { const char* e = getenv("MAMRI_CLUSTER_FENCE"); return e ? atoi(e) : 1; }();
        LabelArgs a;
        a.word_base = c->d_word_base; a.run_pos = c->d_run_pos; a.run_end = c->d_run_end;
        a.parent = c->d_parent; a.run_label = c->d_run_label; a.root_count = c->d_root_count; a.label_count = c->d_label_count;
        a.label_slot = c->d_label_slot; a.cand_label = c->d_cand_label; a.sums = c->d_cand_sums; a.sc = c->d_scalars;
        a.W = W; a.ny = ny; a.nz = nz; a.max_markers = c->max_markers; a.fence = fence;
        double vv = 1.0;
        for (int i = 0; i < 3; ++i) vv *= desc->spacing[i];       // ITK: sizePerPixel *= spacing[i]
        a.voxel_volume = vv; a.min_volume = prm->min_volume; a.max_volume = prm->max_volume;
        const unsigned nc = unsigned(c->label_cluster);
        cudaError_t e = connectivity == 26
            ? launch_kc(k_label_cluster<true>, dim3(nc), dim3(LC_THREADS), 0, nc, s, false, a)
            : launch_kc(k_label_cluster<false>, dim3(nc), dim3(LC_THREADS), 0, nc, s, false, a);
        if (e != cudaSuccess) return e;
        prof_mark(c, s, "label_cluster");
        return cudaGetLastError();
    }
    int radix = 1;
    while (radix * radix < nz) radix <<= 1;
    const int RG = c->run_ctas > 0 ? c->run_ctas : MAMRI_RUN_CTAS;
    // threads per slice CTA: about one per run of an average slice of the scans just processed (slice_threads_class)
    const int slice_threads = c->slice_threads > 0 ? c->slice_threads : (c->slice_threads < 0 ? -c->slice_threads : SLICE_THREADS);
    if (connectivity == 26) {
        LK(k_union_slices<true>, nz, slice_threads, s, false, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_parent, W, ny, nz, c->d_scalars);
        prof_mark(c, s, "union_slices");
        if (nz > 1) {
            LK(k_union_z<true>, RG, 256, s, false, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_parent, W, ny, radix, 0, c->d_scalars);
            prof_mark(c, s, "union_z_within_blocks");
        }
        if (nz > radix) {
            LK(k_union_z<true>, RG, 256, s, false, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_parent, W, ny, radix, 1, c->d_scalars);
            prof_mark(c, s, "union_z_between_blocks");
        }
    } else {
        LK(k_union_slices<false>, nz, slice_threads, s, false, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_parent, W, ny, nz, c->d_scalars);
        prof_mark(c, s, "union_slices");
        if (nz > 1) {
            LK(k_union_z<false>, RG, 256, s, false, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_parent, W, ny, radix, 0, c->d_scalars);
            prof_mark(c, s, "union_z_within_blocks");
        }
        if (nz > radix) {
            LK(k_union_z<false>, RG, 256, s, false, c->d_word_base, c->d_run_pos, c->d_run_end, c->d_parent, W, ny, radix, 1, c->d_scalars);
            prof_mark(c, s, "union_z_between_blocks");
        }
    }
    LK(k_flatten_rank, RG, FR_THREADS, s, false, c->d_parent, c->d_run_pos, c->d_run_end, c->d_scan_rank, c->d_dyn, c->d_run_label,
       c->d_root_count, c->d_scalars);
    prof_mark(c, s, "flatten_rank");
    return launch_select(c, desc, prm, s);
}

// ------------------------------------------------------------------------------------------------
// materialise: closed mask (u8), label volume (u32), body mask (u8) -- the per-voxel outputs
// ------------------------------------------------------------------------------------------------
// Aligned path (nx % 32 == 0): a warp writes 4 words = 128 voxels per iteration; lane l owns word
// l/8, voxels 4*(l%8)..+3 -> one 16-byte label store and one 4-byte mask store per lane, i.e.
// 512 B / 128 B contiguous per warp instruction.
__device__ __forceinline__ void st_stream(uint4* p, uint4 v, bool hint) {
    if (hint) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else *p = v;
}
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v, bool hint) {
    if (hint) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else *p = v;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_materialise(const uint32_t* __restrict__ mask,
                                                     const uint32_t* __restrict__ word_base,
                                                     const uint32_t* __restrict__ run_label, int nx, int W,
                                                     uint32_t n_words, const DynArgs* __restrict__ dyn,
                                                     const DevScalars* sc, int stream_hint) {
    pdl_wait();
    ktrace(KT_MAT);
    const bool hint = stream_hint != 0;
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
            const bool any_fg = __any_sync(0xFFFFFFFFu, m_l != 0u);
            if (!any_fg) {
                // 1024 background voxels (most of an MRI volume is air): nothing to resolve, just the zero stores
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t wi = w0 + k * 4 + (lane >> 3);
                    if (wi >= n_words) continue;
                    const uint32_t v = wi * 32 + (lane & 7u) * 4;
                    if (labels_out) st_stream(reinterpret_cast<uint4*>(labels_out + v), make_uint4(0u, 0u, 0u, 0u), hint);
                    if (mask_out) st_stream(reinterpret_cast<uint32_t*>(mask_out + v), 0u, hint);
                    if (body_out) st_stream(reinterpret_cast<uint32_t*>(body_out + v), 0u, hint);
                }
                continue;
            }
            const bool resolve = need_labels && ok;
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
                // streaming stores: the volumes are written once and not read back by this pipeline, so they
                // should not push the bit-packed intermediates of the scans in flight out of L2
                if (labels_out) st_stream(reinterpret_cast<uint4*>(labels_out + v), make_uint4(lab[0], lab[1], lab[2], lab[3]), hint);
                if (mask_out) {
                    uint32_t mb = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                    st_stream(reinterpret_cast<uint32_t*>(mask_out + v), mb, hint);
                }
                if (body_out) {
                    uint32_t bb = 0;
                    if (has_body) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) bb |= (lab[q] == body ? 1u : 0u) << (8 * q);
                    }
                    st_stream(reinterpret_cast<uint32_t*>(body_out + v), bb, hint);
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
    }    ktrace_last(KT_END);
}

cudaError_t launch_materialise(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int outs_aligned,
                               cudaStream_t s) {
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    const bool aligned = (nx % 32 == 0) && outs_aligned;
    static const int stream_hint = [] { const char* e = getenv("MAMRI_STREAM_HINTS"); return e ? atoi(e) : 1; }();
    if (aligned) {
        // one trip of 32 mask words per warp by default: short-lived CTAs measured faster than a grid-stride loop
        uint32_t blocks = (n_words / 32 + 7) / 8;
        static const int per_sm = [] { const char* e = getenv("MAMRI_MAT_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
        if (per_sm > 0 && blocks > uint32_t(148 * per_sm)) blocks = uint32_t(148 * per_sm);
        if (blocks == 0) blocks = 1;
        LK(k_materialise<true>, blocks, 256, s, true, d_mask, c->d_word_base, c->d_run_label, nx, W, n_words, c->d_dyn, c->d_scalars, stream_hint);
    } else {
        uint32_t blocks = (n_words + 7) / 8;
        if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
        if (blocks == 0) blocks = 1;
        LK(k_materialise<false>, blocks, 256, s, true, d_mask, c->d_word_base, c->d_run_label, nx, W, n_words, c->d_dyn, c->d_scalars, stream_hint);
    }
    prof_mark(c, s, "materialise");
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// body labelmap at 1 bit per voxel (DynArgs::body_bits_out): what a host caller gets back instead of the uint8
// labelmap when the link to the host is the bottleneck (8x fewer bytes; Mamri.py:1323 `largest_object_img`)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_body_bits(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ word_base,
                                                   const uint32_t* __restrict__ run_label, int W, uint32_t n_words,
                                                   const DynArgs* __restrict__ dyn, const DevScalars* sc) {
    pdl_wait();
    uint32_t* __restrict__ bits = dyn->body_bits_out;
    if (!bits) return;
    const unsigned long long bp = sc->body_packed;
    const uint32_t body = (sc->status == MAMRI_OK && (bp >> 32) != 0ull) ? 0xFFFFFFFFu - uint32_t(bp & 0xFFFFFFFFull) : 0u;
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
        const uint32_t m = mask[wi];
        uint32_t out = 0;
        if (m && body) {
            const uint32_t prev = (wi % uint32_t(W)) ? mask[wi - 1] : 0u;
            const uint32_t starts = run_starts(m, prev), base = word_base[wi];
            uint32_t rem = m;
            while (rem) {                                           // one piece of set bits = part of one run
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

cudaError_t launch_body_bits(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, cudaStream_t s) {
    const int W = (nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * ny * nz;
    uint32_t blocks = (n_words + 255u) / 256u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    if (blocks == 0) blocks = 1;
    LK(k_body_bits, blocks, 256, s, false, d_mask, c->d_word_base, c->d_run_label, W, n_words, c->d_dyn, c->d_scalars);
    prof_mark(c, s, "body_bits");
    return cudaGetLastError();
}

KTRACE_TU(ccl)
