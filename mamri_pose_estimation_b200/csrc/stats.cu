// Stages 3 and 4a: label-shape statistics, the candidate-marker volume filter and the body label.
// Replaces sitk.LabelShapeStatisticsImageFilter().Execute + GetLabels/GetPhysicalSize/GetCentroid and
// the list comprehension / max() at Mamri/Mamri.py:1309-1310, 1316-1322.
//
// Statistics are accumulated as exact integers per x-RUN (one thread per entry of the run table), using
// closed forms for sum x and sum x^2 over the run, so the work is O(#runs) instead of O(#voxels).
// Lanes of a warp that hit the same label are combined with warp shuffles before one atomic per
// (warp, label): integer sums are order-independent -> bit-reproducible.  Two phases bound
// the table on noisy scans: voxel counts for every label first, then second moments only for the
// labels that pass the volume filter (plus the body).  Finalisation (centroid, physical size,
// principal moments/axes) is float64 on the device.
#include "common.cuh"

struct GeomArgs {
    double spacing[3], origin[3], dir[9];
    double voxel_volume;     // spacing[0]*spacing[1]*spacing[2], multiplied in that order on the host
    double min_volume, max_volume;
};

// ------------------------------------------------------------------------------------------------
// stage 4a: volume filter (Mamri.py:1310) and body label (Mamri.py:1320-1322)
// ------------------------------------------------------------------------------------------------
// One thread per run.  Roots carry their component's voxel count (k_flatten_rank) and label: they apply
// the volume filter, claim a marker slot and bid for the body; every other run copies its root's label.
__global__ void __launch_bounds__(256) k_select(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ root_count,
                                                uint32_t* run_label, uint32_t* __restrict__ label_count,
                                                uint32_t* __restrict__ label_slot, uint32_t* __restrict__ cand_label,
                                                unsigned long long* __restrict__ sums, uint32_t max_markers, GeomArgs g,
                                                DevScalars* sc) {
    pdl_wait();
    ktrace(KT_SELECT);
    const unsigned lane = lane_id();
    const uint32_t n = sc->status == MAMRI_OK ? sc->n_runs : 0u;
    const uint32_t warp0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5;
    const uint32_t stride = gridDim.x * blockDim.x;
    // per-thread body bid and foreground count over all of the thread's runs: one pair of atomics per warp at the
    // end (a noisy scan has millions of one-voxel components, every one of them a root that bids)
    unsigned long long packed = 0ull, cnt64 = 0ull;
    // SEL_U runs per thread and trip: their parents, then their roots' entries, travel together (a noisy table is ten
    // trips per thread, each one two dependent loads deep)
    constexpr int SEL_U = 4;
    for (uint32_t r0 = warp0; r0 < n; r0 += stride * SEL_U) {
        uint32_t root[SEL_U], cnt[SEL_U], lab[SEL_U];
#pragma unroll
        for (int u = 0; u < SEL_U; ++u) {
            const uint32_t r = r0 + u * stride + lane;
            root[u] = r < n ? parent[r] : MAMRI_NONE;
        }
#pragma unroll
        for (int u = 0; u < SEL_U; ++u)
            if (root[u] != MAMRI_NONE) { cnt[u] = root_count[root[u]]; lab[u] = run_label[root[u]]; }
#pragma unroll
        for (int u = 0; u < SEL_U; ++u) {
            const uint32_t r = r0 + u * stride + lane;
            if (root[u] == MAMRI_NONE) continue;
            const double vol = double(cnt[u]) * g.voxel_volume;           // GetPhysicalSize
            const bool keep = vol >= g.min_volume && vol <= g.max_volume; // inclusive bounds
            if (root[u] != r) {
                run_label[r] = lab[u];                                    // roots were ranked by k_flatten_rank
                // the filter's verdict on the run's component (the count on the root is final): the statistics kernel
                // then knows from the run's own entry whether it has anything to fetch from the root
                label_slot[r] = keep ? MAMRI_SLOT_OF_ROOT : MAMRI_NONE;
            } else {
                const uint32_t label = lab[u];
                label_count[label - 1u] = cnt[u];
                cnt64 += cnt[u];
                uint32_t slot = MAMRI_NONE;
                if (keep) {
                    slot = atomicAdd(&sc->n_cand, 1u);
                    if (slot < max_markers) cand_label[slot] = label; else slot = MAMRI_NONE;
                } else {
                    // max(..., key=GetPhysicalSize) returns the FIRST maximum -> lowest label on ties
                    const unsigned long long bid = ((unsigned long long)cnt[u] << 32) | (unsigned long long)(0xFFFFFFFFu - label);
                    packed = bid > packed ? bid : packed;
                }
                label_slot[r] = slot;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long p2 = __shfl_xor_sync(FULL, packed, o);
        packed = p2 > packed ? p2 : packed;
        cnt64 += __shfl_xor_sync(FULL, cnt64, o);
    }
    if (lane == 0) {
        if (packed) atomicMax(&sc->body_packed, packed);
        if (cnt64) atomicAdd(&sc->n_foreground, cnt64);
    }
    ktrace_both(KT_SEL_END);
    // The last CTA to finish clamps the candidate count, names the body's label in slot `max_markers`
    // and zeroes the moment sums of the slots in use.
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&sc->done_select, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    uint32_t nc = *(volatile unsigned int*)&sc->n_cand;
    if (nc > max_markers) {
        nc = max_markers;
        if (threadIdx.x == 0) sc->status = MAMRI_ERR_CAPACITY;
    }
    const unsigned long long bp = *(volatile unsigned long long*)&sc->body_packed;
    if (threadIdx.x == 0 && (bp >> 32) != 0ull) cand_label[max_markers] = 0xFFFFFFFFu - uint32_t(bp & 0xFFFFFFFFull);
    for (uint32_t i = threadIdx.x; i < nc * 9u; i += blockDim.x) sums[i] = 0ull;
    for (uint32_t i = threadIdx.x; i < 9u; i += blockDim.x) sums[uint32_t(max_markers) * 9u + i] = 0ull;
}

// ------------------------------------------------------------------------------------------------
// phase 2: first and second moments of the kept labels (+ body)
// ------------------------------------------------------------------------------------------------
// sum_{i=0..k} i^2 = k(k+1)(2k+1)/6, k >= -1.  One of k, k+1 is even and one of k, k+1, 2k+1 is a multiple of 3, so
// the divisions are done on the factors first: the product then never exceeds the result (exact in 64 bits for any
// row length a 2^32-voxel volume can have).
__device__ __forceinline__ unsigned long long sum_sq_upto(long long k) {
    if (k <= 0) return 0ull;
    unsigned long long a = (unsigned long long)k, b = a + 1ull, c = 2ull * a + 1ull;
    if (a % 2ull == 0ull) a /= 2ull; else b /= 2ull;
    if (a % 3ull == 0ull) a /= 3ull; else if (b % 3ull == 0ull) b /= 3ull; else c /= 3ull;
    return a * b * c;
}

// ------------------------------------------------------------------------------------------------
// finalisation (float64): ShapeLabelMapFilter's attributes from the exact sums
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_eigen3(double a[3][3], double w[3], double v[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {           // A <- A J
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {           // A <- J^T A
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {           // V <- V J
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];
    for (int i = 0; i < 2; ++i)                          // ascending eigenvalues
        for (int j = 0; j < 2 - i; ++j)
            if (w[j] > w[j + 1]) {
                double t = w[j]; w[j] = w[j + 1]; w[j + 1] = t;
                for (int k = 0; k < 3; ++k) { double u = v[k][j]; v[k][j] = v[k][j + 1]; v[k][j + 1] = u; }
            }
}

__device__ void make_marker(mamri_marker& m, uint32_t label, unsigned long long count, const unsigned long long* s9,
                            const GeomArgs& g) {
    m.label = label;
    m.reserved = 0;
    m.count = count;
    for (int i = 0; i < 3; ++i) m.sum_idx[i] = s9[i];
    for (int i = 0; i < 6; ++i) m.sum_mom[i] = s9[3 + i];
    const double n = double(count);
    m.volume_mm3 = n * g.voxel_volume;
    double c[3];
    for (int i = 0; i < 3; ++i) { c[i] = double(s9[i]) / n; m.centroid_index[i] = c[i]; }
    // A = direction * diag(spacing); physical = origin + A c
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = g.dir[3 * i + j] * g.spacing[j];
    for (int i = 0; i < 3; ++i) m.centroid_lps[i] = g.origin[i] + (A[i][0] * c[0] + A[i][1] * c[1] + A[i][2] * c[2]);
    m.centroid_ras[0] = -m.centroid_lps[0];
    m.centroid_ras[1] = -m.centroid_lps[1];
    m.centroid_ras[2] = m.centroid_lps[2];
    // index-space covariance, then M = A cov A^T + diag(spacing^2 / 12)
    double cov[3][3];
    cov[0][0] = double(s9[3]) / n - c[0] * c[0];
    cov[1][1] = double(s9[4]) / n - c[1] * c[1];
    cov[2][2] = double(s9[5]) / n - c[2] * c[2];
    cov[0][1] = cov[1][0] = double(s9[6]) / n - c[0] * c[1];
    cov[0][2] = cov[2][0] = double(s9[7]) / n - c[0] * c[2];
    cov[1][2] = cov[2][1] = double(s9[8]) / n - c[1] * c[2];
    double t[3][3], M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[i][j] = A[i][0] * cov[0][j] + A[i][1] * cov[1][j] + A[i][2] * cov[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = t[i][0] * A[j][0] + t[i][1] * A[j][1] + t[i][2] * A[j][2];
    for (int i = 0; i < 3; ++i) M[i][i] += g.spacing[i] * g.spacing[i] / 12.0;
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j) M[i][j] = M[j][i] = 0.5 * (M[i][j] + M[j][i]);
    double w[3], v[3][3];
    jacobi_eigen3(M, w, v);
    double ax[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ax[i][j] = v[j][i];              // rows = eigenvectors (V^T)
    double det = ax[0][0] * (ax[1][1] * ax[2][2] - ax[1][2] * ax[2][1]) - ax[0][1] * (ax[1][0] * ax[2][2] - ax[1][2] * ax[2][0]) +
                 ax[0][2] * (ax[1][0] * ax[2][1] - ax[1][1] * ax[2][0]);
    for (int j = 0; j < 3; ++j) ax[2][j] *= det;                      // proper rotation
    for (int i = 0; i < 3; ++i) {
        m.principal_moments[i] = w[i];
        for (int j = 0; j < 3; ++j) m.principal_axes[3 * i + j] = ax[i][j];
    }
}

// Exact integer sums of one x-run [pos .. end] (linear bit positions in the [nz][ny][W*32] mask): x, y, z, xx, yy, zz, xy, xz, yz.
__device__ __forceinline__ void run_sums(uint32_t pos, uint32_t end, int W, int ny, unsigned long long (&v)[9]) {
    const uint32_t wi = pos >> 5, row = wi / W;
    const long long z = row / ny, y = row - uint32_t(z) * ny;
    const long long xs = (long long)(wi - row * W) * 32 + (pos & 31u);
    const long long len = (long long)(end - pos) + 1, xe = xs + len - 1;
    const unsigned long long sx = (unsigned long long)((xs + xe) * len / 2);
    v[0] = sx;                                              // sum x
    v[1] = (unsigned long long)(len * y);                   // sum y
    v[2] = (unsigned long long)(len * z);                   // sum z
    v[3] = sum_sq_upto(xe) - sum_sq_upto(xs - 1);           // sum xx
    v[4] = (unsigned long long)(len * y * y);               // sum yy
    v[5] = (unsigned long long)(len * z * z);               // sum zz
    v[6] = sx * (unsigned long long)y;                      // sum xy
    v[7] = sx * (unsigned long long)z;                      // sum xz
    v[8] = (unsigned long long)(len * y * z);               // sum yz
}

// Turns the sums into the marker table and the summary.  One kernel does three things:
//   ranks    every kept label's place in GetLabels order = number of kept labels below it (slots are claimed by
//            atomics in arbitrary order); spread over the whole grid, it only needs the slot table
//   moments  exact integer sums per run for the kept labels + the body, as described above
//   finalise by the LAST CTA to finish its sums (ticket in DevScalars::done_stats): one thread per kept label turns
//            the sums into a mamri_marker (float64 centroid, physical size, Jacobi eigen-decomposition), another
//            warp's thread does the body, thread 0 the summary scalars.  `markers` and `summary` are the context's
//            PINNED HOST buffers: the records go straight over PCIe as posted writes and are visible to the host when
//            the stream has drained, so no copy nodes follow the kernel.  Also fills the scan's fixed-size table
//            (DynArgs::table_out, rows of {label, count, volume_mm3, RAS x y z, n_labels, body_label}: the layout
//            distributed.pack_table builds on the host) when the caller asked for it.
// It runs beside `materialise` on the second branch of the graph.
// MODE 0: all three in one launch (the usual scan: the kernel hides beside `materialise`).  MODE 1: ranks + moments only,
// compiled for 4 CTAs per SM (<= 64 registers), MODE 2: the finalisation alone, one CTA, launched behind MODE 1 -- for
// run tables of millions of entries, where the one-launch form kept half of every SM's register file (the float64
// eigen-decomposition needs 124 registers per thread) for 0.3 ms and `materialise` beside it ran at half occupancy.
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 1 ? 4 : 1) k_stats(const uint32_t* __restrict__ run_pos, const uint32_t* __restrict__ run_end,
                                               const uint32_t* __restrict__ parent, const uint32_t* __restrict__ run_label,
                                               const uint32_t* __restrict__ label_slot, int W, int ny,
                                               unsigned long long* sums, const uint32_t* __restrict__ cand_label,
                                               uint32_t* cand_rank, const uint32_t* __restrict__ label_count,
                                               uint32_t max_markers, GeomArgs g, mamri_marker* __restrict__ markers,
                                               mamri_summary* summary, DevScalars* sc, const DynArgs* __restrict__ dyn) {
    __shared__ CtaCache<9, unsigned long long, 16> cache;
    __shared__ bool last;
    pdl_wait();
    if (MODE != 2) ktrace(KT_STATS);
    const bool ok = sc->status == MAMRI_OK;
    const uint32_t n = ok ? sc->n_runs : 0u;
    const uint32_t n_all = sc->n_cand;
    const uint32_t n_kept = ok ? (n_all < max_markers ? n_all : max_markers) : 0u;
    const unsigned long long bp = sc->body_packed;
    const uint32_t body = (ok && (bp >> 32) != 0ull) ? 0xFFFFFFFFu - uint32_t(bp & 0xFFFFFFFFull) : 0u;
    if (MODE != 2) {
        cache.init();
        const uint32_t stride = gridDim.x * blockDim.x;
        // ---- ranks of the kept labels
        // one warp per kept label, the lanes share the comparisons (coalesced loads of the slot table): with one thread
        // per label the first few CTAs walked the whole table serially, n_kept dependent iterations on the critical path
        for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_kept; i += stride >> 5) {
            const uint32_t lab = cand_label[i];
            uint32_t below = 0;
            for (uint32_t j = lane_id(); j < n_kept; j += 32) below += cand_label[j] < lab ? 1u : 0u;
            below = __reduce_add_sync(FULL, below);
            if (lane_id() == 0) cand_rank[i] = below;
        }
        // ---- moments: a warp takes ST_U x 32 consecutive runs per trip; whether a run counts is in its own entries
        // (slot, label: two coalesced loads), only the runs of kept labels go to their root for the slot
        // The body's runs (most of a clinical scan's table, every fifth warp trip of a noisy one) are summed in registers
        // over the thread's whole loop and combined once at the end; only the markers' runs go through the warp-level
        // combine for every trip.
        constexpr int ST_U = MODE == 1 ? 8 : 4;        // noisy table: latency-bound on 3 dependent loads per trip, so twice the runs in flight
        unsigned long long bsum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t r0 = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) * ST_U; r0 < n; r0 += stride * ST_U) {
            uint32_t key[ST_U], lab[ST_U];
            if constexpr (MODE == 1) {
                // the parents travel with the slots (coalesced, mostly unused) instead of after them: one dependent load less
                uint32_t par[ST_U];
#pragma unroll
                for (int u = 0; u < ST_U; ++u) {
                    const uint32_t r = r0 + u * 32 + lane_id();
                    key[u] = r < n ? label_slot[r] : MAMRI_NONE;
                    lab[u] = r < n ? run_label[r] : 0u;
                    par[u] = r < n ? parent[r] : 0u;
                }
#pragma unroll
                for (int u = 0; u < ST_U; ++u)
                    if (key[u] == MAMRI_SLOT_OF_ROOT) key[u] = label_slot[par[u]];
            } else {
#pragma unroll
                for (int u = 0; u < ST_U; ++u) {
                    const uint32_t r = r0 + u * 32 + lane_id();
                    key[u] = r < n ? label_slot[r] : MAMRI_NONE;      // a root's slot, SLOT_OF_ROOT for the other runs of a kept label
                    lab[u] = r < n ? run_label[r] : 0u;
                }
#pragma unroll
                for (int u = 0; u < ST_U; ++u)
                    if (key[u] == MAMRI_SLOT_OF_ROOT) key[u] = label_slot[parent[r0 + u * 32 + lane_id()]];
            }
#pragma unroll
            for (int u = 0; u < ST_U; ++u)
                if (key[u] == MAMRI_NONE && body != 0u && lab[u] == body) key[u] = max_markers;   // the body has the extra slot
            if constexpr (MODE == 1) {
                // Noisy table (millions of runs, a thousand kept labels whose runs sit between the specks' runs): the runs
                // that count are ~10 % of the table, a few per 32 consecutive runs and of several labels.  Combining them by
                // shuffles (9 values x 5 rounds per distinct label) cost ~1800 instructions per trip, and the sums of a
                // 32-run row were computed with three or four lanes active.  So: the warp compacts the runs that count of
                // its ST_U x 32 into shared memory (ballot + popc), one lane per run computes the sums (one dense pass),
                // the body's go to the thread's registers and the markers' straight to the table (fire-and-forget
                // reductions in L2, spread over a thousand rows).
                __shared__ uint32_t st_k[8][ST_U * 32], st_r[8][ST_U * 32];
                const unsigned wid = threadIdx.x >> 5, lane = lane_id();
                uint32_t cnt = 0;
#pragma unroll
                for (int u = 0; u < ST_U; ++u) {
                    const bool take = key[u] != MAMRI_NONE;
                    const unsigned bal = __ballot_sync(FULL, take);
                    if (take) {
                        const uint32_t j = cnt + __popc(bal & ((1u << lane) - 1u));
                        st_k[wid][j] = key[u];
                        st_r[wid][j] = r0 + u * 32 + lane;
                    }
                    cnt += __popc(bal);
                }
                __syncwarp();
                for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {                  // cnt is the same in every lane
                    const uint32_t j = j0 + lane;
                    uint32_t k = MAMRI_NONE;
                    unsigned long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    if (j < cnt) {
                        k = st_k[wid][j];
                        const uint32_t r = st_r[wid][j];
                        run_sums(run_pos[r], run_end[r], W, ny, v);
                        if (k == max_markers) {
#pragma unroll
                            for (int i = 0; i < 9; ++i) bsum[i] += v[i];
                            k = MAMRI_NONE;
                        }
                    }
                    // neighbours in the compacted list are mostly runs of ONE marker (its run in consecutive rows): the lanes
                    // of a label are summed by shuffles first.  One reduction per run instead made the few hundred runs of a
                    // big marker queue on the same L2 line, ~10 ns each: 40 us for the CTAs that held the biggest ones
                    warp_agg_add(k, v, cache, sums);
                }
                __syncwarp();
            } else {
                // the extents of the runs that count, all ST_U loads in flight before the first is used
                uint32_t rpos[ST_U], rend[ST_U];
#pragma unroll
                for (int u = 0; u < ST_U; ++u) {
                    const uint32_t r = r0 + u * 32 + lane_id();
                    rpos[u] = rend[u] = 0u;
                    if (key[u] != MAMRI_NONE) { rpos[u] = run_pos[r]; rend[u] = run_end[r]; }
                }
#pragma unroll
                for (int u = 0; u < ST_U; ++u) {
                    uint32_t k = key[u];
                    if (!__any_sync(FULL, k != MAMRI_NONE)) continue;
                    unsigned long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    if (k != MAMRI_NONE) {
                        run_sums(rpos[u], rend[u], W, ny, v);
                        if (k == max_markers) {
#pragma unroll
                            for (int i = 0; i < 9; ++i) bsum[i] += v[i];
                            k = MAMRI_NONE;
                        }
                    }
                    warp_agg_add(k, v, cache, sums);
                }
            }
        }
        warp_agg_add(bsum[4] | bsum[5] | bsum[0] | bsum[1] | bsum[2] ? max_markers : MAMRI_NONE, bsum, cache, sums);
        cache.flush(sums);
    }
    if (MODE == 1) return;
    // ---- the last CTA to get here finalises (MODE 2: the only CTA, behind the kernel that made the sums)
    if (MODE == 0) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(&sc->done_stats, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!last) return;
        __threadfence();
    }
    ktrace(KT_STATS_FIN);
    double* __restrict__ table = dyn->table_out;
    const uint32_t slots = table ? dyn->table_slots : 0u;
    const double body_d = double(body);
    const uint32_t n_labels = sc->n_labels;
    // MODE 2 may be launched with several CTAs: the markers are dealt over all of them, CTA 0 does the rest
    const uint32_t m_first = MODE == 2 ? blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
    const uint32_t m_step = MODE == 2 ? gridDim.x * blockDim.x : blockDim.x;
    for (uint32_t i = m_first; i < n_kept; i += m_step) {
        const uint32_t lab = __ldcg(cand_label + i), rank = __ldcg(cand_rank + i);
        unsigned long long s9[9];
        for (int k = 0; k < 9; ++k) s9[k] = __ldcg(sums + i * 9u + k);
        mamri_marker m;
        make_marker(m, lab, __ldcg(label_count + lab - 1u), s9, g);
        markers[rank] = m;
        if (rank < slots) {
            double* row = table + size_t(rank) * 8;
            row[0] = double(m.label); row[1] = double(m.count); row[2] = m.volume_mm3;
            row[3] = m.centroid_ras[0]; row[4] = m.centroid_ras[1]; row[5] = m.centroid_ras[2];
            row[6] = double(n_labels); row[7] = body_d;
        }
    }
    if (MODE == 2 && blockIdx.x != 0) return;
    for (uint32_t i = n_kept * 8 + threadIdx.x; i < slots * 8; i += blockDim.x) table[i] = 0.0;   // unused rows
    if (threadIdx.x == blockDim.x - 1) {                     // the body: a thread of another warp than marker 0's
        mamri_marker m;
        if (body != 0u) {
            unsigned long long s9[9];
            for (int k = 0; k < 9; ++k) s9[k] = __ldcg(sums + max_markers * 9u + k);
            make_marker(m, body, bp >> 32, s9, g);
        } else {
            memset(&m, 0, sizeof(m));
        }
        summary->body = m;
    }
    if (threadIdx.x == 0) {
        summary->n_labels = n_labels;
        summary->n_runs = sc->n_runs;
        summary->n_markers = n_all;
        summary->n_foreground = sc->n_foreground;
        summary->device_status = sc->status;
        summary->reserved = 0;
        summary->body_label = body;
        summary->body_count = body != 0u ? (bp >> 32) : 0ull;
    }    ktrace_last(KT_FINAL);
}

static GeomArgs geom_args(const mamri_volume_desc* desc, const mamri_params* prm) {
    GeomArgs g;
    for (int i = 0; i < 3; ++i) { g.spacing[i] = desc->spacing[i]; g.origin[i] = desc->origin[i]; }
    for (int i = 0; i < 9; ++i) g.dir[i] = desc->direction[i];
    double vv = 1.0;
    for (int i = 0; i < 3; ++i) vv *= desc->spacing[i];       // ITK: sizePerPixel *= spacing[i]
    g.voxel_volume = vv;
    g.min_volume = prm->min_volume;
    g.max_volume = prm->max_volume;
    return g;
}

// Volume filter + body label + final label of every run: everything `materialise` needs.
cudaError_t launch_select(mamri_ctx* c, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s) {
    const GeomArgs g = geom_args(desc, prm);
    LK(k_select, c->run_ctas > 0 ? c->run_ctas : MAMRI_RUN_CTAS, 256, s, false, c->d_parent, c->d_root_count, c->d_run_label, c->d_label_count, c->d_label_slot,
       c->d_cand_label, c->d_cand_sums, c->max_markers, g, c->d_scalars);
    prof_mark(c, s, "select");
    return cudaGetLastError();
}

// Moments of the kept labels + body, and the marker table / summary (independent of `materialise`).
// Usual scan: ONE launch (k_stats<0>) on the graph's side branch, hidden beside `materialise`.  Run tables of >= 150 k
// entries (noisy scans): the sums are made by a full grid BEFORE `materialise` (launch_stats_early, ~0.03 ms alone) and
// only the finalisation runs beside it -- beside `materialise` the sums' loads queue behind its stores and the kernel
// only finishes when `materialise` does (measured on config C4), while its CTAs cost `materialise` its occupancy.
#define MAMRI_STATS_ARGS c->d_run_pos, c->d_run_end, c->d_parent, c->d_run_label, c->d_label_slot, W, desc->ny, c->d_cand_sums, \
                         c->d_cand_label, c->d_cand_rank, c->d_label_count, c->max_markers, g, c->h_markers, c->h_summary, c->d_scalars, c->d_dyn
static int stats_split_env() {
    static const int split_env = [] { const char* e = getenv("MAMRI_STATS_SPLIT"); return e ? atoi(e) : 1; }();
    return split_env;
}
static bool stats_split(const mamri_ctx* c) {
    const int e = stats_split_env();
    return e == 2 || ((e == 1 || e == 3) && c->run_ctas >= 592);
}
bool stats_early_beside() { return stats_split_env() == 3; }

cudaError_t launch_stats_early(mamri_ctx* c, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s) {
    if (!stats_split(c)) return cudaSuccess;
    const GeomArgs g = geom_args(desc, prm);
    const int W = (desc->nx + 31) / 32;
    static const int early_ctas = [] { const char* e = getenv("MAMRI_STATS_EARLY_CTAS"); return e ? atoi(e) : 0; }();
    const int grid = early_ctas > 0 ? early_ctas : (c->run_ctas > 0 ? c->run_ctas : MAMRI_RUN_CTAS);
    LK(k_stats<1>, grid, 256, s, false, MAMRI_STATS_ARGS);
    prof_mark(c, s, "stats_moments");
    return cudaGetLastError();
}

cudaError_t launch_stats(mamri_ctx* c, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s) {
    const GeomArgs g = geom_args(desc, prm);
    const int W = (desc->nx + 31) / 32;
    if (stats_split(c)) {
        int fin = int((c->max_markers + 127u) / 128u);             // one marker per thread: the float64 eigen-decomposition
        fin = fin < 1 ? 1 : (fin > 64 ? 64 : fin);                 // is a ~20 us dependent chain
        LK(k_stats<2>, fin, 128, s, false, MAMRI_STATS_ARGS);
    } else {
        // at most one CTA per SM: the kernel runs beside `materialise`, which needs the thread slots more
        int grid = c->run_ctas > 0 ? c->run_ctas : MAMRI_RUN_CTAS;
        if (grid > 148) grid = 148;
        LK(k_stats<0>, grid, 256, s, false, MAMRI_STATS_ARGS);
    }
    prof_mark(c, s, "stats");
    return cudaGetLastError();
}
#undef MAMRI_STATS_ARGS

KTRACE_TU(stats)
