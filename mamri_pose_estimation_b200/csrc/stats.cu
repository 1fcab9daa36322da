// Stages 3 and 4a: label-shape statistics, the candidate-marker volume filter and the body label.
// Replaces sitk.LabelShapeStatisticsImageFilter().Execute + GetLabels/GetPhysicalSize/GetCentroid and
// the list comprehension / max() at Mamri/Mamri.py:1309-1310, 1316-1322.
//
// Statistics are accumulated as exact integers per x-run PIECE (the part of a run inside one 32-voxel
// word), using closed forms for sum x and sum x^2 over the piece, so the work is O(#runs) instead of
// O(#voxels).  Lanes of a warp that hit the same label are combined with warp shuffles before one
// atomic per (warp, label): integer sums are order-independent -> bit-reproducible.  Two phases bound
// the table on noisy scans: voxel counts for every label first, then second moments only for the
// labels that pass the volume filter (plus the body).  Finalisation (centroid, physical size,
// principal moments/axes) is float64 on the device.
#include "common.cuh"

#define FULL 0xFFFFFFFFu

struct GeomArgs {
    double spacing[3], origin[3], dir[9];
    double voxel_volume;     // spacing[0]*spacing[1]*spacing[2], multiplied in that order on the host
    double min_volume, max_volume;
};

// ------------------------------------------------------------------------------------------------
// aggregation: warp shuffles first, then a per-CTA shared-memory cache, then global atomics
// ------------------------------------------------------------------------------------------------
// One huge component (the body) makes every warp hit the same table row; per-warp global atomics on
// one address serialise in L2.  Each CTA therefore keeps a small direct-mapped cache of accumulators in
// shared memory (slot = key % SLOTS, claimed by the first key that arrives, never evicted); cached
// keys cost a shared-memory atomic, the rest go to global memory; the cache is flushed once per CTA.
template <int NV, typename V, int SLOTS>
struct CtaCache {
    uint32_t tag[SLOTS];
    V val[SLOTS][NV];

    __device__ void init() {
        for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) tag[i] = MAMRI_NONE;
        for (int i = threadIdx.x; i < SLOTS * NV; i += blockDim.x) (&val[0][0])[i] = V(0);
        __syncthreads();
    }
    __device__ __forceinline__ void add(uint32_t key, const V (&v)[NV], V* table) {
        const int s = int(key % SLOTS);
        uint32_t t = *(volatile uint32_t*)&tag[s];
        if (t == MAMRI_NONE) {
            t = atomicCAS(&tag[s], MAMRI_NONE, key);
            if (t == MAMRI_NONE) t = key;
        }
        if (t == key) {
#pragma unroll
            for (int i = 0; i < NV; ++i) atomicAdd(&val[s][i], v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) atomicAdd(table + uint32_t(key) * NV + i, v[i]);
        }
    }
    __device__ void flush(V* table) {
        __syncthreads();
        for (int i = threadIdx.x; i < SLOTS * NV; i += blockDim.x) {
            const uint32_t key = tag[i / NV];
            const V x = (&val[0][0])[i];
            if (key != MAMRI_NONE && x != V(0)) atomicAdd(table + uint32_t(key) * NV + (i % NV), x);
        }
    }
};

// Lanes with equal keys are summed by shuffles; one lane per distinct key forwards to the CTA cache.
template <int NV, typename V, int SLOTS>
__device__ __forceinline__ void warp_agg_add(uint32_t key, V (&v)[NV], CtaCache<NV, V, SLOTS>& cache, V* table) {
    const unsigned lane = lane_id();
    const bool valid = key != MAMRI_NONE;
    const unsigned peers = __match_any_sync(FULL, key);
    const bool single = valid && peers == (1u << lane);
    if (single) cache.add(key, v, table);
    __syncwarp();
    unsigned remaining = __ballot_sync(FULL, valid && !single);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const uint32_t k = __shfl_sync(FULL, key, leader);
        const bool mine = valid && key == k;
        V x[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            x[i] = mine ? v[i] : V(0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x[i] += __shfl_xor_sync(FULL, x[i], o);
        }
        if (int(lane) == leader) cache.add(k, x, table);
        remaining &= ~__ballot_sync(FULL, mine);
    }
}

// First maximal piece of set bits of `pm`: start bit and length; clears it from pm.
__device__ __forceinline__ void pop_piece(uint32_t& pm, int& b, int& len) {
    b = __ffs(pm) - 1;
    uint32_t t = pm >> b;
    len = (t == 0xFFFFFFFFu) ? 32 : (__ffs(~t) - 1);
    uint32_t bits = (len == 32) ? 0xFFFFFFFFu : ((1u << len) - 1u);
    pm &= ~(bits << b);
}

// ------------------------------------------------------------------------------------------------
// phase 1: voxel count of every label
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_labels(const uint32_t* __restrict__ mask,
                                                      const uint32_t* __restrict__ word_base,
                                                      const uint32_t* __restrict__ run_label, int W, uint32_t n_words,
                                                      uint32_t* label_count, const DevScalars* sc) {
    __shared__ CtaCache<1, uint32_t, 64> cache;
    if (sc->status != MAMRI_OK) return;
    cache.init();
    const unsigned lane = lane_id();
    const uint32_t warp0 = ((uint32_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) << 5;
    const uint32_t stride = uint32_t(gridDim.x) * blockDim.x;
    for (uint32_t w0 = warp0; w0 < n_words; w0 += stride) {
        const uint32_t wi = w0 + lane;
        uint32_t pm = wi < n_words ? mask[wi] : 0u;
        if (!__any_sync(FULL, pm != 0u)) continue;
        uint32_t starts = 0, base = 0;
        if (pm) {
            starts = run_starts(pm, (wi % W) ? mask[wi - 1] : 0u);
            base = word_base[wi];
        }
        while (__any_sync(FULL, pm != 0u)) {
            uint32_t key = MAMRI_NONE;
            uint32_t v[1] = {0u};
            if (pm) {
                int b, len;
                pop_piece(pm, b, len);
                key = run_label[run_id_in_word(base, starts, b)] - 1u;
                v[0] = uint32_t(len);
            }
            warp_agg_add(key, v, cache, label_count);
        }
    }
    cache.flush(label_count);
}

// ------------------------------------------------------------------------------------------------
// stage 4a: volume filter (Mamri.py:1310) and body label (Mamri.py:1320-1322)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_select(const uint32_t* __restrict__ label_count, uint32_t* __restrict__ label_slot,
                                                uint32_t* __restrict__ cand_label, uint32_t max_markers, GeomArgs g,
                                                DevScalars* sc) {
    const unsigned lane = lane_id();
    const uint32_t n = sc->n_labels;
    const uint32_t warp0 = ((uint32_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) << 5;
    const uint32_t stride = uint32_t(gridDim.x) * blockDim.x;
    for (uint32_t l0 = warp0; l0 < n; l0 += stride) {
        const uint32_t l = l0 + lane;
        unsigned long long packed = 0ull, cnt64 = 0ull;
        if (l < n) {
            const uint32_t cnt = label_count[l];
            cnt64 = cnt;
            const double vol = double(cnt) * g.voxel_volume;        // GetPhysicalSize
            if (vol >= g.min_volume && vol <= g.max_volume) {         // inclusive bounds
                uint32_t slot = atomicAdd(&sc->n_cand, 1u);
                if (slot < max_markers) { cand_label[slot] = uint32_t(l) + 1u; label_slot[l] = slot; }
                else label_slot[l] = MAMRI_NONE;
            } else {
                label_slot[l] = MAMRI_NONE;
                // max(..., key=GetPhysicalSize) returns the FIRST maximum -> lowest label on ties
                packed = (cnt64 << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t(l) + 1u));
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
    }
}

// One CTA: clamps the candidate count, gives the body the extra slot `max_markers`, zeroes the sums.
__global__ void __launch_bounds__(256) k_prepare_moments(uint32_t* __restrict__ label_slot, uint32_t* __restrict__ cand_label,
                                                         unsigned long long* __restrict__ sums, uint32_t max_markers,
                                                         DevScalars* sc) {
    uint32_t n = sc->n_cand;
    if (n > max_markers) {
        n = max_markers;
        if (threadIdx.x == 0) sc->status = MAMRI_ERR_CAPACITY;
    }
    const unsigned long long bp = sc->body_packed;
    if (threadIdx.x == 0 && (bp >> 32) != 0ull && sc->n_labels > 0) {
        uint32_t body = 0xFFFFFFFFu - uint32_t(bp & 0xFFFFFFFFull);
        label_slot[body - 1u] = max_markers;
        cand_label[max_markers] = body;
    }
    for (uint32_t i = threadIdx.x; i < n * 9u; i += blockDim.x) sums[i] = 0ull;
    for (uint32_t i = threadIdx.x; i < 9u; i += blockDim.x) sums[uint32_t(max_markers) * 9u + i] = 0ull;
}

// ------------------------------------------------------------------------------------------------
// phase 2: first and second moments of the kept labels (+ body)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sum_sq_upto(long long k) {   // sum_{i=0..k} i^2, k >= -1
    return (unsigned long long)(k * (k + 1) * (2 * k + 1) / 6);
}

__global__ void __launch_bounds__(256) k_moments(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ word_base,
                                                 const uint32_t* __restrict__ run_label,
                                                 const uint32_t* __restrict__ label_slot, int W, int ny, uint32_t n_words,
                                                 unsigned long long* sums, const DevScalars* sc) {
    __shared__ CtaCache<9, unsigned long long, 16> cache;
    if (sc->status != MAMRI_OK) return;
    cache.init();
    const unsigned lane = lane_id();
    const uint32_t warp0 = ((uint32_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) << 5;
    const uint32_t stride = uint32_t(gridDim.x) * blockDim.x;
    for (uint32_t w0 = warp0; w0 < n_words; w0 += stride) {
        const uint32_t wi = w0 + lane;
        uint32_t pm = wi < n_words ? mask[wi] : 0u;
        if (!__any_sync(FULL, pm != 0u)) continue;
        uint32_t starts = 0, base = 0;
        long long x0 = 0, y = 0, z = 0;
        if (pm) {
            const uint32_t row = wi / W;
            const int xw = int(wi - row * W);
            starts = run_starts(pm, xw > 0 ? mask[wi - 1] : 0u);
            base = word_base[wi];
            x0 = 32ll * xw;
            y = (long long)(row % ny);
            z = (long long)(row / ny);
        }
        while (__any_sync(FULL, pm != 0u)) {
            uint32_t key = MAMRI_NONE;
            unsigned long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (pm) {
                int b, len;
                pop_piece(pm, b, len);
                const uint32_t label = run_label[run_id_in_word(base, starts, b)];
                key = label_slot[label - 1u];
                if (key != MAMRI_NONE) {
                    const long long xs = x0 + b, xe = xs + len - 1, n = len;
                    const unsigned long long sx = (unsigned long long)((xs + xe) * n / 2);
                    v[0] = sx;                                              // sum x
                    v[1] = (unsigned long long)(n * y);                     // sum y
                    v[2] = (unsigned long long)(n * z);                     // sum z
                    v[3] = sum_sq_upto(xe) - sum_sq_upto(xs - 1);           // sum xx
                    v[4] = (unsigned long long)(n * y * y);                 // sum yy
                    v[5] = (unsigned long long)(n * z * z);                 // sum zz
                    v[6] = sx * (unsigned long long)y;                      // sum xy
                    v[7] = sx * (unsigned long long)z;                      // sum xz
                    v[8] = (unsigned long long)(n * y * z);                 // sum yz
                }
            }
            warp_agg_add(key, v, cache, sums);
        }
    }
    cache.flush(sums);
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

__device__ void make_marker(mamri_marker* out, uint32_t label, unsigned long long count, const unsigned long long* s9,
                            const GeomArgs& g) {
    mamri_marker m;
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
    *out = m;
}

// One CTA: orders the kept labels ascending (= GetLabels order), emits their records and the summary.
__global__ void __launch_bounds__(256) k_finalize(const uint32_t* __restrict__ cand_label,
                                                  const uint32_t* __restrict__ label_count,
                                                  const unsigned long long* __restrict__ sums, uint32_t max_markers,
                                                  GeomArgs g, mamri_marker* __restrict__ markers, mamri_summary* summary,
                                                  const DevScalars* sc) {
    const bool ok = sc->status == MAMRI_OK;
    const uint32_t n_all = sc->n_cand;
    const uint32_t n = ok ? (n_all < max_markers ? n_all : max_markers) : 0u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t lab = cand_label[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) rank += cand_label[j] < lab;
        make_marker(markers + rank, lab, label_count[lab - 1u], sums + uint32_t(i) * 9u, g);
    }
    if (threadIdx.x == 0) {
        summary->n_labels = sc->n_labels;
        summary->n_runs = sc->n_runs;
        summary->n_markers = n_all;
        summary->n_foreground = sc->n_foreground;
        summary->device_status = sc->status;
        summary->reserved = 0;
        const unsigned long long bp = sc->body_packed;
        if (ok && (bp >> 32) != 0ull) {
            const uint32_t body = 0xFFFFFFFFu - uint32_t(bp & 0xFFFFFFFFull);
            summary->body_label = body;
            summary->body_count = bp >> 32;
            make_marker(&summary->body, body, bp >> 32, sums + uint32_t(max_markers) * 9u, g);
        } else {
            summary->body_label = 0;
            summary->body_count = 0;
            memset(&summary->body, 0, sizeof(mamri_marker));
        }
    }
}

cudaError_t launch_stats(mamri_ctx* c, const uint32_t* d_mask, const mamri_volume_desc* desc, const mamri_params* prm,
                         cudaStream_t s) {
    const int W = (desc->nx + 31) / 32;
    const uint32_t n_words = uint32_t(W) * desc->ny * desc->nz;
    GeomArgs g;
    for (int i = 0; i < 3; ++i) { g.spacing[i] = desc->spacing[i]; g.origin[i] = desc->origin[i]; }
    for (int i = 0; i < 9; ++i) g.dir[i] = desc->direction[i];
    double vv = 1.0;
    for (int i = 0; i < 3; ++i) vv *= desc->spacing[i];       // ITK: sizePerPixel *= spacing[i]
    g.voxel_volume = vv;
    g.min_volume = prm->min_volume;
    g.max_volume = prm->max_volume;
    uint32_t wb = (n_words + 255) / 256;
    if (wb > 148 * 8) wb = 148 * 8;
    if (wb == 0) wb = 1;
    k_count_labels<<<unsigned(wb), 256, 0, s>>>(d_mask, c->d_word_base, c->d_run_label, W, n_words, c->d_label_count,
                                                c->d_scalars);
    k_select<<<148 * 2, 256, 0, s>>>(c->d_label_count, c->d_label_slot, c->d_cand_label, c->max_markers, g, c->d_scalars);
    k_prepare_moments<<<1, 256, 0, s>>>(c->d_label_slot, c->d_cand_label, c->d_cand_sums, c->max_markers, c->d_scalars);
    k_moments<<<unsigned(wb), 256, 0, s>>>(d_mask, c->d_word_base, c->d_run_label, c->d_label_slot, W, desc->ny, n_words,
                                           c->d_cand_sums, c->d_scalars);
    k_finalize<<<1, 256, 0, s>>>(c->d_cand_label, c->d_label_count, c->d_cand_sums, c->max_markers, g, c->d_markers,
                                 c->d_summary, c->d_scalars);
    return cudaGetLastError();
}
