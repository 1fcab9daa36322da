// Stage 4b: closest suitable entry point.  Replaces the Python loop of
// MamriLogic.findAndSetEntryPoint (Mamri/Mamri.py:1008-1023): radius query around the target, the
// normal-direction score |nx| - 2|ny| > -0.5, arg-min of the distance.  Optional needle-path sampling
// against a voxel mask (north_star extension; the reference has no counterpart).
//
// One candidate per thread (24 B read, coalesced float loads), float64 arithmetic with separately
// rounded operations in the oracle's order so distances are bit-identical, lexicographic
// (distance, index) minimum by warp shuffles -> per-CTA partial -> one final CTA: deterministic,
// lowest index wins ties.
#include "common.cuh"

#define FULL 0xFFFFFFFFu

struct EntryArgs {
    double t[3];
    double r2, wx, wy, cutoff;
    int n_samples;
    int mnx, mny, mnz;
    double m[12];          // RAS mm -> voxel index, row-major 3x4
    int free_value;
};

__device__ __forceinline__ bool better(double d, long long i, double bd, long long bi) {
    return d < bd || (d == bd && i < bi);
}

__global__ void __launch_bounds__(256) k_entry_search(const float* __restrict__ pts, const float* __restrict__ nrm,
                                                      long long n, EntryArgs a, const uint8_t* __restrict__ mask,
                                                      double* __restrict__ blk_dist, long long* __restrict__ blk_idx,
                                                      unsigned long long* counters) {
    double best_d = INFINITY;
    long long best_i = -1;
    unsigned n_rad = 0, n_ok = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
        const double dx = __dsub_rn(px, a.t[0]), dy = __dsub_rn(py, a.t[1]), dz = __dsub_rn(pz, a.t[2]);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (!(d2 <= a.r2)) continue;                         // FindPointsWithinRadius
        ++n_rad;
        const double score = __dadd_rn(__dmul_rn(a.wx, fabs(double(nrm[3 * i]))), __dmul_rn(a.wy, fabs(double(nrm[3 * i + 1]))));
        if (!(score > a.cutoff)) continue;                   // suitability_score > -0.5
        bool clear = true;
        for (int k = 1; k <= a.n_samples && clear; ++k) {
            const double f = double(k) / double(a.n_samples + 1);
            const double qx = px + f * (a.t[0] - px), qy = py + f * (a.t[1] - py), qz = pz + f * (a.t[2] - pz);
            const long long ix = llrint(a.m[0] * qx + a.m[1] * qy + a.m[2] * qz + a.m[3]);
            const long long iy = llrint(a.m[4] * qx + a.m[5] * qy + a.m[6] * qz + a.m[7]);
            const long long iz = llrint(a.m[8] * qx + a.m[9] * qy + a.m[10] * qz + a.m[11]);
            int val = 0;                                      // outside the mask volume reads as 0
            if (ix >= 0 && ix < a.mnx && iy >= 0 && iy < a.mny && iz >= 0 && iz < a.mnz)
                val = mask[(size_t(iz) * a.mny + iy) * a.mnx + ix];
            clear = (val == a.free_value);
        }
        if (!clear) continue;
        ++n_ok;
        const double d = sqrt(d2);
        if (better(d, i, best_d, best_i) || best_i < 0) { best_d = d; best_i = i; }
    }
    // warp then block reduction of (distance, index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double od = __shfl_xor_sync(FULL, best_d, o);
        long long oi = __shfl_xor_sync(FULL, best_i, o);
        if (oi >= 0 && (best_i < 0 || better(od, oi, best_d, best_i))) { best_d = od; best_i = oi; }
        n_rad += __shfl_xor_sync(FULL, n_rad, o);
        n_ok += __shfl_xor_sync(FULL, n_ok, o);
    }
    __shared__ double sd[8];
    __shared__ long long si[8];
    __shared__ unsigned sr[8], so[8];
    const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
    if (lane == 0) { sd[wid] = best_d; si[wid] = best_i; sr[wid] = n_rad; so[wid] = n_ok; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tr = 0, to = 0;
        for (int w = 0; w < 8; ++w) {
            if (si[w] >= 0 && (best_i < 0 || better(sd[w], si[w], best_d, best_i))) { best_d = sd[w]; best_i = si[w]; }
            tr += sr[w];
            to += so[w];
        }
        blk_dist[blockIdx.x] = best_d;
        blk_idx[blockIdx.x] = best_i;
        if (tr) atomicAdd(counters, tr);
        if (to) atomicAdd(counters + 1, to);
    }
}

__global__ void k_entry_final(const double* __restrict__ blk_dist, const long long* __restrict__ blk_idx, int n_blocks,
                              const float* __restrict__ pts, const unsigned long long* counters, mamri_entry_result* res) {
    if (threadIdx.x != 0) return;
    double bd = INFINITY;
    long long bi = -1;
    for (int b = 0; b < n_blocks; ++b)
        if (blk_idx[b] >= 0 && (bi < 0 || better(blk_dist[b], blk_idx[b], bd, bi))) { bd = blk_dist[b]; bi = blk_idx[b]; }
    res->index = bi;
    res->distance = bd;
    for (int k = 0; k < 3; ++k) res->point[k] = bi >= 0 ? double(pts[3 * bi + k]) : 0.0;
    res->n_in_radius = counters[0];
    res->n_suitable = counters[1];
}

cudaError_t launch_entry_search(mamri_ctx* c, const float* d_points, const float* d_normals, long long n,
                                const double target[3], double radius, double wx, double wy, double cutoff,
                                int n_path_samples, const uint8_t* d_path_mask, int mnx, int mny, int mnz,
                                const double ras_to_index[12], int path_free_value, cudaStream_t s) {
    EntryArgs a;
    for (int i = 0; i < 3; ++i) a.t[i] = target[i];
    a.r2 = radius * radius;
    a.wx = wx; a.wy = wy; a.cutoff = cutoff;
    a.n_samples = (d_path_mask && n_path_samples > 0) ? n_path_samples : 0;
    a.mnx = mnx; a.mny = mny; a.mnz = mnz;
    for (int i = 0; i < 12; ++i) a.m[i] = ras_to_index ? ras_to_index[i] : 0.0;
    a.free_value = path_free_value;
    long long blocks = (n + 255) / 256;
    if (blocks > MAMRI_SCAN_CTAS) blocks = MAMRI_SCAN_CTAS;
    if (blocks < 1) blocks = 1;
    cudaError_t e = cudaMemsetAsync(c->d_entry_cnt, 0, 2 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    k_entry_search<<<unsigned(blocks), 256, 0, s>>>(d_points, d_normals, n, a, d_path_mask, c->d_entry_dist, c->d_entry_idx,
                                                    c->d_entry_cnt);
    k_entry_final<<<1, 32, 0, s>>>(c->d_entry_dist, c->d_entry_idx, int(blocks), d_points, c->d_entry_cnt, c->d_entry_res);
    return cudaGetLastError();
}
