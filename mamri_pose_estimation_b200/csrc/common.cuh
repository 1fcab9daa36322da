// Internal declarations shared by the kernels and the C ABI (include/mamri_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mamri_b200.h"

#define MAMRI_RMAX 3                 // largest closing radius the scratch layout is sized for
#define MAMRI_SCAN_CTAS 592          // 148 SMs x 4: CTA count of the chunked scans
#define MAMRI_RUN_CTAS (148 * 4)     // CTA count of the per-run kernels (grid-stride over the run table)
#define MAMRI_NONE 0xFFFFFFFFu

// Device-side scalars of one scan (one cudaMemsetAsync clears them).
struct DevScalars {
    unsigned int n_runs;
    unsigned int n_labels;
    unsigned int n_cand;             // labels passing the volume filter (may exceed max_markers)
    int          status;             // MAMRI_OK / MAMRI_ERR_CAPACITY
    unsigned long long body_packed;  // (count << 32) | (0xFFFFFFFF - label): atomicMax picks largest, lowest label
    unsigned long long n_foreground;
    unsigned int done_select;        // CTAs of k_select that have finished (last one prepares the moment table)
    unsigned int done_moments;       // CTAs of k_moments that have finished (last one finalises)
};

// Per-call pointers, read by the kernels from device memory so that the captured CUDA graph of the
// pipeline stays valid when only the caller's buffers change.
struct DynArgs {
    const void* vol;
    uint8_t* mask_out;
    uint32_t* labels_out;
    uint8_t* body_out;
};

// Everything else a captured pipeline depends on.
struct GraphKey {
    mamri_volume_desc desc;
    mamri_params prm;
    int vol_aligned, outs_aligned, has_mask, has_labels, has_body;
};

// Bit-packed volume: `w` 32-voxel words per row, `h` rows per slice, `d` slices.
struct BitVol {
    uint32_t* p;
    int w, h, d;
};

struct mamri_ctx {
    int device;
    int max_nx, max_ny, max_nz;
    uint32_t max_runs, max_markers;
    size_t cap_words, cap_pad_words;

    // scratch (device)
    uint32_t* d_raw;        // thresholded mask, bit-packed, zero-apron padded layout   [cap_pad_words]
    uint32_t* d_planes;     // morphology pass-A planes (x-z part of the ball)          [3 * cap_pad_words]
    uint32_t* d_dil;        // dilation on the r-grown domain, padded layout            [cap_pad_words]
    uint32_t* d_closed;     // closed mask, bit-packed [nz][ny][W]                      [cap_words]
    int raw_nx, raw_ny, raw_nz, raw_r;   // geometry d_raw's zero apron was last cleared for
    uint32_t* d_word_base;  // runs that start before each word         [cap_words]
    uint32_t* d_run_pos;    // word*32 + bit of each run's first voxel   [max_runs]
    uint32_t* d_run_len;    // voxels in each run                        [max_runs]
    uint32_t* d_parent;     // union-find over runs                     [max_runs]
    uint32_t* d_run_label;  // final label of each run                  [max_runs]
    uint32_t* d_label_count;// voxels per label                         [max_runs]
    uint32_t* d_label_slot; // marker-table slot of each label / NONE   [max_runs]
    uint32_t* d_block_sums; // scan partials                            [2 * 1024]
    uint32_t* d_cand_label; // label of each slot                       [max_markers + 1]
    unsigned long long* d_cand_sums; // 9 sums per slot (sx sy sz xx yy zz xy xz yz) [(max_markers+1)*9]
    mamri_marker* d_markers;         // sorted marker table              [max_markers]
    mamri_summary* d_summary;
    DevScalars* d_scalars;
    void* d_stage_in;       // staging for mamri_detect_host_async (lazy)
    size_t stage_in_bytes;
    uint8_t* d_stage_body;  // staging for the body mask (lazy)
    size_t stage_body_bytes;
    // entry search scratch
    double* d_entry_dist;   // per-CTA best                              [MAMRI_SCAN_CTAS]
    long long* d_entry_idx;
    unsigned long long* d_entry_cnt; // [2]
    mamri_entry_result* d_entry_res;

    // pinned host mirrors
    mamri_marker* h_markers;
    mamri_summary* h_summary;
    mamri_entry_result* h_entry_res;

    // CUDA graph of the whole pipeline (captured on first use of a configuration, relaunched afterwards)
    DynArgs* d_dyn;
    DynArgs* h_dyn;                  // pinned; copied to d_dyn by the graph's first node
    cudaStream_t cap_stream;
    cudaGraphExec_t gexec;
    GraphKey gkey;
    bool gvalid;
    bool use_graph;

    // optional per-stage timing (mamri_set_profiling): events recorded between the stage launches
    bool profile;
    cudaEvent_t ev[6];
    cudaEvent_t ev_fine[48];          // one event after every kernel launch (profile mode only)
    const char* fine_name[48];
    int n_fine;

    // state of the pending scan
    bool pending;
    cudaStream_t pending_stream;
    mamri_volume_desc last_desc;
    uint32_t last_n_labels;

    char err[512];
};

// Profile mode: mark the end of the kernel just launched.
inline void prof_mark(mamri_ctx* c, cudaStream_t s, const char* name) {
    if (!c->profile || c->n_fine >= 48) return;
    if (!c->ev_fine[c->n_fine] && cudaEventCreate(&c->ev_fine[c->n_fine]) != cudaSuccess) return;
    cudaEventRecord(c->ev_fine[c->n_fine], s);
    c->fine_name[c->n_fine++] = name;
}

// ---- stage launchers (each enqueues on `s`; returns cudaError_t) -------------------------------
cudaError_t prepare_raw_apron(mamri_ctx* c, int nx, int ny, int nz, int radius, cudaStream_t s);
cudaError_t launch_threshold_pack(mamri_ctx* c, int vol_aligned16, int dtype, int nx, int ny, int nz, double lo,
                                  double hi, int radius, cudaStream_t s);
cudaError_t launch_closing(mamri_ctx* c, int nx, int ny, int nz, int radius, cudaStream_t s);
cudaError_t launch_ccl(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int connectivity, cudaStream_t s);
cudaError_t launch_stats(mamri_ctx* c, const uint32_t* d_mask, const mamri_volume_desc* desc,
                         const mamri_params* prm, cudaStream_t s);
cudaError_t launch_materialise(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int outs_aligned,
                               cudaStream_t s);
cudaError_t launch_entry_search(mamri_ctx* c, const float* d_points, const float* d_normals, long long n,
                                const double target[3], double radius, double wx, double wy, double cutoff,
                                int n_path_samples, const uint8_t* d_path_mask, int mnx, int mny, int mnz,
                                const double ras_to_index[12], int path_free_value, cudaStream_t s);
cudaError_t launch_phantom(uint16_t* d_volume, int nx, int ny, int nz, const float* h_ell, int n_ell, float sigma,
                           unsigned long long seed, unsigned int scan_index, cudaStream_t s);

// ---- kernel launches ---------------------------------------------------------------------------
// Every pipeline kernel goes through launch_k: it attaches (a) a launch priority -- the two DRAM-bound
// kernels (threshold+pack, materialise) run at the lowest priority and the short latency-bound kernels
// between them at the highest, so that in a pipelined batch the short kernels of one scan are
// scheduled ahead of the queued CTAs of another scan's streaming kernel -- and (b) programmatic
// dependent launch: the next kernel's CTAs are set up while the current kernel drains; every kernel
// therefore starts with pdl_wait() (griddepcontrol.wait) before it touches memory.
// MAMRI_PRIO_SMALL / MAMRI_PRIO_BIG / MAMRI_PDL override the defaults (experiments only).
struct LaunchTuning { int prio_small, prio_big, pdl; };
const LaunchTuning& launch_tuning();

template <typename... KA, typename... A>
inline cudaError_t launch_ks(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool big, A&&... args) {
    const LaunchTuning& t = launch_tuning();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    at[n].id = cudaLaunchAttributePriority;
    at[n].val.priority = big ? t.prio_big : t.prio_small;
    ++n;
    if (t.pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}

template <typename... KA, typename... A>
inline cudaError_t launch_k(void (*kern)(KA...), dim3 grid, dim3 block, cudaStream_t s, bool big, A&&... args) {
    return launch_ks(kern, grid, block, 0, s, big, static_cast<A&&>(args)...);
}

#define LKS(...)                                                   \
    do {                                                           \
        cudaError_t _lk_e = launch_ks(__VA_ARGS__);                \
        if (_lk_e != cudaSuccess) return _lk_e;                    \
    } while (0)

#define LK(...)                                                    \
    do {                                                           \
        cudaError_t _lk_e = launch_k(__VA_ARGS__);                 \
        if (_lk_e != cudaSuccess) return _lk_e;                    \
    } while (0)

// ---- small device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Programmatic dependent launch: wait for the preceding kernel's memory to be visible, then let the
// following kernel's CTAs be scheduled as soon as every CTA of this one has got this far.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Run-start bits of word `m` given the previous word of the same row (0 at the row start).
__device__ __forceinline__ uint32_t run_starts(uint32_t m, uint32_t prev) { return m & ~((m << 1) | (prev >> 31)); }

// Id of the run that contains bit `bit` of a word: `base` runs start before the word.
__device__ __forceinline__ uint32_t run_id_in_word(uint32_t base, uint32_t starts, int bit) {
    return base + __popc(starts & (0xFFFFFFFFu >> (31 - bit))) - 1u;
}
#endif
