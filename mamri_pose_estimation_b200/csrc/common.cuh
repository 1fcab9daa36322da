// Internal declarations shared by the kernels and the C ABI (include/mamri_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mamri_b200.h"

#define MAMRI_RMAX 3                 // largest closing radius the scratch layout is sized for
#define MAMRI_SCAN_CTAS 592          // 148 SMs x 4: CTA count of the chunked scans
#define MAMRI_RUN_CTAS (148 * 8)     // largest CTA count of the per-run kernels (grid-stride over the run table)
#define MAMRI_NONE 0xFFFFFFFFu
#define MAMRI_SLOT_OF_ROOT 0xFFFFFFFEu   // label_slot of a non-root run whose label passed the volume filter: the slot is on its root

// Device-side scalars of one scan (one cudaMemsetAsync clears them).
struct DevScalars {
    unsigned int n_runs;
    unsigned int n_labels;
    unsigned int n_cand;             // labels passing the volume filter (may exceed max_markers)
    int          status;             // MAMRI_OK / MAMRI_ERR_CAPACITY
    unsigned long long body_packed;  // (count << 32) | (0xFFFFFFFF - label): atomicMax picks largest, lowest label
    unsigned long long n_foreground;
    unsigned int ticket_runs;        // tile tickets of k_runs_scan / k_flatten_rank
    unsigned int ticket_rank;
    unsigned int done_select;        // CTAs of k_select that have finished (last one prepares the moment table)
    unsigned int done_stats;         // CTAs of k_stats that have finished their moment sums (last one finalises)
    unsigned int ticket_close;       // tile tickets of k_close_fused
    unsigned int reserved_[3];
    unsigned int cl_roots[16];       // k_label_cluster: roots found by each CTA of the cluster (rank exchange)
};

// Device-side scalars of one body-surface extraction (surface.cu).
struct SurfScalars {
    unsigned int ticket;             // tile tickets of k_surface
    unsigned int reserved;
    unsigned long long n_points;     // surface voxels found (may exceed the caller's capacity)
    unsigned long long n_body;       // voxels of the body
};

// Per-call pointers, read by the kernels from device memory so that the captured CUDA graph of the
// pipeline stays valid when only the caller's buffers change.
struct DynArgs {
    const void* vol;
    uint8_t* mask_out;
    uint32_t* labels_out;
    uint8_t* body_out;
    uint32_t* body_bits_out;         // NULL or the body labelmap at 1 bit per voxel, [nz][ny][ceil(nx/32)] words
    uint32_t gen;                    // launch generation (tags the look-back states of the single-pass scans)
    uint32_t table_slots;            // rows of table_out
    double* table_out;               // NULL or [table_slots][8]: fixed-size marker table of the scan, written by the
                                     // finaliser so that a collective can be enqueued right behind the scan
};

// One device block per context: the per-call pointers followed by the scan's scalars.  The host keeps a pinned
// image whose scalar part is all zero, so ONE host-to-device copy per scan (per wave for a pool) both sets
// the pointers and resets the scalars.
struct ScanArgs {
    DynArgs dyn;
    DevScalars sc;
};

// Everything else a captured pipeline depends on.
struct GraphKey {
    mamri_volume_desc desc;
    mamri_params prm;
    int vol_aligned, outs_aligned, has_mask, has_labels, has_body, has_body_bits;
    int run_ctas;                    // grid of the per-run kernels (sized from the previous scans' run counts)
    int slice_threads;               // threads of the per-slice union-find CTAs (same source)
    int label_cluster;               // 0: labelling by the scalable multi-kernel path; N > 0: one cluster of N CTAs
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
    uint32_t* d_open;       // erosion of an opening, padded layout; its apron is zero and never written [cap_pad_words]
    uint32_t* d_closed;     // closed mask, bit-packed [nz][ny][W]                      [cap_words]
    int raw_nx, raw_ny, raw_nz, raw_r;   // geometry d_raw's zero apron was last cleared for
    uint8_t* d_occ_raw;     // occupancy cells of the raw mask (set by threshold+pack, cleared by the erosion)  [occ_cap]
    uint8_t* d_occ_dil;     // per-tile occupancy of the dilated mask (written by the dilation)                 [occ_cap]
    size_t occ_cap;
    uint32_t* d_word_base;  // runs that start before each word         [cap_words]
    uint32_t* d_run_pos;    // word*32 + bit of each run's first voxel   [max_runs]
    uint32_t* d_run_end;    // word*32 + bit of each run's last voxel    [max_runs]
    uint32_t* d_parent;     // union-find over runs                     [max_runs]
    uint32_t* d_run_label;  // final label of each run                  [max_runs]
    uint32_t* d_root_count; // voxels of the component rooted at each run (roots only) [max_runs]
    uint32_t* d_label_count;// voxels per label                         [max_runs]
    uint32_t* d_label_slot; // marker-table slot of each ROOT run / NONE; other runs: SLOT_OF_ROOT / NONE [max_runs]
    unsigned long long* d_scan_runs;  // look-back states of k_runs_scan     [tiles of cap_words]
    unsigned long long* d_scan_rank;  // look-back states of k_flatten_rank  [tiles of max_runs]
    uint32_t gen;
    uint32_t* d_cand_label; // label of each slot                       [max_markers + 1]
    uint32_t* d_cand_rank;  // place of each slot in ascending label order [max_markers]
    unsigned long long* d_cand_sums; // 9 sums per slot (sx sy sz xx yy zz xy xz yz) [(max_markers+1)*9]
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
    SurfScalars* d_surf;    // body-surface scalars
    void* d_pose_buf;       // pose stage: robot + points + counts + poses (grown on demand)
    size_t pose_buf_bytes;

    // pinned host mirrors
    mamri_marker* h_markers;
    mamri_summary* h_summary;
    mamri_entry_result* h_entry_res;
    SurfScalars* h_surf;

    // CUDA graph of the whole pipeline (captured on first use of a configuration, relaunched afterwards)
    ScanArgs* d_args;                // {per-call pointers, scalars}: d_dyn / d_scalars point into it
    ScanArgs* h_args;                // pinned image (scalars zero); copied to d_args by the graph's first node
    DynArgs* d_dyn;
    DynArgs* h_dyn;
    bool shared_args;                // d_args / h_args are slices of a pool's arrays (not owned)
    cudaStream_t cap_stream;
    cudaStream_t cap_stream2;        // second branch of the captured graph (materialise || moments + table copies)
    cudaEvent_t ev_fork, ev_join;
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
    uint32_t last_n_runs;             // x-runs of the last collected scan: sizes the next scan's per-run grids
    int run_ctas;                     // CTAs of the per-run kernels for the scan being enqueued
    int slice_threads;                // threads of the per-slice union-find CTAs (negative: default, nothing known yet)
    int label_cluster;                // CTAs of the labelling cluster for the scan being enqueued (0 = scalable path)
    int max_cluster;                  // largest cluster size k_label_cluster can be launched with on this device
    int n_launches;                   // kernels of one scan on the path last enqueued / captured

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
// geom_r = morph_geom_radius(open, close): the radius the padded layout (apron 2 geom_r) is built for
int morph_geom_radius(int open_radius, int close_radius);
cudaError_t launch_opening(mamri_ctx* c, int nx, int ny, int nz, int radius, int geom_r, cudaStream_t s);
cudaError_t launch_closing(mamri_ctx* c, int nx, int ny, int nz, int radius, int geom_r, cudaStream_t s);
// run numbering + union-find + ranking + volume filter + body label (one cluster kernel or the scalable kernels)
cudaError_t launch_label(mamri_ctx* c, const uint32_t* d_mask, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s);
cudaError_t launch_select(mamri_ctx* c, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s);
cudaError_t launch_stats_early(mamri_ctx* c, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s);   // before the fork (big run tables)
cudaError_t launch_stats(mamri_ctx* c, const mamri_volume_desc* desc, const mamri_params* prm, cudaStream_t s);
bool stats_early_beside();     // MAMRI_STATS_SPLIT=3: the early sums run beside `materialise` instead of before it
// per-device function attributes (dynamic shared memory opt-in, non-portable cluster size); called by mamri_create
cudaError_t segment_init_device();
int ccl_init_device();               // returns the largest cluster size the labelling kernel can use (0 = none)
cudaError_t launch_materialise(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, int outs_aligned,
                               cudaStream_t s);
cudaError_t launch_body_bits(mamri_ctx* c, const uint32_t* d_mask, int nx, int ny, int nz, cudaStream_t s);
cudaError_t launch_entry_search(mamri_ctx* c, const float* d_points, const float* d_normals, long long n,
                                const double target[3], double radius, double wx, double wy, double cutoff,
                                int n_path_samples, const uint8_t* d_path_mask, int mnx, int mny, int mnz,
                                const double ras_to_index[12], int path_free_value, cudaStream_t s);
cudaError_t launch_body_surface(mamri_ctx* c, const mamri_volume_desc* desc, const uint8_t* d_body_mask, uint32_t body_label,
                                float* d_points, float* d_normals, unsigned long long capacity, cudaStream_t s);
cudaError_t launch_pose(const mamri_robot* d_robot, const double* d_points, const int32_t* d_counts, int n_scans,
                        int max_points, const double* d_initial, const double* d_saved_base, int prefer_saved,
                        mamri_pose* d_poses, cudaStream_t s);
cudaError_t launch_collision(const mamri_robot* d_robot, const double base[16], const double m[12], int nx, int ny, int nz,
                             const int* offsets, int n_links, const double* d_angles, int n_configs, const float* d_points,
                             const uint8_t* d_mask, mamri_collision_result* d_out, cudaStream_t s);
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
struct LaunchTuning { int prio_small, prio_big, pdl, run_ctas; };
const LaunchTuning& launch_tuning();
// The run-table kernels (boundary merge, ranking, filter, moments) are latency-bound; CTAs they do not need
// still take SM slots away from the streaming kernels of the scans running beside them, so their grid follows
// the work: two runs per thread, from the run count of the scans just processed (`hint`, 0 = unknown), in
// power-of-two classes with hysteresis so that a captured graph is only re-captured when the load really changes.
int run_grid_class(int current, uint32_t hint);
// Same idea for the per-slice union-find CTAs: 128..512 threads, about one per run of an average slice.  A negative
// value means "default, nothing known yet" (the first hint then sets the class without hysteresis).
int slice_threads_class(int current, uint32_t hint, int nz);

// Kernels launched through launch_kc on this thread since the counter was last reset: the enqueue functions read it
// to report how many kernels one scan takes on the path it was given (mamri_kernel_launches).
inline int& launch_counter() { static thread_local int n = 0; return n; }

// `cluster` > 1: the grid is launched as thread-block clusters of that many CTAs along x.
template <typename... KA, typename... A>
inline cudaError_t launch_kc(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, unsigned cluster, cudaStream_t s, bool big,
                             A&&... args) {
    const LaunchTuning& t = launch_tuning();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[3];
    unsigned n = 0;
    at[n].id = cudaLaunchAttributePriority;
    at[n].val.priority = big ? t.prio_big : t.prio_small;
    ++n;
    if (t.pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at; cfg.numAttrs = n;
    ++launch_counter();
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}

template <typename... KA, typename... A>
inline cudaError_t launch_ks(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool big, A&&... args) {
    return launch_kc(kern, grid, block, smem, 1u, s, big, static_cast<A&&>(args)...);
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


#define FULL 0xFFFFFFFFu

// ---- in-pipeline kernel timeline (trace build only: -DMAMRI_KTRACE, libmamri_b200_trace.so) -------
// Every kernel stamps %globaltimer when its first CTA passes griddepcontrol.wait (= the kernel before it
// has completed) into slot `id`; ktrace_mark adds in-kernel phase stamps.  mamri_ktrace_read() returns
// the earliest stamp per slot.  One scan in flight at a time; the production library compiles this away.
enum KId { KT_THRESHOLD = 0, KT_CLOSE, KT_ERODE, KT_RUNS, KT_USLICE, KT_UZ1, KT_UZ2, KT_RANK, KT_SELECT, KT_LABEL,
           KT_STATS, KT_FINAL, KT_MAT, KT_END, KT_L_U1, KT_L_U2, KT_L_F, KT_L_FIX, KT_L_S, KT_L_END, KT_RUNS_LB, KT_RUNS_WR,
           KT_CLOSE_LD, KT_CLOSE_DIL, KT_CLOSE_ERO, KT_STATS_FIN, KT_CLOSE_LAST, KT_RUNS_LAST, KT_THR_LAST,
           // pairs (earliest CTA, latest CTA) of in-kernel phase stamps: ktrace_both(id) fills id and id + 1
           KT_US_N = 29, KT_US_JOIN = 31, KT_US_FLAT = 33, KT_US_END = 35, KT_UZ1_END = 37, KT_UZ2_END = 39, KT_RANK_END = 41,
           KT_SEL_END = 43, KT_RUNS_LB2 = 45, KT_SLOTS = 48 };
#ifdef MAMRI_KTRACE
static __device__ unsigned long long g_ktrace[KT_SLOTS];      // one copy per translation unit (no -rdc)
__device__ __forceinline__ void ktrace(int id) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(&g_ktrace[id], t);
    }
}
// latest stamp of a slot instead of the earliest (stored complemented; tools/ktrace.py complements KT_END back)
__device__ __forceinline__ void ktrace_last(int id) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(&g_ktrace[id], ~t);
    }
}
__device__ __forceinline__ void ktrace_both(int id) { ktrace(id); ktrace_last(id + 1); }
// reset / min-merge of this translation unit's stamps (called by mamri_ktrace_reset / mamri_ktrace_read)
#define KTRACE_TU(name)                                                                         \
    void ktrace_reset_##name() {                                                                \
        unsigned long long v[KT_SLOTS];                                                         \
        memset(v, 0xFF, sizeof(v));                                                             \
        cudaMemcpyToSymbol(g_ktrace, v, sizeof(v));                                             \
    }                                                                                           \
    void ktrace_merge_##name(unsigned long long* out) {                                         \
        unsigned long long v[KT_SLOTS];                                                         \
        if (cudaMemcpyFromSymbol(v, g_ktrace, sizeof(v)) != cudaSuccess) return;                \
        for (int i = 0; i < KT_SLOTS; ++i) out[i] = v[i] < out[i] ? v[i] : out[i];              \
    }
#else
__device__ __forceinline__ void ktrace(int) {}
__device__ __forceinline__ void ktrace_last(int) {}
__device__ __forceinline__ void ktrace_both(int) {}
#define KTRACE_TU(name)
#endif

// ---- single-pass prefix sums over tiles (decoupled look-back) ------------------------------------
// One 64-bit word per tile: [63:34] generation of the launch that wrote it, [33:2] value, [1:0] flag
// (1 = the tile's own total, 2 = inclusive prefix).  Words of older generations read as "empty", so the
// array never needs clearing; tiles are handed out by an atomic ticket, hence a tile only ever waits
// for tiles whose CTAs are already running.
__device__ __forceinline__ unsigned long long scan_pack(uint32_t gen, uint32_t value, uint32_t flag) {
    return ((unsigned long long)(gen & 0x3FFFFFFFu) << 34) | ((unsigned long long)value << 2) | flag;
}

// Called by the whole first warp of the CTA that owns `tile`: publishes `total`, returns the sum of the
// totals of all earlier tiles.
__device__ __forceinline__ uint32_t scan_lookback(volatile unsigned long long* state, uint32_t tile, uint32_t total,
                                                  uint32_t gen) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0) {
        if (lane == 0) state[0] = scan_pack(gen, total, 2u);
        return 0u;
    }
    if (lane == 0) state[tile] = scan_pack(gen, total, 1u);
    const unsigned long long g30 = gen & 0x3FFFFFFFu;
    uint32_t excl = 0;
    int pos = int(tile);
    while (true) {
        const int idx = pos - 1 - int(lane);
        uint32_t flag = 2u, val = 0u;                    // before the first tile: inclusive prefix 0
        if (idx >= 0) {
            unsigned long long v;
            do {
                v = state[idx];
                flag = ((v >> 34) == g30) ? uint32_t(v & 3u) : 0u;
            } while (flag == 0u);
            val = uint32_t(v >> 2);
        }
        const unsigned incl = __ballot_sync(FULL, flag == 2u);
        const int first = incl ? (__ffs(incl) - 1) : 31; // nearest earlier tile with an inclusive prefix
        uint32_t c = int(lane) <= first ? val : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        excl += c;
        if (incl) break;
        pos -= 32;
    }
    if (lane == 0) state[tile] = scan_pack(gen, excl + total, 2u);
    return excl;
}

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
    if (!__any_sync(FULL, valid)) return;                    // nothing to add in this warp
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

// Run-start bits of word `m` given the previous word of the same row (0 at the row start).
__device__ __forceinline__ uint32_t run_starts(uint32_t m, uint32_t prev) { return m & ~((m << 1) | (prev >> 31)); }

// Id of the run that contains bit `bit` of a word: `base` runs start before the word.
__device__ __forceinline__ uint32_t run_id_in_word(uint32_t base, uint32_t starts, int bit) {
    return base + __popc(starts & (0xFFFFFFFFu >> (31 - bit))) - 1u;
}
#endif
