// C ABI of the fiducial-detection path (include/mamri_b200.h).  Host-side orchestration only: it
// validates arguments, enqueues the stage kernels on the caller's stream and moves the small result
// tables.  No allocation on the hot calls; errors become status codes, never aborts.
#include "common.cuh"

#include <new>
#include <stdlib.h>

static char g_create_err[512] = "";

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                              \
            return MAMRI_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)

// Launch priorities and programmatic dependent launch (common.cuh: launch_k).
const LaunchTuning& launch_tuning() {
    static const LaunchTuning t = [] {
        LaunchTuning v;
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { least = greatest = 0; cudaGetLastError(); }
        v.prio_small = greatest;     // numerically lowest = scheduled first
        v.prio_big = least;
        v.pdl = 1;
        v.run_ctas = 0;                 // 0 = follow the work (run_grid_class)
        if (const char* e = getenv("MAMRI_RUN_CTAS")) v.run_ctas = atoi(e) < 1 ? 1 : atoi(e);
        if (const char* e = getenv("MAMRI_PRIO_SMALL")) v.prio_small = atoi(e);
        if (const char* e = getenv("MAMRI_PRIO_BIG")) v.prio_big = atoi(e);
        if (const char* e = getenv("MAMRI_PDL")) v.pdl = atoi(e);
        return v;
    }();
    return t;
}

int run_grid_class(int current, uint32_t hint) {
    const int forced = launch_tuning().run_ctas;
    if (forced > 0) return forced > MAMRI_RUN_CTAS ? MAMRI_RUN_CTAS : forced;
    if (hint == 0) return current > 0 ? current : MAMRI_RUN_CTAS;       // nothing known yet: the full grid
    const unsigned long long need = (hint + 511ull) / 512ull;           // CTAs of 256 threads at two runs per thread
    int want = 37;                                                      // classes 37, 74, ..., 1184 (148 SMs / 4 ... x 8)
    while (want < MAMRI_RUN_CTAS && (unsigned long long)want < need) want *= 2;
    if (current <= 0 || want > current) return want;                    // grow at once
    if (want * 4 <= current) return want * 2;                           // shrink only on a 4x drop, keep 2x headroom
    return current;
}

int slice_threads_class(int current, uint32_t hint, int nz) {
    static const int forced = [] { const char* e = getenv("MAMRI_SLICE_THREADS"); return e ? atoi(e) : 0; }();
    if (forced >= 32 && forced <= 512) return forced;
    if (hint == 0 || nz <= 0) return current != 0 ? current : -512;
    const unsigned long long per_slice = hint / (unsigned long long)nz;
    int want = 128;
    while (want < 512 && (unsigned long long)want < per_slice) want *= 2;
    if (current <= 0 || want > current) return want;
    if (want * 4 <= current) return want * 2;
    return current;
}

// Labelling path for the scan being enqueued: one cluster kernel while the run table is small (about 8 K runs per CTA
// of the cluster, so that every chunk's parents fit shared memory), the scalable kernels above that.  `hint` = runs of
// the scans just processed (0 = nothing known yet: decided by the volume size).
int label_cluster_class(int current, uint32_t hint, unsigned long long voxels, int max_cluster) {
    if (max_cluster < 2) return 0;
    static const long long limit_env = [] { const char* e = getenv("MAMRI_LABEL_CLUSTER_RUNS"); return e ? atoll(e) : 0ll; }();
    const unsigned long long limit = limit_env > 0 ? (unsigned long long)limit_env : (unsigned long long)max_cluster * 8192ull;
    if (hint == 0) return voxels <= (1ull << 27) ? max_cluster : 0;
    if (current > 0) return hint <= limit + limit / 4 ? max_cluster : 0;   // hysteresis: leave the path only on a clear excess
    return hint <= limit ? max_cluster : 0;
}

namespace {
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

size_t dtype_size(int dtype) {
    switch (dtype) {
        case MAMRI_U8: return 1;
        case MAMRI_I16: case MAMRI_U16: return 2;
        case MAMRI_I32: case MAMRI_F32: return 4;
        case MAMRI_F64: return 8;
        default: return 0;
    }
}

int fail(mamri_ctx* ctx, int code, const char* msg) {
    snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return code;
}
}  // namespace

#ifdef MAMRI_KTRACE
void ktrace_reset_segment(); void ktrace_merge_segment(unsigned long long*);
void ktrace_reset_ccl(); void ktrace_merge_ccl(unsigned long long*);
void ktrace_reset_stats(); void ktrace_merge_stats(unsigned long long*);
#endif
// Trace build only (libmamri_b200_trace.so): reset / read the in-pipeline kernel timeline (common.cuh: ktrace).
extern "C" int mamri_ktrace_reset(void) {
#ifdef MAMRI_KTRACE
    ktrace_reset_segment(); ktrace_reset_ccl(); ktrace_reset_stats();
    return cudaDeviceSynchronize() == cudaSuccess ? MAMRI_OK : MAMRI_ERR_CUDA;
#else
    return MAMRI_ERR_STATE;
#endif
}
extern "C" int mamri_ktrace_read(unsigned long long* out_ns, int n) {
#ifdef MAMRI_KTRACE
    if (!out_ns || n < KT_SLOTS) return MAMRI_ERR_INVALID_ARG;
    if (cudaDeviceSynchronize() != cudaSuccess) return MAMRI_ERR_CUDA;
    for (int i = 0; i < KT_SLOTS; ++i) out_ns[i] = ~0ull;
    ktrace_merge_segment(out_ns); ktrace_merge_ccl(out_ns); ktrace_merge_stats(out_ns);
    return KT_SLOTS;
#else
    (void)out_ns; (void)n;
    return MAMRI_ERR_STATE;
#endif
}

extern "C" const char* mamri_version(void) { return "mamri_b200 0.1 (sm_100a)"; }

extern "C" void mamri_default_params(mamri_params* p) {
    if (!p) return;
    p->lower = 65.0;          // Mamri.py:810
    p->upper = 65535.0;       // Mamri.py:1308
    p->close_radius = 2;      // Mamri.py:1308
    p->open_radius = 0;       // the reference does not open
    p->reserved = 0;
    p->connectivity = 6;      // Mamri.py:1309 (SimpleITK default fullyConnected=False)
    p->min_volume = 50.0;     // Mamri.py:811
    p->max_volume = 1500.0;   // Mamri.py:812
}

extern "C" const char* mamri_last_error(const mamri_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int mamri_destroy(mamri_ctx* ctx) {
    if (!ctx) return MAMRI_OK;
    DeviceGuard g(ctx->device);
    cudaFree(ctx->d_occ_raw); cudaFree(ctx->d_occ_dil);
    cudaFree(ctx->d_raw); cudaFree(ctx->d_planes); cudaFree(ctx->d_dil); cudaFree(ctx->d_open); cudaFree(ctx->d_closed); cudaFree(ctx->d_word_base);
    cudaFree(ctx->d_run_pos); cudaFree(ctx->d_run_end); cudaFree(ctx->d_parent); cudaFree(ctx->d_run_label); cudaFree(ctx->d_label_count); cudaFree(ctx->d_label_slot);
    cudaFree(ctx->d_root_count); cudaFree(ctx->d_scan_runs); cudaFree(ctx->d_scan_rank); cudaFree(ctx->d_cand_label); cudaFree(ctx->d_cand_rank); cudaFree(ctx->d_cand_sums);
    cudaFree(ctx->d_stage_in); cudaFree(ctx->d_stage_body);
    cudaFree(ctx->d_entry_dist); cudaFree(ctx->d_entry_idx); cudaFree(ctx->d_entry_cnt); cudaFree(ctx->d_entry_res);
    cudaFree(ctx->d_surf); cudaFreeHost(ctx->h_surf); cudaFree(ctx->d_pose_buf);
    for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 48; ++i) if (ctx->ev_fine[i]) cudaEventDestroy(ctx->ev_fine[i]);
    if (ctx->gexec) cudaGraphExecDestroy(ctx->gexec);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    if (ctx->cap_stream2) cudaStreamDestroy(ctx->cap_stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (!ctx->shared_args) { cudaFree(ctx->d_args); cudaFreeHost(ctx->h_args); }
    cudaFreeHost(ctx->h_markers); cudaFreeHost(ctx->h_summary); cudaFreeHost(ctx->h_entry_res);
    delete ctx;
    return MAMRI_OK;
}

extern "C" int mamri_create(mamri_ctx** out, int device, int32_t max_nx, int32_t max_ny, int32_t max_nz,
                            uint32_t max_runs, uint32_t max_markers) {
    if (!out) { snprintf(g_create_err, sizeof(g_create_err), "ctx output pointer is NULL"); return MAMRI_ERR_INVALID_ARG; }
    *out = nullptr;
    if (max_nx <= 0 || max_ny <= 0 || max_nz <= 0) {
        snprintf(g_create_err, sizeof(g_create_err), "max dimensions must be positive");
        return MAMRI_ERR_INVALID_ARG;
    }
    const unsigned long long voxels = (unsigned long long)max_nx * max_ny * max_nz;
    if (voxels >= (1ull << 32)) {
        snprintf(g_create_err, sizeof(g_create_err), "volumes of 2^32 voxels or more are not supported");
        return MAMRI_ERR_INVALID_ARG;
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) {
        snprintf(g_create_err, sizeof(g_create_err), "no usable CUDA device %d (%d visible); there is no CPU fallback",
                 device, n_dev);
        cudaGetLastError();
        return MAMRI_ERR_NO_DEVICE;
    }
    mamri_ctx* ctx = new (std::nothrow) mamri_ctx();
    if (!ctx) { snprintf(g_create_err, sizeof(g_create_err), "out of host memory"); return MAMRI_ERR_CUDA; }
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->max_nx = max_nx; ctx->max_ny = max_ny; ctx->max_nz = max_nz;
    if (max_runs == 0) {
        unsigned long long d = voxels / 16;
        if (d < (1u << 20)) d = 1u << 20;
        max_runs = uint32_t(d);
    }
    if (max_markers == 0) max_markers = 4096;
    ctx->max_runs = max_runs;
    ctx->max_markers = max_markers;
    const size_t W = (size_t(max_nx) + 31) / 32;
    ctx->cap_words = W * max_ny * max_nz;
    ctx->cap_pad_words = ((W + 2 + 3) & ~size_t(3)) * (size_t(max_ny) + 4 * MAMRI_RMAX) * (size_t(max_nz) + 4 * MAMRI_RMAX);
    if (3 * ctx->cap_pad_words >= (1ull << 32)) {
        snprintf(g_create_err, sizeof(g_create_err), "volume too large for 32-bit word indexing");
        delete ctx;
        return MAMRI_ERR_INVALID_ARG;
    }
    DeviceGuard g(device);
    auto bail = [&](cudaError_t e, const char* what) {
        snprintf(g_create_err, sizeof(g_create_err), "allocating %s failed: %s", what, cudaGetErrorString(e));
        mamri_destroy(ctx);
        return MAMRI_ERR_CUDA;
    };
    cudaError_t e;
#define ALLOC(ptr, bytes, what) if ((e = cudaMalloc((void**)&(ptr), (bytes))) != cudaSuccess) return bail(e, what)
    ALLOC(ctx->d_raw, ctx->cap_pad_words * 4, "raw mask");
    ALLOC(ctx->d_planes, 3 * ctx->cap_pad_words * 4, "morphology planes");
    ALLOC(ctx->d_dil, ctx->cap_pad_words * 4, "dilated mask");
    ALLOC(ctx->d_open, ctx->cap_pad_words * 4, "opening scratch");
    if ((e = cudaMemset(ctx->d_open, 0, ctx->cap_pad_words * 4)) != cudaSuccess) return bail(e, "opening scratch");
    ALLOC(ctx->d_closed, ctx->cap_words * 4, "closed mask");
    ctx->occ_cap = (size_t(max_ny) / 8 + 2) * (size_t(max_nz) / 4 + 2);
    ALLOC(ctx->d_occ_raw, ctx->occ_cap, "occupancy cells");
    ALLOC(ctx->d_occ_dil, ctx->occ_cap, "occupancy cells");
    if ((e = cudaMemset(ctx->d_occ_raw, 0, ctx->occ_cap)) != cudaSuccess) return bail(e, "occupancy cells");
    ALLOC(ctx->d_word_base, ctx->cap_words * 4, "run bases");
    ALLOC(ctx->d_run_pos, size_t(max_runs) * 4, "run positions");
    ALLOC(ctx->d_run_end, size_t(max_runs) * 4, "run lengths");
    ALLOC(ctx->d_parent, size_t(max_runs) * 4, "union-find parents");
    ALLOC(ctx->d_run_label, size_t(max_runs) * 4, "run labels");
    ALLOC(ctx->d_label_count, size_t(max_runs) * 4, "label counts");
    ALLOC(ctx->d_label_slot, size_t(max_runs) * 4, "label slots");
    ALLOC(ctx->d_root_count, size_t(max_runs) * 4, "root counts");
    {   // look-back states of the two single-pass scans (smallest tiles: ccl.cu k_runs_scan<4> = 2048 words, FR_TILE = 1024 runs)
        const size_t t_runs = ctx->cap_words / 2048 + 2, t_rank = size_t(max_runs) / 1024 + 2;
        ALLOC(ctx->d_scan_runs, t_runs * 8, "scan states");
        ALLOC(ctx->d_scan_rank, t_rank * 8, "scan states");
        if ((e = cudaMemset(ctx->d_scan_runs, 0, t_runs * 8)) != cudaSuccess) return bail(e, "scan states");
        if ((e = cudaMemset(ctx->d_scan_rank, 0, t_rank * 8)) != cudaSuccess) return bail(e, "scan states");
    }
    ALLOC(ctx->d_cand_label, (size_t(max_markers) + 1) * 4, "candidate labels");
    ALLOC(ctx->d_cand_rank, size_t(max_markers) * 4, "candidate ranks");
    ALLOC(ctx->d_cand_sums, (size_t(max_markers) + 1) * 9 * 8, "candidate sums");
    ALLOC(ctx->d_entry_dist, MAMRI_SCAN_CTAS * sizeof(double), "entry partials");
    ALLOC(ctx->d_entry_idx, MAMRI_SCAN_CTAS * sizeof(long long), "entry partials");
    ALLOC(ctx->d_entry_cnt, 2 * sizeof(unsigned long long), "entry counters");
    ALLOC(ctx->d_entry_res, sizeof(mamri_entry_result), "entry result");
    ALLOC(ctx->d_surf, sizeof(SurfScalars), "surface scalars");
#undef ALLOC
    if ((e = cudaMallocHost((void**)&ctx->h_markers, size_t(max_markers) * sizeof(mamri_marker))) != cudaSuccess)
        return bail(e, "pinned marker table");
    if ((e = cudaMallocHost((void**)&ctx->h_summary, sizeof(mamri_summary))) != cudaSuccess) return bail(e, "pinned summary");
    if ((e = cudaMallocHost((void**)&ctx->h_entry_res, sizeof(mamri_entry_result))) != cudaSuccess)
        return bail(e, "pinned entry result");
    if ((e = cudaMallocHost((void**)&ctx->h_surf, sizeof(SurfScalars))) != cudaSuccess) return bail(e, "pinned surface scalars");
    for (int i = 0; i < 6; ++i)
        if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) return bail(e, "events");
    if ((e = cudaMalloc((void**)&ctx->d_args, sizeof(ScanArgs))) != cudaSuccess) return bail(e, "scan arguments");
    if ((e = cudaMallocHost((void**)&ctx->h_args, sizeof(ScanArgs))) != cudaSuccess) return bail(e, "pinned scan arguments");
    memset(ctx->h_args, 0, sizeof(ScanArgs));
    ctx->d_dyn = &ctx->d_args->dyn; ctx->d_scalars = &ctx->d_args->sc; ctx->h_dyn = &ctx->h_args->dyn;
    // per-device function attributes (dynamic shared memory above 48 KB, cluster size): every context sets them for
    // its own device, so one process can drive several GPUs
    if ((e = segment_init_device()) != cudaSuccess) return bail(e, "kernel attributes");
    // The single-cluster labelling kernel (ccl.cu: k_label_cluster) is kept as an experiment: on the B200 it measured
    // slower than the scalable kernels even on small run tables (its 16 CTAs give every thread several runs to hook one
    // after the other, and chains grow between them), so it is off unless MAMRI_LABEL_CLUSTER=<cluster size> asks for it.
    ctx->max_cluster = 0;
    {
        const int probe = ccl_init_device();
        if (const char* lc = getenv("MAMRI_LABEL_CLUSTER")) {
            const int want = atoi(lc);
            ctx->max_cluster = want < probe ? (want < 0 ? 0 : want) : probe;
        }
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "capture stream");
    if ((e = cudaStreamCreateWithFlags(&ctx->cap_stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "capture stream");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "events");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "events");
    {
        const char* ng = getenv("MAMRI_NO_GRAPH");
        ctx->use_graph = !(ng && ng[0] == '1');
    }
    *out = ctx;
    return MAMRI_OK;
}

extern "C" int mamri_kernel_launches(const mamri_ctx* ctx) { return ctx ? ctx->n_launches : MAMRI_ERR_INVALID_ARG; }

extern "C" int mamri_set_profiling(mamri_ctx* ctx, int enable) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    ctx->profile = enable != 0;
    return MAMRI_OK;
}

extern "C" int mamri_stage_times(mamri_ctx* ctx, float* ms) {
    if (!ctx || !ms) return MAMRI_ERR_INVALID_ARG;
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "collect the pending detect first");
    DeviceGuard g(ctx->device);
    for (int i = 0; i < 5; ++i) CK(cudaEventElapsedTime(ms + i, ctx->ev[i], ctx->ev[i + 1]));
    return MAMRI_OK;
}

extern "C" int mamri_kernel_times(mamri_ctx* ctx, float* ms, const char** names, int max_n) {
    if (!ctx || !ms || !names) return MAMRI_ERR_INVALID_ARG;
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "collect the pending detect first");
    DeviceGuard g(ctx->device);
    int n = ctx->n_fine < max_n ? ctx->n_fine : max_n;
    for (int i = 0; i < n; ++i) {
        CK(cudaEventElapsedTime(ms + i, i == 0 ? ctx->ev[0] : ctx->ev_fine[i - 1], ctx->ev_fine[i]));
        names[i] = ctx->fine_name[i];
    }
    return n;
}

static int validate(mamri_ctx* ctx, const mamri_volume_desc* d, const mamri_params* p) {
    if (!d || !p) return fail(ctx, MAMRI_ERR_INVALID_ARG, "desc/params is NULL");
    if (d->nx <= 0 || d->ny <= 0 || d->nz <= 0) return fail(ctx, MAMRI_ERR_INVALID_ARG, "volume dimensions must be positive");
    if (d->nx > ctx->max_nx || d->ny > ctx->max_ny || d->nz > ctx->max_nz)
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "volume larger than the context was created for");
    if (dtype_size(d->dtype) == 0) return fail(ctx, MAMRI_ERR_INVALID_ARG, "unsupported voxel type");
    if (p->close_radius < 0 || p->close_radius > MAMRI_RMAX)
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "close_radius must be 0..3");
    if (p->open_radius < 0 || p->open_radius > MAMRI_RMAX)
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "open_radius must be 0..3");
    if (p->connectivity != 6 && p->connectivity != 26) return fail(ctx, MAMRI_ERR_INVALID_ARG, "connectivity must be 6 or 26");
    for (int i = 0; i < 3; ++i)
        if (!(d->spacing[i] > 0.0)) return fail(ctx, MAMRI_ERR_INVALID_ARG, "spacing must be positive");
    return MAMRI_OK;
}

// Profiling mode only (mamri_set_profiling): keeps the GPU busy for `ns` nanoseconds ahead of the first stage event, so
// that the host has every stage queued before the device gets there -- without it the first span (threshold) included
// the host's launch latency (~10 us of a 26 us kernel).  Never launched on the product path (graphs).
__global__ void k_prof_delay(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}

// Enqueues the stage kernels and the result copies on `s` (a real stream or one in capture mode).
// `fork` (capture only): once the labels are final, the per-voxel outputs are written on a second
// branch while the first computes the moments and copies the tables -- the two do not depend on each other.
static int enqueue_pipeline(mamri_ctx* ctx, const GraphKey& k, bool prof, bool fork, cudaStream_t s) {
    const mamri_volume_desc* desc = &k.desc;
    const mamri_params* params = &k.prm;
    const int nx = desc->nx, ny = desc->ny, nz = desc->nz;
    launch_counter() = 0;
    CK(cudaMemcpyAsync(ctx->d_args, ctx->h_args, sizeof(ScanArgs), cudaMemcpyHostToDevice, s));   // pointers + zeroed scalars
    if (prof) {
        k_prof_delay<<<1, 1, 0, s>>>(200000ull);          // 0.2 ms: the whole scan is queued by then
        ctx->n_fine = 0;
        CK(cudaEventRecord(ctx->ev[0], s));
    }
    const int geom_r = morph_geom_radius(params->open_radius, params->close_radius);
    CK(launch_threshold_pack(ctx, k.vol_aligned, desc->dtype, nx, ny, nz, params->lower, params->upper, geom_r, s));
    if (prof) CK(cudaEventRecord(ctx->ev[1], s));
    const uint32_t* mask = ctx->d_closed;
    CK(launch_opening(ctx, nx, ny, nz, params->open_radius, geom_r, s));
    CK(launch_closing(ctx, nx, ny, nz, params->close_radius, geom_r, s));
    if (prof) CK(cudaEventRecord(ctx->ev[2], s));
    CK(launch_label(ctx, mask, desc, params, s));
    if (prof) CK(cudaEventRecord(ctx->ev[3], s));
    const bool outputs = k.has_mask || k.has_labels || k.has_body;
    const bool forked = fork && (outputs || k.has_body_bits);
    // Statistics (marker table + summary, written straight into the pinned host buffers) and the per-voxel outputs
    // both depend on the labels only.  The statistics kernel goes FIRST, on the second branch: its few CTAs take their
    // slots before `materialise` floods the machine with thousands of CTAs (launched after it, a small kernel only
    // gets in when that grid has drained -- measured: it then ran after materialise instead of beside it).
    const bool early_beside = stats_early_beside();        // experiment knob (stats.cu): the sums of a noisy scan on the side branch too
    if (!early_beside) CK(launch_stats_early(ctx, desc, params, s));          // run tables of a noisy scan: the sums first (stats.cu)
    if (forked) {
        CK(cudaEventRecord(ctx->ev_fork, s));
        CK(cudaStreamWaitEvent(ctx->cap_stream2, ctx->ev_fork, 0));
        if (early_beside) CK(launch_stats_early(ctx, desc, params, ctx->cap_stream2));
        CK(launch_stats(ctx, desc, params, ctx->cap_stream2));
        CK(cudaEventRecord(ctx->ev_join, ctx->cap_stream2));
    } else {
        if (early_beside) CK(launch_stats_early(ctx, desc, params, s));
        CK(launch_stats(ctx, desc, params, s));
    }
    if (prof) CK(cudaEventRecord(ctx->ev[4], s));
    if (k.has_body_bits) CK(launch_body_bits(ctx, mask, nx, ny, nz, s));
    if (outputs) CK(launch_materialise(ctx, mask, nx, ny, nz, k.outs_aligned, s));
    if (prof) CK(cudaEventRecord(ctx->ev[5], s));
    if (forked) CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    ctx->n_launches = launch_counter();
    return MAMRI_OK;
}

static int detect_async_impl(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* d_volume,
                             const mamri_params* params, uint8_t* d_mask_out, uint32_t* d_labels_out,
                             uint8_t* d_body_out, uint32_t* d_body_bits_out, double* d_table, uint32_t table_slots, void* stream);

extern "C" int mamri_detect_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* d_volume,
                                  const mamri_params* params, uint8_t* d_mask_out, uint32_t* d_labels_out,
                                  uint8_t* d_body_out, void* stream) {
    return detect_async_impl(ctx, desc, d_volume, params, d_mask_out, d_labels_out, d_body_out, nullptr, nullptr, 0, stream);
}

extern "C" int mamri_detect_bits_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* d_volume,
                                       const mamri_params* params, uint8_t* d_mask_out, uint32_t* d_labels_out,
                                       uint32_t* d_body_bits_out, void* stream) {
    return detect_async_impl(ctx, desc, d_volume, params, d_mask_out, d_labels_out, nullptr, d_body_bits_out, nullptr, 0, stream);
}

static int detect_async_impl(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* d_volume,
                             const mamri_params* params, uint8_t* d_mask_out, uint32_t* d_labels_out,
                             uint8_t* d_body_out, uint32_t* d_body_bits_out, double* d_table, uint32_t table_slots, void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    int rc = validate(ctx, desc, params);
    if (rc != MAMRI_OK) return rc;
    if (!d_volume) return fail(ctx, MAMRI_ERR_INVALID_ARG, "volume pointer is NULL");
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "a detect is already pending on this context; collect it first");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GraphKey k;
    memset(&k, 0, sizeof(k));
    k.desc = *desc;
    k.prm = *params;
    k.vol_aligned = (reinterpret_cast<uintptr_t>(d_volume) & 15u) ? 0 : ((reinterpret_cast<uintptr_t>(d_volume) & 31u) ? 1 : 2);   // 16 / 32 bytes
    k.outs_aligned = ((reinterpret_cast<uintptr_t>(d_labels_out) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(d_mask_out) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_body_out) & 15u) == 0);
    k.has_mask = d_mask_out != nullptr; k.has_labels = d_labels_out != nullptr; k.has_body = d_body_out != nullptr;
    k.has_body_bits = d_body_bits_out != nullptr;
    k.run_ctas = ctx->run_ctas = run_grid_class(ctx->run_ctas, ctx->last_n_runs);
    ctx->slice_threads = slice_threads_class(ctx->slice_threads, ctx->last_n_runs, desc->nz);
    k.slice_threads = ctx->slice_threads < 0 ? -ctx->slice_threads : ctx->slice_threads;
    k.label_cluster = ctx->label_cluster = label_cluster_class(ctx->label_cluster, ctx->last_n_runs,
                                                               (unsigned long long)desc->nx * desc->ny * desc->nz, ctx->max_cluster);
    ctx->h_dyn->vol = d_volume;
    ctx->h_dyn->mask_out = d_mask_out;
    ctx->h_dyn->labels_out = d_labels_out;
    ctx->h_dyn->body_out = d_body_out;
    ctx->h_dyn->body_bits_out = d_body_bits_out;
    ctx->h_dyn->table_out = d_table;
    ctx->h_dyn->table_slots = d_table ? table_slots : 0u;
    ctx->h_dyn->gen = ++ctx->gen;            // generation 0 is the cleared state: never used
    if ((ctx->gen & 0x3FFFFFFFu) == 0) ctx->h_dyn->gen = ++ctx->gen;
    CK(prepare_raw_apron(ctx, desc->nx, desc->ny, desc->nz, morph_geom_radius(params->open_radius, params->close_radius), s));
    if (ctx->profile || !ctx->use_graph) {
        rc = enqueue_pipeline(ctx, k, ctx->profile, false, s);
        if (rc != MAMRI_OK) return rc;
    } else {
        if (!ctx->gvalid || memcmp(&k, &ctx->gkey, sizeof(k)) != 0) {
            // (re)capture: the pipeline is a fixed kernel sequence for a given geometry and parameter set
            if (ctx->gexec) { cudaGraphExecDestroy(ctx->gexec); ctx->gexec = nullptr; }
            ctx->gvalid = false;
            CK(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_pipeline(ctx, k, false, true, ctx->cap_stream);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(ctx->cap_stream, &graph);
            if (rc != MAMRI_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "graph capture failed: %s", cudaGetErrorString(e));
                return MAMRI_ERR_CUDA;
            }
            e = cudaGraphInstantiate(&ctx->gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "graph instantiation failed: %s", cudaGetErrorString(e));
                return MAMRI_ERR_CUDA;
            }
            ctx->gkey = k;
            ctx->gvalid = true;
        }
        CK(cudaGraphLaunch(ctx->gexec, s));
    }
    ctx->pending = true;
    ctx->pending_stream = s;
    ctx->last_desc = *desc;
    return MAMRI_OK;
}

// Host-buffer form: H2D copy of the volume into the context's staging buffer, the scan, D2H copy of the body labelmap
// (uint8, or 1 bit per voxel).  The staging buffers are sized on the first call at a size (mamri_reserve_staging does
// it ahead of time); a call that has to grow them synchronises the stream first.
static int detect_host_impl(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* h_volume, const mamri_params* params,
                            uint8_t* h_body_out, uint32_t* h_body_bits_out, void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    int rc = validate(ctx, desc, params);
    if (rc != MAMRI_OK) return rc;
    if (!h_volume) return fail(ctx, MAMRI_ERR_INVALID_ARG, "volume pointer is NULL");
    if (h_body_out && h_body_bits_out) return fail(ctx, MAMRI_ERR_INVALID_ARG, "ask for the body labelmap as uint8 or as bits, not both");
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "a detect is already pending on this context; collect it first");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = size_t(desc->nx) * desc->ny * desc->nz;
    const size_t in_bytes = n * dtype_size(desc->dtype);
    const size_t body_bytes = h_body_out ? n : (h_body_bits_out ? size_t((desc->nx + 31) / 32) * desc->ny * desc->nz * 4 : 0);
    if (ctx->stage_in_bytes < in_bytes || ctx->stage_body_bytes < body_bytes) {
        CK(cudaStreamSynchronize(s));
        rc = mamri_reserve_staging(ctx, in_bytes, body_bytes);
        if (rc != MAMRI_OK) return rc;
    }
    CK(cudaMemcpyAsync(ctx->d_stage_in, h_volume, in_bytes, cudaMemcpyHostToDevice, s));
    rc = detect_async_impl(ctx, desc, ctx->d_stage_in, params, nullptr, nullptr, h_body_out ? ctx->d_stage_body : nullptr,
                           h_body_bits_out ? reinterpret_cast<uint32_t*>(ctx->d_stage_body) : nullptr, nullptr, 0, stream);
    if (rc != MAMRI_OK) return rc;
    if (body_bytes) CK(cudaMemcpyAsync(h_body_out ? static_cast<void*>(h_body_out) : static_cast<void*>(h_body_bits_out), ctx->d_stage_body,
                                       body_bytes, cudaMemcpyDeviceToHost, s));
    return MAMRI_OK;
}

extern "C" int mamri_reserve_staging(mamri_ctx* ctx, size_t volume_bytes, size_t body_bytes) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "collect the pending detect first");
    DeviceGuard g(ctx->device);
    if (ctx->stage_in_bytes < volume_bytes) {
        cudaFree(ctx->d_stage_in);
        ctx->d_stage_in = nullptr; ctx->stage_in_bytes = 0;
        CK(cudaMalloc(&ctx->d_stage_in, volume_bytes));
        ctx->stage_in_bytes = volume_bytes;
    }
    if (ctx->stage_body_bytes < body_bytes) {
        cudaFree(ctx->d_stage_body);
        ctx->d_stage_body = nullptr; ctx->stage_body_bytes = 0;
        CK(cudaMalloc((void**)&ctx->d_stage_body, body_bytes));
        ctx->stage_body_bytes = body_bytes;
    }
    return MAMRI_OK;
}

extern "C" int mamri_detect_host_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* h_volume,
                                       const mamri_params* params, uint8_t* h_body_out, void* stream) {
    return detect_host_impl(ctx, desc, h_volume, params, h_body_out, nullptr, stream);
}

extern "C" int mamri_detect_host_bits_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* h_volume,
                                            const mamri_params* params, uint32_t* h_body_bits_out, void* stream) {
    return detect_host_impl(ctx, desc, h_volume, params, nullptr, h_body_bits_out, stream);
}

extern "C" int mamri_detect_collect(mamri_ctx* ctx, mamri_summary* summary, mamri_marker* h_markers, uint32_t max_markers) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (!ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "no detect pending on this context");
    DeviceGuard g(ctx->device);
    ctx->pending = false;
    CK(cudaStreamSynchronize(ctx->pending_stream));
    const mamri_summary* hs = ctx->h_summary;
    ctx->last_n_labels = hs->n_labels;
    ctx->last_n_runs = hs->device_status == MAMRI_OK ? hs->n_runs : ctx->max_runs;
    if (summary) *summary = *hs;
    if (hs->device_status != MAMRI_OK) {
        snprintf(ctx->err, sizeof(ctx->err),
                 "scan exceeds the context's capacity (runs limit %u, markers limit %u, %u markers found); "
                 "create the context with larger max_runs/max_markers",
                 ctx->max_runs, ctx->max_markers, hs->n_markers);
        return MAMRI_ERR_CAPACITY;
    }
    const uint32_t n = hs->n_markers;
    if (n > max_markers || (n && !h_markers)) {
        if (h_markers && max_markers) memcpy(h_markers, ctx->h_markers, size_t(max_markers) * sizeof(mamri_marker));
        return fail(ctx, MAMRI_ERR_CAPACITY, "caller's marker array is smaller than n_markers");
    }
    if (n) memcpy(h_markers, ctx->h_markers, size_t(n) * sizeof(mamri_marker));
    return MAMRI_OK;
}

// ------------------------------------------------------------------------------------------------
// pool: a batch of independent scans pipelined over a few contexts/streams (host loop in C++)
// ------------------------------------------------------------------------------------------------
struct WaveGraph {
    cudaGraphExec_t exec;
    GraphKey key;
    bool valid;
};

struct mamri_pool {
    int device, k;
    mamri_ctx** ctx;
    cudaStream_t* streams;
    cudaEvent_t fork;
    cudaEvent_t* join;
    // device-resident batches run as ONE captured graph per wave of up to k scans (see enqueue_wave)
    ScanArgs* d_args_all;            // [k], the contexts' d_args point into it -> one copy per wave sets the
    ScanArgs* h_args_all;            // [k] pinned                                 pointers and resets the scalars
    cudaStream_t hbm;                // capture origin
    cudaStream_t chain[4];           // the streaming kernels of scan i run on chain i % n_chains, one after another
    int n_chains;
    cudaEvent_t ev_chain[4];
    cudaEvent_t* ev_thr;             // [k] threshold of scan i done
    cudaEvent_t* ev_sel;             // [k] labels of scan i final
    cudaEvent_t* ev_done;            // [k] tables of scan i copied
    WaveGraph* waves;                // [k + 1], indexed by the number of scans in the wave
    bool use_wave_graph;
    bool trace;                      // MAMRI_WAVE_TRACE=1: timing events inside the wave graph, timeline on stderr
    cudaEvent_t* ev_trace;           // [k * 8]
    int pending_n;                   // scans enqueued by mamri_pool_detect_begin and not yet collected
    char err[512];
};

static char g_pool_err[512] = "";

extern "C" const char* mamri_pool_last_error(const mamri_pool* pool) { return pool ? pool->err : g_pool_err; }

extern "C" int mamri_pool_destroy(mamri_pool* pool) {
    if (!pool) return MAMRI_OK;
    DeviceGuard g(pool->device);
    for (int i = 0; i < pool->k; ++i) {
        if (pool->ctx && pool->ctx[i]) mamri_destroy(pool->ctx[i]);
        if (pool->streams && pool->streams[i]) cudaStreamDestroy(pool->streams[i]);
        if (pool->join && pool->join[i]) cudaEventDestroy(pool->join[i]);
    }
    for (int i = 0; i < pool->k; ++i) {
        if (pool->ev_thr && pool->ev_thr[i]) cudaEventDestroy(pool->ev_thr[i]);
        if (pool->ev_sel && pool->ev_sel[i]) cudaEventDestroy(pool->ev_sel[i]);
        if (pool->ev_done && pool->ev_done[i]) cudaEventDestroy(pool->ev_done[i]);
    }
    for (int i = 0; pool->waves && i <= pool->k; ++i)
        if (pool->waves[i].exec) cudaGraphExecDestroy(pool->waves[i].exec);
    if (pool->hbm) cudaStreamDestroy(pool->hbm);
    for (int i = 0; i < 4; ++i) {
        if (pool->chain[i]) cudaStreamDestroy(pool->chain[i]);
        if (pool->ev_chain[i]) cudaEventDestroy(pool->ev_chain[i]);
    }
    cudaFree(pool->d_args_all); cudaFreeHost(pool->h_args_all);
    if (pool->fork) cudaEventDestroy(pool->fork);
    delete[] pool->ctx; delete[] pool->streams; delete[] pool->join;
    delete[] pool->ev_thr; delete[] pool->ev_sel; delete[] pool->ev_done; delete[] pool->waves;
    delete pool;
    return MAMRI_OK;
}

extern "C" int mamri_pool_create(mamri_pool** out, int device, int32_t n_contexts, int32_t max_nx, int32_t max_ny,
                                 int32_t max_nz, uint32_t max_runs, uint32_t max_markers) {
    if (!out) { snprintf(g_pool_err, sizeof(g_pool_err), "pool output pointer is NULL"); return MAMRI_ERR_INVALID_ARG; }
    *out = nullptr;
    if (n_contexts < 1 || n_contexts > 64) {
        snprintf(g_pool_err, sizeof(g_pool_err), "n_contexts must be 1..64");
        return MAMRI_ERR_INVALID_ARG;
    }
    mamri_pool* p = new (std::nothrow) mamri_pool();
    if (!p) { snprintf(g_pool_err, sizeof(g_pool_err), "out of host memory"); return MAMRI_ERR_CUDA; }
    memset(p, 0, sizeof(*p));
    p->device = device; p->k = n_contexts;
    p->ctx = new (std::nothrow) mamri_ctx*[n_contexts]();
    p->streams = new (std::nothrow) cudaStream_t[n_contexts]();
    p->join = new (std::nothrow) cudaEvent_t[n_contexts]();
    p->ev_thr = new (std::nothrow) cudaEvent_t[n_contexts]();
    p->ev_sel = new (std::nothrow) cudaEvent_t[n_contexts]();
    p->ev_done = new (std::nothrow) cudaEvent_t[n_contexts]();
    p->waves = new (std::nothrow) WaveGraph[n_contexts + 1]();
    if (!p->ctx || !p->streams || !p->join || !p->ev_thr || !p->ev_sel || !p->ev_done || !p->waves) {
        snprintf(g_pool_err, sizeof(g_pool_err), "out of host memory");
        mamri_pool_destroy(p);
        return MAMRI_ERR_CUDA;
    }
    for (int i = 0; i < n_contexts; ++i) {
        int rc = mamri_create(&p->ctx[i], device, max_nx, max_ny, max_nz, max_runs, max_markers);
        if (rc != MAMRI_OK) {
            snprintf(g_pool_err, sizeof(g_pool_err), "context %d: %s", i, mamri_last_error(nullptr));
            mamri_pool_destroy(p);
            return rc;
        }
    }
    DeviceGuard g(device);
    cudaError_t e = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming);
    for (int i = 0; i < n_contexts && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->join[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_thr[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_sel[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->hbm, cudaStreamNonBlocking);
    p->n_chains = 2;
    if (const char* nc = getenv("MAMRI_HBM_CHAINS")) p->n_chains = atoi(nc) < 1 ? 1 : (atoi(nc) > 4 ? 4 : atoi(nc));
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&p->chain[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_chain[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_args_all, sizeof(ScanArgs) * n_contexts);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_args_all, sizeof(ScanArgs) * n_contexts);
    if (e == cudaSuccess) memset(p->h_args_all, 0, sizeof(ScanArgs) * n_contexts);
    if (e == cudaSuccess) {
        // the contexts' per-call arguments and scalars become slices of the pool's arrays
        for (int i = 0; i < n_contexts; ++i) {
            mamri_ctx* c = p->ctx[i];
            cudaFree(c->d_args); cudaFreeHost(c->h_args);
            c->d_args = p->d_args_all + i; c->h_args = p->h_args_all + i;
            c->d_dyn = &c->d_args->dyn; c->d_scalars = &c->d_args->sc; c->h_dyn = &c->h_args->dyn;
            c->shared_args = true;
        }
        const char* tr = getenv("MAMRI_WAVE_TRACE");
        p->trace = tr && tr[0] == '1';
        if (p->trace) {
            p->ev_trace = new (std::nothrow) cudaEvent_t[n_contexts * 8]();
            for (int i = 0; p->ev_trace && i < n_contexts * 8; ++i) cudaEventCreate(&p->ev_trace[i]);
        }
        const char* ng = getenv("MAMRI_NO_WAVE_GRAPH");
        p->use_wave_graph = !(ng && ng[0] == '1') && p->ctx[0]->use_graph;
    }
    if (e != cudaSuccess) {
        snprintf(g_pool_err, sizeof(g_pool_err), "creating the pool's streams failed: %s", cudaGetErrorString(e));
        mamri_pool_destroy(p);
        return MAMRI_ERR_CUDA;
    }
    *out = p;
    return MAMRI_OK;
}

extern "C" mamri_ctx* mamri_pool_context(mamri_pool* pool, int32_t k) {
    return (pool && k >= 0 && k < pool->k) ? pool->ctx[k] : nullptr;
}

#define CKP(call)                                                                                      \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            snprintf(pool->err, sizeof(pool->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                              \
            return MAMRI_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)

// One wave = up to k device-resident scans, scan i on context i, as ONE graph.  The two DRAM-bound
// kernels of every scan sit on a few chains (chain c: threshold of scans c, c + n_chains, ..., then their
// materialise kernels): with one or two of them in flight they reach close to their stand-alone bandwidth
// and fill each other's ramp-up / tail, whereas many of them sharing HBM slow each other down.  The
// latency-bound middle of scan i (closing, labelling, filter) branches off after its threshold and runs
// beside the chains; a chain reaches materialise(i) only after the thresholds of its later scans, by
// which time labels(i) are final, so the chains do not wait.  Moments + table copies of scan i form a
// further branch off its labels.
static int enqueue_wave(mamri_pool* pool, const GraphKey& k, int m) {
#define TRACE(i, j, st) do { if (pool->trace && pool->ev_trace) CKP(cudaEventRecordWithFlags(pool->ev_trace[(i) * 8 + (j)], st, cudaEventRecordExternal)); } while (0)
    const mamri_volume_desc* desc = &k.desc;
    const mamri_params* prm = &k.prm;
    const int nx = desc->nx, ny = desc->ny, nz = desc->nz;
    cudaStream_t H0 = pool->hbm;
    const int NC = pool->n_chains;
    // MAMRI_WAVE_MAT_CHAINS=1 (experiment): the materialise kernels get chains of their own (NC more streams) instead of
    // queueing behind the wave's thresholds on the same chains
    static const int mat_chains = [] { const char* e = getenv("MAMRI_WAVE_MAT_CHAINS"); return e ? atoi(e) : 0; }();
    const int MC = (mat_chains && NC <= 2) ? NC : 0;                   // first materialise chain = chain[MC ? NC : 0]
    const int geom_r = morph_geom_radius(prm->open_radius, prm->close_radius);
    launch_counter() = 0;
    CKP(cudaMemcpyAsync(pool->d_args_all, pool->h_args_all, sizeof(ScanArgs) * m, cudaMemcpyHostToDevice, H0));
    CKP(cudaEventRecord(pool->fork, H0));
    for (int c = 0; c < NC + MC; ++c) CKP(cudaStreamWaitEvent(pool->chain[c], pool->fork, 0));
    for (int i = 0; i < m; ++i) {
        cudaStream_t H = pool->chain[i % NC];
        TRACE(i, 0, H);
        CKP(launch_threshold_pack(pool->ctx[i], k.vol_aligned, desc->dtype, nx, ny, nz, prm->lower, prm->upper, geom_r, H));
        TRACE(i, 1, H);
        CKP(cudaEventRecord(pool->ev_thr[i], H));
    }
    const bool outputs = k.has_mask || k.has_labels || k.has_body;
    for (int i = 0; i < m; ++i) {
        mamri_ctx* c = pool->ctx[i];
        // MAMRI_WAVE_MID_CHAINS=n: the latency-bound middles of the wave's scans queue on n streams in scan order instead
        // of all sharing the machine, so that the first scans' labels are final early (0 = one stream per scan)
        static const int mid_chains = [] { const char* e = getenv("MAMRI_WAVE_MID_CHAINS"); return e ? atoi(e) : 0; }();
        cudaStream_t s = pool->streams[mid_chains > 0 ? i % mid_chains : i];
        CKP(cudaStreamWaitEvent(s, pool->ev_thr[i], 0));
        CKP(launch_opening(c, nx, ny, nz, prm->open_radius, geom_r, s));
        CKP(launch_closing(c, nx, ny, nz, prm->close_radius, geom_r, s));
        TRACE(i, 2, s);
        TRACE(i, 3, s);
        CKP(launch_label(c, c->d_closed, desc, prm, s));
        CKP(launch_stats_early(c, desc, prm, s));
        TRACE(i, 4, s);
        CKP(cudaEventRecord(pool->ev_sel[i], s));
        // statistics + tables (straight into the pinned host buffers) right behind the labels, before the wave's
        // materialise kernels are queued: a small kernel launched after those only gets in when they have drained
        CKP(launch_stats(c, desc, prm, s));
        TRACE(i, 7, s);
        CKP(cudaEventRecord(pool->ev_done[i], s));
    }
    if (outputs)
        for (int i = 0; i < m; ++i) {
            cudaStream_t H = pool->chain[(MC ? NC : 0) + i % NC];
            CKP(cudaStreamWaitEvent(H, pool->ev_sel[i], 0));
            TRACE(i, 5, H);
            CKP(launch_materialise(pool->ctx[i], pool->ctx[i]->d_closed, nx, ny, nz, k.outs_aligned, H));
            TRACE(i, 6, H);
        }
    for (int i = 0; i < m; ++i) CKP(cudaStreamWaitEvent(H0, pool->ev_done[i], 0));
    for (int c = 0; c < NC + MC; ++c) {
        CKP(cudaEventRecord(pool->ev_chain[c], pool->chain[c]));
        CKP(cudaStreamWaitEvent(H0, pool->ev_chain[c], 0));
    }
#undef TRACE
    for (int i = 0; i < m; ++i) pool->ctx[i]->n_launches = launch_counter() / m;
    return MAMRI_OK;
}

// Enqueues scans [first, first + m) as one wave on `cur` (scan first + i on context i).
static int pool_wave_launch(mamri_pool* pool, const GraphKey& k, const void* const* volumes, int first, int m,
                            uint8_t* const* mask_out, uint32_t* const* labels_out, uint8_t* const* body_out,
                            double* tables, uint32_t table_slots, cudaStream_t cur) {
    for (int i = 0; i < m; ++i) {
        mamri_ctx* c = pool->ctx[i];
        DynArgs* d = c->h_dyn;
        d->vol = volumes[first + i];
        d->mask_out = mask_out ? mask_out[first + i] : nullptr;
        d->labels_out = labels_out ? labels_out[first + i] : nullptr;
        d->body_out = body_out ? body_out[first + i] : nullptr;
        d->body_bits_out = nullptr;
        d->table_out = tables ? tables + size_t(first + i) * table_slots * 8 : nullptr;
        d->table_slots = tables ? table_slots : 0u;
        d->gen = ++c->gen;
        if ((c->gen & 0x3FFFFFFFu) == 0) d->gen = ++c->gen;
        CKP(prepare_raw_apron(c, k.desc.nx, k.desc.ny, k.desc.nz, morph_geom_radius(k.prm.open_radius, k.prm.close_radius), cur));
    }
    WaveGraph& w = pool->waves[m];
    if (!w.valid || memcmp(&w.key, &k, sizeof(k)) != 0) {
        if (w.exec) { cudaGraphExecDestroy(w.exec); w.exec = nullptr; }
        w.valid = false;
        CKP(cudaStreamBeginCapture(pool->hbm, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_wave(pool, k, m);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(pool->hbm, &graph);
        if (rc != MAMRI_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
        if (e != cudaSuccess) {
            snprintf(pool->err, sizeof(pool->err), "wave graph capture failed: %s", cudaGetErrorString(e));
            return MAMRI_ERR_CUDA;
        }
        const cudaError_t e2 = cudaGraphInstantiate(&w.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            snprintf(pool->err, sizeof(pool->err), "wave graph instantiation failed: %s", cudaGetErrorString(e2));
            return MAMRI_ERR_CUDA;
        }
        w.key = k;
        w.valid = true;
    }
    CKP(cudaGraphLaunch(w.exec, cur));
    if (pool->trace && pool->ev_trace) {
        static int dumps = 0;
        CKP(cudaStreamSynchronize(cur));
        if (++dumps == 12) {                         // one steady-state wave
            static const char* names[8] = {"thr>", "thr<", "clo<", "ccl<", "sel<", "mat>", "mat<", "mom<"};
            for (int i = 0; i < m; ++i) {
                fprintf(stderr, "scan %d:", i);
                for (int j = 0; j < 8; ++j) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, pool->ev_trace[0], pool->ev_trace[i * 8 + j]);
                    fprintf(stderr, " %s%7.1f", names[j], ms * 1e3f);
                }
                fprintf(stderr, "  us\n");
            }
        }
    }
    for (int i = 0; i < m; ++i) {
        mamri_ctx* c = pool->ctx[i];
        c->pending = true;
        c->pending_stream = cur;
        c->last_desc = k.desc;
    }
    return MAMRI_OK;
}

// Collects the m scans of the wave in flight into summaries[first ..] / markers.
static void pool_wave_collect(mamri_pool* pool, int first, int m, mamri_summary* summaries, mamri_marker* markers,
                              uint32_t max_m, int* first_err) {
    for (int i = 0; i < m; ++i) {
        mamri_ctx* c = pool->ctx[i];
        const int rc = mamri_detect_collect(c, &summaries[first + i], markers ? markers + size_t(first + i) * max_m : nullptr, max_m);
        if (rc != MAMRI_OK && *first_err == MAMRI_OK) {
            *first_err = rc;
            snprintf(pool->err, sizeof(pool->err), "scan %d: %s", first + i, mamri_last_error(c));
        }
    }
}

// Validates a device-resident batch and fills the key of its wave graph.
static int wave_key(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* volumes, int n, const mamri_params* params,
                    uint8_t* const* mask_out, uint32_t* const* labels_out, uint8_t* const* body_out, GraphKey& k) {
    auto pfail = [&](int code, const char* msg) { snprintf(pool->err, sizeof(pool->err), "%s", msg); return code; };
    int rc = validate(pool->ctx[0], desc, params);
    if (rc != MAMRI_OK) return pfail(rc, mamri_last_error(pool->ctx[0]));
    memset(&k, 0, sizeof(k));
    k.desc = *desc;
    k.prm = *params;
    k.has_mask = mask_out != nullptr; k.has_labels = labels_out != nullptr; k.has_body = body_out != nullptr;
    k.vol_aligned = 2; k.outs_aligned = 1;
    {   // one grid class for the wave, from the largest run count the pool's contexts have seen last
        uint32_t hint = 0;
        for (int j = 0; j < pool->k; ++j) hint = pool->ctx[j]->last_n_runs > hint ? pool->ctx[j]->last_n_runs : hint;
        k.run_ctas = run_grid_class(pool->ctx[0]->run_ctas, hint);
        const int st = slice_threads_class(pool->ctx[0]->slice_threads, hint, desc->nz);
        k.slice_threads = st < 0 ? -st : st;
        k.label_cluster = label_cluster_class(pool->ctx[0]->label_cluster, hint, (unsigned long long)desc->nx * desc->ny * desc->nz,
                                              pool->ctx[0]->max_cluster);
        for (int j = 0; j < pool->k; ++j) {
            pool->ctx[j]->run_ctas = k.run_ctas; pool->ctx[j]->slice_threads = st; pool->ctx[j]->label_cluster = k.label_cluster;
        }
    }
    for (int i = 0; i < n; ++i) {
        if (!volumes[i]) return pfail(MAMRI_ERR_INVALID_ARG, "volume pointer is NULL");
        if (reinterpret_cast<uintptr_t>(volumes[i]) & 15u) k.vol_aligned = 0;
        else if ((reinterpret_cast<uintptr_t>(volumes[i]) & 31u) && k.vol_aligned > 1) k.vol_aligned = 1;
        if ((mask_out && (reinterpret_cast<uintptr_t>(mask_out[i]) & 15u)) ||
            (labels_out && (reinterpret_cast<uintptr_t>(labels_out[i]) & 15u)) ||
            (body_out && (reinterpret_cast<uintptr_t>(body_out[i]) & 15u)))
            k.outs_aligned = 0;
    }
    return MAMRI_OK;
}

static int pool_run(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* volumes, bool host, int32_t n,
                    const mamri_params* params, uint8_t* const* mask_out, uint32_t* const* labels_out,
                    uint8_t* const* body_out, mamri_summary* summaries, mamri_marker* markers, uint32_t max_m,
                    void* stream) {
    if (!pool) return MAMRI_ERR_INVALID_ARG;
    auto pfail = [&](int code, const char* msg) { snprintf(pool->err, sizeof(pool->err), "%s", msg); return code; };
    if (n < 0 || (n > 0 && (!volumes || !summaries))) return pfail(MAMRI_ERR_INVALID_ARG, "bad batch arguments");
    if (max_m > 0 && !markers) return pfail(MAMRI_ERR_INVALID_ARG, "markers array is NULL");
    if (n == 0) return MAMRI_OK;
    DeviceGuard g(pool->device);
    cudaStream_t cur = static_cast<cudaStream_t>(stream);
    const int K = pool->k;
    int first_err = MAMRI_OK;
    bool profiling = false;
    for (int j = 0; j < K; ++j) profiling = profiling || pool->ctx[j]->profile || pool->ctx[j]->pending;
    if (pool->pending_n) return pfail(MAMRI_ERR_STATE, "a batch begun with mamri_pool_detect_begin is pending; end it first");
    if (!host && pool->use_wave_graph && !profiling) {
        GraphKey k;
        int rc = wave_key(pool, desc, volumes, n, params, mask_out, labels_out, body_out, k);
        if (rc != MAMRI_OK) return rc;
        for (int first = 0; first < n; first += K) {
            const int m = n - first < K ? n - first : K;
            rc = pool_wave_launch(pool, k, volumes, first, m, mask_out, labels_out, body_out, nullptr, 0, cur);
            if (rc != MAMRI_OK) return rc;
            pool_wave_collect(pool, first, m, summaries, markers, max_m, &first_err);
        }
        return first_err;
    }
    cudaError_t e = cudaEventRecord(pool->fork, cur);
    for (int j = 0; j < K && j < n && e == cudaSuccess; ++j) e = cudaStreamWaitEvent(pool->streams[j], pool->fork, 0);
    if (e != cudaSuccess) return pfail(MAMRI_ERR_CUDA, cudaGetErrorString(e));
    auto note = [&](int rc, int scan, mamri_ctx* c) {
        if (rc != MAMRI_OK && first_err == MAMRI_OK) {
            first_err = rc;
            snprintf(pool->err, sizeof(pool->err), "scan %d: %s", scan, mamri_last_error(c));
        }
    };
    auto collect = [&](int i) {
        mamri_ctx* c = pool->ctx[i % K];
        if (!c->pending) { memset(&summaries[i], 0, sizeof(mamri_summary)); summaries[i].device_status = MAMRI_ERR_STATE; return; }
        int rc = mamri_detect_collect(c, &summaries[i], markers ? markers + size_t(i) * max_m : nullptr, max_m);
        note(rc, i, c);
    };
    for (int i = 0; i < n; ++i) {
        const int j = i % K;
        if (i >= K) collect(i - K);
        mamri_ctx* c = pool->ctx[j];
        int rc;
        if (host)
            rc = mamri_detect_host_async(c, desc, volumes[i], params, body_out ? body_out[i] : nullptr, pool->streams[j]);
        else
            rc = mamri_detect_async(c, desc, volumes[i], params, mask_out ? mask_out[i] : nullptr,
                                    labels_out ? labels_out[i] : nullptr, body_out ? body_out[i] : nullptr, pool->streams[j]);
        note(rc, i, c);
    }
    for (int i = (n > K ? n - K : 0); i < n; ++i) collect(i);
    for (int j = 0; j < K && j < n; ++j) {
        if (cudaEventRecord(pool->join[j], pool->streams[j]) != cudaSuccess ||
            cudaStreamWaitEvent(cur, pool->join[j], 0) != cudaSuccess)
            return pfail(MAMRI_ERR_CUDA, "joining the pool's streams failed");
    }
    return first_err;
}

extern "C" int mamri_pool_detect(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* d_volumes, int32_t n,
                                 const mamri_params* params, uint8_t* const* d_mask_out, uint32_t* const* d_labels_out,
                                 uint8_t* const* d_body_out, mamri_summary* summaries, mamri_marker* markers,
                                 uint32_t max_markers_per_scan, void* stream) {
    return pool_run(pool, desc, d_volumes, false, n, params, d_mask_out, d_labels_out, d_body_out, summaries, markers,
                    max_markers_per_scan, stream);
}

extern "C" int mamri_pool_detect_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* d_volumes, int32_t n,
                                       const mamri_params* params, uint8_t* const* d_mask_out, uint32_t* const* d_labels_out,
                                       uint8_t* const* d_body_out, double* d_tables, uint32_t table_slots, void* stream) {
    if (!pool) return MAMRI_ERR_INVALID_ARG;
    auto pfail = [&](int code, const char* msg) { snprintf(pool->err, sizeof(pool->err), "%s", msg); return code; };
    if (n < 1 || n > pool->k || !d_volumes) return pfail(MAMRI_ERR_INVALID_ARG, "begin/end handles 1..n_contexts scans per call");
    if (d_tables && table_slots == 0) return pfail(MAMRI_ERR_INVALID_ARG, "table_slots must be positive");
    if (pool->pending_n) return pfail(MAMRI_ERR_STATE, "a batch is already pending; end it first");
    for (int j = 0; j < pool->k; ++j)
        if (pool->ctx[j]->pending) return pfail(MAMRI_ERR_STATE, "a context of the pool has a detect pending");
    DeviceGuard g(pool->device);
    cudaStream_t cur = static_cast<cudaStream_t>(stream);
    bool profiling = false;
    for (int j = 0; j < pool->k; ++j) profiling = profiling || pool->ctx[j]->profile;
    if (pool->use_wave_graph && !profiling) {
        GraphKey k;
        int rc = wave_key(pool, desc, d_volumes, n, params, d_mask_out, d_labels_out, d_body_out, k);
        if (rc != MAMRI_OK) return rc;
        rc = pool_wave_launch(pool, k, d_volumes, 0, n, d_mask_out, d_labels_out, d_body_out, d_tables, table_slots, cur);
        if (rc != MAMRI_OK) return rc;
    } else {                                          // no wave graph: one context and stream per scan
        if (cudaEventRecord(pool->fork, cur) != cudaSuccess) return pfail(MAMRI_ERR_CUDA, "recording the fork event failed");
        for (int i = 0; i < n; ++i) {
            if (cudaStreamWaitEvent(pool->streams[i], pool->fork, 0) != cudaSuccess) return pfail(MAMRI_ERR_CUDA, "forking failed");
            const int rc = detect_async_impl(pool->ctx[i], desc, d_volumes[i], params, d_mask_out ? d_mask_out[i] : nullptr,
                                             d_labels_out ? d_labels_out[i] : nullptr, d_body_out ? d_body_out[i] : nullptr, nullptr,
                                             d_tables ? d_tables + size_t(i) * table_slots * 8 : nullptr, table_slots, pool->streams[i]);
            if (rc != MAMRI_OK) {
                snprintf(pool->err, sizeof(pool->err), "scan %d: %s", i, mamri_last_error(pool->ctx[i]));
                for (int j = 0; j < i; ++j) { mamri_summary tmp; mamri_detect_collect(pool->ctx[j], &tmp, nullptr, 0); }
                return rc;
            }
            if (cudaEventRecord(pool->join[i], pool->streams[i]) != cudaSuccess ||
                cudaStreamWaitEvent(cur, pool->join[i], 0) != cudaSuccess)
                return pfail(MAMRI_ERR_CUDA, "joining the pool's streams failed");
        }
    }
    pool->pending_n = n;
    return MAMRI_OK;
}

static int pool_host_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes, int32_t n,
                           const mamri_params* params, uint8_t* const* h_body_out, uint32_t* const* h_body_bits_out, void* stream) {
    if (!pool) return MAMRI_ERR_INVALID_ARG;
    auto pfail = [&](int code, const char* msg) { snprintf(pool->err, sizeof(pool->err), "%s", msg); return code; };
    if (n < 1 || n > pool->k || !h_volumes) return pfail(MAMRI_ERR_INVALID_ARG, "begin/end handles 1..n_contexts scans per call");
    if (pool->pending_n) return pfail(MAMRI_ERR_STATE, "a batch is already pending; end it first");
    for (int j = 0; j < pool->k; ++j)
        if (pool->ctx[j]->pending) return pfail(MAMRI_ERR_STATE, "a context of the pool has a detect pending");
    DeviceGuard g(pool->device);
    cudaStream_t cur = static_cast<cudaStream_t>(stream);
    if (cudaEventRecord(pool->fork, cur) != cudaSuccess) return pfail(MAMRI_ERR_CUDA, "recording the fork event failed");
    for (int i = 0; i < n; ++i) {                     // scan i: H2D, kernels, body-mask D2H on stream i
        if (cudaStreamWaitEvent(pool->streams[i], pool->fork, 0) != cudaSuccess) return pfail(MAMRI_ERR_CUDA, "forking failed");
        const int rc = detect_host_impl(pool->ctx[i], desc, h_volumes[i], params, h_body_out ? h_body_out[i] : nullptr,
                                        h_body_bits_out ? h_body_bits_out[i] : nullptr, pool->streams[i]);
        if (rc != MAMRI_OK) {
            snprintf(pool->err, sizeof(pool->err), "scan %d: %s", i, mamri_last_error(pool->ctx[i]));
            for (int j = 0; j < i; ++j) { mamri_summary tmp; mamri_detect_collect(pool->ctx[j], &tmp, nullptr, 0); }
            return rc;
        }
        if (cudaEventRecord(pool->join[i], pool->streams[i]) != cudaSuccess ||
            cudaStreamWaitEvent(cur, pool->join[i], 0) != cudaSuccess)
            return pfail(MAMRI_ERR_CUDA, "joining the pool's streams failed");
    }
    pool->pending_n = n;
    return MAMRI_OK;
}

extern "C" int mamri_pool_detect_host_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes,
                                            int32_t n, const mamri_params* params, uint8_t* const* h_body_out, void* stream) {
    return pool_host_begin(pool, desc, h_volumes, n, params, h_body_out, nullptr, stream);
}

extern "C" int mamri_pool_detect_host_bits_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes,
                                                 int32_t n, const mamri_params* params, uint32_t* const* h_body_bits_out, void* stream) {
    return pool_host_begin(pool, desc, h_volumes, n, params, nullptr, h_body_bits_out, stream);
}

extern "C" int mamri_pool_detect_end(mamri_pool* pool, mamri_summary* summaries, mamri_marker* markers,
                                     uint32_t max_markers_per_scan) {
    if (!pool) return MAMRI_ERR_INVALID_ARG;
    auto pfail = [&](int code, const char* msg) { snprintf(pool->err, sizeof(pool->err), "%s", msg); return code; };
    if (!pool->pending_n) return pfail(MAMRI_ERR_STATE, "no batch pending on this pool");
    if (!summaries || (max_markers_per_scan > 0 && !markers)) return pfail(MAMRI_ERR_INVALID_ARG, "bad result arrays");
    DeviceGuard g(pool->device);
    int first_err = MAMRI_OK;
    const int n = pool->pending_n;
    pool->pending_n = 0;
    pool_wave_collect(pool, 0, n, summaries, markers, max_markers_per_scan, &first_err);
    return first_err;
}

extern "C" int mamri_pool_detect_host(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes,
                                      int32_t n, const mamri_params* params, uint8_t* const* h_body_out,
                                      mamri_summary* summaries, mamri_marker* markers, uint32_t max_markers_per_scan,
                                      void* stream) {
    return pool_run(pool, desc, h_volumes, true, n, params, nullptr, nullptr, h_body_out, summaries, markers,
                    max_markers_per_scan, stream);
}

extern "C" int mamri_label_counts(mamri_ctx* ctx, uint32_t* h_counts, uint32_t max_labels) {
    if (!ctx || !h_counts) return MAMRI_ERR_INVALID_ARG;
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "collect the pending detect first");
    DeviceGuard g(ctx->device);
    uint32_t n = ctx->last_n_labels < max_labels ? ctx->last_n_labels : max_labels;
    if (n) CK(cudaMemcpy(h_counts, ctx->d_label_count, size_t(n) * 4, cudaMemcpyDeviceToHost));
    return ctx->last_n_labels <= max_labels ? MAMRI_OK : fail(ctx, MAMRI_ERR_CAPACITY, "h_counts smaller than n_labels");
}

extern "C" int mamri_entry_search(mamri_ctx* ctx, const float* d_points, const float* d_normals, int64_t n,
                                  const double target[3], double radius, double wx, double wy, double cutoff,
                                  int32_t n_path_samples, const uint8_t* d_path_mask, const mamri_volume_desc* mask_desc,
                                  const double ras_to_index[12], int32_t path_free_value, mamri_entry_result* result,
                                  void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (!result || !target) return fail(ctx, MAMRI_ERR_INVALID_ARG, "result/target is NULL");
    if (n < 0 || (n > 0 && (!d_points || !d_normals))) return fail(ctx, MAMRI_ERR_INVALID_ARG, "bad candidate arrays");
    if (n_path_samples < 0) return fail(ctx, MAMRI_ERR_INVALID_ARG, "n_path_samples must not be negative");
    if (n_path_samples > 0 && (!d_path_mask || !mask_desc || !ras_to_index))
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "path sampling (n_path_samples > 0) needs d_path_mask, mask_desc and ras_to_index");
    if (n_path_samples > 0 && (mask_desc->nx <= 0 || mask_desc->ny <= 0 || mask_desc->nz <= 0))
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "path mask dimensions must be positive");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        memset(result, 0, sizeof(*result));
        result->index = -1;
        result->distance = INFINITY;
        return MAMRI_OK;
    }
    const bool sample = n_path_samples > 0;
    CK(launch_entry_search(ctx, d_points, d_normals, n, target, radius, wx, wy, cutoff, sample ? n_path_samples : 0,
                           sample ? d_path_mask : nullptr, sample ? mask_desc->nx : 0, sample ? mask_desc->ny : 0,
                           sample ? mask_desc->nz : 0, ras_to_index, path_free_value, s));
    CK(cudaMemcpyAsync(ctx->h_entry_res, ctx->d_entry_res, sizeof(mamri_entry_result), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *result = *ctx->h_entry_res;
    return MAMRI_OK;
}

extern "C" int mamri_body_surface(mamri_ctx* ctx, const mamri_volume_desc* desc, const uint8_t* d_body_mask,
                                  float* d_points_out, float* d_normals_out, int64_t capacity, int64_t* n_points,
                                  int64_t* n_body_voxels, void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (!desc || !n_points) return fail(ctx, MAMRI_ERR_INVALID_ARG, "desc/n_points is NULL");
    if (desc->nx <= 0 || desc->ny <= 0 || desc->nz <= 0 || desc->nx > ctx->max_nx || desc->ny > ctx->max_ny || desc->nz > ctx->max_nz)
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "volume dimensions outside what the context was created for");
    for (int i = 0; i < 3; ++i)
        if (!(desc->spacing[i] > 0.0)) return fail(ctx, MAMRI_ERR_INVALID_ARG, "spacing must be positive");
    if (capacity < 0 || (capacity > 0 && (!d_points_out || !d_normals_out)))
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "capacity > 0 needs both output arrays");
    if (ctx->pending) return fail(ctx, MAMRI_ERR_STATE, "collect the pending detect first");
    uint32_t body_label = 0;
    if (!d_body_mask) {                                   // body of the last collected scan, from its run table
        const mamri_volume_desc& l = ctx->last_desc;
        if (ctx->last_n_labels == 0 && ctx->h_summary->body_label == 0) { /* empty scan: no candidates */ }
        else if (l.nx != desc->nx || l.ny != desc->ny || l.nz != desc->nz)
            return fail(ctx, MAMRI_ERR_STATE, "d_body_mask == NULL needs the geometry of the last collected scan");
        body_label = ctx->last_n_labels ? ctx->h_summary->body_label : 0;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(launch_body_surface(ctx, desc, d_body_mask, body_label, capacity > 0 ? d_points_out : nullptr,
                           capacity > 0 ? d_normals_out : nullptr, (unsigned long long)capacity, s));
    CK(cudaMemcpyAsync(ctx->h_surf, ctx->d_surf, sizeof(SurfScalars), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *n_points = (int64_t)ctx->h_surf->n_points;
    if (n_body_voxels) *n_body_voxels = (int64_t)ctx->h_surf->n_body;
    if (capacity > 0 && *n_points > capacity)
        return fail(ctx, MAMRI_ERR_CAPACITY, "more surface voxels than the output arrays hold (n_points is the number needed)");
    return MAMRI_OK;
}

extern "C" void mamri_default_robot(mamri_robot* r) {
    if (!r) return;
    memset(r, 0, sizeof(*r));
    // Mamri/Resources/Robot/robot_config.json, in file order
    struct Row { int parent, axis, markers, chain; double t[3], mc[9], arm[2], lim[2]; };
    static const Row rows[8] = {
        {-1, MAMRI_AXIS_NONE, 1, -1, {0, 0, 0}, {-10, 20, 5, 10, 20, 5, -10, -20, 5}, {40, 20}, {0, 0}},                 // Baseplate :2-17
        {0, MAMRI_AXIS_IS, 0, 0, {0, 0, 20}, {0}, {0, 0}, {-180, 180}},                                                  // Joint1   :18-30
        {1, MAMRI_AXIS_PA, 1, 1, {0, 0, 30}, {12.5, 45, 110, -12.5, 45, 110, 12.5, 45, 40}, {70, 25}, {-120, 120}},      // Joint2   :31-49
        {2, MAMRI_AXIS_PA, 0, 2, {0, 0, 150}, {0}, {0, 0}, {-120, 120}},                                                 // Joint3   :50-62
        {3, MAMRI_AXIS_IS, 1, 3, {0, 0, 0}, {-10, 35, 90, 10, 35, 90, -10, -35, 90}, {70, 20}, {-180, 180}},             // Joint4   :63-81
        {4, MAMRI_AXIS_PA, 0, 4, {0, 0, 155}, {0}, {0, 0}, {-120, 120}},                                                 // Joint5   :82-94
        {5, MAMRI_AXIS_IS, 1, 5, {0, 0, 13}, {-10, 22.5, 26, 10, 22.5, 26, -10, -22.5, 26}, {45, 20}, {-270, 270}},      // Joint6   :95-115
        {6, MAMRI_AXIS_TRANS, 0, -1, {-50, 0, 71}, {0}, {0, 0}, {0, 0}},                                                 // Needle   :116-130
    };
    r->n_links = 8;
    r->base_link = 0; r->effector_link = 6; r->secondary_link = 4;
    r->distance_tolerance = 5.0;     // Mamri.py:813
    r->secondary_weight = 0.05;      // Mamri.py:1507
    r->apply_correction = 0;
    for (int i = 0; i < 8; ++i) {
        mamri_link& l = r->links[i];
        l.parent = rows[i].parent; l.axis = rows[i].axis; l.has_markers = rows[i].markers; l.chain_index = rows[i].chain;
        for (int k = 0; k < 3; ++k) l.translate[k] = rows[i].t[k];
        for (int k = 0; k < 9; ++k) l.marker_coords[k] = rows[i].mc[k];
        for (int k = 0; k < 2; ++k) { l.arm_lengths[k] = rows[i].arm[k]; l.limits_deg[k] = rows[i].lim[k]; }
    }
}

extern "C" int mamri_pose_estimate_ex(mamri_ctx* ctx, const mamri_robot* robot, const double* h_points_ras,
                                      const int32_t* h_counts, int32_t n_scans, int32_t max_points,
                                      const mamri_pose_options* options, mamri_pose* h_poses, void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (!robot || n_scans < 0 || max_points < 0 || (n_scans > 0 && (!h_counts || !h_poses)) ||
        (n_scans > 0 && max_points > 0 && !h_points_ras))
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "bad pose arguments");
    if (robot->n_links < 1 || robot->n_links > MAMRI_MAX_LINKS) return fail(ctx, MAMRI_ERR_INVALID_ARG, "robot: n_links out of range");
    for (int l = 0; l < robot->n_links; ++l)
        if (robot->links[l].parent >= l || robot->links[l].chain_index >= MAMRI_MAX_CHAIN)
            return fail(ctx, MAMRI_ERR_INVALID_ARG, "robot: a link's parent must precede it; chain_index < MAMRI_MAX_CHAIN");
    if (n_scans == 0) return MAMRI_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double* h_saved = options ? options->h_saved_base : nullptr;
    const double* h_init = options ? options->h_initial_angles : nullptr;
    const size_t off_pts = (sizeof(mamri_robot) + 255) & ~size_t(255);
    const size_t pts_bytes = size_t(n_scans) * max_points * 3 * sizeof(double);
    const size_t off_cnt = (off_pts + pts_bytes + 255) & ~size_t(255);
    const size_t off_pose = (off_cnt + size_t(n_scans) * sizeof(int32_t) + 255) & ~size_t(255);
    const size_t off_base = (off_pose + size_t(n_scans) * sizeof(mamri_pose) + 255) & ~size_t(255);
    const size_t off_init = off_base + 16 * sizeof(double);
    const size_t init_bytes = size_t(n_scans) * MAMRI_MAX_CHAIN * sizeof(double);
    const size_t total = off_init + init_bytes;
    if (ctx->pose_buf_bytes < total) {                   // first call at this batch size: grow the scratch
        CK(cudaStreamSynchronize(s));
        cudaFree(ctx->d_pose_buf);
        ctx->d_pose_buf = nullptr; ctx->pose_buf_bytes = 0;
        CK(cudaMalloc(&ctx->d_pose_buf, total));
        ctx->pose_buf_bytes = total;
    }
    char* base = static_cast<char*>(ctx->d_pose_buf);
    CK(cudaMemcpyAsync(base, robot, sizeof(mamri_robot), cudaMemcpyHostToDevice, s));
    if (pts_bytes) CK(cudaMemcpyAsync(base + off_pts, h_points_ras, pts_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(base + off_cnt, h_counts, size_t(n_scans) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (h_saved) CK(cudaMemcpyAsync(base + off_base, h_saved, 16 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h_init) CK(cudaMemcpyAsync(base + off_init, h_init, init_bytes, cudaMemcpyHostToDevice, s));
    CK(launch_pose(reinterpret_cast<const mamri_robot*>(base), reinterpret_cast<const double*>(base + off_pts),
                   reinterpret_cast<const int32_t*>(base + off_cnt), n_scans, max_points,
                   h_init ? reinterpret_cast<const double*>(base + off_init) : nullptr,
                   h_saved ? reinterpret_cast<const double*>(base + off_base) : nullptr,
                   options ? options->prefer_saved_base : 0, reinterpret_cast<mamri_pose*>(base + off_pose), s));
    CK(cudaMemcpyAsync(h_poses, base + off_pose, size_t(n_scans) * sizeof(mamri_pose), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MAMRI_OK;
}

extern "C" int mamri_pose_estimate(mamri_ctx* ctx, const mamri_robot* robot, const double* h_points_ras,
                                   const int32_t* h_counts, int32_t n_scans, int32_t max_points, mamri_pose* h_poses,
                                   void* stream) {
    return mamri_pose_estimate_ex(ctx, robot, h_points_ras, h_counts, n_scans, max_points, nullptr, h_poses, stream);
}

extern "C" int mamri_pose_from_tables(mamri_ctx* ctx, const mamri_robot* robot, const double* d_tables, int32_t n_scans,
                                      uint32_t table_slots, mamri_pose* h_poses, void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (!robot || n_scans < 0 || (n_scans > 0 && (!d_tables || !h_poses || table_slots == 0)))
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "bad pose arguments");
    if (robot->n_links < 1 || robot->n_links > MAMRI_MAX_LINKS) return fail(ctx, MAMRI_ERR_INVALID_ARG, "robot: n_links out of range");
    for (int l = 0; l < robot->n_links; ++l)
        if (robot->links[l].parent >= l || robot->links[l].chain_index >= MAMRI_MAX_CHAIN)
            return fail(ctx, MAMRI_ERR_INVALID_ARG, "robot: a link's parent must precede it; chain_index < MAMRI_MAX_CHAIN");
    if (n_scans == 0) return MAMRI_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t off_pose = (sizeof(mamri_robot) + 255) & ~size_t(255);
    const size_t total = off_pose + size_t(n_scans) * sizeof(mamri_pose);
    if (ctx->pose_buf_bytes < total) {
        CK(cudaStreamSynchronize(s));
        cudaFree(ctx->d_pose_buf);
        ctx->d_pose_buf = nullptr; ctx->pose_buf_bytes = 0;
        CK(cudaMalloc(&ctx->d_pose_buf, total));
        ctx->pose_buf_bytes = total;
    }
    char* base = static_cast<char*>(ctx->d_pose_buf);
    CK(cudaMemcpyAsync(base, robot, sizeof(mamri_robot), cudaMemcpyHostToDevice, s));
    CK(launch_pose(reinterpret_cast<const mamri_robot*>(base), d_tables, nullptr, n_scans, int(table_slots), nullptr, nullptr, 0,
                   reinterpret_cast<mamri_pose*>(base + off_pose), s));
    CK(cudaMemcpyAsync(h_poses, base + off_pose, size_t(n_scans) * sizeof(mamri_pose), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MAMRI_OK;
}

extern "C" int mamri_collision_check(mamri_ctx* ctx, const mamri_robot* robot, const double base_matrix[16],
                                     const double* h_joint_angles, int32_t n_configs, const float* d_part_points,
                                     const int32_t* h_part_offsets, const uint8_t* d_body_mask,
                                     const mamri_volume_desc* mask_desc, const double ras_to_index[12],
                                     mamri_collision_result* h_results, void* stream) {
    if (!ctx) return MAMRI_ERR_INVALID_ARG;
    if (!robot || !base_matrix || n_configs < 0 || !h_part_offsets || !d_body_mask || !mask_desc || !ras_to_index ||
        (n_configs > 0 && (!h_joint_angles || !h_results)))
        return fail(ctx, MAMRI_ERR_INVALID_ARG, "bad collision arguments");
    if (robot->n_links < 1 || robot->n_links > MAMRI_MAX_LINKS) return fail(ctx, MAMRI_ERR_INVALID_ARG, "robot: n_links out of range");
    for (int l = 0; l < robot->n_links; ++l) {
        if (robot->links[l].parent >= l) return fail(ctx, MAMRI_ERR_INVALID_ARG, "robot: a link's parent must precede it");
        if (h_part_offsets[l + 1] < h_part_offsets[l] || h_part_offsets[l] < 0)
            return fail(ctx, MAMRI_ERR_INVALID_ARG, "part offsets must be non-negative and non-decreasing");
    }
    if (h_part_offsets[robot->n_links] > 0 && !d_part_points) return fail(ctx, MAMRI_ERR_INVALID_ARG, "part points are NULL");
    if (mask_desc->nx <= 0 || mask_desc->ny <= 0 || mask_desc->nz <= 0) return fail(ctx, MAMRI_ERR_INVALID_ARG, "mask dimensions must be positive");
    if (n_configs == 0) return MAMRI_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t off_ang = (sizeof(mamri_robot) + 255) & ~size_t(255);
    const size_t ang_bytes = size_t(n_configs) * MAMRI_MAX_CHAIN * sizeof(double);
    const size_t off_res = (off_ang + ang_bytes + 255) & ~size_t(255);
    const size_t total = off_res + size_t(n_configs) * sizeof(mamri_collision_result);
    if (ctx->pose_buf_bytes < total) {
        CK(cudaStreamSynchronize(s));
        cudaFree(ctx->d_pose_buf);
        ctx->d_pose_buf = nullptr; ctx->pose_buf_bytes = 0;
        CK(cudaMalloc(&ctx->d_pose_buf, total));
        ctx->pose_buf_bytes = total;
    }
    char* base = static_cast<char*>(ctx->d_pose_buf);
    CK(cudaMemcpyAsync(base, robot, sizeof(mamri_robot), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(base + off_ang, h_joint_angles, ang_bytes, cudaMemcpyHostToDevice, s));
    CK(launch_collision(reinterpret_cast<const mamri_robot*>(base), base_matrix, ras_to_index, mask_desc->nx, mask_desc->ny,
                        mask_desc->nz, h_part_offsets, robot->n_links, reinterpret_cast<const double*>(base + off_ang), n_configs,
                        d_part_points, d_body_mask, reinterpret_cast<mamri_collision_result*>(base + off_res), s));
    CK(cudaMemcpyAsync(h_results, base + off_res, size_t(n_configs) * sizeof(mamri_collision_result), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MAMRI_OK;
}

extern "C" int mamri_phantom_generate(uint16_t* d_volume, int32_t nx, int32_t ny, int32_t nz, const float* h_ellipsoids,
                                      int32_t n_ellipsoids, float sigma, uint64_t seed, uint32_t scan_index, void* stream) {
    if (!d_volume || nx <= 0 || ny <= 0 || nz <= 0 || n_ellipsoids < 0 || (n_ellipsoids && !h_ellipsoids)) {
        snprintf(g_create_err, sizeof(g_create_err), "mamri_phantom_generate: invalid argument");
        return MAMRI_ERR_INVALID_ARG;
    }
    cudaError_t e = launch_phantom(d_volume, nx, ny, nz, h_ellipsoids, n_ellipsoids, sigma, seed, scan_index,
                                   static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "mamri_phantom_generate: %s", cudaGetErrorString(e));
        return MAMRI_ERR_CUDA;
    }
    return MAMRI_OK;
}
