/*
 * mamri_b200.h -- C ABI of the B200-native fiducial-detection path.
 *
 * Drop-in boundary for ONE path of PaulSchlabach/mamri-pose-estimation: what
 * MamriLogic.volume_threshold_segmentation (Mamri/Mamri.py:1304-1323) asks
 * SimpleITK to do, and the candidate loop of MamriLogic.findAndSetEntryPoint
 * (Mamri/Mamri.py:1008-1023).  The reference has no FFI of its own (it is a
 * Python method body calling SimpleITK); each entry point below cites the
 * reference lines it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer would put behind those two methods.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types;
 *   - every function returns an int status (MAMRI_OK == 0, negative = error) and
 *     never aborts the host process; mamri_last_error() gives the text;
 *   - volumes are x-fastest: linear index = x + nx*(y + ny*z) (ITK buffer order);
 *   - "d_" pointers are device memory owned by the caller, "h_" pointers host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - a context is not thread-safe and serves one scan at a time; use one per
 *     (device, stream).  The device-buffer calls never allocate.  The host-buffer
 *     calls keep device staging buffers that grow on the first call at a size
 *     (mamri_reserve_staging sizes them ahead of time); the pose / collision calls
 *     keep a scratch buffer that grows with the batch size the same way.
 */
#ifndef MAMRI_B200_H
#define MAMRI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MAMRI_API __attribute__((visibility("default")))
#else
#define MAMRI_API
#endif

#define MAMRI_OK                0
#define MAMRI_ERR_INVALID_ARG  -1
#define MAMRI_ERR_CUDA         -2
#define MAMRI_ERR_CAPACITY     -3   /* run table or marker table too small for this scan */
#define MAMRI_ERR_NO_DEVICE    -4
#define MAMRI_ERR_STATE        -5   /* collect without a pending detect, ... */

/* voxel types PullVolumeFromSlicer can hand over (Mamri.py:1306) */
#define MAMRI_U8   0
#define MAMRI_I16  1
#define MAMRI_U16  2
#define MAMRI_I32  3
#define MAMRI_F32  4
#define MAMRI_F64  5

typedef struct mamri_ctx mamri_ctx;

/* Geometry of a SimpleITK image (LPS): physical = origin + direction * (spacing .* index). */
typedef struct mamri_volume_desc {
    int32_t nx, ny, nz;
    int32_t dtype;            /* MAMRI_U8 .. MAMRI_F64 */
    double  spacing[3];
    double  origin[3];
    double  direction[9];     /* row-major 3x3 */
} mamri_volume_desc;

/* The constants of MamriLogic.__init__ (Mamri.py:810-812) and of the call at :1308-1309.
 * mamri_default_params() fills in the reference's values. */
typedef struct mamri_params {
    double  lower;            /* INTENSITY_THRESHOLD = 65.0            Mamri.py:810  */
    double  upper;            /* 65535                                 Mamri.py:1308 */
    int32_t close_radius;     /* [2]*3 sitkBall -> 2 (0..3 supported)  Mamri.py:1308 */
    int32_t connectivity;     /* 6 = SimpleITK default, or 26          Mamri.py:1309 */
    double  min_volume;       /* MIN_VOLUME_THRESHOLD = 50.0 mm^3      Mamri.py:811  */
    double  max_volume;       /* MAX_VOLUME_THRESHOLD = 1500.0 mm^3    Mamri.py:812  */
    int32_t open_radius;      /* 0 = off (the reference only closes); 1..3: the thresholded mask is opened with the
                                 ITK ball first (sitk.BinaryMorphologicalOpening: north_star "open/close") */
    int32_t reserved;
} mamri_params;

/* One label kept by the list comprehension at Mamri.py:1310, in ascending label order
 * (= the control-point order of "DetectedFiducials", :1316-1317), plus what
 * LabelShapeStatisticsImageFilter computes for it. */
typedef struct mamri_marker {
    uint32_t label;                 /* "id": ITK-consecutive label (rank of the min linear index) */
    uint32_t reserved;
    uint64_t count;                 /* voxels */
    uint64_t sum_idx[3];            /* exact: sum x, sum y, sum z */
    uint64_t sum_mom[6];            /* exact: sum xx, yy, zz, xy, xz, yz */
    double   volume_mm3;            /* "vol": GetPhysicalSize */
    double   centroid_index[3];     /* mean voxel index (x,y,z) */
    double   centroid_lps[3];       /* "centroid": GetCentroid (LPS) */
    double   centroid_ras[3];       /* [-x,-y,z], Mamri.py:1317 */
    double   principal_moments[3];  /* ascending */
    double   principal_axes[9];     /* row-major, rows = axes (ITK GetPrincipalAxes) */
} mamri_marker;

typedef struct mamri_summary {
    uint32_t n_labels;        /* K = len(stats.GetLabels())                         */
    uint32_t n_runs;          /* x-runs of the closed mask (internal CCL nodes)      */
    uint32_t n_markers;       /* labels kept by the volume filter (may exceed the
                                 caller's array: then MAMRI_ERR_CAPACITY)            */
    uint32_t body_label;      /* largest non-fiducial label, 0 = none (Mamri.py:1318-1322) */
    uint64_t body_count;      /* its voxel count                                     */
    uint64_t n_foreground;    /* voxels set in the closed mask                       */
    int32_t  device_status;   /* MAMRI_OK or MAMRI_ERR_CAPACITY raised on the device  */
    int32_t  reserved;
    mamri_marker body;        /* shape statistics of the body label (count == 0 if none) */
} mamri_summary;

/* Result of the entry-point search (Mamri.py:1023-1024). index == -1: no suitable point
 * (the reference's warning + return at :1020-1022). */
typedef struct mamri_entry_result {
    int64_t index;
    double  distance;
    double  point[3];
    uint64_t n_in_radius;     /* candidates inside the search radius                */
    uint64_t n_suitable;      /* ... that also pass the normal score (and the path check) */
} mamri_entry_result;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* Allocates every scratch buffer for volumes up to max_nx x max_ny x max_nz.  max_runs bounds the
 * number of x-runs of the closed mask (0 = default: voxels/16, at least 1M); max_markers bounds the
 * labels passing the volume filter (0 = default 4096). */
MAMRI_API int mamri_create(mamri_ctx** ctx, int device, int32_t max_nx, int32_t max_ny, int32_t max_nz,
                 uint32_t max_runs, uint32_t max_markers);
MAMRI_API int mamri_destroy(mamri_ctx* ctx);
MAMRI_API const char* mamri_last_error(const mamri_ctx* ctx);   /* ctx may be NULL: last create() error */
MAMRI_API const char* mamri_version(void);
MAMRI_API void mamri_default_params(mamri_params* p);

/* ---- stages 1-4a: replaces Mamri.py:1308-1323 ------------------------------------------- */
/* Enqueues threshold -> ball closing -> connected components -> label statistics -> volume
 * filter -> body selection on `stream`.  Optional device outputs (NULL = not materialised):
 *   d_mask_out   uint8  [nz*ny*nx]  closed binary mask            (`closed`,  :1308)
 *   d_labels_out uint32 [nz*ny*nx]  ITK-consecutive label volume  (`labeled`, :1309)
 *   d_body_out   uint8  [nz*ny*nx]  1 where label == body label   (`largest_object_img`, :1323)
 * Returns as soon as the work is enqueued. */
MAMRI_API int mamri_detect_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* d_volume,
                       const mamri_params* params, uint8_t* d_mask_out, uint32_t* d_labels_out,
                       uint8_t* d_body_out, void* stream);
/* Same, from/to HOST buffers (pinned for full PCIe speed): copies the volume in, and the body mask
 * out when h_body_out != NULL.  This is the call a MamriLogic adapter makes. */
MAMRI_API int mamri_detect_host_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* h_volume,
                            const mamri_params* params, uint8_t* h_body_out, void* stream);
/* The body labelmap at 1 bit per voxel instead of uint8 (8x fewer bytes over the link to the host, which is what bounds
 * the host-buffer calls): [nz][ny][W] 32-bit words, W = ceil(nx / 32), bit k of word w of a row = voxel x = 32 w + k.
 * _bits_async: device output (d_body_bits_out may be NULL); _host_bits_async: host output. */
MAMRI_API int mamri_detect_bits_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* d_volume,
                            const mamri_params* params, uint8_t* d_mask_out, uint32_t* d_labels_out,
                            uint32_t* d_body_bits_out, void* stream);
MAMRI_API int mamri_detect_host_bits_async(mamri_ctx* ctx, const mamri_volume_desc* desc, const void* h_volume,
                                 const mamri_params* params, uint32_t* h_body_bits_out, void* stream);
/* Sizes the device staging buffers of the host-buffer calls (volume in, body labelmap out) ahead of the first call. */
MAMRI_API int mamri_reserve_staging(mamri_ctx* ctx, size_t volume_bytes, size_t body_bytes);
/* Waits for the pending detect, then writes the summary and up to max_markers markers
 * (ascending label).  "No markers" is MAMRI_OK with n_markers == 0 (Mamri.py:1312). */
MAMRI_API int mamri_detect_collect(mamri_ctx* ctx, mamri_summary* summary, mamri_marker* h_markers,
                         uint32_t max_markers);
/* Voxel count of every label 1..K of the last collected scan (GetPhysicalSize / voxel volume). */
MAMRI_API int mamri_label_counts(mamri_ctx* ctx, uint32_t* h_counts, uint32_t max_labels);

/* ---- batches of independent scans ------------------------------------------------------- */
/* MamriLogic.process handles one inputVolume per call (Mamri.py:850-858), so scans are independent.
 * A pool owns n_contexts contexts and as many streams on one device and pipelines a batch over
 * them: scan i runs on context i % n_contexts, and its enqueue + collect overlap the kernels of the
 * scans in flight on the other contexts.  The host loop lives in the library (no per-scan call
 * overhead in the caller's language). */
typedef struct mamri_pool mamri_pool;
MAMRI_API int mamri_pool_create(mamri_pool** pool, int device, int32_t n_contexts, int32_t max_nx, int32_t max_ny,
                      int32_t max_nz, uint32_t max_runs, uint32_t max_markers);
MAMRI_API int mamri_pool_destroy(mamri_pool* pool);
MAMRI_API const char* mamri_pool_last_error(const mamri_pool* pool);
/* Context k of the pool (0 <= k < n_contexts), e.g. for mamri_set_profiling; NULL if out of range. */
MAMRI_API mamri_ctx* mamri_pool_context(mamri_pool* pool, int32_t k);
/* n device-resident scans of one geometry.  d_volumes[i] is scan i; d_mask_out / d_labels_out /
 * d_body_out are NULL or arrays of n device pointers (entries may repeat when the caller only wants
 * them as temporaries, provided entries i and i + n_contexts are the only ones that alias).
 * summaries[n] and markers[n * max_markers_per_scan] (host) receive what mamri_detect_collect
 * returns per scan.  The pool's streams fork from and join back into `stream`; the call returns
 * when every scan has been collected.  Returns MAMRI_OK or the first failing scan's status (the
 * other scans are still processed; per-scan status is in summaries[i].device_status). */
MAMRI_API int mamri_pool_detect(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* d_volumes, int32_t n,
                      const mamri_params* params, uint8_t* const* d_mask_out, uint32_t* const* d_labels_out,
                      uint8_t* const* d_body_out, mamri_summary* summaries, mamri_marker* markers,
                      uint32_t max_markers_per_scan, void* stream);
/* The same batch in two halves, for callers that have device work to queue behind the scans (the NCCL
 * gather of the marker tables): _begin enqueues n <= n_contexts scans on `stream` and returns at once;
 * _end waits and fills summaries / markers as mamri_pool_detect does.  d_tables (optional): device
 * float64 [n][table_slots][8]; scan i's first table_slots markers are written there as rows of
 * {label, count, volume_mm3, RAS x, y, z, n_labels, body_label} (unused rows zero) by the scan's last
 * kernel, so a collective on d_tables can be enqueued on `stream` before _end is called. */
MAMRI_API int mamri_pool_detect_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* d_volumes, int32_t n,
                            const mamri_params* params, uint8_t* const* d_mask_out, uint32_t* const* d_labels_out,
                            uint8_t* const* d_body_out, double* d_tables, uint32_t table_slots, void* stream);
/* _begin for HOST buffers (as mamri_pool_detect_host; n <= n_contexts): every scan's H2D copy, kernels and
 * body-mask D2H are enqueued before the call returns, so a second pool can start feeding the link while
 * this one drains.  Ended by mamri_pool_detect_end. */
MAMRI_API int mamri_pool_detect_host_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes,
                                 int32_t n, const mamri_params* params, uint8_t* const* h_body_out, void* stream);
/* Same with the body labelmaps at 1 bit per voxel (see mamri_detect_host_bits_async). */
MAMRI_API int mamri_pool_detect_host_bits_begin(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes,
                                      int32_t n, const mamri_params* params, uint32_t* const* h_body_bits_out, void* stream);
MAMRI_API int mamri_pool_detect_end(mamri_pool* pool, mamri_summary* summaries, mamri_marker* markers,
                          uint32_t max_markers_per_scan);
/* Same from/to HOST buffers (pinned for full PCIe speed): the H2D copy of scan i+1 overlaps the
 * kernels and the body-mask D2H of scan i.  h_body_out is NULL or an array of n host pointers. */
MAMRI_API int mamri_pool_detect_host(mamri_pool* pool, const mamri_volume_desc* desc, const void* const* h_volumes,
                           int32_t n, const mamri_params* params, uint8_t* const* h_body_out,
                           mamri_summary* summaries, mamri_marker* markers, uint32_t max_markers_per_scan,
                           void* stream);

/* ---- measurement hooks (no reference counterpart) ----------------------------------------------- */
/* With profiling on, detect records CUDA events between its stages on the caller's stream;
 * mamri_stage_times() then returns, for the last collected scan, the milliseconds of
 * {threshold+pack, closing, connected components, statistics+filter, materialise}. */
MAMRI_API int mamri_set_profiling(mamri_ctx* ctx, int enable);
MAMRI_API int mamri_stage_times(mamri_ctx* ctx, float ms[5]);
/* Per-kernel milliseconds of the last profiled scan, in launch order; returns the number of entries
 * written (<= max_n) or a negative status. */
MAMRI_API int mamri_kernel_times(mamri_ctx* ctx, float* ms, const char** names, int max_n);

/* Kernels one scan took on the path it was last enqueued / captured with (6 with the cluster labelling kernel that
 * small run tables get, 10 with the scalable kernels of large ones). */
MAMRI_API int mamri_kernel_launches(const mamri_ctx* ctx);
/* In-pipeline kernel timeline, trace build only (libmamri_b200_trace.so, -DMAMRI_KTRACE; MAMRI_ERR_STATE otherwise):
 * every kernel stamps the GPU's nanosecond timer when its first CTA gets past the dependency on the kernel before it.
 * _read fills out_ns[32] (slot order: csrc/common.cuh KId; ~0 = not stamped since the last _reset). */
MAMRI_API int mamri_ktrace_reset(void);
MAMRI_API int mamri_ktrace_read(unsigned long long* out_ns, int n);

/* ---- stage 4b: replaces the loop at Mamri.py:1008-1023 ---------------------------------- */
/* d_points / d_normals: float32 [n][3] (RAS mm / unit normals), as vtkPolyData stores them.
 * Keeps points with |p - target|^2 <= radius^2 and wx*|nx| + wy*|ny| > cutoff (reference:
 * radius 80, wx 1, wy -2, cutoff -0.5), returns the closest (lowest index on ties).
 * Needle-path sampling (north_star extension, no reference counterpart; off when
 * n_path_samples == 0): a candidate is also rejected when any of n_path_samples points spaced
 * evenly strictly between it and the target lies on a voxel of d_path_mask whose value is
 * != path_free_value.  ras_to_index is the row-major 3x4 affine from RAS mm to voxel index of
 * that mask (nx,ny,nz from mask_desc). */
MAMRI_API int mamri_entry_search(mamri_ctx* ctx, const float* d_points, const float* d_normals, int64_t n,
                       const double target[3], double radius, double wx, double wy, double cutoff,
                       int32_t n_path_samples, const uint8_t* d_path_mask,
                       const mamri_volume_desc* mask_desc, const double ras_to_index[12],
                       int32_t path_free_value, mamri_entry_result* result, void* stream);

/* ---- skin-surface candidates: stands in for Mamri.py:994-1003 ---------------------------- */
/* The reference takes its entry-point candidates from Slicer's closed-surface representation of
 * "AutoBodySegmentation" (Mamri.py:1338-1339, _get_body_polydata at :994) and vtkPolyDataNormals
 * (:997-1003).  Neither exists outside Slicer, so the candidate set is defined on the voxel grid:
 * every body voxel with a face neighbour outside the body, in ascending linear index; point = its
 * physical centre in RAS (float32 [n][3]); normal = outward unit normal (float32 [n][3], RAS) from
 * the first moment of the body inside the radius-2 ball around the voxel (zero vector where that
 * moment vanishes).  The arrays feed mamri_entry_search directly.
 *   d_body_mask  uint8 [nz*ny*nx] (non-zero = body: `largest_object_img`, :1323), or NULL = the body
 *                of the scan last collected on this context (no per-voxel pass at all);
 *   capacity     points the output arrays hold; 0 (outputs may be NULL) only counts.
 * *n_points receives the number of surface voxels, *n_body_voxels (optional) the body size.
 * Returns MAMRI_ERR_CAPACITY when capacity > 0 is too small (the first `capacity` points are written). */
MAMRI_API int mamri_body_surface(mamri_ctx* ctx, const mamri_volume_desc* desc, const uint8_t* d_body_mask,
                       float* d_points_out, float* d_normals_out, int64_t capacity, int64_t* n_points,
                       int64_t* n_body_voxels, void* stream);

/* ---- marker table -> robot pose: replaces Mamri.py:1343-1363, 1371-1373, 1771-1792, 1410-1447 ---- */
/* What MamriLogic.process does with "DetectedFiducials" after the segmentation (Mamri.py:858-870),
 * batched on the device: L-shape triplet matching per marker-bearing link (joint_detection), the
 * baseplate y-flatten, the rigid landmark registration of the baseplate (vtkLandmarkTransform, float32
 * landmarks) and the full-chain IK on the effector (+ weighted secondary) markers.  The IK is the reference's own
 * solver restated: Trust Region Reflective bounded least squares with a 2-point finite-difference Jacobian and
 * ftol = xtol = 1e-6 (scipy.optimize.least_squares defaults otherwise), so it takes the same iterates and stops in
 * the same minimum as the reference, to rounding. */
#define MAMRI_MAX_LINKS        16
#define MAMRI_MAX_CHAIN         8
#define MAMRI_POSE_MAX_POINTS  64    /* control points per scan the matcher accepts */
#define MAMRI_AXIS_NONE  0
#define MAMRI_AXIS_IS    1           /* RotateZ(angle)    Mamri.py:1763 */
#define MAMRI_AXIS_PA    2           /* RotateY(-angle)   Mamri.py:1765 */
#define MAMRI_AXIS_LR    3           /* RotateX(angle)    Mamri.py:1767 */
#define MAMRI_AXIS_TRANS 4           /* "TRANS_X": no rotation (Mamri.py:1498) */
#define MAMRI_IK_NOT_RUN   0         /* baseplate or effector markers not identified */
#define MAMRI_IK_CONVERGED 1         /* the solver met a stopping criterion (res.success, Mamri.py:1434) */
#define MAMRI_IK_MAX_ITER  2         /* evaluation budget spent on every initial guess: the reference would report "IK failed" */

/* One entry of robot_config.json, in file order (joint_detection iterates in this order, :1349). */
typedef struct mamri_link {
    int32_t parent;              /* index of the parent link, -1 = root (its parent transform is the baseplate registration) */
    int32_t axis;                /* MAMRI_AXIS_* ("articulation_axis") */
    int32_t has_markers;
    int32_t chain_index;         /* position in the articulated chain (Mamri.py:819), -1 = not an IK unknown */
    double  translate[3];        /* "fixed_offset_to_parent": translate */
    double  marker_coords[9];    /* "local_marker_coords": 3 points */
    double  arm_lengths[2];
    double  limits_deg[2];       /* "joint_limits" */
} mamri_link;

typedef struct mamri_robot {
    int32_t n_links;
    int32_t base_link;           /* "Baseplate" */
    int32_t effector_link;       /* "Joint6" */
    int32_t secondary_link;      /* "Joint4", -1 = none */
    double  distance_tolerance;  /* DISTANCE_TOLERANCE = 5.0, Mamri.py:813 */
    double  secondary_weight;    /* joint4_weight = 0.05, Mamri.py:1507 */
    int32_t apply_correction;    /* effector markers turned 180 deg about z (Mamri.py:1511-1514) */
    int32_t reserved;
    mamri_link links[MAMRI_MAX_LINKS];
} mamri_robot;

typedef struct mamri_pose {
    int32_t n_points;                        /* control points of the scan */
    int32_t status;                          /* MAMRI_OK, or MAMRI_ERR_CAPACITY: more than MAMRI_POSE_MAX_POINTS points */
    int32_t matched[MAMRI_MAX_LINKS][3];     /* per link: control-point indices (corner, short arm, long arm), -1 = not identified */
    int32_t has_base;                        /* 0 = none; 1 = baseplate registered from the scan; 2 = the saved transform */
    int32_t ik_status;                       /* MAMRI_IK_* */
    int32_t ik_iterations;                   /* residual evaluations of the solver (SciPy's nfev), all initial guesses */
    int32_t ik_termination;                  /* SciPy's status of the run that was kept: 1 gtol, 2 ftol, 3 xtol, 4 both, 0 = budget spent */
    double  base_matrix[16];                 /* row-major 4x4, baseplate model -> world (RAS) */
    double  joint_angles[MAMRI_MAX_CHAIN];   /* rad, articulated-chain order */
    double  ik_cost;                         /* 0.5 * sum of squared residuals */
    double  ik_rms_error;                    /* last_ik_error (Mamri.py:1443-1444): rms of the effector residuals */
} mamri_pose;

/* Fills in robot_config.json (Baseplate, Joint1..Joint6, Needle) and the reference's constants. */
MAMRI_API void mamri_default_robot(mamri_robot* robot);
/* h_points_ras: float64 [n_scans][max_points][3], the control points of each scan's "DetectedFiducials"
 * in node order (= ascending label; mamri_marker.centroid_ras); h_counts[n_scans] their numbers.
 * One warp per scan on the device; h_poses[n_scans] receives the results.  A scan with fewer than three
 * points, or without baseplate / effector markers, is MAMRI_OK with the corresponding fields unset. */
MAMRI_API int mamri_pose_estimate(mamri_ctx* ctx, const mamri_robot* robot, const double* h_points_ras,
                        const int32_t* h_counts, int32_t n_scans, int32_t max_points, mamri_pose* h_poses,
                        void* stream);

/* What MamriLogic holds besides the scan when it estimates a pose (all optional; NULL options = a fresh scene):
 *   h_saved_base      row-major 4x4 "MamriSavedBaseplateTransform" (Mamri.py:820, 1376-1408): used instead of the scan's
 *                     baseplate when prefer_saved_base is set (pNode.useSavedBaseplate), and as the fall-back when the
 *                     scan shows no baseplate; mamri_pose.has_base tells which one a scan got
 *   h_initial_angles  [n_scans][MAMRI_MAX_CHAIN] current joint angles (rad): the first of the two initial guesses of
 *                     _solve_full_chain_ik (:1425), the second being zeros; the lower cost of the successful runs is kept */
typedef struct mamri_pose_options {
    const double* h_saved_base;
    const double* h_initial_angles;
    int32_t prefer_saved_base;
    int32_t reserved;
} mamri_pose_options;
MAMRI_API int mamri_pose_estimate_ex(mamri_ctx* ctx, const mamri_robot* robot, const double* h_points_ras,
                           const int32_t* h_counts, int32_t n_scans, int32_t max_points,
                           const mamri_pose_options* options, mamri_pose* h_poses, void* stream);

/* Same, fed from the device-written marker tables of mamri_pool_detect_begin (float64
 * [n_scans][table_slots][8]; a row is a control point while its label is non-zero): enqueued on `stream`
 * right behind the scans, no host round trip of the tables.  Waits for the poses. */
MAMRI_API int mamri_pose_from_tables(mamri_ctx* ctx, const mamri_robot* robot, const double* d_tables, int32_t n_scans,
                           uint32_t table_slots, mamri_pose* h_poses, void* stream);

/* ---- robot-vs-body collision sampling: stands in for Mamri.py:1555-1575 ------------------- */
/* _check_collision runs vtkCollisionDetectionFilter between each link's collision mesh and the body
 * mesh, one joint configuration per call (per configuration of a planned path, :976-982; inside the
 * trajectory IK's error function, :1541).  Here a batch of configurations is tested on the device:
 * every sample point of every link (float32 [total][3], link-local coordinates, e.g. the vertices of
 * the *_collision.STL meshes; h_part_offsets[n_links + 1] delimits each link's points, empty ranges
 * for links that are not checked) goes through the link's forward-kinematics transform
 * (base_matrix = baseplate -> world, row-major 4x4) and the RAS -> voxel affine, and hits if it lands
 * on a non-zero voxel of d_body_mask (uint8 [nz*ny*nx]; outside the volume = free).  Points inside the
 * body instead of surface intersection: decisions agree with the mesh test on clear and on colliding
 * poses, not on grazing contacts. */
typedef struct mamri_collision_result {
    uint32_t link_mask;        /* bit l set: link l has a sample point inside the body */
    uint32_t n_points_inside;
    int32_t  first_link;       /* lowest colliding link (the reference returns at the first contact), -1 = clear */
    int32_t  reserved;
} mamri_collision_result;
MAMRI_API int mamri_collision_check(mamri_ctx* ctx, const mamri_robot* robot, const double base_matrix[16],
                          const double* h_joint_angles /* [n_configs][MAMRI_MAX_CHAIN] rad */, int32_t n_configs,
                          const float* d_part_points, const int32_t* h_part_offsets, const uint8_t* d_body_mask,
                          const mamri_volume_desc* mask_desc, const double ras_to_index[12],
                          mamri_collision_result* h_results, void* stream);

/* ---- synthetic phantoms (benchmark utility, not part of the reference path) -------------- */
/* Paints `n_ellipsoids` (7 floats each: cx,cy,cz,ax,ay,az,intensity; index units; in order)
 * into a zeroed uint16 volume, then adds Rician noise (Philox4x32-10; see phantom.py). */
MAMRI_API int mamri_phantom_generate(uint16_t* d_volume, int32_t nx, int32_t ny, int32_t nz,
                           const float* h_ellipsoids, int32_t n_ellipsoids, float sigma,
                           uint64_t seed, uint32_t scan_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAMRI_B200_H */
