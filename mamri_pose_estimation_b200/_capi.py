"""ctypes binding of ``include/mamri_b200.h``.  Thin by design: structures mirror
the header field for field, every call goes straight to ``libmamri_b200.so``.
There is no CPU fallback: a missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

MAMRI_OK = 0
MAMRI_ERR_INVALID_ARG = -1
MAMRI_ERR_CUDA = -2
MAMRI_ERR_CAPACITY = -3
MAMRI_ERR_NO_DEVICE = -4
MAMRI_ERR_STATE = -5

DTYPE_CODES = {"uint8": 0, "int16": 1, "uint16": 2, "int32": 3, "float32": 4, "float64": 5}

import os

# MAMRI_LIB: another build of the same library (the trace build libmamri_b200_trace.so of tools/ktrace.py)
LIB_PATH = Path(os.environ.get("MAMRI_LIB") or Path(__file__).resolve().parent / "libmamri_b200.so")


class VolumeDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dtype", C.c_int32),
                ("spacing", C.c_double * 3), ("origin", C.c_double * 3), ("direction", C.c_double * 9)]


class Params(C.Structure):
    _fields_ = [("lower", C.c_double), ("upper", C.c_double), ("close_radius", C.c_int32),
                ("connectivity", C.c_int32), ("min_volume", C.c_double), ("max_volume", C.c_double),
                ("open_radius", C.c_int32), ("reserved", C.c_int32)]


class Marker(C.Structure):
    _fields_ = [("label", C.c_uint32), ("reserved", C.c_uint32), ("count", C.c_uint64),
                ("sum_idx", C.c_uint64 * 3), ("sum_mom", C.c_uint64 * 6), ("volume_mm3", C.c_double),
                ("centroid_index", C.c_double * 3), ("centroid_lps", C.c_double * 3),
                ("centroid_ras", C.c_double * 3), ("principal_moments", C.c_double * 3),
                ("principal_axes", C.c_double * 9)]


class Summary(C.Structure):
    _fields_ = [("n_labels", C.c_uint32), ("n_runs", C.c_uint32), ("n_markers", C.c_uint32),
                ("body_label", C.c_uint32), ("body_count", C.c_uint64), ("n_foreground", C.c_uint64),
                ("device_status", C.c_int32), ("reserved", C.c_int32), ("body", Marker)]


class EntryResult(C.Structure):
    _fields_ = [("index", C.c_int64), ("distance", C.c_double), ("point", C.c_double * 3),
                ("n_in_radius", C.c_uint64), ("n_suitable", C.c_uint64)]


MAX_LINKS, MAX_CHAIN, POSE_MAX_POINTS = 16, 8, 64
AXIS_CODES = {None: 0, "IS": 1, "PA": 2, "LR": 3, "TRANS_X": 4}
IK_NOT_RUN, IK_CONVERGED, IK_MAX_ITER = 0, 1, 2


class Link(C.Structure):
    _fields_ = [("parent", C.c_int32), ("axis", C.c_int32), ("has_markers", C.c_int32), ("chain_index", C.c_int32),
                ("translate", C.c_double * 3), ("marker_coords", C.c_double * 9), ("arm_lengths", C.c_double * 2),
                ("limits_deg", C.c_double * 2)]


class Robot(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("base_link", C.c_int32), ("effector_link", C.c_int32),
                ("secondary_link", C.c_int32), ("distance_tolerance", C.c_double), ("secondary_weight", C.c_double),
                ("apply_correction", C.c_int32), ("reserved", C.c_int32), ("links", Link * MAX_LINKS)]


class Pose(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("status", C.c_int32), ("matched", (C.c_int32 * 3) * MAX_LINKS),
                ("has_base", C.c_int32), ("ik_status", C.c_int32), ("ik_iterations", C.c_int32), ("ik_termination", C.c_int32),
                ("base_matrix", C.c_double * 16), ("joint_angles", C.c_double * MAX_CHAIN), ("ik_cost", C.c_double),
                ("ik_rms_error", C.c_double)]


class PoseOptions(C.Structure):
    _fields_ = [("h_saved_base", C.POINTER(C.c_double)), ("h_initial_angles", C.POINTER(C.c_double)),
                ("prefer_saved_base", C.c_int32), ("reserved", C.c_int32)]


class CollisionResult(C.Structure):
    _fields_ = [("link_mask", C.c_uint32), ("n_points_inside", C.c_uint32), ("first_link", C.c_int32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/mamri_b200.h declares
SIGNATURES = {
    "mamri_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32]),
    "mamri_destroy": (C.c_int, [C.c_void_p]),
    "mamri_last_error": (C.c_char_p, [C.c_void_p]),
    "mamri_version": (C.c_char_p, []),
    "mamri_default_params": (None, [C.POINTER(Params)]),
    "mamri_detect_async": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.c_void_p, C.POINTER(Params), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "mamri_detect_host_async": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.c_void_p, C.POINTER(Params),
                                          C.c_void_p, C.c_void_p]),
    "mamri_detect_collect": (C.c_int, [C.c_void_p, C.POINTER(Summary), C.POINTER(Marker), C.c_uint32]),
    "mamri_pool_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32,
                                    C.c_uint32]),
    "mamri_pool_destroy": (C.c_int, [C.c_void_p]),
    "mamri_pool_last_error": (C.c_char_p, [C.c_void_p]),
    "mamri_pool_context": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "mamri_pool_detect": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.POINTER(C.c_void_p), C.c_int32, C.POINTER(Params),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(Summary), C.POINTER(Marker), C.c_uint32, C.c_void_p]),
    "mamri_pool_detect_begin": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.POINTER(C.c_void_p), C.c_int32, C.POINTER(Params),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                          C.c_uint32, C.c_void_p]),
    "mamri_pool_detect_host_begin": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.POINTER(C.c_void_p), C.c_int32,
                                               C.POINTER(Params), C.POINTER(C.c_void_p), C.c_void_p]),
    "mamri_pool_detect_host_bits_begin": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.POINTER(C.c_void_p), C.c_int32,
                                                    C.POINTER(Params), C.POINTER(C.c_void_p), C.c_void_p]),
    "mamri_detect_bits_async": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "mamri_detect_host_bits_async": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p]),
    "mamri_reserve_staging": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "mamri_pool_detect_end": (C.c_int, [C.c_void_p, C.POINTER(Summary), C.POINTER(Marker), C.c_uint32]),
    "mamri_pool_detect_host": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.POINTER(C.c_void_p), C.c_int32,
                                         C.POINTER(Params), C.POINTER(C.c_void_p), C.POINTER(Summary), C.POINTER(Marker),
                                         C.c_uint32, C.c_void_p]),
    "mamri_label_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "mamri_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "mamri_kernel_launches": (C.c_int, [C.c_void_p]),
    "mamri_ktrace_reset": (C.c_int, []),
    "mamri_ktrace_read": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "mamri_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mamri_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.c_int]),
    "mamri_entry_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.POINTER(VolumeDesc),
                                     C.POINTER(C.c_double), C.c_int32, C.POINTER(EntryResult), C.c_void_p]),
    "mamri_body_surface": (C.c_int, [C.c_void_p, C.POINTER(VolumeDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]),
    "mamri_default_robot": (None, [C.POINTER(Robot)]),
    "mamri_pose_estimate": (C.c_int, [C.c_void_p, C.POINTER(Robot), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.POINTER(Pose), C.c_void_p]),
    "mamri_pose_estimate_ex": (C.c_int, [C.c_void_p, C.POINTER(Robot), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.POINTER(PoseOptions), C.POINTER(Pose), C.c_void_p]),
    "mamri_pose_from_tables": (C.c_int, [C.c_void_p, C.POINTER(Robot), C.c_void_p, C.c_int32, C.c_uint32, C.POINTER(Pose),
                                         C.c_void_p]),
    "mamri_collision_check": (C.c_int, [C.c_void_p, C.POINTER(Robot), C.POINTER(C.c_double), C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.POINTER(VolumeDesc), C.POINTER(C.c_double),
                                        C.POINTER(CollisionResult), C.c_void_p]),
    "mamri_phantom_generate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int32,
                                         C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libmamri_b200.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first "
                           "(python -m mamri_pose_estimation_b200.build); there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MamriError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mamri_b200 error {code}: {message}")
        self.code = code


def check(rc: int, ctx=None) -> None:
    if rc != MAMRI_OK:
        msg = load().mamri_last_error(ctx)
        raise MamriError(rc, msg.decode() if msg else "unknown error")


def check_pool(rc: int, pool=None) -> None:
    if rc != MAMRI_OK:
        msg = load().mamri_pool_last_error(pool)
        raise MamriError(rc, msg.decode() if msg else "unknown error")
