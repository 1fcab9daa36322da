"""ctypes front-end of the plain-C oracle (oracle/c/mamri_oracle.c).  TEST INFRASTRUCTURE
(see oracle/__init__.py): used by tests for full-size parity and by bench.py as the timed CPU arm.
Returns the same `Detection` structure as oracle.segmentation.detect_fiducials."""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np

from . import build_c
from . import segmentation as seg

_DT = {"uint8": 0, "int16": 1, "uint16": 2, "int32": 3, "float32": 4, "float64": 5}
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build_c.build()))
        lib.oracle_detect.restype = C.c_int
        lib.oracle_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
        lib.oracle_detect2.restype = C.c_int
        lib.oracle_detect2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
        lib.oracle_opening.restype = C.c_int
        lib.oracle_opening.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_closing.restype = C.c_int
        lib.oracle_closing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_ccl.restype = C.c_int
        lib.oracle_ccl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint32)]
        lib.oracle_label_sums.restype = C.c_int
        lib.oracle_label_sums.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_set_num_threads.restype = None
        lib.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(load().oracle_num_threads())


def use_all_cores() -> int:
    """Sets the OpenMP team to every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().oracle_set_num_threads(n)
    return num_threads()


def run_pipeline(vol: np.ndarray, lo=seg.INTENSITY_THRESHOLD, hi=seg.UPPER_THRESHOLD, close_radius=seg.CLOSE_RADIUS,
                 connectivity=6, with_sums=True, open_radius=0):
    """threshold -> closing -> CCL (-> integer sums).  Returns (closed u8, labels u32, K, sums[K,10] or None, seconds)."""
    lib = load()
    vol = np.ascontiguousarray(vol)
    nz, ny, nx = vol.shape
    closed = np.empty(vol.shape, dtype=np.uint8)
    labels = np.empty(vol.shape, dtype=np.uint32)
    k = C.c_uint32(0)
    t0 = time.perf_counter()
    rc = lib.oracle_detect2(vol.ctypes.data, _DT[vol.dtype.name], nx, ny, nz, float(lo), float(hi), int(open_radius),
                            int(close_radius), int(connectivity), closed.ctypes.data, labels.ctypes.data, C.byref(k), None, 0)
    if rc != 0:
        raise RuntimeError(f"oracle_detect failed: {rc}")
    sums = None
    if with_sums:
        sums = np.zeros((max(int(k.value), 1), 10), dtype=np.uint64)
        rc = lib.oracle_label_sums(labels.ctypes.data, nx, ny, nz, k.value, sums.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"oracle_label_sums failed: {rc}")
        sums = sums[:k.value]
    dt = time.perf_counter() - t0
    return closed, labels, int(k.value), sums, dt


def detect_fiducials(vol: np.ndarray, geom: seg.Geometry, lo=seg.INTENSITY_THRESHOLD, hi=seg.UPPER_THRESHOLD,
                     close_radius=seg.CLOSE_RADIUS, connectivity=6, min_vol=seg.MIN_VOLUME_THRESHOLD,
                     max_vol=seg.MAX_VOLUME_THRESHOLD, want_body_mask=True, open_radius=0) -> seg.Detection:
    """Mamri.py:1308-1323 via the C oracle; statistics finalised exactly like oracle.segmentation."""
    closed, labels, k, sums, _ = run_pipeline(vol, lo, hi, close_radius, connectivity, open_radius=open_radius)
    counts = sums[:, 0].astype(np.int64) if k else np.zeros(0, dtype=np.int64)
    kept, body = seg.select_candidates(counts, geom.voxel_volume(), min_vol, max_vol)
    stats = []
    for l in sorted(set(kept) | ({body} if body else set())):
        s = sums[l - 1]
        n = int(s[0])
        sidx = tuple(int(v) for v in s[1:4])
        smom = tuple(int(v) for v in s[4:10])
        cidx = np.array(sidx, dtype=np.float64) / n
        pm, pa = seg.moments_from_sums(n, sidx, smom, geom)
        stats.append(seg.LabelStats(label=l, count=n, sum_idx=sidx, sum_mom=smom, physical_size=n * geom.voxel_volume(),
                                    centroid_index=cidx, centroid=geom.index_to_physical(cidx),
                                    principal_moments=pm, principal_axes=pa))
    by = {s.label: s for s in stats}
    fid = [{"vol": by[l].physical_size, "centroid": tuple(float(c) for c in by[l].centroid), "id": int(l)} for l in kept]
    ras = np.array([[-f["centroid"][0], -f["centroid"][1], f["centroid"][2]] for f in fid], dtype=np.float64).reshape(-1, 3)
    names = [f"M_{f['id']}_{f['vol']:.0f}mm³" for f in fid]
    body_mask = (labels == body).astype(np.uint8) if (body and want_body_mask) else None
    return seg.Detection(binary=None, closed=closed, labels=labels, n_labels=k, counts=counts, fiducials=fid,
                         ras_points=ras, marker_labels=names, body_label=body, body_mask=body_mask, stats=stats)
