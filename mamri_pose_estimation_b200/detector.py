"""Host-side driver of the CUDA fiducial-detection path.

`FiducialDetector` wraps one ``mamri_ctx`` (one GPU, one scan in flight).  PyTorch
is used only for device memory and streams: tensors are passed to the C ABI as raw
pointers.  Results come back as plain Python/NumPy objects shaped like what
``MamriLogic.volume_threshold_segmentation`` builds (Mamri/Mamri.py:1310-1323).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _capi
from ._capi import EntryResult, Marker, Params, Pose, Robot, Summary, VolumeDesc, check

_TORCH_DTYPES = {torch.uint8: "uint8", torch.int16: "int16", torch.uint16: "uint16",
                 torch.int32: "int32", torch.float32: "float32", torch.float64: "float64"}

IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)


@dataclasses.dataclass
class DetectParams:
    """Constants of the reference (Mamri.py:810-812, :1308-1309)."""
    lower: float = 65.0
    upper: float = 65535.0
    close_radius: int = 2
    connectivity: int = 6
    min_volume: float = 50.0
    max_volume: float = 1500.0
    open_radius: int = 0            # 0 = the reference (closing only); > 0: sitk.BinaryMorphologicalOpening first

    def to_c(self) -> Params:
        return Params(self.lower, self.upper, self.close_radius, self.connectivity, self.min_volume, self.max_volume,
                      self.open_radius, 0)


@dataclasses.dataclass
class MarkerStats:
    label: int
    count: int
    sum_idx: tuple
    sum_mom: tuple
    volume_mm3: float
    centroid_index: np.ndarray
    centroid_lps: np.ndarray
    centroid_ras: np.ndarray
    principal_moments: np.ndarray
    principal_axes: np.ndarray

    @staticmethod
    def from_c(m: Marker) -> "MarkerStats":
        return MarkerStats(label=int(m.label), count=int(m.count), sum_idx=tuple(int(v) for v in m.sum_idx),
                           sum_mom=tuple(int(v) for v in m.sum_mom), volume_mm3=float(m.volume_mm3),
                           centroid_index=np.array(m.centroid_index[:]), centroid_lps=np.array(m.centroid_lps[:]),
                           centroid_ras=np.array(m.centroid_ras[:]), principal_moments=np.array(m.principal_moments[:]),
                           principal_axes=np.array(m.principal_axes[:]).reshape(3, 3))


class _LazyMarkers:
    """list-like view of a marker table copied out of the C array: MarkerStats objects are built on first access
    (a noisy high-resolution scan keeps ~1000 labels; building them eagerly would cost more than the scan)."""

    def __init__(self, table: np.ndarray):
        self.table = table                       # structured array, dtype = np.dtype(Marker), own copy
        self._items: Optional[List[MarkerStats]] = None

    def _all(self) -> List[MarkerStats]:
        if self._items is None:
            mk = (Marker * len(self.table)).from_buffer_copy(self.table.tobytes()) if len(self.table) else []
            self._items = [MarkerStats.from_c(m) for m in mk]
        return self._items

    def __len__(self):
        return len(self.table)

    def __iter__(self):
        return iter(self._all())

    def __getitem__(self, i):
        return self._all()[i]

    def __eq__(self, other):
        return list(self) == list(other)

    def __add__(self, other):
        return list(self) + list(other)

    def __radd__(self, other):
        return list(other) + list(self)


def _marker_table(c_markers, n: int, first: int = 0) -> np.ndarray:
    """Copies markers [first, first + n) of a ctypes Marker array into a NumPy structured array."""
    if n <= 0:
        return np.zeros(0, dtype=np.dtype(Marker))
    view = np.frombuffer(c_markers, dtype=np.dtype(Marker), count=n, offset=first * C.sizeof(Marker))
    return view.copy()


@dataclasses.dataclass
class DetectionResult:
    n_labels: int
    n_runs: int
    n_foreground: int
    markers: Sequence[MarkerStats]               # ascending label = "DetectedFiducials" control-point order
    body_label: int
    body_count: int
    body: Optional[MarkerStats]
    mask: Optional[torch.Tensor] = None          # uint8 [nz,ny,nx] closed mask (device)
    labels: Optional[torch.Tensor] = None        # uint32-valued int32 tensor [nz,ny,nx] (device)
    body_mask: Optional[object] = None           # uint8 [nz,ny,nx] (device tensor, or host array for detect_host)
    status: int = 0                              # MAMRI_OK; batches: MAMRI_ERR_CAPACITY for a scan beyond the pool's capacity

    @property
    def marker_table(self) -> Optional[np.ndarray]:
        """The markers as a NumPy structured array (fields of mamri_marker), when they came from the library."""
        return self.markers.table if isinstance(self.markers, _LazyMarkers) else None

    @property
    def fiducials_data(self) -> List[dict]:
        """The list the reference builds at Mamri.py:1310."""
        t = self.marker_table
        if t is not None:
            return [{"vol": float(v), "centroid": (float(c[0]), float(c[1]), float(c[2])), "id": int(l)}
                    for v, c, l in zip(t["volume_mm3"], t["centroid_lps"], t["label"])]
        return [{"vol": m.volume_mm3, "centroid": tuple(float(c) for c in m.centroid_lps), "id": m.label}
                for m in self.markers]

    @property
    def ras_points(self) -> np.ndarray:
        t = self.marker_table
        if t is not None:
            return np.array(t["centroid_ras"], dtype=np.float64).reshape(-1, 3)
        return np.array([m.centroid_ras for m in self.markers], dtype=np.float64).reshape(-1, 3)

    @property
    def marker_labels(self) -> List[str]:
        t = self.marker_table
        if t is not None:
            return [f"M_{int(l)}_{float(v):.0f}mm³" for l, v in zip(t["label"], t["volume_mm3"])]      # Mamri.py:1317
        return [f"M_{m.label}_{m.volume_mm3:.0f}mm³" for m in self.markers]   # Mamri.py:1317


ROBOT_LINK_NAMES = ("Baseplate", "Joint1", "Joint2", "Joint3", "Joint4", "Joint5", "Joint6", "Needle")   # robot_config.json order


@dataclasses.dataclass
class PoseResult:
    """What MamriLogic.process derives from "DetectedFiducials" (Mamri.py:858-870) for one scan."""
    n_points: int
    identified: dict                             # link name -> [control-point ids] (corner, short arm, long arm): joint_detection
    base_matrix: Optional[np.ndarray]            # 4x4 baseplate model -> world (RAS), None if no baseplate in the scan
    joint_angles: Optional[np.ndarray]           # rad, articulated-chain order (Joint1..Joint6), None if the IK did not run
    ik_converged: bool
    ik_iterations: int                           # residual evaluations (SciPy's nfev) over all initial guesses
    ik_cost: float
    ik_rms_error: float                          # last_ik_error (Mamri.py:1443-1444)
    ik_termination: int = 0                      # SciPy's status of the kept run: 1 gtol, 2 ftol, 3 xtol, 4 both, 0 budget spent
    base_source: str = ""                        # "scan", "saved" or "" (no baseplate transform at all)

    @staticmethod
    def from_c(p: Pose, names=ROBOT_LINK_NAMES, n_chain: int = 6) -> "PoseResult":
        ident = {names[l]: [int(v) for v in p.matched[l]] for l in range(len(names)) if p.matched[l][0] >= 0}
        ran = p.ik_status != _capi.IK_NOT_RUN
        return PoseResult(n_points=int(p.n_points), identified=ident,
                          base_matrix=np.array(p.base_matrix[:]).reshape(4, 4) if p.has_base else None,
                          joint_angles=np.array(p.joint_angles[:n_chain]) if ran else None,
                          ik_converged=p.ik_status == _capi.IK_CONVERGED, ik_iterations=int(p.ik_iterations),
                          ik_cost=float(p.ik_cost), ik_rms_error=float(p.ik_rms_error), ik_termination=int(p.ik_termination),
                          base_source={0: "", 1: "scan", 2: "saved"}.get(int(p.has_base), ""))


def _desc(shape_zyx, dtype_name, spacing, origin, direction) -> VolumeDesc:
    nz, ny, nx = (int(v) for v in shape_zyx)
    d = VolumeDesc()
    d.nx, d.ny, d.nz = nx, ny, nz
    d.dtype = _capi.DTYPE_CODES[dtype_name]
    d.spacing[:] = [float(v) for v in spacing]
    d.origin[:] = [float(v) for v in origin]
    d.direction[:] = [float(v) for v in direction]
    return d


class FiducialDetector:
    """One detection context on one GPU.  Not thread-safe; one scan in flight."""

    def __init__(self, max_dims_xyz: Sequence[int], device: int = 0, max_runs: int = 0, max_markers: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("mamri_pose_estimation_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = _capi.load()
        self.device = int(device)
        self.max_dims = tuple(int(v) for v in max_dims_xyz)
        self._ctx = C.c_void_p()
        rc = self._lib.mamri_create(C.byref(self._ctx), self.device, *self.max_dims, int(max_runs), int(max_markers))
        check(rc, None)
        self.max_markers = int(max_markers) if max_markers else 4096
        self._markers = (Marker * self.max_markers)()
        self._pending = None

    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            if getattr(self, "_owned", True):            # views of a pool's contexts are destroyed by the pool
                self._lib.mamri_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    # ------------------------------------------------------------------ device-resident path
    def detect_async(self, volume: torch.Tensor, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=IDENTITY,
                     params: Optional[DetectParams] = None, want_mask=False, want_labels=False, want_body=False,
                     out_mask: Optional[torch.Tensor] = None, out_labels: Optional[torch.Tensor] = None,
                     out_body: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Enqueues one scan.  `volume`: contiguous CUDA tensor [nz, ny, nx] (x fastest)."""
        if not (volume.is_cuda and volume.is_contiguous() and volume.dim() == 3):
            raise ValueError("volume must be a contiguous CUDA tensor [nz, ny, nx]")
        if volume.dtype not in _TORCH_DTYPES:
            raise ValueError(f"unsupported voxel type {volume.dtype}")
        if volume.device.index != self.device:
            raise ValueError(f"volume is on {volume.device}, this context runs on cuda:{self.device}")
        params = params or DetectParams()
        dev = volume.device
        shape = tuple(volume.shape)
        if want_mask and out_mask is None:
            out_mask = torch.empty(shape, dtype=torch.uint8, device=dev)
        if want_labels and out_labels is None:
            out_labels = torch.empty(shape, dtype=torch.int32, device=dev)
        if want_body and out_body is None:
            out_body = torch.empty(shape, dtype=torch.uint8, device=dev)
        s = stream or torch.cuda.current_stream(dev)
        d = _desc(shape, _TORCH_DTYPES[volume.dtype], spacing, origin, direction)
        p = params.to_c()
        rc = self._lib.mamri_detect_async(self._ctx, C.byref(d), volume.data_ptr(), C.byref(p),
                                          out_mask.data_ptr() if out_mask is not None else None,
                                          out_labels.data_ptr() if out_labels is not None else None,
                                          out_body.data_ptr() if out_body is not None else None, s.cuda_stream)
        check(rc, self._ctx)
        self._pending = (out_mask, out_labels, out_body, volume)   # keep the buffers alive until collect

    def collect(self) -> DetectionResult:
        summ = Summary()
        rc = self._lib.mamri_detect_collect(self._ctx, C.byref(summ), self._markers, self.max_markers)
        pend, self._pending = self._pending, None
        check(rc, self._ctx)
        markers = _LazyMarkers(_marker_table(self._markers, int(summ.n_markers)))
        body = MarkerStats.from_c(summ.body) if summ.body_label else None
        out_mask, out_labels, out_body = (pend[0], pend[1], pend[2]) if pend else (None, None, None)
        return DetectionResult(n_labels=int(summ.n_labels), n_runs=int(summ.n_runs), n_foreground=int(summ.n_foreground),
                               markers=markers, body_label=int(summ.body_label), body_count=int(summ.body_count),
                               body=body, mask=out_mask, labels=out_labels, body_mask=out_body)

    def detect(self, volume: torch.Tensor, **kw) -> DetectionResult:
        self.detect_async(volume, **kw)
        return self.collect()

    # ------------------------------------------------------------------ host-buffer path (the drop-in call)
    def reserve_staging(self, volume_bytes: int, body_bytes: int = 0) -> None:
        """Sizes the device staging buffers of the host-buffer calls ahead of the first call (mamri_reserve_staging)."""
        check(self._lib.mamri_reserve_staging(self._ctx, int(volume_bytes), int(body_bytes)), self._ctx)

    def detect_host_async(self, volume, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=IDENTITY,
                          params: Optional[DetectParams] = None, body_out=None, body_bits_out=None,
                          stream: Optional[torch.cuda.Stream] = None) -> None:
        """`volume`: C-contiguous host array [nz, ny, nx] (numpy, or a pinned CPU torch tensor for full PCIe
        speed).  `body_out`: optional host uint8 buffer of the same shape receiving the body mask;
        `body_bits_out`: instead, a host int32/uint32 buffer [nz, ny, ceil(nx/32)] receiving it at 1 bit per voxel
        (8x fewer bytes back over PCIe; `unpack_body_bits` expands it)."""
        arr, keep = _host_view(volume)
        params = params or DetectParams()
        d = _desc(arr["shape"], arr["dtype"], spacing, origin, direction)
        p = params.to_c()
        body_ptr, keep_b = (None, None)
        if body_out is not None and body_bits_out is not None:
            raise ValueError("ask for the body mask as uint8 (body_out) or bit-packed (body_bits_out), not both")
        if body_out is not None:
            b, keep_b = _host_view(body_out)
            if b["dtype"] != "uint8" or tuple(b["shape"]) != tuple(arr["shape"]):
                raise ValueError("body_out must be uint8 with the volume's shape")
            body_ptr = b["ptr"]
        s = stream or torch.cuda.current_stream(self.device)
        if body_bits_out is not None:
            bptr, keep_b = _bits_view(body_bits_out, arr["shape"])
            rc = self._lib.mamri_detect_host_bits_async(self._ctx, C.byref(d), arr["ptr"], C.byref(p), bptr, s.cuda_stream)
        else:
            rc = self._lib.mamri_detect_host_async(self._ctx, C.byref(d), arr["ptr"], C.byref(p), body_ptr, s.cuda_stream)
        check(rc, self._ctx)
        self._pending = (None, None, body_out if body_out is not None else body_bits_out, (keep, keep_b))

    def detect_host(self, volume, **kw) -> DetectionResult:
        self.detect_host_async(volume, **kw)
        return self.collect()

    STAGES = ("threshold_pack", "closing", "ccl", "stats_filter", "materialise")

    @property
    def kernel_launches(self) -> int:
        """Kernels per scan on the path the last scan was enqueued with (6: cluster labelling, 10: scalable kernels)."""
        return int(self._lib.mamri_kernel_launches(self._ctx))

    def set_profiling(self, enable: bool) -> None:
        check(self._lib.mamri_set_profiling(self._ctx, int(bool(enable))), self._ctx)

    def stage_times_ms(self) -> dict:
        ms = (C.c_float * 5)()
        check(self._lib.mamri_stage_times(self._ctx, ms), self._ctx)
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def kernel_times_ms(self) -> list:
        """[(kernel, ms)] of the last profiled scan, in launch order."""
        ms = (C.c_float * 48)()
        names = (C.c_char_p * 48)()
        n = self._lib.mamri_kernel_times(self._ctx, ms, names, 48)
        if n < 0:
            check(n, self._ctx)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def label_counts(self, n_labels: int) -> np.ndarray:
        out = np.zeros(max(int(n_labels), 1), dtype=np.uint32)
        rc = self._lib.mamri_label_counts(self._ctx, out.ctypes.data, out.size)
        check(rc, self._ctx)
        return out[:n_labels]

    # ------------------------------------------------------------------ entry-point search
    def entry_search(self, points: torch.Tensor, normals: torch.Tensor, target, radius=80.0, wx=1.0, wy=-2.0,
                     cutoff=-0.5, n_path_samples=0, path_mask: Optional[torch.Tensor] = None, ras_to_index=None,
                     path_free_value=1, stream: Optional[torch.cuda.Stream] = None) -> dict:
        """Closest suitable entry point among float32 CUDA `points`/`normals` [n,3] (RAS).  Constants default to
        the reference's (Mamri.py:1009, :1015-1016)."""
        for t in (points, normals):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == 3):
                raise ValueError("points/normals must be contiguous float32 CUDA tensors [n, 3]")
        n = int(points.shape[0])
        tgt = (C.c_double * 3)(*[float(v) for v in target])
        res = EntryResult()
        md, m2i, mptr = None, None, None
        if n_path_samples:
            if path_mask is None or ras_to_index is None:
                raise ValueError("needle-path sampling (n_path_samples > 0) needs path_mask and ras_to_index")
            if not (path_mask.is_cuda and path_mask.is_contiguous() and path_mask.dtype == torch.uint8 and path_mask.dim() == 3):
                raise ValueError("path_mask must be a contiguous uint8 CUDA tensor [nz, ny, nx]")
            md = C.byref(_desc(tuple(path_mask.shape), "uint8", (1, 1, 1), (0, 0, 0), IDENTITY))
            m2i = (C.c_double * 12)(*[float(v) for v in np.asarray(ras_to_index, dtype=np.float64).reshape(12)])
            mptr = path_mask.data_ptr()
        s = stream or torch.cuda.current_stream(points.device)
        rc = self._lib.mamri_entry_search(self._ctx, points.data_ptr(), normals.data_ptr(), n, tgt, float(radius),
                                          float(wx), float(wy), float(cutoff), int(n_path_samples), mptr, md, m2i,
                                          int(path_free_value), C.byref(res), s.cuda_stream)
        check(rc, self._ctx)
        return {"index": int(res.index), "distance": float(res.distance), "point": np.array(res.point[:]),
                "n_in_radius": int(res.n_in_radius), "n_suitable": int(res.n_suitable)}


    # ------------------------------------------------------------------ marker table -> robot pose
    def default_robot(self, apply_correction: bool = False) -> Robot:
        r = Robot()
        self._lib.mamri_default_robot(C.byref(r))
        r.apply_correction = int(bool(apply_correction))
        return r

    def pose_estimate(self, ras_points: Sequence, robot: Optional[Robot] = None, apply_correction: bool = False,
                      stream: Optional[torch.cuda.Stream] = None, saved_base=None, prefer_saved_base: bool = False,
                      initial_angles=None) -> List[PoseResult]:
        """L-shape matching, baseplate registration and full-chain IK (Mamri.py:1343-1447) for a batch of scans on the
        device, one warp per scan.  `ras_points[i]`: [n_i, 3] control points of scan i in node order
        (DetectionResult.ras_points).  `saved_base`: 4x4 "MamriSavedBaseplateTransform" -- used instead of the scan's
        baseplate when `prefer_saved_base` (pNode.useSavedBaseplate) and as the fall-back when a scan shows no
        baseplate (Mamri.py:1376-1408).  `initial_angles`: [n, <= 8] current joint angles, the first initial guess of
        the IK (:1425)."""
        n = len(ras_points)
        if n == 0:
            return []
        arrs = [np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in ras_points]
        max_pts = max(1, max(a.shape[0] for a in arrs))
        pts = np.zeros((n, max_pts, 3), dtype=np.float64)
        cnt = np.zeros(n, dtype=np.int32)
        for i, a in enumerate(arrs):
            pts[i, :a.shape[0]] = a
            cnt[i] = a.shape[0]
        robot = robot if robot is not None else self.default_robot(apply_correction)
        poses = (Pose * n)()
        s = stream or torch.cuda.current_stream(self.device)
        opt = _capi.PoseOptions()
        keep = []
        if saved_base is not None:
            sb = np.ascontiguousarray(np.asarray(saved_base, dtype=np.float64).reshape(16))
            keep.append(sb)
            opt.h_saved_base = sb.ctypes.data_as(C.POINTER(C.c_double))
        if initial_angles is not None:
            ia = np.zeros((n, _capi.MAX_CHAIN), dtype=np.float64)
            a = np.asarray(initial_angles, dtype=np.float64).reshape(n, -1)
            ia[:, :a.shape[1]] = a
            keep.append(ia)
            opt.h_initial_angles = ia.ctypes.data_as(C.POINTER(C.c_double))
        opt.prefer_saved_base = int(bool(prefer_saved_base))
        rc = self._lib.mamri_pose_estimate_ex(self._ctx, C.byref(robot), pts.ctypes.data, cnt.ctypes.data, n, max_pts,
                                              C.byref(opt), poses, s.cuda_stream)
        check(rc, self._ctx)
        out = []
        for i in range(n):
            if poses[i].status != _capi.MAMRI_OK:
                raise _capi.MamriError(poses[i].status, f"scan {i}: {cnt[i]} control points exceed the matcher's limit "
                                                        f"of {_capi.POSE_MAX_POINTS}")
            out.append(PoseResult.from_c(poses[i]))
        return out

    def pose_from_tables(self, tables: torch.Tensor, n_scans: Optional[int] = None, robot: Optional[Robot] = None,
                         apply_correction: bool = False, stream: Optional[torch.cuda.Stream] = None) -> List[PoseResult]:
        """`pose_estimate` fed from the device-written marker tables of `BatchDetector.begin(tables=...)`: queued on the
        stream right behind the scans (call it between begin() and end()); only the poses travel to the host."""
        if not (tables.is_cuda and tables.is_contiguous() and tables.dtype == torch.float64 and tables.dim() == 3
                and tables.shape[2] == 8):
            raise ValueError("tables must be a contiguous float64 CUDA tensor [n, slots, 8]")
        n = int(tables.shape[0]) if n_scans is None else int(n_scans)
        if n == 0:
            return []
        robot = robot if robot is not None else self.default_robot(apply_correction)
        poses = (Pose * n)()
        s = stream or torch.cuda.current_stream(tables.device)
        check(self._lib.mamri_pose_from_tables(self._ctx, C.byref(robot), tables.data_ptr(), n, int(tables.shape[1]), poses,
                                               s.cuda_stream), self._ctx)
        out = []
        for i in range(n):
            if poses[i].status != _capi.MAMRI_OK:
                raise _capi.MamriError(poses[i].status, f"scan {i}: more control points than the matcher's limit of "
                                                        f"{_capi.POSE_MAX_POINTS}")
            out.append(PoseResult.from_c(poses[i]))
        return out

    # ------------------------------------------------------------------ robot-vs-body collision sampling
    def collision_check(self, part_points: dict, joint_angles, base_matrix, body_mask: torch.Tensor, ras_to_index,
                        robot: Optional[Robot] = None, stream: Optional[torch.cuda.Stream] = None) -> List[dict]:
        """Stands in for MamriLogic._check_collision (Mamri.py:1555-1575) for a batch of joint configurations.
        `part_points`: link name -> float32 [n,3] sample points in the link's own frame (e.g. the vertices of its
        *_collision.STL); `joint_angles`: [n_configs, 6] rad; `body_mask`: uint8 CUDA tensor [nz,ny,nx];
        `ras_to_index`: 3x4 affine RAS mm -> voxel index.  Returns per configuration
        {"collision": bool, "links": [names], "first_link": name or None, "n_points_inside": int}."""
        if not (body_mask.is_cuda and body_mask.is_contiguous() and body_mask.dtype == torch.uint8 and body_mask.dim() == 3):
            raise ValueError("body_mask must be a contiguous uint8 CUDA tensor [nz, ny, nx]")
        if len(joint_angles) == 0:
            return []
        robot = robot if robot is not None else self.default_robot()
        names = ROBOT_LINK_NAMES[:robot.n_links]
        unknown = set(part_points) - set(names)
        if unknown:
            raise ValueError(f"unknown links: {sorted(unknown)}")
        offs = (C.c_int32 * (robot.n_links + 1))()
        chunks = []
        for l, nm in enumerate(names):
            p = np.asarray(part_points.get(nm, np.zeros((0, 3))), dtype=np.float32).reshape(-1, 3)
            chunks.append(p)
            offs[l + 1] = offs[l] + p.shape[0]
        allp = np.ascontiguousarray(np.concatenate(chunks)) if offs[robot.n_links] else np.zeros((1, 3), np.float32)
        pts_d = torch.from_numpy(allp).to(body_mask.device)
        ang = np.zeros((len(joint_angles), _capi.MAX_CHAIN), dtype=np.float64)
        ja = np.asarray(joint_angles, dtype=np.float64).reshape(len(joint_angles), -1)
        ang[:, :ja.shape[1]] = ja
        n = ang.shape[0]
        res = (_capi.CollisionResult * max(n, 1))()
        base = (C.c_double * 16)(*[float(v) for v in np.asarray(base_matrix, dtype=np.float64).reshape(16)])
        m2i = (C.c_double * 12)(*[float(v) for v in np.asarray(ras_to_index, dtype=np.float64).reshape(12)])
        d = _desc(tuple(body_mask.shape), "uint8", (1, 1, 1), (0, 0, 0), IDENTITY)
        s = stream or torch.cuda.current_stream(body_mask.device)
        rc = self._lib.mamri_collision_check(self._ctx, C.byref(robot), base, ang.ctypes.data, n, pts_d.data_ptr(), offs,
                                             body_mask.data_ptr(), C.byref(d), m2i, res, s.cuda_stream)
        check(rc, self._ctx)
        out = []
        for i in range(n):
            r = res[i]
            links = [names[l] for l in range(robot.n_links) if (r.link_mask >> l) & 1]
            out.append({"collision": bool(r.link_mask), "links": links,
                        "first_link": names[r.first_link] if r.first_link >= 0 else None,
                        "n_points_inside": int(r.n_points_inside)})
        return out

    # ------------------------------------------------------------------ skin-surface candidates
    def body_surface(self, body_mask: Optional[torch.Tensor] = None, shape_zyx=None, spacing=(1.0, 1.0, 1.0),
                     origin=(0.0, 0.0, 0.0), direction=IDENTITY, stream: Optional[torch.cuda.Stream] = None):
        """Entry-point candidates of a body labelmap (stands in for Mamri.py:994-1003): float32 CUDA tensors
        (points [n,3] RAS, normals [n,3] RAS), surface voxels in ascending linear index.  `body_mask`: uint8 CUDA
        tensor [nz,ny,nx], or None = the body of the scan last collected on this detector (then pass `shape_zyx`).
        Two calls into the library: count, then emit into exactly-sized arrays."""
        if body_mask is not None:
            if not (body_mask.is_cuda and body_mask.is_contiguous() and body_mask.dim() == 3 and body_mask.dtype == torch.uint8):
                raise ValueError("body_mask must be a contiguous uint8 CUDA tensor [nz, ny, nx]")
            shape_zyx, ptr, dev = tuple(body_mask.shape), body_mask.data_ptr(), body_mask.device
        else:
            if shape_zyx is None:
                raise ValueError("shape_zyx is needed when the body of the last scan is used")
            ptr, dev = None, torch.device(f"cuda:{self.device}")
        d = _desc(shape_zyx, "uint8", spacing, origin, direction)
        s = stream or torch.cuda.current_stream(dev)
        n, n_body = C.c_int64(0), C.c_int64(0)
        check(self._lib.mamri_body_surface(self._ctx, C.byref(d), ptr, None, None, 0, C.byref(n), C.byref(n_body),
                                           s.cuda_stream), self._ctx)
        pts = torch.empty((n.value, 3), dtype=torch.float32, device=dev)
        nrm = torch.empty((n.value, 3), dtype=torch.float32, device=dev)
        if n.value:
            check(self._lib.mamri_body_surface(self._ctx, C.byref(d), ptr, pts.data_ptr(), nrm.data_ptr(), n.value,
                                               C.byref(n), C.byref(n_body), s.cuda_stream), self._ctx)
        self.last_body_voxels = int(n_body.value)
        return pts, nrm


def _host_view(a):
    """(ptr, shape, dtype-name) of a C-contiguous host numpy array or CPU torch tensor, plus a keep-alive ref."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda or not a.is_contiguous() or a.dim() != 3:
            raise ValueError("host volume must be a contiguous CPU tensor [nz, ny, nx]")
        if a.dtype not in _TORCH_DTYPES:
            raise ValueError(f"unsupported voxel type {a.dtype}")
        return {"ptr": a.data_ptr(), "shape": tuple(a.shape), "dtype": _TORCH_DTYPES[a.dtype]}, a
    arr = np.asarray(a)
    if arr.ndim != 3 or not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("host volume must be a C-contiguous array [nz, ny, nx]")
    if arr.dtype.name not in _capi.DTYPE_CODES:
        raise ValueError(f"unsupported voxel type {arr.dtype}")
    return {"ptr": arr.ctypes.data, "shape": arr.shape, "dtype": arr.dtype.name}, arr


def body_bits_shape(shape_zyx) -> tuple:
    nz, ny, nx = (int(v) for v in shape_zyx)
    return (nz, ny, (nx + 31) // 32)


def _bits_view(a, shape_zyx):
    """Pointer of a host buffer for the bit-packed body mask: 32-bit words [nz, ny, ceil(nx/32)]."""
    want = body_bits_shape(shape_zyx)
    if isinstance(a, torch.Tensor):
        if a.is_cuda or not a.is_contiguous() or a.dtype not in (torch.int32, torch.uint32) or tuple(a.shape) != want:
            raise ValueError(f"body_bits_out must be a contiguous CPU int32 tensor of shape {want}")
        return a.data_ptr(), a
    arr = np.asarray(a)
    if not arr.flags["C_CONTIGUOUS"] or arr.dtype.itemsize != 4 or arr.dtype.kind not in "iu" or tuple(arr.shape) != want:
        raise ValueError(f"body_bits_out must be a C-contiguous 32-bit integer array of shape {want}")
    return arr.ctypes.data, arr


def unpack_body_bits(bits, shape_zyx) -> np.ndarray:
    """uint8 [nz, ny, nx] labelmap (what `largest_object_img` is, Mamri.py:1323) from the bit-packed form the
    `*_bits` calls return: bit k of word w of a row is voxel x = 32 w + k."""
    nz, ny, nx = (int(v) for v in shape_zyx)
    words = bits.numpy() if isinstance(bits, torch.Tensor) else np.asarray(bits)
    by = np.ascontiguousarray(words).view(np.uint8).reshape(nz, ny, -1)
    return np.unpackbits(by, axis=2, bitorder="little")[:, :, :nx]


def generate_phantom_cuda(ph, device: int = 0, out: Optional[torch.Tensor] = None,
                          stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Creates the phantom described by `ph` (phantom.Phantom) directly in HBM; uint16 [nz, ny, nx]."""
    lib = _capi.load()
    nx, ny, nz = ph.dims
    if out is None:
        out = torch.empty((nz, ny, nx), dtype=torch.uint16, device=f"cuda:{device}")
    ell = np.ascontiguousarray(ph.ellipsoids, dtype=np.float32)
    s = stream or torch.cuda.current_stream(out.device)
    with torch.cuda.device(out.device):
        rc = lib.mamri_phantom_generate(out.data_ptr(), nx, ny, nz, ell.ctypes.data_as(C.POINTER(C.c_float)),
                                        int(ell.shape[0]), float(ph.sigma), int(ph.seed), int(ph.scan_index), s.cuda_stream)
    check(rc, None)
    return out


class BatchResult:
    """Results of one batch: per-scan summaries and marker tables copied out of the C arrays `mamri_pool_detect` filled
    (only the records in use, so a pool may keep a large marker capacity without a large per-batch cost); a sequence of
    `DetectionResult` built on demand.  `status[i]` is the scan's device status: MAMRI_OK, or MAMRI_ERR_CAPACITY when
    the scan found more runs / kept labels than the pool was created for -- that scan's result carries the status and
    no markers, the other scans of the batch are unaffected (the reference never fails on many candidates, a noisy
    scan just yields more control points; Mamri.py:1310-1317)."""

    def __init__(self, n, summaries, markers, max_m, outs):
        self.n, self._max_m, self._outs = n, max_m, outs
        self._summ = np.frombuffer(summaries, dtype=np.dtype(Summary), count=n).copy()
        mk = np.frombuffer(markers, dtype=np.dtype(Marker), count=n * max_m).reshape(n, max_m) if max_m else None
        self.status = [int(v) for v in self._summ["device_status"]]
        self._tables = [mk[i, :min(int(self._summ["n_markers"][i]), max_m)].copy() if (mk is not None and self.status[i] == 0)
                        else np.zeros(0, dtype=np.dtype(Marker)) for i in range(n)]
        self._cache = {}

    def __len__(self):
        return self.n

    def __getitem__(self, i) -> DetectionResult:
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        if i not in self._cache:
            summ = self._summ[i]
            body = None
            if int(summ["body_label"]) and self.status[i] == 0:
                body = MarkerStats.from_c(Marker.from_buffer_copy(summ["body"].tobytes()))
            m, l, b = self._outs(i)
            self._cache[i] = DetectionResult(n_labels=int(summ["n_labels"]), n_runs=int(summ["n_runs"]),
                                             n_foreground=int(summ["n_foreground"]), markers=_LazyMarkers(self._tables[i]),
                                             body_label=int(summ["body_label"]), body_count=int(summ["body_count"]), body=body,
                                             mask=m, labels=l, body_mask=b, status=self.status[i])
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(self.n))

    def ras_points(self) -> List[np.ndarray]:
        """Per scan: [n_markers, 3] control points of "DetectedFiducials" (node order)."""
        return [np.array(t["centroid_ras"], dtype=np.float64).reshape(-1, 3) for t in self._tables]

    def table(self, slots: int = 32) -> np.ndarray:
        """[n, slots, 8] float64: label, count, volume_mm3, RAS x y z, n_labels, body_label (distributed.pack_table)."""
        t = np.zeros((self.n, slots, 8), dtype=np.float64)
        for i, mk in enumerate(self._tables):
            k = min(slots, len(mk))
            t[i, :k, 0] = mk["label"][:k]
            t[i, :k, 1] = mk["count"][:k]
            t[i, :k, 2] = mk["volume_mm3"][:k]
            t[i, :k, 3:6] = mk["centroid_ras"][:k]
            t[i, :k, 6] = self._summ["n_labels"][i]
            t[i, :k, 7] = self._summ["body_label"][i]
        return t


class BatchDetector:
    """A batch of independent scans pipelined over a pool of contexts/streams on one GPU (`mamri_pool_*`): the
    enqueue + collect latency of one scan hides behind the kernels of the others (scans are independent:
    ``MamriLogic.process`` handles exactly one inputVolume, Mamri.py:850-858).  The per-scan loop runs inside the
    library; one call per batch crosses the ctypes boundary.

    The pool's streams fork from and join back into the caller's current stream, so CUDA events recorded on the
    current stream around `run` / `run_host` bracket all the work."""

    def __init__(self, dims_xyz: Sequence[int], device: int = 0, n_contexts: int = 4, max_runs: int = 0,
                 max_markers: int = 4096, materialise: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("mamri_pose_estimation_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = _capi.load()
        self.device = int(device)
        self.dims = tuple(int(v) for v in dims_xyz)
        self.n_contexts = int(n_contexts)
        self.max_markers = int(max_markers)
        nx, ny, nz = self.dims
        dev = torch.device(f"cuda:{self.device}")
        self._pool = C.c_void_p()
        _capi.check_pool(self._lib.mamri_pool_create(C.byref(self._pool), self.device, self.n_contexts, nx, ny, nz,
                                                     int(max_runs), self.max_markers), None)
        self.materialise = materialise
        self.masks = [torch.empty((nz, ny, nx), dtype=torch.uint8, device=dev) if materialise else None
                      for _ in range(self.n_contexts)]
        self.labels = [torch.empty((nz, ny, nx), dtype=torch.int32, device=dev) if materialise else None
                       for _ in range(self.n_contexts)]

    @property
    def kernel_launches_per_scan(self) -> int:
        """Kernels per scan on the path of the last batch (counted by the library while it enqueues / captures)."""
        return int(self._lib.mamri_kernel_launches(C.c_void_p(self._lib.mamri_pool_context(self._pool, 0))))

    def close(self):
        if getattr(self, "_pool", None) and self._pool.value:
            self._lib.mamri_pool_destroy(self._pool)
            self._pool = C.c_void_p()

    __del__ = close

    def context(self, k: int = 0) -> "FiducialDetector":
        """Context k of the pool as a FiducialDetector view (profiling hooks, single scans); owned by the pool."""
        view = FiducialDetector.__new__(FiducialDetector)
        view._lib, view.device, view.max_dims = self._lib, self.device, self.dims
        view._ctx = C.c_void_p(self._lib.mamri_pool_context(self._pool, int(k)))
        if not view._ctx.value:
            raise IndexError(k)
        view.max_markers = self.max_markers
        view._markers = (Marker * self.max_markers)()
        view._pending = None
        view._owned = False                  # the pool destroys its contexts
        return view

    def _scratch(self, n):
        """Result arrays for n scans, allocated once per size (ctypes zero-fills what it allocates; at the default
        capacity of 4096 markers per scan that would cost more than a batch of scans).  BatchResult copies out what
        is in use before the arrays are reused."""
        cache = self.__dict__.setdefault("_scratch_cache", {})
        if n not in cache:
            cache[n] = ((Summary * n)(), (Marker * (n * self.max_markers))())
        return cache[n]

    def _check_volumes(self, volumes):
        v0 = volumes[0]
        for v in volumes:
            if not (v.is_cuda and v.is_contiguous() and v.dim() == 3 and v.dtype == v0.dtype and v.shape == v0.shape):
                raise ValueError("volumes must be contiguous CUDA tensors [nz, ny, nx] of one shape and type")
            if v.device.index != self.device:
                raise ValueError(f"volume is on {v.device}, this pool runs on cuda:{self.device}")
        return v0

    def _finish(self, rc):
        """A scan beyond the pool's capacity is reported per scan (BatchResult.status), everything else raises."""
        if rc != _capi.MAMRI_ERR_CAPACITY:
            _capi.check_pool(rc, self._pool)

    @staticmethod
    def _ptrs(items, n):
        arr = (C.c_void_p * n)()
        for i in range(n):
            arr[i] = items[i]
        return arr

    def run(self, volumes: Sequence[torch.Tensor], spacing, origin, direction=IDENTITY,
            params: Optional[DetectParams] = None) -> BatchResult:
        """Device-resident scans in, marker tables out (mask + label volumes are materialised into the pool's
        ring of buffers, as the reference's `closed` / `labeled` temporaries are)."""
        n, k = len(volumes), self.n_contexts
        v0 = self._check_volumes(volumes)
        d = _desc(tuple(v0.shape), _TORCH_DTYPES[v0.dtype], spacing, origin, direction)
        p = (params or DetectParams()).to_c()
        summ, mk = self._scratch(n)
        vp = self._ptrs([v.data_ptr() for v in volumes], n)
        mp = self._ptrs([self.masks[i % k].data_ptr() for i in range(n)], n) if self.materialise else None
        lp = self._ptrs([self.labels[i % k].data_ptr() for i in range(n)], n) if self.materialise else None
        s = torch.cuda.current_stream(self.device)
        rc = self._lib.mamri_pool_detect(self._pool, C.byref(d), vp, n, C.byref(p), mp, lp, None, summ, mk,
                                         self.max_markers, s.cuda_stream)
        self._finish(rc)
        masks, labels = self.masks, self.labels
        # the mask / label volumes live in a ring of n_contexts buffers: only the last n_contexts scans still own theirs
        return BatchResult(n, summ, mk, self.max_markers,
                           lambda i: (masks[i % k], labels[i % k], None) if i >= n - k else (None, None, None))

    def begin(self, volumes: Sequence[torch.Tensor], spacing, origin, direction=IDENTITY,
              params: Optional[DetectParams] = None, tables: Optional[torch.Tensor] = None) -> None:
        """First half of `run` for up to n_contexts scans: enqueues them on the current stream and returns at once.
        `tables` (optional): float64 CUDA tensor [n, slots, 8] that the scans' last kernels fill with the fixed-size
        marker tables (distributed.pack_table layout), so that the gather of the tables can be queued behind the
        scans before `end()` makes the host wait."""
        n, k = len(volumes), self.n_contexts
        if not 1 <= n <= k:
            raise ValueError(f"begin/end handles 1..{k} scans per call")
        v0 = self._check_volumes(volumes)
        tptr, slots = None, 0
        if tables is not None:
            if not (tables.is_cuda and tables.is_contiguous() and tables.dtype == torch.float64 and tables.dim() == 3
                    and tables.shape[0] >= n and tables.shape[2] == 8):
                raise ValueError("tables must be a contiguous float64 CUDA tensor [n, slots, 8]")
            tptr, slots = tables.data_ptr(), int(tables.shape[1])
        d = _desc(tuple(v0.shape), _TORCH_DTYPES[v0.dtype], spacing, origin, direction)
        p = (params or DetectParams()).to_c()
        vp = self._ptrs([v.data_ptr() for v in volumes], n)
        mp = self._ptrs([self.masks[i % k].data_ptr() for i in range(n)], n) if self.materialise else None
        lp = self._ptrs([self.labels[i % k].data_ptr() for i in range(n)], n) if self.materialise else None
        s = torch.cuda.current_stream(self.device)
        rc = self._lib.mamri_pool_detect_begin(self._pool, C.byref(d), vp, n, C.byref(p), mp, lp, None, tptr, slots,
                                               s.cuda_stream)
        _capi.check_pool(rc, self._pool)
        self._begun = (n, volumes, tables)           # keep the buffers alive until end()

    def begin_host(self, volumes: Sequence, spacing, origin, direction=IDENTITY, params: Optional[DetectParams] = None,
                   body_out: Optional[Sequence] = None, body_bits_out: Optional[Sequence] = None) -> None:
        """First half of `run_host` for up to n_contexts scans: every scan's H2D copy, kernels and body-mask D2H are
        enqueued; `end()` waits.  `body_bits_out`: host buffers for the body masks at 1 bit per voxel instead of
        `body_out`'s uint8 (see FiducialDetector.detect_host_async)."""
        n = len(volumes)
        if not 1 <= n <= self.n_contexts:
            raise ValueError(f"begin/end handles 1..{self.n_contexts} scans per call")
        views = [_host_view(v) for v in volumes]
        a0 = views[0][0]
        for a, _ in views:
            if a["dtype"] != a0["dtype"] or tuple(a["shape"]) != tuple(a0["shape"]):
                raise ValueError("host volumes must share one shape and type")
        d = _desc(a0["shape"], a0["dtype"], spacing, origin, direction)
        p = (params or DetectParams()).to_c()
        vp = self._ptrs([a["ptr"] for a, _ in views], n)
        bp, bviews = None, None
        if body_out is not None:
            bviews = [_host_view(b) for b in body_out]
            for b, _ in bviews:
                if b["dtype"] != "uint8" or tuple(b["shape"]) != tuple(a0["shape"]):
                    raise ValueError("body_out must be uint8 with the volume's shape")
            bp = self._ptrs([b["ptr"] for b, _ in bviews], n)
        s = torch.cuda.current_stream(self.device)
        if body_bits_out is not None:
            if body_out is not None:
                raise ValueError("ask for the body masks as uint8 or bit-packed, not both")
            bviews = [_bits_view(b, a0["shape"]) for b in body_bits_out]
            bp = self._ptrs([ptr for ptr, _ in bviews], n)
            rc = self._lib.mamri_pool_detect_host_bits_begin(self._pool, C.byref(d), vp, n, C.byref(p), bp, s.cuda_stream)
            body_out = body_bits_out
        else:
            rc = self._lib.mamri_pool_detect_host_begin(self._pool, C.byref(d), vp, n, C.byref(p), bp, s.cuda_stream)
        _capi.check_pool(rc, self._pool)
        self._begun = (n, views, bviews)
        self._begun_body = body_out if body_out is not None else ()

    def end(self) -> BatchResult:
        """Second half of `run`: waits for the scans begun with `begin` and returns their results."""
        if getattr(self, "_begun", None) is None:
            raise RuntimeError("no batch pending: call begin() first")
        n = self._begun[0]
        self._begun = None
        summ, mk = self._scratch(n)
        rc = self._lib.mamri_pool_detect_end(self._pool, summ, mk, self.max_markers)
        self._finish(rc)
        masks, labels, k = self.masks, self.labels, self.n_contexts
        body, self._begun_body = getattr(self, "_begun_body", None), None
        if body is not None:                         # begun with begin_host: no device volumes, host body masks if asked for
            return BatchResult(n, summ, mk, self.max_markers, lambda i: (None, None, body[i] if len(body) else None))
        return BatchResult(n, summ, mk, self.max_markers, lambda i: (masks[i % k], labels[i % k], None))

    def estimate_poses(self, results: BatchResult, apply_correction: bool = False) -> List[PoseResult]:
        """Matching + baseplate registration + IK for every scan of a batch in one device call (one warp per scan)."""
        return self.context(0).pose_estimate(results.ras_points(), apply_correction=apply_correction)

    def run_host(self, volumes: Sequence, spacing, origin, direction=IDENTITY, params: Optional[DetectParams] = None,
                 body_out: Optional[Sequence] = None) -> BatchResult:
        """Host buffers in (pinned CPU tensors / numpy), marker tables + optional host body masks out: the
        drop-in call for a batch, with the H2D copy of scan i+1 overlapping the kernels and D2H of scan i."""
        n = len(volumes)
        views = [_host_view(v) for v in volumes]
        a0 = views[0][0]
        for a, _ in views:
            if a["dtype"] != a0["dtype"] or tuple(a["shape"]) != tuple(a0["shape"]):
                raise ValueError("host volumes must share one shape and type")
        d = _desc(a0["shape"], a0["dtype"], spacing, origin, direction)
        p = (params or DetectParams()).to_c()
        summ, mk = self._scratch(n)
        vp = self._ptrs([a["ptr"] for a, _ in views], n)
        bp = None
        if body_out is not None:
            bviews = [_host_view(b) for b in body_out]
            for b, _ in bviews:
                if b["dtype"] != "uint8" or tuple(b["shape"]) != tuple(a0["shape"]):
                    raise ValueError("body_out must be uint8 with the volume's shape")
            bp = self._ptrs([b["ptr"] for b, _ in bviews], n)
        s = torch.cuda.current_stream(self.device)
        rc = self._lib.mamri_pool_detect_host(self._pool, C.byref(d), vp, n, C.byref(p), bp, summ, mk, self.max_markers,
                                              s.cuda_stream)
        self._finish(rc)
        return BatchResult(n, summ, mk, self.max_markers,
                           lambda i: (None, None, body_out[i] if body_out is not None else None))


class BatchPipeline:
    """A stream of batches through `depth` pools that alternate: batch k+1 is enqueued (on its own stream) before
    the results of batch k are collected, so the thresholds of one batch overlap the materialise kernels of the
    previous one and the host-side collection of one batch hides behind the kernels of the next.  Same kernels, same
    results as `BatchDetector.run`; every batch still has at most n_contexts scans."""

    def __init__(self, dims_xyz: Sequence[int], device: int = 0, n_contexts: int = 8, depth: int = 2, **kw):
        self.device = int(device)
        self.pools = [BatchDetector(dims_xyz, device=device, n_contexts=n_contexts, **kw) for _ in range(int(depth))]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.pools]
        self.n_contexts = n_contexts
        self._in_flight: List[int] = []
        self._next = 0

    @property
    def kernel_launches_per_scan(self) -> int:
        return self.pools[0].kernel_launches_per_scan

    def close(self):
        for p in self.pools:
            p.close()

    def submit(self, volumes: Sequence[torch.Tensor], spacing, origin, direction=IDENTITY,
               params: Optional[DetectParams] = None, tables: Optional[torch.Tensor] = None) -> torch.cuda.Stream:
        """Enqueues one batch; returns the stream it runs on (queue device work that consumes `tables` there)."""
        i = self._next
        if i in self._in_flight:
            raise RuntimeError("pipeline full: collect a result() first")
        self._next = (i + 1) % len(self.pools)
        s = self.streams[i]
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.pools[i].begin(volumes, spacing, origin, direction, params, tables=tables)
        self._in_flight.append(i)
        return s

    def submit_host(self, volumes: Sequence, spacing, origin, direction=IDENTITY, params: Optional[DetectParams] = None,
                    body_out: Optional[Sequence] = None, body_bits_out: Optional[Sequence] = None) -> torch.cuda.Stream:
        """`submit` for HOST buffers (pinned for full PCIe speed): the next batch's copies start feeding the link while
        the previous batch's last scans still compute and drain."""
        i = self._next
        if i in self._in_flight:
            raise RuntimeError("pipeline full: collect a result() first")
        self._next = (i + 1) % len(self.pools)
        s = self.streams[i]
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.pools[i].begin_host(volumes, spacing, origin, direction, params, body_out=body_out, body_bits_out=body_bits_out)
        self._in_flight.append(i)
        return s

    def result(self) -> BatchResult:
        """Results of the oldest batch in flight; the current stream then also waits for that batch's stream."""
        if not self._in_flight:
            raise RuntimeError("no batch in flight")
        i = self._in_flight.pop(0)
        res = self.pools[i].end()
        torch.cuda.current_stream(self.device).wait_stream(self.streams[i])
        return res

    def pending(self) -> int:
        return len(self._in_flight)
