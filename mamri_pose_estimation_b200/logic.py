"""Slicer-free mirror of the two `MamriLogic` methods this package replaces, with the reference's
names, arguments and error behaviour, delegating to the CUDA path through the C ABI:

  * ``MamriLogic.volume_threshold_segmentation(pNode)``   Mamri/Mamri.py:1304-1341
  * ``MamriLogic.findAndSetEntryPoint(pNode)``            Mamri/Mamri.py:987-1033

The reference stores its results in the MRML scene (markups node "DetectedFiducials", segmentation
node "AutoBodySegmentation", markups node "ClosestSuitableEntryPoint").  Here the scene is a plain
dict of tiny node objects with the accessor names the reference's downstream code uses
(`GetNumberOfControlPoints`, `GetNthControlPointPositionWorld`, ...), so `joint_detection`-style
consumers read them the same way.  There is no CPU fallback: without the CUDA library / a GPU the
constructor raises.
"""
from __future__ import annotations

import dataclasses
import logging
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .detector import IDENTITY, DetectParams, DetectionResult, FiducialDetector


# ----------------------------------------------------------------------------- scene stand-ins
@dataclasses.dataclass
class ScalarVolumeNode:
    """What `sitkUtils.PullVolumeFromSlicer(pNode.inputVolume)` yields (Mamri.py:1306): voxels [nz,ny,nx]
    (numpy array, CPU tensor or CUDA tensor) plus the image geometry in LPS."""
    array: object
    spacing: Sequence[float] = (1.0, 1.0, 1.0)
    origin: Sequence[float] = (0.0, 0.0, 0.0)
    direction: Sequence[float] = IDENTITY
    name: str = "inputVolume"

    @staticmethod
    def from_ijk_to_ras(array, ijk_to_ras, name: str = "inputVolume") -> "ScalarVolumeNode":
        """The geometry conversion `PullVolumeFromSlicer` performs (Mamri.py:1306): a MRML volume node carries a 4x4
        IJK-to-RAS matrix; the SimpleITK image it becomes has spacing = the column norms, direction = the
        normalised columns and origin = the translation, both turned RAS -> LPS (x and y negated)."""
        m = np.asarray(ijk_to_ras, dtype=np.float64).reshape(4, 4)
        spacing = np.linalg.norm(m[:3, :3], axis=0)
        if not np.all(spacing > 0):
            raise ValueError("IJK-to-RAS matrix has a zero column")
        flip = np.diag([-1.0, -1.0, 1.0])
        direction = flip @ (m[:3, :3] / spacing)
        origin = flip @ m[:3, 3]
        return ScalarVolumeNode(array, tuple(float(v) for v in spacing), tuple(float(v) for v in origin),
                                tuple(float(v) for v in direction.reshape(9)), name)


class MarkupsFiducialNode:
    """vtkMRMLMarkupsFiducialNode subset used by the reference (Mamri.py:1313-1317, 1345-1347)."""

    def __init__(self, name: str):
        self.name = name
        self._points: List[List[float]] = []
        self._labels: List[str] = []

    def AddControlPoint(self, ras) -> int:
        self._points.append([float(v) for v in ras])
        self._labels.append("")
        return len(self._points) - 1

    AddControlPointWorld = AddControlPoint

    def SetNthControlPointLabel(self, idx: int, label: str) -> None:
        self._labels[idx] = label

    def GetNthControlPointLabel(self, idx: int) -> str:
        return self._labels[idx]

    def GetNumberOfControlPoints(self) -> int:
        return len(self._points)

    def GetNthControlPointPositionWorld(self, idx: int):
        return tuple(self._points[idx])

    def points(self) -> np.ndarray:
        return np.asarray(self._points, dtype=np.float64).reshape(-1, 3)


@dataclasses.dataclass
class SegmentationNode:
    """Stand-in for the "AutoBodySegmentation" node (Mamri.py:1328-1341): the uint8 body labelmap the
    reference imports, plus -- optionally -- the closed-surface points/normals Slicer would derive from
    it (supplied by the host application; the voxel-to-surface conversion is Slicer's)."""
    name: str
    body_mask: object                       # uint8 [nz,ny,nx] (host array or CUDA tensor)
    body_label: int
    spacing: Sequence[float]
    origin: Sequence[float]
    direction: Sequence[float]
    surface_points: Optional[object] = None     # float32 [n,3] RAS
    surface_normals: Optional[object] = None    # float32 [n,3]


@dataclasses.dataclass
class MamriParameterNode:
    """Fields of the reference's parameter node that the two methods read or write (Mamri.py:50-61)."""
    inputVolume: Optional[ScalarVolumeNode] = None
    segmentationNode: Optional[SegmentationNode] = None
    targetFiducialNode: Optional[MarkupsFiducialNode] = None
    entryPointFiducialNode: Optional[MarkupsFiducialNode] = None
    safetyDistance: float = 5.0
    useSavedBaseplate: bool = False              # Mamri.py:60: prefer "MamriSavedBaseplateTransform" over the scan's baseplate


class MamriLogic:
    """The fiducial-detection and entry-point parts of the reference's MamriLogic."""

    def __init__(self, device: int = 0) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("MamriLogic (B200 path) needs a CUDA device; there is no CPU fallback")
        # constants of MamriLogic.__init__ (Mamri.py:810-813) and of findAndSetEntryPoint (:1009, :1015-1016)
        self.INTENSITY_THRESHOLD = 65.0
        self.MIN_VOLUME_THRESHOLD = 50.0
        self.MAX_VOLUME_THRESHOLD = 1500.0
        self.DISTANCE_TOLERANCE = 5.0
        self.SEARCH_RADIUS = 80.0
        self.device = int(device)
        self.SAVED_BASEPLATE_TRANSFORM_NODE_NAME = "MamriSavedBaseplateTransform"      # Mamri.py:820
        self.scene: Dict[str, object] = {}
        self._detector: Optional[FiducialDetector] = None
        self.last_detection: Optional[DetectionResult] = None

    # -- helpers ---------------------------------------------------------------------------------
    def _clear_node_by_name(self, name: str) -> None:
        self.scene.pop(name, None)

    def _detector_for(self, dims_xyz) -> FiducialDetector:
        d = self._detector
        if d is None or any(a > b for a, b in zip(dims_xyz, d.max_dims)):
            if d is not None:
                d.close()
            self._detector = d = FiducialDetector(dims_xyz, device=self.device)
        return d

    # -- Mamri.py:1304-1341 ------------------------------------------------------------------------
    def volume_threshold_segmentation(self, pNode: MamriParameterNode) -> None:
        '''Segments the input MRI volume to isolate fiducials and the main anatomical structure.'''
        try:
            vol = pNode.inputVolume
            arr = vol.array
            shape = tuple(arr.shape)
            if len(shape) != 3:
                raise ValueError(f"volume must be 3-D, got shape {shape}")
        except Exception as e:                       # the reference's guarded pull (Mamri.py:1306-1307)
            logging.error(f"Failed to pull volume: {e}")
            return
        nz, ny, nx = shape
        det = self._detector_for((nx, ny, nz))
        params = DetectParams(lower=self.INTENSITY_THRESHOLD, upper=65535.0, close_radius=2, connectivity=6,
                              min_volume=self.MIN_VOLUME_THRESHOLD, max_volume=self.MAX_VOLUME_THRESHOLD)
        if isinstance(arr, torch.Tensor) and arr.is_cuda:
            res = det.detect(arr, spacing=vol.spacing, origin=vol.origin, direction=vol.direction, params=params,
                             want_body=True)
            body_mask = res.body_mask
        else:
            body_host = np.empty(shape, dtype=np.uint8)
            res = det.detect_host(arr, spacing=vol.spacing, origin=vol.origin, direction=vol.direction, params=params,
                                  body_out=body_host)
            body_mask = body_host
        self.last_detection = res
        fiducials_data = res.fiducials_data
        self._clear_node_by_name("DetectedFiducials")
        if fiducials_data:
            node = MarkupsFiducialNode("DetectedFiducials")
            self.scene["DetectedFiducials"] = node
            for fd in fiducials_data:
                lps = fd["centroid"]
                idx = node.AddControlPoint([-lps[0], -lps[1], lps[2]])
                node.SetNthControlPointLabel(idx, f"M_{fd['id']}_{fd['vol']:.0f}mm³")
        if res.n_labels == 0:
            return                                   # `if not all_labels: return`        (Mamri.py:1319)
        if not res.body_label:
            return                                   # `if not non_fiducial_labels: return` (Mamri.py:1321)
        self._clear_node_by_name("AutoBodySegmentation")
        seg = SegmentationNode(name="AutoBodySegmentation", body_mask=body_mask, body_label=res.body_label,
                               spacing=tuple(vol.spacing), origin=tuple(vol.origin), direction=tuple(vol.direction))
        self.scene["AutoBodySegmentation"] = seg
        pNode.segmentationNode = seg

    # -- Mamri.py:1343-1363 -----------------------------------------------------------------------
    def joint_detection(self, pNode: MamriParameterNode = None) -> Dict[str, List[Dict]]:
        '''Identifies known L-shaped fiducial patterns from a list of detected points.'''
        all_node = self.scene.get("DetectedFiducials")
        if not (all_node and all_node.GetNumberOfControlPoints() >= 3):
            return {}
        det = self._detector or self._detector_for((32, 32, 32))
        pts = all_node.points()
        self.last_pose = pose = det.pose_estimate([pts])[0]
        return {jn: [{"id": i, "ras_coords": [float(c) for c in pts[i]]} for i in ids] for jn, ids in pose.identified.items()}

    # -- Mamri.py:858-870 (the part of process() after the segmentation, on the device) ------------
    def estimate_pose(self, pNode: MamriParameterNode = None, apply_correction: bool = False):
        '''Matching, baseplate registration and full-chain IK; returns a detector.PoseResult (or None when fewer
        than three fiducials were detected).  The reference's SciPy solver remains the parity path for the
        1e-6 rad criterion; this is its batched on-device counterpart.'''
        all_node = self.scene.get("DetectedFiducials")
        if not (all_node and all_node.GetNumberOfControlPoints() >= 3):
            return None
        det = self._detector or self._detector_for((32, 32, 32))
        # _get_baseplate_transform (Mamri.py:1376-1408): the saved transform node first when the parameter node says so,
        # the scan's baseplate otherwise, the saved node again as the fall-back; the IK starts from the current joint
        # angles and from zeros (:1425)
        saved = self.scene.get(self.SAVED_BASEPLATE_TRANSFORM_NODE_NAME)
        prefer = bool(pNode is not None and getattr(pNode, "useSavedBaseplate", False))
        if prefer and saved is None:
            logging.warning(f"'Use Saved Transform' is checked, but node '{self.SAVED_BASEPLATE_TRANSFORM_NODE_NAME}' was not "
                            "found. Will attempt detection from scan.")
        current = getattr(self, "current_joint_angles", None)
        self.last_pose = det.pose_estimate([all_node.points()], apply_correction=apply_correction, saved_base=saved,
                                           prefer_saved_base=prefer, initial_angles=None if current is None else [current])[0]
        if self.last_pose.base_source == "saved" and not prefer:
            logging.info("Baseplate not found in scan; successfully used saved transform instead.")
        elif self.last_pose.base_matrix is None:
            logging.error("Pose estimation failed. A scan containing the baseplate is required, or a previously saved "
                          "baseplate transform must exist.")
        return self.last_pose

    def save_baseplate_transform(self, matrix=None) -> None:
        '''Keeps a baseplate transform in the scene under SAVED_BASEPLATE_TRANSFORM_NODE_NAME (the reference's
        saveBaseplateTransform, Mamri.py:883-907): the given 4x4, or the one of the last estimated pose.'''
        if matrix is None:
            pose = getattr(self, "last_pose", None)
            matrix = None if pose is None else pose.base_matrix
        if matrix is None:
            raise ValueError("no baseplate transform to save: estimate a pose from a scan with the baseplate first")
        self.scene[self.SAVED_BASEPLATE_TRANSFORM_NODE_NAME] = np.array(matrix, dtype=np.float64).reshape(4, 4)

    # -- Mamri.py:850-881 (without the scene/model building and the motor-step conversion) ----------
    def process(self, pNode: MamriParameterNode, apply_correction: bool = False):
        '''Executes the pipeline: segmentation, fiducial detection, baseplate registration and inverse kinematics.
        Returns the joint angles (rad, Joint1..Joint6) or None, like the first element of the reference's result:
        None when there is no baseplate transform, neither from the scan nor saved (`:861-864`), or the Joint6 markers were not found
        (`:868-875`).  Details of the run stay in `last_detection` / `last_pose`.'''
        self.volume_threshold_segmentation(pNode)
        pose = self.estimate_pose(pNode, apply_correction=apply_correction)
        if pose is None or pose.base_matrix is None:
            return None
        if pose.joint_angles is None:
            logging.info("Prerequisites for full-chain IK not met (e.g., Joint6 markers not found). Cannot estimate pose.")
            return None
        self.current_joint_angles = np.array(pose.joint_angles)        # the next IK starts here (Mamri.py:1425)
        return pose.joint_angles

    # -- Mamri.py:987-1033 -------------------------------------------------------------------------
    def findAndSetEntryPoint(self, pNode: MamriParameterNode) -> None:
        '''Finds and marks the closest suitable entry point on the body surface for the biopsy needle.'''
        targetNode = pNode.targetFiducialNode
        segmentationNode = self.scene.get("AutoBodySegmentation")
        if not (targetNode and targetNode.GetNumberOfControlPoints() > 0 and segmentationNode):
            logging.error("Please place a target marker and ensure 'AutoBodySegmentation' exists.")
            return
        pts, nrm = segmentationNode.surface_points, segmentationNode.surface_normals
        dev = torch.device(f"cuda:{self.device}")
        if pts is None or nrm is None:
            # no closed-surface representation supplied by the host application: derive the candidates from
            # the body labelmap on the GPU (mamri_body_surface; stands in for Mamri.py:994-1003)
            bm = segmentationNode.body_mask
            if bm is None:
                return                               # `if not body_poly: return` (Mamri.py:995-996)
            bm_t = bm if isinstance(bm, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bm))
            bm_t = bm_t.to(dev).contiguous()
            nz, ny, nx = bm_t.shape
            det = self._detector_for((nx, ny, nz))
            pts_t, nrm_t = det.body_surface(bm_t, spacing=segmentationNode.spacing, origin=segmentationNode.origin,
                                            direction=segmentationNode.direction)
            segmentationNode.surface_points, segmentationNode.surface_normals = pts_t, nrm_t
        else:
            pts_t = torch.as_tensor(pts, dtype=torch.float32).to(dev).contiguous()
            nrm_t = torch.as_tensor(nrm, dtype=torch.float32).to(dev).contiguous()
        if len(pts_t) == 0:
            return                                   # `if not body_poly: return` (Mamri.py:995-996)
        target_pos = np.array(targetNode.GetNthControlPointPositionWorld(0), dtype=np.float64)
        det = self._detector or self._detector_for((32, 32, 32))
        best = det.entry_search(pts_t, nrm_t, target_pos, radius=self.SEARCH_RADIUS, wx=1.0, wy=-2.0, cutoff=-0.5)
        if best["index"] < 0:
            logging.warning(f"Could not find a suitable side-entry point within {self.SEARCH_RADIUS}mm of the target.")
            return
        newNodeName = "ClosestSuitableEntryPoint"
        self._clear_node_by_name(newNodeName)
        entryNode = MarkupsFiducialNode(newNodeName)
        self.scene[newNodeName] = entryNode
        entryNode.AddControlPointWorld(best["point"])
        entryNode.SetNthControlPointLabel(0, "Suitable Entry")
        pNode.entryPointFiducialNode = entryNode
