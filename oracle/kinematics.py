"""Oracle for the consumer chain behind the hot path: L-shape marker matching,
baseplate registration, forward kinematics and the bounded least-squares IK.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Needed only to evaluate the
north_star's "identical joint angles within 1e-6 rad" criterion: the same chain
consumes the CUDA path's markers and the oracle's markers.  VTK is absent, so
vtkTransform / vtkLandmarkTransform are restated in NumPy.

ROBOT restates the values of Mamri/Resources/Robot/robot_config.json that the
chain reads (names, parents, translate offsets, marker coordinates, arm lengths,
articulation axes, limits); tests/test_oracle_kinematics.py compares it with the
JSON when /root/reference is present.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.optimize

DISTANCE_TOLERANCE = 5.0           # Mamri/Mamri.py:813
ARTICULATED_CHAIN = ["Joint1", "Joint2", "Joint3", "Joint4", "Joint5", "Joint6"]  # Mamri.py:819

# robot_config.json, in file order (order matters: joint_detection iterates it, Mamri.py:1349)
ROBOT: List[dict] = [
    dict(name="Baseplate", parent=None, translate=None, has_markers=True,                 # json:2-17
         local_marker_coords=[[-10.0, 20.0, 5.0], [10.0, 20.0, 5.0], [-10.0, -20.0, 5.0]],
         arm_lengths=[40.0, 20.0], articulation_axis=None),
    dict(name="Joint1", parent="Baseplate", translate=[0, 0, 20.0], has_markers=False,     # json:18-30
         articulation_axis="IS", joint_limits=[-180, 180], steps_per_rev=3332),
    dict(name="Joint2", parent="Joint1", translate=[0, 0, 30], has_markers=True,           # json:31-49
         local_marker_coords=[[12.5, 45.0, 110.0], [-12.5, 45.0, 110.0], [12.5, 45.0, 40.0]],
         arm_lengths=[70.0, 25.0], articulation_axis="PA", joint_limits=[-120, 120], steps_per_rev=3332),
    dict(name="Joint3", parent="Joint2", translate=[0, 0, 150], has_markers=False,         # json:50-62
         articulation_axis="PA", joint_limits=[-120, 120], steps_per_rev=3332),
    dict(name="Joint4", parent="Joint3", translate=[0, 0, 0], has_markers=True,            # json:63-81
         local_marker_coords=[[-10, 35.0, 90], [10, 35.0, 90], [-10, -35.0, 90]],
         arm_lengths=[70.0, 20.0], articulation_axis="IS", joint_limits=[-180, 180], steps_per_rev=3332),
    dict(name="Joint5", parent="Joint4", translate=[0, 0, 155], has_markers=False,         # json:82-94
         articulation_axis="PA", joint_limits=[-120, 120], steps_per_rev=3332),
    dict(name="Joint6", parent="Joint5", translate=[0, 0, 13], has_markers=True,           # json:95-115
         local_marker_coords=[[-10, 22.5, 26], [10, 22.5, 26], [-10, -22.5, 26]],
         arm_lengths=[45.0, 20.0], articulation_axis="IS", joint_limits=[-270, 270], steps_per_rev=3332),
    dict(name="Needle", parent="Joint6", translate=[-50, 0, 71], has_markers=False,        # json:116-130
         articulation_axis="TRANS_X", joint_limits=[0, 0]),
]
ROBOT_BY_NAME: Dict[str, dict] = {j["name"]: j for j in ROBOT}


# ------------------------------ vtkTransform bits ------------------------------
def translation(t: Sequence[float]) -> np.ndarray:
    m = np.eye(4)
    m[:3, 3] = np.asarray(t, dtype=np.float64)
    return m


def rotation(angle_deg: float, axis: str) -> np.ndarray:
    """vtkTransform.RotateX/Y/Z(angle_deg) as a 4x4 (right-handed, degrees)."""
    a = math.radians(angle_deg)
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4)
    if axis == "X":
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    elif axis == "Y":
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    elif axis == "Z":
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    else:
        raise ValueError(axis)
    return m


def articulation(angle_deg: float, axis_str: Optional[str]) -> np.ndarray:
    """MamriLogic._get_rotation_transform (Mamri.py:1760-1769): IS -> RotateZ(a),
    PA -> RotateY(-a), LR -> RotateX(a), anything else identity."""
    if axis_str == "IS":
        return rotation(angle_deg, "Z")
    if axis_str == "PA":
        return rotation(-angle_deg, "Y")
    if axis_str == "LR":
        return rotation(angle_deg, "X")
    return np.eye(4)


def fixed_offset(joint: dict) -> np.ndarray:
    """_load_robot_definition (Mamri.py:1602-1612): only 'translate' occurs in the JSON."""
    t = joint.get("translate")
    return translation(t) if t is not None else np.eye(4)


# ------------------------------ forward kinematics ------------------------------
def world_transform_for_joint(joint_angles_rad: Dict[str, float], target: str,
                              base: np.ndarray) -> Optional[np.ndarray]:
    """MamriLogic._get_world_transform_for_joint (Mamri.py:1486-1505):
    world = parent_world @ fixed_offset @ articulation; the Baseplate's own
    world transform is its local transform (identity) because a falsy parent_world
    takes the DeepCopy(local) branch only when base is None -- with a base matrix
    given, Baseplate = base @ I."""
    world: Dict[str, np.ndarray] = {}
    for j in ROBOT:
        name, parent = j["name"], j.get("parent")
        parent_world = base if not parent else world.get(parent)
        if parent_world is None and parent is not None:
            return None
        art = np.eye(4)
        axis = j.get("articulation_axis")
        if axis:
            ang = joint_angles_rad.get(name, 0.0)
            if "TRANS" not in axis:
                art = articulation(math.degrees(ang), axis)
        local = fixed_offset(j) @ art
        world[name] = parent_world @ local if parent_world is not None else local
        if name == target:
            return world[name]
    return world.get(target)


def marker_world_positions(angles_rad: Sequence[float], base: np.ndarray, joint: str) -> np.ndarray:
    tf = world_transform_for_joint(dict(zip(ARTICULATED_CHAIN, angles_rad)), joint, base)
    loc = np.asarray(ROBOT_BY_NAME[joint]["local_marker_coords"], dtype=np.float64)
    return (tf[:3, :3] @ loc.T).T + tf[:3, 3]


# ------------------------------ marker matching ------------------------------
def sort_l_shaped_markers(markers: List[dict], len1: float, len2: float, tol: float):
    """MamriLogic._sort_l_shaped_markers (Mamri.py:1782-1792): corner, short arm, long arm."""
    if len(markers) != 3:
        return None
    pts = [tuple(m["ras_coords"]) for m in markers]
    l_short, l_long = sorted((len1, len2))
    for i in range(3):
        c, p1, p2 = i, (i + 1) % 3, (i + 2) % 3
        d1, d2 = math.dist(pts[c], pts[p1]), math.dist(pts[c], pts[p2])
        if abs(d1 - l_short) <= tol and abs(d2 - l_long) <= tol:
            return [markers[c], markers[p1], markers[p2]]
        if abs(d1 - l_long) <= tol and abs(d2 - l_short) <= tol:
            return [markers[c], markers[p2], markers[p1]]
    return None


def joint_detection(ras_points: np.ndarray, tol: float = DISTANCE_TOLERANCE) -> Dict[str, List[dict]]:
    """MamriLogic.joint_detection (Mamri.py:1343-1363) on the control points of
    "DetectedFiducials" (node order = ascending label).  First matching
    3-combination per marker-bearing link wins; its ids are then consumed."""
    pts = np.asarray(ras_points, dtype=np.float64).reshape(-1, 3)
    if pts.shape[0] < 3:
        return {}
    fids = [{"id": i, "ras_coords": [float(c) for c in pts[i]]} for i in range(pts.shape[0])]
    identified: Dict[str, List[dict]] = {}
    used = set()
    for jc in ROBOT:
        if not jc.get("has_markers"):
            continue
        arm = jc.get("arm_lengths")
        if not arm or len(arm) != 2:
            continue
        l1, l2 = arm
        expected = sorted([l1, l2, math.hypot(l1, l2)])
        avail = [f for f in fids if f["id"] not in used]
        if len(avail) < 3:
            continue
        for combo in itertools.combinations(avail, 3):
            p = [c["ras_coords"] for c in combo]
            d = sorted([math.dist(p[0], p[1]), math.dist(p[0], p[2]), math.dist(p[1], p[2])])
            if all(abs(a - e) <= tol for a, e in zip(d, expected)):
                matched = [dict(c, ras_coords=list(c["ras_coords"])) for c in combo]
                srt = sort_l_shaped_markers(matched, l1, l2, tol)
                identified[jc["name"]] = srt if srt else matched
                used.update(c["id"] for c in combo)
                break
    return identified


def flatten_baseplate(identified: Dict[str, List[dict]]) -> None:
    """_handle_joint_detection_results (Mamri.py:1371-1373): baseplate markers get their mean y."""
    m = identified.get("Baseplate")
    if m and len(m) == 3:
        avg = sum(q["ras_coords"][1] for q in m) / 3.0
        for q in m:
            q["ras_coords"][1] = avg


# ------------------------------ vtkLandmarkTransform (rigid) ------------------------------
def landmark_rigid(source: np.ndarray, target: np.ndarray) -> np.ndarray:
    """vtkLandmarkTransform in RigidBody mode (Horn 1987, unit quaternion), as
    called by _calculate_fiducial_alignment_matrix (Mamri.py:1771-1780).
    vtkPoints stores float32, so both landmark sets are rounded to float32 first."""
    s = np.asarray(source, dtype=np.float32).astype(np.float64)
    t = np.asarray(target, dtype=np.float32).astype(np.float64)
    sc, tc = s.mean(axis=0), t.mean(axis=0)
    a, b = s - sc, t - tc
    m = a.T @ b                                   # M[i][j] = sum a_i b_j
    n = np.array([
        [m[0, 0] + m[1, 1] + m[2, 2], m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]],
        [m[1, 2] - m[2, 1], m[0, 0] - m[1, 1] - m[2, 2], m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]],
        [m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], -m[0, 0] + m[1, 1] - m[2, 2], m[1, 2] + m[2, 1]],
        [m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], -m[0, 0] - m[1, 1] + m[2, 2]]])
    w, v = np.linalg.eigh(n)
    q = v[:, np.argmax(w)]
    qw, qx, qy, qz = q
    r = np.array([
        [qw * qw + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)],
        [2 * (qx * qy + qw * qz), qw * qw - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - qw * qx)],
        [2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz]])
    out = np.eye(4)
    out[:3, :3] = r
    out[:3, 3] = tc - r @ sc
    return out


# ------------------------------ IK ------------------------------
def ik_error(angles_rad, j6_target, base, apply_correction=False, j4_target=None, j4_weight=0.05):
    """MamriLogic._full_chain_ik_error_function (Mamri.py:1507-1536)."""
    vals = dict(zip(ARTICULATED_CHAIN, angles_rad))
    j6_local = [list(p) for p in ROBOT_BY_NAME["Joint6"]["local_marker_coords"]]
    if apply_correction:
        rz = rotation(180.0, "Z")
        j6_local = [list((rz @ np.array(p + [1.0]))[:3]) for p in j6_local]
    tf6 = world_transform_for_joint(vals, "Joint6", base)
    err = []
    for i, p in enumerate(j6_local):
        ph = tf6 @ np.array(list(p) + [1.0])
        pred = ph[:3] / ph[3]
        err.extend(pred[j] - j6_target[i][j] for j in range(3))
    if j4_target is not None:
        tf4 = world_transform_for_joint(vals, "Joint4", base)
        for i, p in enumerate(ROBOT_BY_NAME["Joint4"]["local_marker_coords"]):
            ph = tf4 @ np.array(list(p) + [1.0])
            pred = ph[:3] / ph[3]
            err.extend(j4_weight * (pred[j] - j4_target[i][j]) for j in range(3))
    return err


def solve_full_chain_ik(j6_target, base, apply_correction=False, j4_target=None):
    """MamriLogic._solve_full_chain_ik (Mamri.py:1410-1447): TRF, bounds = joint
    limits, ftol = xtol = 1e-6, two initial guesses (both zero on a fresh scene,
    :1425), best = lowest cost among successes."""
    chain = [ROBOT_BY_NAME[n] for n in ARTICULATED_CHAIN]
    lo = [math.radians(j["joint_limits"][0]) for j in chain]
    hi = [math.radians(j["joint_limits"][1]) for j in chain]
    best, lowest = None, float("inf")
    for guess in ([0.0] * 6, [0.0] * 6):
        res = scipy.optimize.least_squares(ik_error, guess, bounds=(lo, hi),
                                           args=(j6_target, base, apply_correction, j4_target),
                                           method="trf", ftol=1e-6, xtol=1e-6, verbose=0)
        if res.success and res.cost < lowest:
            lowest, best = res.cost, res
    return None if best is None else best.x


def pose_from_markers(ras_points: np.ndarray, apply_correction: bool = False):
    """MamriLogic.process after the segmentation (Mamri.py:858-870): matching ->
    baseplate y-flatten -> rigid registration -> IK.  Returns (angles or None,
    identified joints, baseplate matrix or None)."""
    ident = joint_detection(ras_points)
    flatten_baseplate(ident)
    if "Baseplate" not in ident:
        return None, ident, None
    tgt = np.array([m["ras_coords"] for m in ident["Baseplate"]], dtype=np.float64)
    base = landmark_rigid(np.asarray(ROBOT_BY_NAME["Baseplate"]["local_marker_coords"], dtype=np.float64), tgt)
    if "Joint6" not in ident:
        return None, ident, base
    j6 = [tuple(m["ras_coords"]) for m in ident["Joint6"]]
    j4 = [tuple(m["ras_coords"]) for m in ident["Joint4"]] if "Joint4" in ident else None
    ang = solve_full_chain_ik(j6, base, apply_correction, j4)
    return ang, ident, base


# ------------------------------ entry-point search ------------------------------
def find_entry_point(points: np.ndarray, normals: np.ndarray, target: Sequence[float],
                     radius: float = 80.0, wx: float = 1.0, wy: float = -2.0, cutoff: float = -0.5):
    """The loop of MamriLogic.findAndSetEntryPoint (Mamri.py:1008-1023) over given
    surface points/normals (RAS).  Candidates within ``radius`` of the target
    (vtkStaticPointLocator.FindPointsWithinRadius: distance^2 <= r^2), kept if
    |nx| - 2|ny| > -0.5, winner = minimum distance; ties resolved by the lowest
    point id (the reference's tie order is the locator's bucket order, unpinned).
    Returns (index, distance) or (-1, inf) when nothing is suitable (:1020-1022)."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    n = np.asarray(normals, dtype=np.float64).reshape(-1, 3)
    t = np.asarray(target, dtype=np.float64)
    d = p - t[None, :]
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    dist = np.sqrt(d2)
    score = wx * np.abs(n[:, 0]) + wy * np.abs(n[:, 1])
    ok = (d2 <= radius * radius) & (score > cutoff)
    if not ok.any():
        return -1, float("inf")
    cand = np.flatnonzero(ok)
    best = cand[np.argmin(dist[cand])]
    return int(best), float(dist[best])


# ------------------------------ robot-vs-body collision sampling ------------------------------
def check_collision_voxel(angles_rad: Sequence[float], base: np.ndarray, part_points: Dict[str, np.ndarray],
                          body_mask: np.ndarray, ras_to_index: np.ndarray):
    """Voxel-sampling stand-in for MamriLogic._check_collision (Mamri.py:1555-1575), the definition
    ``pose.cu:k_collision`` implements: every sample point of every link (link-local float32 coordinates) goes
    through the link's world transform (_get_world_transform_for_joint, :1486-1505) and the RAS -> voxel affine
    (nearest voxel, ties to even) and hits if the body labelmap ``body_mask[z, y, x]`` is non-zero there; outside
    the volume is free.  Returns (colliding link names in robot order, number of points inside)."""
    vals = dict(zip(ARTICULATED_CHAIN, angles_rad))
    m = np.asarray(ras_to_index, dtype=np.float64).reshape(3, 4)
    nz, ny, nx = body_mask.shape
    links, inside = [], 0
    for j in ROBOT:
        p = part_points.get(j["name"])
        if p is None or len(p) == 0:
            continue
        tf = world_transform_for_joint(vals, j["name"], base)
        loc = np.asarray(p, dtype=np.float32).astype(np.float64).reshape(-1, 3)
        w = np.stack([((tf[k, 0] * loc[:, 0] + tf[k, 1] * loc[:, 1]) + tf[k, 2] * loc[:, 2]) + tf[k, 3] for k in range(3)], axis=1)
        idx = np.stack([np.rint(((m[k, 0] * w[:, 0] + m[k, 1] * w[:, 1]) + m[k, 2] * w[:, 2]) + m[k, 3]) for k in range(3)], axis=1).astype(np.int64)
        ok = (idx[:, 0] >= 0) & (idx[:, 0] < nx) & (idx[:, 1] >= 0) & (idx[:, 1] < ny) & (idx[:, 2] >= 0) & (idx[:, 2] < nz)
        hit = np.zeros(len(loc), dtype=bool)
        ii = idx[ok]
        hit[ok] = body_mask[ii[:, 2], ii[:, 1], ii[:, 0]] != 0
        if hit.any():
            links.append(j["name"])
            inside += int(hit.sum())
    return links, inside
