"""Robot geometry the fiducial path's callers need (marker coordinates, link
offsets, articulation axes), restated from the reference's
``Mamri/Resources/Robot/robot_config.json`` (lines cited per entry) and its
forward-kinematics convention (``MamriLogic._get_world_transform_for_joint``,
``Mamri/Mamri.py:1486-1505``; ``_get_rotation_transform`` ``:1760-1769``).

Used by the synthetic-phantom generator to put fiducials where a posed robot
would carry them.  Host-side float64 NumPy; nothing here is on the GPU path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

ARTICULATED_CHAIN = ("Joint1", "Joint2", "Joint3", "Joint4", "Joint5", "Joint6")  # Mamri.py:819

# (name, parent, translate, axis, markers, arm_lengths)   robot_config.json line ranges
LINKS: List[dict] = [
    dict(name="Baseplate", parent=None, translate=(0.0, 0.0, 0.0), axis=None,           # :2-17
         markers=((-10.0, 20.0, 5.0), (10.0, 20.0, 5.0), (-10.0, -20.0, 5.0)), arm_lengths=(40.0, 20.0)),
    dict(name="Joint1", parent="Baseplate", translate=(0.0, 0.0, 20.0), axis="IS"),      # :18-30
    dict(name="Joint2", parent="Joint1", translate=(0.0, 0.0, 30.0), axis="PA",          # :31-49
         markers=((12.5, 45.0, 110.0), (-12.5, 45.0, 110.0), (12.5, 45.0, 40.0)), arm_lengths=(70.0, 25.0)),
    dict(name="Joint3", parent="Joint2", translate=(0.0, 0.0, 150.0), axis="PA"),        # :50-62
    dict(name="Joint4", parent="Joint3", translate=(0.0, 0.0, 0.0), axis="IS",           # :63-81
         markers=((-10.0, 35.0, 90.0), (10.0, 35.0, 90.0), (-10.0, -35.0, 90.0)), arm_lengths=(70.0, 20.0)),
    dict(name="Joint5", parent="Joint4", translate=(0.0, 0.0, 155.0), axis="PA"),        # :82-94
    dict(name="Joint6", parent="Joint5", translate=(0.0, 0.0, 13.0), axis="IS",          # :95-115
         markers=((-10.0, 22.5, 26.0), (10.0, 22.5, 26.0), (-10.0, -22.5, 26.0)), arm_lengths=(45.0, 20.0)),
]
LINK_BY_NAME: Dict[str, dict] = {l["name"]: l for l in LINKS}
MARKER_LINKS = tuple(l["name"] for l in LINKS if "markers" in l)


def _rot(axis: str, deg: float) -> np.ndarray:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4)
    if axis == "Z":
        m[:2, :2] = [[c, -s], [s, c]]
    elif axis == "Y":
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    elif axis == "X":
        m[1:3, 1:3] = [[c, -s], [s, c]]
    return m


def articulation(axis: str, angle_rad: float) -> np.ndarray:
    deg = math.degrees(angle_rad)
    if axis == "IS":
        return _rot("Z", deg)
    if axis == "PA":
        return _rot("Y", -deg)
    if axis == "LR":
        return _rot("X", deg)
    return np.eye(4)


def link_world_transforms(angles_rad: Sequence[float], base: np.ndarray) -> Dict[str, np.ndarray]:
    """World (RAS) transform of every link: parent @ translate @ articulation."""
    ang = dict(zip(ARTICULATED_CHAIN, angles_rad))
    world: Dict[str, np.ndarray] = {}
    for l in LINKS:
        t = np.eye(4)
        t[:3, 3] = l["translate"]
        art = articulation(l["axis"], ang.get(l["name"], 0.0)) if l["axis"] else np.eye(4)
        parent = world[l["parent"]] if l["parent"] else np.asarray(base, dtype=np.float64)
        world[l["name"]] = parent @ t @ art
    return world


def marker_positions_ras(angles_rad: Sequence[float], base: np.ndarray,
                         links: Sequence[str] = MARKER_LINKS) -> Dict[str, np.ndarray]:
    world = link_world_transforms(angles_rad, base)
    out = {}
    for name in links:
        loc = np.asarray(LINK_BY_NAME[name]["markers"], dtype=np.float64)
        tf = world[name]
        out[name] = loc @ tf[:3, :3].T + tf[:3, 3]
    return out
