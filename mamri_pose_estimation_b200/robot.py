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


# ------------------------------------------------------------------------------------------------
# robot_config.json -> mamri_robot (Mamri.py:1577-1613 _load_robot_definition, :816-819)
# ------------------------------------------------------------------------------------------------
def robot_from_definition(definition: Sequence[dict], articulated_chain: Sequence[str] = ARTICULATED_CHAIN,
                          base_link: str = "Baseplate", effector_link: str = "Joint6", secondary_link: str = "Joint4",
                          distance_tolerance: float = 5.0, secondary_weight: float = 0.05, apply_correction: bool = False):
    """Fills a ``mamri_robot`` (ctypes mirror ``_capi.Robot``) from the list of link dictionaries the reference keeps in
    ``robot_config.json`` -- same file order (``joint_detection`` iterates the links in this order, Mamri.py:1349), the
    same keys (``name, parent, fixed_offset_to_parent{translate}, has_markers, local_marker_coords, arm_lengths,
    articulation_axis, joint_limits``).  The articulated chain, the effector / secondary links and the two constants are
    the reference's (Mamri.py:813, 819, 1414-1424, 1507).  A ``rotate`` entry in ``fixed_offset_to_parent`` (which
    _load_robot_definition would apply, and the shipped file does not use) is rejected: ``mamri_link`` carries a
    translation only."""
    from . import _capi
    if not 1 <= len(definition) <= _capi.MAX_LINKS:
        raise ValueError(f"robot definition has {len(definition)} links; mamri_robot holds 1..{_capi.MAX_LINKS}")
    names = [d["name"] for d in definition]
    r = _capi.Robot()
    r.n_links = len(definition)
    for key, want in (("base_link", base_link), ("effector_link", effector_link)):
        if want not in names:
            raise ValueError(f"{key} {want!r} is not in the robot definition")
        setattr(r, key, names.index(want))
    r.secondary_link = names.index(secondary_link) if secondary_link in names else -1
    r.distance_tolerance, r.secondary_weight = float(distance_tolerance), float(secondary_weight)
    r.apply_correction = int(bool(apply_correction))
    chain = list(articulated_chain)
    if len(chain) > _capi.MAX_CHAIN:
        raise ValueError(f"articulated chain longer than {_capi.MAX_CHAIN}")
    for i, d in enumerate(definition):
        l = r.links[i]
        parent = d.get("parent")
        if parent is None:
            l.parent = -1
        else:
            if parent not in names[:i]:
                raise ValueError(f"link {d['name']!r}: parent {parent!r} must come earlier in the file")
            l.parent = names.index(parent)
        axis = d.get("articulation_axis")
        if axis not in _capi.AXIS_CODES:
            raise ValueError(f"link {d['name']!r}: unknown articulation_axis {axis!r}")
        l.axis = _capi.AXIS_CODES[axis]
        off = d.get("fixed_offset_to_parent")
        if isinstance(off, dict):
            if off.get("rotate"):
                raise ValueError(f"link {d['name']!r}: fixed_offset_to_parent.rotate is not supported (translation only)")
            l.translate[:] = [float(v) for v in off.get("translate", (0.0, 0.0, 0.0))]
        markers = d.get("local_marker_coords") if d.get("has_markers") else None
        l.has_markers = int(bool(markers))
        if markers:
            if len(markers) != 3:
                raise ValueError(f"link {d['name']!r}: exactly three markers per link (L-shape), got {len(markers)}")
            l.marker_coords[:] = [float(v) for m in markers for v in m]
            l.arm_lengths[:] = [float(v) for v in d["arm_lengths"]]
        l.chain_index = chain.index(d["name"]) if d["name"] in chain else -1
        lim = d.get("joint_limits")
        if lim is not None:
            l.limits_deg[:] = [float(lim[0]), float(lim[1])]
    return r


def load_robot_config(path: str, **kw):
    """``robot_config.json`` (the reference's ``Mamri/Resources/Robot/robot_config.json``) -> ``mamri_robot``."""
    import json
    with open(path) as f:
        return robot_from_definition(json.load(f), **kw)


def default_definition() -> List[dict]:
    """The constants of this module in robot_config.json's own shape (what ``mamri_default_robot`` hard-codes)."""
    limits = {"Joint1": (-180, 180), "Joint2": (-120, 120), "Joint3": (-120, 120), "Joint4": (-180, 180),
              "Joint5": (-120, 120), "Joint6": (-270, 270)}
    out = []
    for l in LINKS:
        d = {"name": l["name"], "parent": l["parent"],
             "fixed_offset_to_parent": {"translate": list(l["translate"])} if l["parent"] else None,
             "has_markers": "markers" in l, "articulation_axis": l["axis"]}
        if "markers" in l:
            d["local_marker_coords"] = [list(m) for m in l["markers"]]
            d["arm_lengths"] = list(l["arm_lengths"])
        if l["name"] in limits:
            d["joint_limits"] = list(limits[l["name"]])
        out.append(d)
    out.append({"name": "Needle", "parent": "Joint6", "fixed_offset_to_parent": {"translate": [-50.0, 0.0, 71.0]},
                "has_markers": False, "articulation_axis": "TRANS_X", "joint_limits": [0, 0]})
    return out
