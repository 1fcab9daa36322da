"""Golden fixtures of the stages around the segmentation (tests/golden/stages/*.npz), generated from the oracles:
skin-surface candidates (oracle/surface.py), the marker-table consumers (oracle/kinematics.py: matching, registration,
SciPy IK), the entry search and the voxel collision sampling.  The reference ships no golden vectors and its
SimpleITK / VTK / Slicer stack cannot run here: these pin OUR oracles against regressions and give
the GPU tests fixed inputs; they are not outputs of the reference.

    python tests/golden/make_golden_stages.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mamri_pose_estimation_b200 import phantom      # noqa: E402
from mamri_pose_estimation_b200 import robot as rb  # noqa: E402
from oracle import kinematics as kin                # noqa: E402
from oracle import segmentation as seg              # noqa: E402
from oracle import surface as srf                   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stages")


def surface_case():
    ph = phantom.small_phantom(dims=(72, 56, 40), n_fiducials=4, n_blobs=3, seed=31, spacing=(1.4, 1.4, 2.8), flip_lps=True)
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    det = seg.detect_fiducials(vol, geom, min_vol=20.0, max_vol=600.0)
    pts, nrm, lin = srf.body_surface(det.body_mask, geom)
    assert len(pts) > 500
    # the entry search over exactly these candidates
    tgt = pts.astype(np.float64).mean(axis=0) + np.array([12.0, 3.0, -4.0])
    wi, wd = kin.find_entry_point(pts, nrm, tgt)
    assert wi >= 0
    np.savez_compressed(os.path.join(OUT, "s1_surface_72x56x40.npz"), body_bits=np.packbits(det.body_mask), shape=np.array(det.body_mask.shape),
                        spacing=np.array(ph.spacing), origin=np.array(ph.origin), direction=np.array(ph.direction),
                        points=pts, normals=nrm, linear_index=lin, target=tgt, entry_index=np.int64(wi), entry_distance=np.float64(wd))
    print("surface", len(pts), wi, wd)


def pose_case():
    rng = np.random.default_rng(4242)
    scenes, idents, bases, angles, has_base, has_ik = [], [], [], [], [], []
    for i in range(12):
        a = rng.uniform(-0.4, 0.4)
        turn = np.eye(4)
        turn[:3, :3] = [[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]]
        turn[:3, 3] = rng.uniform(-40, 40, 3)
        base = turn @ rb._rot("X", -90.0)
        theta = np.radians(rng.uniform(-30, 30, 6))
        links = ("Baseplate", "Joint4", "Joint6") if i % 3 else ("Baseplate", "Joint6")
        if i == 11:
            links = ("Joint4",)                                   # no baseplate in the scan
        pos = rb.marker_positions_ras(theta, base, links)
        pts = np.concatenate([pos[l][rng.permutation(3)] + rng.normal(0, 0.12, (3, 3)) for l in links]
                             + [rng.uniform(-300, 300, (2, 3)) + np.array([0, 0, 900.0])])
        ang, ident, b = kin.pose_from_markers(pts)
        scenes.append(pts)
        m = -np.ones((16, 3), dtype=np.int32)
        names = [j["name"] for j in kin.ROBOT]
        for jn, ms in ident.items():
            m[names.index(jn)] = [q["id"] for q in ms]
        idents.append(m)
        bases.append(b if b is not None else np.eye(4))
        has_base.append(b is not None)
        angles.append(ang if ang is not None else np.zeros(6))
        has_ik.append(ang is not None)
    mx = max(len(s) for s in scenes)
    P = np.zeros((len(scenes), mx, 3))
    cnt = np.array([len(s) for s in scenes], dtype=np.int32)
    for i, s in enumerate(scenes):
        P[i, :len(s)] = s
    np.savez_compressed(os.path.join(OUT, "p1_pose_12_scans.npz"), points=P, counts=cnt, matched=np.array(idents), base=np.array(bases),
                        has_base=np.array(has_base), scipy_angles=np.array(angles), has_ik=np.array(has_ik))
    print("pose", cnt.tolist(), sum(has_base), sum(has_ik))


def collision_case():
    rng = np.random.default_rng(77)
    dims = (80, 72, 48)
    sp = np.array([5.0, 5.0, 8.0])
    org = -0.5 * sp * (np.array(dims) - 1)
    z, y, x = np.meshgrid(np.arange(dims[2]), np.arange(dims[1]), np.arange(dims[0]), indexing="ij")
    body = ((((x - 39.5) / 22.0) ** 2 + ((y - 20.0) / 13.0) ** 2 + ((z - 23.5) / 18.0) ** 2) <= 1.0).astype(np.uint8)
    m = np.array([[-1 / sp[0], 0, 0, -org[0] / sp[0]], [0, -1 / sp[1], 0, -org[1] / sp[1]], [0, 0, 1 / sp[2], -org[2] / sp[2]]])
    base = phantom.robot_base_matrix()
    base[:3, 3] = [20.0, -170.0, -60.0]
    names = ["Joint1", "Joint2", "Joint3", "Joint4", "Joint5", "Joint6"]
    spans = [(0, 30), (0, 150), (0, 10), (0, 155), (0, 13), (0, 60)]
    parts = {n: np.stack([rng.uniform(-15, 15, 200), rng.uniform(-15, 15, 200), rng.uniform(lo, hi, 200)], axis=1).astype(np.float32)
             for n, (lo, hi) in zip(names, spans)}
    configs = np.radians(rng.uniform(-100, 100, (24, 6)))
    configs[0] = 0.0
    masks, inside = [], []
    all_names = [j["name"] for j in kin.ROBOT]
    for c in configs:
        links, n_in = kin.check_collision_voxel(c, base, parts, body, m)
        masks.append(sum(1 << all_names.index(l) for l in links))
        inside.append(n_in)
    assert 0 < sum(1 for v in masks if v) < len(masks)
    np.savez_compressed(os.path.join(OUT, "k1_collision_24_configs.npz"), body_bits=np.packbits(body), shape=np.array(body.shape),
                        ras_to_index=m, base=base, configs=configs, part_names=np.array(names),
                        part_points=np.stack([parts[n] for n in names]), link_mask=np.array(masks, dtype=np.int64),
                        n_inside=np.array(inside, dtype=np.int64))
    print("collision", masks)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    surface_case()
    pose_case()
    collision_case()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
