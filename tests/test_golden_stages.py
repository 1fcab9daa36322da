"""Golden fixtures of the stages around the segmentation (tests/golden/stages/*.npz, made by
tests/golden/make_golden_stages.py): the oracles on CPU, and the CUDA path on the GPU, must reproduce them."""
import math
import os

import numpy as np
import pytest

from oracle import kinematics as kin
from oracle import segmentation as seg
from oracle import surface as srf

DIR = os.path.join(os.path.dirname(__file__), "golden", "stages")
NAMES = [j["name"] for j in kin.ROBOT]


def _unpack(bits, shape):
    shape = tuple(int(v) for v in shape)
    return np.unpackbits(bits)[:int(np.prod(shape))].reshape(shape)


def _surface():
    g = np.load(os.path.join(DIR, "s1_surface_72x56x40.npz"))
    body = _unpack(g["body_bits"], g["shape"])
    geom = seg.Geometry(tuple(g["spacing"]), tuple(g["origin"]), tuple(g["direction"]))
    return g, body, geom


def test_surface_oracle_reproduces_golden():
    g, body, geom = _surface()
    pts, nrm, lin = srf.body_surface(body, geom)
    assert np.array_equal(pts, g["points"]) and np.array_equal(nrm, g["normals"]) and np.array_equal(lin, g["linear_index"])
    wi, wd = kin.find_entry_point(pts, nrm, g["target"])
    assert wi == int(g["entry_index"]) and wd == float(g["entry_distance"])


def test_pose_oracle_reproduces_golden():
    g = np.load(os.path.join(DIR, "p1_pose_12_scans.npz"))
    for i in range(len(g["counts"])):
        pts = g["points"][i, :g["counts"][i]]
        ang, ident, base = kin.pose_from_markers(pts)
        m = -np.ones((16, 3), dtype=np.int32)
        for jn, ms in ident.items():
            m[NAMES.index(jn)] = [q["id"] for q in ms]
        assert np.array_equal(m, g["matched"][i])
        assert (base is not None) == bool(g["has_base"][i])
        if base is not None:
            assert np.abs(base - g["base"][i]).max() < 1e-12
        assert (ang is not None) == bool(g["has_ik"][i])
        if ang is not None:                                   # SciPy's iterate: stable to its own stopping tolerance across builds
            assert np.abs(ang - g["scipy_angles"][i]).max() < 1e-6


def test_collision_oracle_reproduces_golden():
    g = np.load(os.path.join(DIR, "k1_collision_24_configs.npz"))
    body = _unpack(g["body_bits"], g["shape"])
    parts = {str(n): p for n, p in zip(g["part_names"], g["part_points"])}
    for c, want_mask, want_n in zip(g["configs"], g["link_mask"], g["n_inside"]):
        links, n_in = kin.check_collision_voxel(c, g["base"], parts, body, g["ras_to_index"])
        assert sum(1 << NAMES.index(l) for l in links) == int(want_mask) and n_in == int(want_n)


@pytest.mark.gpu
def test_cuda_reproduces_stage_goldens(cuda_lib):
    import torch
    from mamri_pose_estimation_b200.detector import FiducialDetector
    det = FiducialDetector((72, 56, 40))
    # skin-surface candidates + entry search
    g, body, geom = _surface()
    pts, nrm = det.body_surface(torch.from_numpy(body).cuda(), spacing=geom.spacing, origin=geom.origin, direction=geom.direction)
    assert np.array_equal(pts.cpu().numpy(), g["points"]) and np.array_equal(nrm.cpu().numpy(), g["normals"])
    r = det.entry_search(pts, nrm, g["target"])
    assert r["index"] == int(g["entry_index"]) and r["distance"] == float(g["entry_distance"])
    # matching + registration + IK (the device solver is SciPy's algorithm restated: it ends where SciPy does)
    g = np.load(os.path.join(DIR, "p1_pose_12_scans.npz"))
    poses = det.pose_estimate([g["points"][i, :g["counts"][i]] for i in range(len(g["counts"]))])
    agree, n_ik = 0, 0
    for i, p in enumerate(poses):
        m = -np.ones((16, 3), dtype=np.int32)
        for jn, ids in p.identified.items():
            m[NAMES.index(jn)] = ids
        assert np.array_equal(m, g["matched"][i])
        assert (p.base_matrix is not None) == bool(g["has_base"][i])
        if p.base_matrix is not None:
            assert np.abs(p.base_matrix - g["base"][i]).max() < 1e-9
        assert (p.joint_angles is not None) == bool(g["has_ik"][i])
        if p.joint_angles is not None:
            n_ik += 1
            # same iterates as SciPy: 1e-5 rad where the markers fit the model, 1e-2 in the flat valleys of a misfit
            agree += bool(np.abs(p.joint_angles - g["scipy_angles"][i]).max() < (1e-5 if p.ik_cost < 10.0 else 1e-2))
    assert agree >= math.ceil(0.95 * n_ik), (agree, n_ik)
    # collision sampling
    g = np.load(os.path.join(DIR, "k1_collision_24_configs.npz"))
    body = _unpack(g["body_bits"], g["shape"])
    parts = {str(n): p for n, p in zip(g["part_names"], g["part_points"])}
    got = det.collision_check(parts, g["configs"], g["base"], torch.from_numpy(body).cuda(), g["ras_to_index"])
    for r, want_mask, want_n in zip(got, g["link_mask"], g["n_inside"]):
        assert sum(1 << NAMES.index(l) for l in r["links"]) == int(want_mask) and r["n_points_inside"] == int(want_n)
    det.close()
