"""Host-side logic that needs no GPU: phantom generator, Philox, sharding, table packing."""
import os
import types

import numpy as np
import pytest

from mamri_pose_estimation_b200 import distributed as mdist
from mamri_pose_estimation_b200 import phantom


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    def h(c, k):
        return [int(x[0]) for x in phantom.philox4x32_10([c[0]], [c[1]], [c[2]], [c[3]], k[0], k[1])]
    assert h((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert h((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_phantom_is_deterministic_and_chunk_independent():
    ph = phantom.small_phantom(dims=(33, 21, 9), seed=4)
    a = phantom.generate(ph)
    b = phantom.add_rician_noise(phantom.paint_signal(ph), ph.sigma, ph.seed, ph.scan_index, chunk=1001)
    assert np.array_equal(a, b)
    c = phantom.generate(phantom.small_phantom(dims=(33, 21, 9), seed=4, scan_index=1))
    assert not np.array_equal(a, c), "scan_index selects a different noise stream"


def test_noise_statistics():
    sig = np.zeros((16, 64, 64), dtype=np.uint16)
    v = phantom.add_rician_noise(sig, 20.0, 7).astype(np.float64)
    # Rayleigh: mean sigma*sqrt(pi/2), P(v > 65) = exp(-65^2 / (2*400)) ~ 5.1e-3
    assert abs(v.mean() - 20.0 * np.sqrt(np.pi / 2)) < 0.3
    assert abs((v > 65).mean() - np.exp(-65 ** 2 / 800.0)) < 1.5e-3


def test_baseline_configs_have_the_documented_content():
    c1, c2 = phantom.config_c1(), phantom.config_c2()
    assert c1.dims == (256, 256, 128) and c1.truth["n_fiducials"] == 9 and c1.sigma == 10.0
    assert c2.dims == (512, 512, 256) and c2.truth["n_fiducials"] == 6
    assert phantom.config_c3(5).scan_index == 5 and phantom.config_c3(5).sigma == 15.0
    c4 = phantom.config_c4()
    assert c4.dims == (1024, 1024, 512) and c4.truth["n_fiducials"] == 32 and c4.sigma == 20.0
    pts, nrm, tgt = phantom.surface_candidates(1000)
    assert pts.shape == (1000, 3) and pts.dtype == np.float32 and np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-5)


def test_sharding_is_a_partition():
    for n, w in ((64, 8), (10, 4), (3, 8)):
        seen = sorted(i for r in range(w) for i in mdist.shard_indices(n, r, w))
        assert seen == list(range(n))


def test_pack_table_layout():
    m = types.SimpleNamespace(label=3, count=100, volume_mm3=51.2, centroid_ras=np.array([1.0, 2.0, 3.0]))
    r = types.SimpleNamespace(markers=[m], n_labels=7, body_label=1)
    t = mdist.pack_table([r, r])
    assert t.shape == (2, mdist.TABLE_SLOTS, mdist.TABLE_FIELDS)
    assert t[1, 0].tolist() == [3, 100, 51.2, 1.0, 2.0, 3.0, 7, 1] and not t[:, 1:].any()


def test_ijk_to_ras_geometry_conversion():
    """ScalarVolumeNode.from_ijk_to_ras == the RAS->LPS conversion of PullVolumeFromSlicer (Mamri.py:1306): the
    physical point of every index, computed the ITK way from (spacing, origin, direction), is the LPS image of the
    RAS point the MRML matrix gives."""
    import torch  # noqa: F401  (logic imports torch)
    from mamri_pose_estimation_b200.logic import ScalarVolumeNode
    from oracle import segmentation as seg
    rng = np.random.default_rng(0)
    a = rng.uniform(-0.4, 0.4)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    m = np.eye(4)
    m[:3, :3] = (rot @ np.diag([-1.0, -1.0, 1.0])) * np.array([0.8, 0.9, 1.6])     # an axial MR volume, LPS-ish axes
    m[:3, 3] = [110.0, 95.0, -60.0]
    node = ScalarVolumeNode.from_ijk_to_ras(np.zeros((4, 5, 6), np.uint16), m)
    assert np.allclose(node.spacing, [0.8, 0.9, 1.6])
    geom = seg.Geometry(node.spacing, node.origin, node.direction)
    idx = rng.integers(0, 50, (20, 3)).astype(np.float64)
    ras = idx @ m[:3, :3].T + m[:3, 3]
    lps = np.array([geom.index_to_physical(i) for i in idx]) if hasattr(geom, "index_to_physical") else \
        np.array(node.origin) + (idx * np.array(node.spacing)) @ np.array(node.direction).reshape(3, 3).T
    assert np.allclose(lps, ras * np.array([-1.0, -1.0, 1.0]), atol=1e-9)
    d = np.array(node.direction).reshape(3, 3)
    assert np.allclose(d @ d.T, np.eye(3), atol=1e-12)


def _robots_equal(a, b):
    import ctypes as C
    return bytes(C.string_at(C.addressof(a), C.sizeof(a))) == bytes(C.string_at(C.addressof(b), C.sizeof(b)))


def test_robot_config_loader_reproduces_the_default_robot(cuda_lib, tmp_path):
    """mamri_robot filled from a robot_config.json (Mamri.py:1577-1613) == mamri_default_robot, field for field: from
    this module's own constants written out in the file's shape, and -- where the reference tree is present (the build
    container; never at run time on the GPU box) -- from the reference's file itself."""
    import ctypes as C
    import json
    from mamri_pose_estimation_b200 import _capi, robot
    ref = _capi.Robot()
    cuda_lib.mamri_default_robot(C.byref(ref))
    path = tmp_path / "robot_config.json"
    path.write_text(json.dumps(robot.default_definition()))
    mine = robot.load_robot_config(str(path))
    assert _robots_equal(mine, ref)
    assert robot.load_robot_config(str(path), apply_correction=True).apply_correction == 1
    real = "/root/reference/Mamri/Resources/Robot/robot_config.json"
    if os.path.exists(real):
        assert _robots_equal(robot.load_robot_config(real), ref)
    # malformed files are rejected, not guessed at
    bad = robot.default_definition()
    bad[1]["fixed_offset_to_parent"]["rotate"] = [["x", 90.0]]
    with pytest.raises(ValueError):
        robot.robot_from_definition(bad)
    bad = robot.default_definition()
    bad[2]["parent"] = "Joint5"
    with pytest.raises(ValueError):
        robot.robot_from_definition(bad)
