"""The plain-C oracle (oracle/c) against the NumPy oracle: same masks, labels, sums."""
import numpy as np
import pytest

from mamri_pose_estimation_b200 import phantom
from oracle import c_oracle
from oracle import segmentation as seg


@pytest.mark.parametrize("seed,conn,dims", [(1, 6, (70, 45, 33)), (2, 26, (64, 40, 20)), (3, 6, (33, 31, 9))])
def test_c_oracle_matches_numpy_oracle(seed, conn, dims):
    ph = phantom.small_phantom(dims=dims, seed=seed, touch_border=True)
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    a = seg.detect_fiducials(vol, geom, connectivity=conn, min_vol=20, max_vol=600)
    b = c_oracle.detect_fiducials(vol, geom, connectivity=conn, min_vol=20, max_vol=600)
    assert np.array_equal(a.closed, b.closed)
    assert np.array_equal(a.labels, b.labels)
    assert a.n_labels == b.n_labels and np.array_equal(a.counts, b.counts)
    assert a.fiducials == b.fiducials and a.body_label == b.body_label
    for sa, sb in zip(a.stats, b.stats):
        assert sa.sum_idx == sb.sum_idx and sa.sum_mom == sb.sum_mom


@pytest.mark.parametrize("radius", [0, 1, 2, 3])
def test_c_closing_all_radii(radius):
    rng = np.random.default_rng(radius)
    m = (rng.random((20, 24, 28)) < 0.3).astype(np.uint8)
    out = np.empty_like(m)
    assert c_oracle.load().oracle_closing(m.ctypes.data, 28, 24, 20, radius, out.ctypes.data) == 0
    assert np.array_equal(out, seg.binary_closing_safe_border(m, radius))


@pytest.mark.parametrize("dtype", ["uint8", "int16", "uint16", "int32", "float32", "float64"])
def test_c_threshold_types(dtype):
    rng = np.random.default_rng(11)
    vol = rng.integers(0, 200, size=(5, 6, 40)).astype(dtype)
    closed, labels, k, sums, _ = c_oracle.run_pipeline(vol, lo=64.5, hi=65535, close_radius=0)
    assert np.array_equal(closed, seg.binary_threshold(vol, 64.5, 65535))
