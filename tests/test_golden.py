"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle on CPU, and
the CUDA path on the GPU, must reproduce them bit for bit."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import segmentation as seg

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    g = np.load(path)
    vol = g["volume"]
    closed = np.unpackbits(g["closed_bits"])[:vol.size].reshape(vol.shape)
    geom = seg.Geometry(tuple(g["spacing"]), tuple(g["origin"]), tuple(g["direction"]))
    return g, vol, closed, geom


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracles_reproduce_golden(path):
    g, vol, closed, geom = _load(path)
    kw = dict(close_radius=int(g["radius"]), connectivity=int(g["conn"]), min_vol=float(g["min_vol"]), max_vol=float(g["max_vol"]))
    for det in (seg.detect_fiducials(vol, geom, **kw), c_oracle.detect_fiducials(vol, geom, **kw)):
        assert np.array_equal(det.closed, closed)
        assert np.array_equal(det.labels, g["labels"].astype(np.uint32))
        assert np.array_equal(det.counts, g["counts"])
        assert det.body_label == int(g["body_label"])
        assert [f["id"] for f in det.fiducials] == g["markers"][:, 0].astype(int).tolist()
        for f, row in zip(det.fiducials, g["markers"]):
            assert f["vol"] == row[2] and np.allclose(f["centroid"], row[3:6], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_reproduces_golden(cuda_lib, path):
    import torch
    from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector
    g, vol, closed, geom = _load(path)
    nz, ny, nx = vol.shape
    det = FiducialDetector((nx, ny, nz))
    res = det.detect(torch.from_numpy(vol).cuda(), spacing=geom.spacing, origin=geom.origin, direction=geom.direction,
                     params=DetectParams(close_radius=int(g["radius"]), connectivity=int(g["conn"]),
                                         min_volume=float(g["min_vol"]), max_volume=float(g["max_vol"])),
                     want_mask=True, want_labels=True)
    assert np.array_equal(res.mask.cpu().numpy(), closed)
    assert np.array_equal(res.labels.cpu().numpy().view(np.uint32), g["labels"].astype(np.uint32))
    assert np.array_equal(det.label_counts(res.n_labels).astype(np.int64), g["counts"])
    assert res.body_label == int(g["body_label"])
    assert [m.label for m in res.markers] == g["markers"][:, 0].astype(int).tolist()
    for m, row in zip(res.markers, g["markers"]):
        assert m.count == int(row[1]) and m.volume_mm3 == row[2]
        assert np.abs(np.array(m.centroid_lps) - row[3:6]).max() <= 1e-9
        assert list(m.sum_idx) == row[6:9].astype(np.int64).tolist()
        assert list(m.sum_mom) == row[9:15].astype(np.int64).tolist()
    det.close()
