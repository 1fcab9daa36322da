"""Pins the oracle AND the CUDA path to real SimpleITK output -- when fixtures made by tools/make_itk_golden.py with
SimpleITK are present under tests/golden/itk/ (they cannot be produced in the build container: SimpleITK is not
installable there).  Without them the ITK comparisons are skipped and only the kit's plumbing is exercised, on
fixtures the kit writes from the oracle (`--backend oracle`; those prove nothing about ITK and say so).

Bars (BASELINE.json north_star): masks and label volumes bit-exact, centroids within 1e-4 voxel, principal axes
within 1e-3 (up to sign, where the principal moments are separated)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from oracle import segmentation as seg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ITK_DIR = os.path.join(ROOT, "tests", "golden", "itk")
ITK_FIXTURES = sorted(glob.glob(os.path.join(ITK_DIR, "*.npz")))


def _kit():
    spec = importlib.util.spec_from_file_location("make_itk_golden", os.path.join(ROOT, "tools", "make_itk_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _bits(g, key, shape):
    return np.unpackbits(g[key])[:int(np.prod(shape))].reshape(shape)


def _check_axes(axes, moments, ref_axes, ref_moments, tol=1e-3):
    assert np.allclose(moments, ref_moments, rtol=1e-9, atol=1e-9)
    a, r = np.asarray(axes).reshape(3, 3), np.asarray(ref_axes).reshape(3, 3)
    gaps = np.diff(ref_moments)
    for i in range(3):
        separated = (i == 0 or gaps[i - 1] > 1e-6 * max(1.0, abs(ref_moments[i]))) and \
                    (i == 2 or gaps[i] > 1e-6 * max(1.0, abs(ref_moments[i])))
        if separated:
            assert min(np.abs(a[i] - r[i]).max(), np.abs(a[i] + r[i]).max()) <= tol


def compare_oracle(path):
    """The oracle against one fixture, stage by stage."""
    g = np.load(path)
    vol = g["volume"]
    shape = vol.shape
    geom = seg.Geometry(tuple(g["spacing"]), tuple(g["origin"]), tuple(g["direction"]))
    # the structuring element ITK used
    for r in (1, 2, 3):
        b = g[f"ball{r}"]
        assert np.array_equal(b[r:3 * r + 1, r:3 * r + 1, r:3 * r + 1].astype(bool), seg.ball_structure(r)), f"ball {r}"
        assert b.sum() == seg.ball_structure(r).sum()
    thr = seg.binary_threshold(vol, float(g["lo"]), float(g["hi"]))
    assert np.array_equal(thr, _bits(g, "threshold_bits", shape)), "BinaryThreshold"
    opened = seg.binary_opening(thr, int(g["open_radius"]))
    assert np.array_equal(opened, _bits(g, "opened_bits", shape)), "BinaryMorphologicalOpening"
    closed = seg.binary_closing_safe_border(opened, int(g["radius"]))
    assert np.array_equal(closed, _bits(g, "closed_bits", shape)), "BinaryMorphologicalClosing"
    labels, k = seg.connected_components(closed, int(g["conn"]))
    assert k == len(g["label_ids"]) and np.array_equal(g["label_ids"], np.arange(1, k + 1))
    assert np.array_equal(labels, g["labels"]), "ConnectedComponent numbering"
    stats = seg.label_shape_statistics(labels, k, geom)
    a_inv = np.linalg.inv(geom.matrix())
    for s, n, size, cen, pm, pa in zip(stats, g["counts"], g["physical_size"], g["centroid"], g["principal_moments"],
                                       g["principal_axes"]):
        assert s.count == int(n)
        assert s.physical_size == pytest.approx(float(size), rel=1e-12)
        assert np.abs(a_inv @ (s.centroid - cen)).max() <= 1e-4, "centroid (voxel units)"
        if s.count >= 8:
            _check_axes(s.principal_axes, s.principal_moments, pa, pm)
    kept, body = seg.select_candidates(np.asarray(g["counts"]), geom.voxel_volume(), float(g["min_vol"]), float(g["max_vol"]))
    assert kept == g["kept"].tolist() and body == int(g["body_label"])
    return g, closed, labels


def compare_cuda(path):
    import torch
    from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector
    g = np.load(path)
    vol = g["volume"]
    nz, ny, nx = vol.shape
    geom = seg.Geometry(tuple(g["spacing"]), tuple(g["origin"]), tuple(g["direction"]))
    det = FiducialDetector((nx, ny, nz))
    res = det.detect(torch.from_numpy(vol).cuda(), spacing=geom.spacing, origin=geom.origin, direction=geom.direction,
                     params=DetectParams(lower=float(g["lo"]), upper=float(g["hi"]), close_radius=int(g["radius"]),
                                         open_radius=int(g["open_radius"]), connectivity=int(g["conn"]),
                                         min_volume=float(g["min_vol"]), max_volume=float(g["max_vol"])),
                     want_mask=True, want_labels=True)
    assert np.array_equal(res.mask.cpu().numpy(), _bits(g, "closed_bits", vol.shape)), "closed mask"
    assert np.array_equal(res.labels.cpu().numpy().view(np.uint32), g["labels"]), "label volume"
    assert np.array_equal(det.label_counts(res.n_labels).astype(np.int64), g["counts"])
    assert [m.label for m in res.markers] == g["kept"].tolist()
    assert res.body_label == int(g["body_label"])
    a_inv = np.linalg.inv(geom.matrix())
    for m in res.markers:
        i = m.label - 1
        assert m.volume_mm3 == pytest.approx(float(g["physical_size"][i]), rel=1e-12)
        assert np.abs(a_inv @ (np.array(m.centroid_lps) - g["centroid"][i])).max() <= 1e-4
        if m.count >= 8:
            _check_axes(m.principal_axes, m.principal_moments, g["principal_axes"][i], g["principal_moments"][i])
    det.close()


# ---- real ITK fixtures (present only after someone ran the kit where SimpleITK exists)
@pytest.mark.skipif(not ITK_FIXTURES, reason="no tests/golden/itk/*.npz: run tools/make_itk_golden.py where SimpleITK is installed")
@pytest.mark.parametrize("path", ITK_FIXTURES, ids=[os.path.basename(p) for p in ITK_FIXTURES])
def test_oracle_matches_simpleitk(path):
    g, _, _ = compare_oracle(path)
    assert str(g["backend"]).startswith("SimpleITK"), "fixtures under tests/golden/itk must come from SimpleITK"


@pytest.mark.gpu
@pytest.mark.skipif(not ITK_FIXTURES, reason="no tests/golden/itk/*.npz: run tools/make_itk_golden.py where SimpleITK is installed")
@pytest.mark.parametrize("path", ITK_FIXTURES, ids=[os.path.basename(p) for p in ITK_FIXTURES])
def test_cuda_matches_simpleitk(cuda_lib, path):
    compare_cuda(path)


# ---- the kit itself, on fixtures it writes from the oracle
@pytest.fixture(scope="module")
def oracle_fixtures(tmp_path_factory):
    out = tmp_path_factory.mktemp("itk_kit")
    _kit().main(["--backend", "oracle", "--out", str(out)])
    files = sorted(glob.glob(os.path.join(str(out), "*.npz")))
    assert len(files) == len(_kit().CASES)
    return files


def test_kit_round_trip_with_oracle_backend(oracle_fixtures):
    for path in oracle_fixtures:
        g, closed, labels = compare_oracle(path)
        assert str(g["backend"]) == "oracle"
    # the cases cover what they claim to
    names = " ".join(os.path.basename(p) for p in oracle_fixtures)
    for need in ("c26", "r1", "r3", "open", "int16", "float32", "flip", "ragged"):
        assert need in names


@pytest.mark.gpu
def test_kit_fixtures_through_cuda(cuda_lib, oracle_fixtures):
    for path in oracle_fixtures:
        compare_cuda(path)
