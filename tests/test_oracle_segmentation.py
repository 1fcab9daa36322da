"""The oracle against itself: two independent restatements (scipy-based `segmentation`, definitional
`bruteforce`) must agree bit for bit, plus the set-theoretic properties the domain offers.  PARITY
UNPINNED at the SimpleITK boundary (see oracle/__init__.py): this is what pins the oracle instead."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from scipy import ndimage

from mamri_pose_estimation_b200 import phantom
from oracle import bruteforce as bf
from oracle import segmentation as seg


def test_ball_sizes_match_itk_counts():
    # FlatStructuringElement<3>::Ball: r=1 -> 19, r=2 -> 81, r=3 -> 179 voxels
    for r, n in ((1, 19), (2, 81), (3, 179)):
        assert len(seg.ball_offsets(r)) == n
        assert len(bf.ball(r)) == n
        assert sorted(map(tuple, seg.ball_offsets(r).tolist())) == sorted(bf.ball(r))
    # radius 2 on bit rows: 9 rows with half-width 2, 12 rows with half-width 1 (9*5 + 12*3 = 81)
    off = seg.ball_offsets(2)
    rows = {}
    for dz, dy, dx in off:
        rows[(dz, dy)] = max(rows.get((dz, dy), 0), abs(dx))
    assert sorted(rows.values()) == [1] * 12 + [2] * 9


@pytest.mark.parametrize("radius", [1, 2, 3])
@pytest.mark.parametrize("p", [0.02, 0.3, 0.7])
def test_closing_two_restatements_agree(radius, p):
    rng = np.random.default_rng(radius * 10 + int(p * 100))
    m = (rng.random((14, 17, 21)) < p).astype(np.uint8)
    a = seg.binary_closing_safe_border(m, radius)
    b = bf.closing_itk_pipeline(m, radius)
    assert np.array_equal(a, b)
    assert (a >= m).all(), "closing is extensive"
    assert np.array_equal(seg.binary_closing_safe_border(a, radius), a), "closing is idempotent"


def test_safe_border_differs_from_naive_closing():
    # objects touching the border must NOT be eroded; scipy's plain binary_closing erodes them
    m = np.zeros((12, 12, 12), np.uint8)
    m[0:3, 4:8, 4:8] = 1
    ours = seg.binary_closing_safe_border(m, 2)
    naive = ndimage.binary_closing(m, structure=seg.ball_structure(2)).astype(np.uint8)
    assert (ours >= m).all()
    assert not (naive >= m).all()


@pytest.mark.parametrize("conn", [6, 26])
def test_ccl_two_restatements_agree(conn):
    rng = np.random.default_rng(conn)
    for p in (0.1, 0.3, 0.5):
        m = (rng.random((11, 13, 18)) < p).astype(np.uint8)
        a, ka = seg.connected_components(m, conn)
        b, kb = bf.flood_fill_labels(m, conn)
        assert ka == kb and np.array_equal(a, b)
        # ITK numbering: label k's minimum linear index is strictly increasing in k
        first = [np.flatnonzero(a.ravel() == k)[0] for k in range(1, ka + 1)]
        assert all(x < y for x, y in zip(first, first[1:]))
        assert ((a > 0) == (m > 0)).all(), "labels partition the mask"


def test_26_components_are_unions_of_6_components():
    rng = np.random.default_rng(5)
    m = (rng.random((10, 12, 14)) < 0.25).astype(np.uint8)
    l6, _ = seg.connected_components(m, 6)
    l26, _ = seg.connected_components(m, 26)
    for k in range(1, int(l6.max()) + 1):
        assert len(np.unique(l26[l6 == k])) == 1


def test_shape_statistics_match_direct_sums():
    ph = phantom.small_phantom(dims=(40, 30, 22), seed=3, flip_lps=True)
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    det = seg.detect_fiducials(vol, geom, min_vol=20, max_vol=600, full_stats=True)
    direct = bf.shape_stats_direct(det.labels, det.n_labels, ph.spacing, ph.origin, ph.direction)
    assert len(direct) == len(det.stats) == det.n_labels
    for s, d in zip(det.stats, direct):
        assert s.count == d["count"]
        assert np.isclose(s.physical_size, d["physical_size"], rtol=1e-14)
        assert np.allclose(s.centroid, d["centroid"], atol=1e-9)
        assert np.allclose(s.principal_moments, d["principal_moments"], rtol=1e-7, atol=1e-7)
    assert sum(s.count for s in det.stats) == int(det.closed.sum())


def test_threshold_casts():
    v = np.array([[[64, 65, 66, 65535]]], dtype=np.uint16)
    assert seg.binary_threshold(v, 65.0, 65535).ravel().tolist() == [0, 1, 1, 1]
    assert seg.binary_threshold(v, 65.9, 65535).ravel().tolist() == [0, 1, 1, 1]     # truncation toward zero
    i16 = np.array([[[-5, 64, 65, 32767]]], dtype=np.int16)
    assert seg.binary_threshold(i16, 65.0, 65535).ravel().tolist() == [0, 0, 1, 1]   # upper bound clamped
    f = np.array([[[np.nan, 64.9, 65.0, 1e9]]], dtype=np.float32)
    assert seg.binary_threshold(f, 65.0, 65535).ravel().tolist() == [0, 0, 1, 0]


def test_filter_and_body_selection_rules():
    counts = np.array([10, 100, 5000, 5000, 300])
    kept, body = seg.select_candidates(counts, 1.0, 50.0, 1500.0)
    assert kept == [2, 5]
    assert body == 3, "first maximum (lowest label) wins ties"
    kept, body = seg.select_candidates(np.array([50, 1500]), 1.0, 50.0, 1500.0)
    assert kept == [1, 2] and body == 0, "bounds are inclusive; nothing left for the body"


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.sampled_from([6, 26]))
def test_properties_random(seed, radius, conn):
    rng = np.random.default_rng(seed)
    dims = tuple(int(v) for v in rng.integers(3, 12, size=3))
    m = (rng.random(dims) < rng.uniform(0.05, 0.6)).astype(np.uint8)
    c = seg.binary_closing_safe_border(m, radius)
    assert (c >= m).all()
    assert np.array_equal(c, bf.closing_itk_pipeline(m, radius))
    lab, k = seg.connected_components(c, conn)
    counts = np.bincount(lab.ravel(), minlength=k + 1)[1:]
    assert counts.sum() == c.sum() and (counts > 0).all()
    if k:
        cnt, sums, moms = seg.integer_sums(lab, k)
        nz, ny, nx = dims
        assert (sums[:, 0] <= cnt * (nx - 1)).all() and (sums[:, 1] <= cnt * (ny - 1)).all()


@pytest.mark.parametrize("radius", [1, 2, 3])
@pytest.mark.parametrize("p", [0.5, 0.8, 0.97])
def test_opening_three_restatements_agree(radius, p):
    """sitk.BinaryMorphologicalOpening (mamri_params.open_radius): scipy-based, definitional and C restatements agree;
    opening is anti-extensive and idempotent; the outside of the image counts as foreground for its erosion, so a full
    volume stays full (plain scipy binary_opening with border_value=0 would eat its rim)."""
    import ctypes as C
    from oracle import c_oracle
    rng = np.random.default_rng(radius * 17 + int(p * 100))
    m = (rng.random((13, 18, 23)) < p).astype(np.uint8)
    a = seg.binary_opening(m, radius)
    assert np.array_equal(a, bf.opening_itk_pipeline(m, radius))
    c = np.empty_like(m)
    assert c_oracle.load().oracle_opening(m.ctypes.data, 23, 18, 13, radius, c.ctypes.data) == 0
    assert np.array_equal(a, c)
    assert (a <= m).all()
    assert np.array_equal(seg.binary_opening(a, radius), a)
    full = np.ones((9, 10, 11), dtype=np.uint8)
    assert seg.binary_opening(full, radius).all()
    assert not ndimage.binary_opening(full, structure=seg.ball_structure(radius)).all()


def test_detect_with_opening_matches_c_oracle():
    from oracle import c_oracle
    ph = phantom.small_phantom(dims=(48, 40, 32), seed=77, sigma=24.0, touch_border=True)
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    for orad in (1, 2):
        a = seg.detect_fiducials(vol, geom, open_radius=orad, min_vol=20.0, max_vol=600.0)
        b = c_oracle.detect_fiducials(vol, geom, open_radius=orad, min_vol=20.0, max_vol=600.0)
        assert np.array_equal(a.closed, b.closed) and np.array_equal(a.labels, b.labels) and a.fiducials == b.fiducials
        assert (a.closed <= seg.detect_fiducials(vol, geom, min_vol=20.0, max_vol=600.0).closed).all()   # both operators are increasing
