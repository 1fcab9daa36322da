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


@pytest.mark.parametrize("conn", [6, 26])
@pytest.mark.parametrize("threads", [2, 3, 8])
def test_slab_parallel_labelling_equals_the_sequential_scan(conn, threads):
    """oracle_ccl threads the scan over slabs of slices and merges them along their faces (as ITK does); the labels must
    be the sequential scan's, bit for bit, for any number of slabs -- noisy masks, objects crossing every slab face."""
    import ctypes as C
    lib = c_oracle.load()
    lib.oracle_ccl_serial.restype = C.c_int
    lib.oracle_ccl_serial.argtypes = lib.oracle_ccl.argtypes
    before = c_oracle.num_threads()
    try:
        lib.oracle_set_num_threads(threads)
        rng = np.random.default_rng(100 * conn + threads)
        for dims, p in (((41, 37, 45), 0.32), ((64, 48, 40), 0.18), ((33, 20, 64), 0.5), ((130, 9, 57), 0.27)):
            nz, ny, nx = dims[2], dims[1], dims[0]
            m = (rng.random((nz, ny, nx)) < p).astype(np.uint8)
            m[:, ny // 2, nx // 3] = 1                                   # a column through every slab face
            a, b = np.empty(m.shape, np.uint32), np.empty(m.shape, np.uint32)
            ka, kb = C.c_uint32(0), C.c_uint32(0)
            assert lib.oracle_ccl(m.ctypes.data, nx, ny, nz, conn, a.ctypes.data, C.byref(ka)) == 0
            assert lib.oracle_ccl_serial(m.ctypes.data, nx, ny, nz, conn, b.ctypes.data, C.byref(kb)) == 0
            assert ka.value == kb.value and np.array_equal(a, b)
            ref, k_ref = seg.connected_components(m, conn)
            assert k_ref == ka.value and np.array_equal(a, ref)
            # threaded sums against a direct NumPy evaluation
            sums = np.zeros((ka.value, 10), np.uint64)
            assert lib.oracle_label_sums(a.ctypes.data, nx, ny, nz, ka.value, sums.ctypes.data) == 0
            cnt, sidx, smom = seg.integer_sums(a, ka.value)
            want = np.concatenate([cnt[:, None], sidx, smom], axis=1).astype(np.uint64)
            assert np.array_equal(sums, want)
    finally:
        lib.oracle_set_num_threads(before)
