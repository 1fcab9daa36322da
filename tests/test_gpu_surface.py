"""GPU tests of the skin-surface candidate stage (mamri_body_surface, surface.cu) against oracle/surface.py,
and of the self-contained chain segmentation -> candidates -> closest suitable entry point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from mamri_pose_estimation_b200 import phantom
from oracle import kinematics as kin
from oracle import segmentation as seg
from oracle import surface as srf


def _check(det, body, geom):
    pts, nrm = det.body_surface(torch.from_numpy(body).cuda(), spacing=geom.spacing, origin=geom.origin,
                                direction=geom.direction)
    o_pts, o_nrm, _ = srf.body_surface(body, geom)
    assert tuple(pts.shape) == o_pts.shape
    assert np.array_equal(pts.cpu().numpy(), o_pts), "candidate points differ (bit-exact float32 expected)"
    assert np.array_equal(nrm.cpu().numpy(), o_nrm), "candidate normals differ (bit-exact float32 expected)"
    assert det.last_body_voxels == int((body != 0).sum())
    return pts, nrm


@pytest.mark.parametrize("dims,p,seed", [((37, 29, 23), 0.6, 1), ((64, 33, 17), 0.9, 2), ((96, 40, 12), 0.3, 3),
                                         ((31, 5, 3), 0.5, 4), ((160, 16, 9), 0.97, 5), ((1, 1, 1), 1.0, 6)])
def test_random_bodies_bit_exact(cuda_lib, dims, p, seed):
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    body = (rng.random((nz, ny, nx)) < p).astype(np.uint8) * np.uint8(rng.integers(1, 255))
    geom = seg.Geometry((0.7, 1.3, 2.1), (-12.5, 40.25, 7.0), (-1, 0, 0, 0, -1, 0, 0, 0, 1) if seed % 2 else (1, 0, 0, 0, 1, 0, 0, 0, 1))
    det = FiducialDetector((max(nx, 32), max(ny, 8), max(nz, 8)))
    _check(det, body, geom)
    det.close()


def test_unaligned_body_pointer_and_empty_body(cuda_lib):
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(9)
    det = FiducialDetector((64, 32, 16))
    geom = seg.Geometry((1, 1, 1), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    body = (rng.random((16, 32, 64)) < 0.8).astype(np.uint8)
    flat = torch.zeros(body.size + 3, dtype=torch.uint8, device="cuda")
    view = flat[3:].view(16, 32, 64)                       # nx % 32 == 0 but not 16-byte aligned -> generic packer
    view.copy_(torch.from_numpy(body))
    pts, nrm = det.body_surface(view, spacing=geom.spacing, origin=geom.origin, direction=geom.direction)
    o_pts, o_nrm, _ = srf.body_surface(body, geom)
    assert np.array_equal(pts.cpu().numpy(), o_pts) and np.array_equal(nrm.cpu().numpy(), o_nrm)
    pts, nrm = det.body_surface(torch.zeros((16, 32, 64), dtype=torch.uint8, device="cuda"))
    assert pts.shape == (0, 3) and nrm.shape == (0, 3)
    det.close()


def test_capacity_is_reported_not_fatal(cuda_lib):
    import ctypes as C
    from mamri_pose_estimation_b200 import _capi
    from mamri_pose_estimation_b200.detector import FiducialDetector, _desc
    det = FiducialDetector((64, 32, 16))
    body = torch.ones((16, 32, 64), dtype=torch.uint8, device="cuda")
    d = _desc((16, 32, 64), "uint8", (1, 1, 1), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    pts = torch.full((10, 3), -1.0, dtype=torch.float32, device="cuda")
    nrm = torch.full((10, 3), -1.0, dtype=torch.float32, device="cuda")
    n = C.c_int64(0)
    rc = cuda_lib.mamri_body_surface(det._ctx, C.byref(d), body.data_ptr(), pts.data_ptr(), nrm.data_ptr(), 10, C.byref(n),
                                     None, None)
    assert rc == _capi.MAMRI_ERR_CAPACITY and n.value == 16 * 32 * 64 - 14 * 30 * 62
    o_pts, _, _ = srf.body_surface(np.ones((16, 32, 64), np.uint8), seg.Geometry((1, 1, 1), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1)))
    assert np.array_equal(pts.cpu().numpy(), o_pts[:10])   # the first `capacity` candidates are still written
    rc = cuda_lib.mamri_body_surface(det._ctx, C.byref(d), body.data_ptr(), None, None, 5, C.byref(n), None, None)
    assert rc == _capi.MAMRI_ERR_INVALID_ARG
    det.close()


def test_body_of_last_scan_equals_body_mask_path(cuda_lib):
    """d_body_mask == NULL: candidates straight from the last scan's run table, no per-voxel pass."""
    from mamri_pose_estimation_b200.detector import FiducialDetector
    ph = phantom.small_phantom(dims=(96, 80, 48), n_fiducials=6, n_blobs=3, seed=21, spacing=(1.2, 1.2, 2.4))
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    ora = seg.detect_fiducials(vol, geom)
    det = FiducialDetector(ph.dims)
    res = det.detect(torch.from_numpy(vol).cuda(), spacing=ph.spacing, origin=ph.origin, direction=ph.direction, want_body=True)
    assert np.array_equal(res.body_mask.cpu().numpy(), ora.body_mask)
    p1, n1 = det.body_surface(None, shape_zyx=vol.shape, spacing=ph.spacing, origin=ph.origin, direction=ph.direction)
    p2, n2 = _check(det, ora.body_mask, geom)
    assert torch.equal(p1, p2) and torch.equal(n1, n2)
    det.close()


def test_segmentation_to_entry_point_on_device(cuda_lib):
    """MamriLogic mirror with no surface supplied: body labelmap -> GPU candidates -> closest suitable entry,
    equal to the oracle's entry loop (Mamri.py:1011-1023) over the oracle's candidates."""
    from mamri_pose_estimation_b200.logic import MamriLogic, MamriParameterNode, MarkupsFiducialNode, ScalarVolumeNode
    ph = phantom.small_phantom(dims=(128, 112, 64), n_fiducials=6, seed=33, spacing=(1.6, 1.6, 3.2))
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    ora = seg.detect_fiducials(vol, geom)
    o_pts, o_nrm, _ = srf.body_surface(ora.body_mask, geom)
    for dev_input in (False, True):
        logic = MamriLogic()
        arr = torch.from_numpy(vol).cuda() if dev_input else vol
        pNode = MamriParameterNode(inputVolume=ScalarVolumeNode(arr, ph.spacing, ph.origin, ph.direction))
        logic.volume_threshold_segmentation(pNode)
        centre_ras = o_pts.astype(np.float64).mean(axis=0)
        tgt = centre_ras + np.array([25.0, 4.0, -6.0])
        t = MarkupsFiducialNode("target")
        t.AddControlPoint(tgt)
        pNode.targetFiducialNode = t
        logic.findAndSetEntryPoint(pNode)
        wi, wd = kin.find_entry_point(o_pts, o_nrm, tgt)
        assert wi >= 0
        got = np.array(pNode.entryPointFiducialNode.GetNthControlPointPositionWorld(0))
        assert np.array_equal(got, o_pts[wi].astype(np.float64))
        assert len(pNode.segmentationNode.surface_points) == len(o_pts)


def test_c2_body_surface_full_size(cuda_lib):
    """Config C2's body ellipsoid at full size (512x512x256): candidates of the last scan equal the oracle's."""
    from mamri_pose_estimation_b200.detector import FiducialDetector, generate_phantom_cuda
    ph = phantom.config_c2()
    vol = generate_phantom_cuda(ph)
    det = FiducialDetector(ph.dims)
    res = det.detect(vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction, want_body=True)
    pts, nrm = det.body_surface(None, shape_zyx=tuple(vol.shape), spacing=ph.spacing, origin=ph.origin, direction=ph.direction)
    body = res.body_mask.cpu().numpy()
    o_pts, o_nrm, _ = srf.body_surface(body, seg.Geometry(ph.spacing, ph.origin, ph.direction))
    assert len(o_pts) > 100000
    assert np.array_equal(pts.cpu().numpy(), o_pts) and np.array_equal(nrm.cpu().numpy(), o_nrm)
    p2, n2 = det.body_surface(res.body_mask, spacing=ph.spacing, origin=ph.origin, direction=ph.direction)
    assert torch.equal(pts, p2) and torch.equal(nrm, n2)
    det.close()
