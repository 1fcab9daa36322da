"""GPU tests beyond the per-stage parity: BASELINE configs at full size against the C oracle, the
joint-angle criterion, the host-buffer (drop-in) call, the MamriLogic mirror, batching, entry search,
capacity errors and size-independent properties."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from mamri_pose_estimation_b200 import phantom
from oracle import c_oracle
from oracle import kinematics as kin
from oracle import segmentation as seg


def _geom(ph):
    return seg.Geometry(ph.spacing, ph.origin, ph.direction)


def _assert_equal_detection(res, ora, mask=None, labels=None):
    if mask is not None:
        assert np.array_equal(mask, ora.closed)
    if labels is not None:
        assert np.array_equal(labels, ora.labels)
    assert res.n_labels == ora.n_labels
    assert [m.label for m in res.markers] == [f["id"] for f in ora.fiducials]
    assert res.body_label == ora.body_label
    for m, f in zip(res.markers, ora.fiducials):
        assert m.volume_mm3 == f["vol"]
        assert np.abs(np.array(m.centroid_lps) - np.array(f["centroid"])).max() <= 1e-9


def test_c1_end_to_end_joint_angles(cuda_lib):
    """Config C1: 256x256x128, 9 fiducials.  Masks/labels bit-exact, centroids <= 1e-4 voxel, and the
    reference's matching + registration + IK chain gives the same joint angles (<= 1e-6 rad) from the CUDA
    markers as from the oracle markers."""
    from mamri_pose_estimation_b200.detector import FiducialDetector
    ph = phantom.config_c1()
    vol = phantom.generate(ph)
    geom = _geom(ph)
    ora = c_oracle.detect_fiducials(vol, geom)
    det = FiducialDetector(ph.dims)
    res = det.detect(torch.from_numpy(vol).cuda(), spacing=ph.spacing, origin=ph.origin, direction=ph.direction,
                     want_mask=True, want_labels=True, want_body=True)
    _assert_equal_detection(res, ora, res.mask.cpu().numpy(), res.labels.cpu().numpy().view(np.uint32))
    assert np.array_equal(res.body_mask.cpu().numpy(), ora.body_mask)
    assert len(res.markers) == 9
    inv = np.linalg.inv(geom.matrix())
    for m, f in zip(res.markers, ora.fiducials):
        assert np.abs(inv @ (np.array(m.centroid_lps) - np.array(f["centroid"]))).max() <= 1e-4
    ang_gpu, ident_gpu, _ = kin.pose_from_markers(res.ras_points)
    ang_ora, ident_ora, _ = kin.pose_from_markers(ora.ras_points)
    assert ang_gpu is not None and ang_ora is not None
    assert {k: [m["id"] for m in v] for k, v in ident_gpu.items()} == {k: [m["id"] for m in v] for k, v in ident_ora.items()}
    assert np.abs(ang_gpu - ang_ora).max() <= 1e-6
    # and the oracle chain recovers the phantom's pose (voxel quantisation limits it to ~1-2 degrees)
    assert np.abs(np.degrees(ang_ora) - np.array(phantom.ROBOT_POSE_DEG)).max() < 2.0
    det.close()


def test_c2_full_size_device_phantom(cuda_lib):
    """Config C2 at full size, phantom generated on the device: bit-exact against the C oracle."""
    from mamri_pose_estimation_b200.detector import FiducialDetector, generate_phantom_cuda
    ph = phantom.config_c2()
    d_vol = generate_phantom_cuda(ph)
    det = FiducialDetector(ph.dims)
    res = det.detect(d_vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction, want_mask=True, want_labels=True)
    host = d_vol.cpu().numpy()
    ora = c_oracle.detect_fiducials(host, _geom(ph), want_body_mask=False)
    _assert_equal_detection(res, ora, res.mask.cpu().numpy(), res.labels.cpu().numpy().view(np.uint32))
    assert len(res.markers) == 6
    det.close()


@pytest.mark.parametrize("conn", [6, 26])
def test_c4_merge_stress_reduced(cuda_lib, conn):
    """Config C4's ingredients (sigma 20 specks, 2000 blobs incl. tile-straddling ones and 26-only diagonal
    chains, 32 fiducials) at 512x512x256 so the C oracle finishes in seconds."""
    from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda
    ph = phantom.config_c4(dims=(512, 512, 256), n_blobs=2000)
    d_vol = generate_phantom_cuda(ph)
    det = FiducialDetector(ph.dims, max_runs=ph.n_voxels // 4, max_markers=16384)
    res = det.detect(d_vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction,
                     params=DetectParams(connectivity=conn), want_mask=True, want_labels=True)
    ora = c_oracle.detect_fiducials(d_vol.cpu().numpy(), _geom(ph), connectivity=conn, want_body_mask=False)
    _assert_equal_detection(res, ora, res.mask.cpu().numpy(), res.labels.cpu().numpy().view(np.uint32))
    assert np.array_equal(det.label_counts(res.n_labels).astype(np.int64), ora.counts)
    assert res.n_labels > 1000
    det.close()


def test_c4_full_size(cuda_lib):
    """Config C4 at its full 1024x1024x512 (32 fiducials, 2000 blobs, sigma 20: ~3.2 M runs, ~2.4 M labels):
    closed mask, label volume, per-label counts and the marker table bit-exact against the C oracle."""
    from mamri_pose_estimation_b200.detector import FiducialDetector, generate_phantom_cuda
    ph = phantom.config_c4()
    d_vol = generate_phantom_cuda(ph)
    det = FiducialDetector(ph.dims, max_runs=ph.n_voxels // 8, max_markers=16384)
    res = det.detect(d_vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction, want_mask=True, want_labels=True)
    host = d_vol.cpu().numpy()
    del d_vol
    ora = c_oracle.detect_fiducials(host, _geom(ph), want_body_mask=False)
    assert np.array_equal(res.mask.cpu().numpy(), ora.closed)
    assert np.array_equal(res.labels.cpu().numpy().view(np.uint32), ora.labels)
    _assert_equal_detection(res, ora)
    assert np.array_equal(det.label_counts(res.n_labels).astype(np.int64), ora.counts)
    assert res.n_labels > 100000 and len(res.markers) >= 32
    det.close()


def test_device_phantom_matches_numpy_twin(cuda_lib):
    from mamri_pose_estimation_b200.detector import generate_phantom_cuda
    ph = phantom.small_phantom(dims=(96, 64, 40), seed=8, sigma=15.0)
    host = phantom.generate(ph)
    dev = generate_phantom_cuda(ph).cpu().numpy()
    clean = dataclass_replace(ph, sigma=0.0)
    assert np.array_equal(generate_phantom_cuda(clean).cpu().numpy(), phantom.paint_signal(ph)), "noise-free signal is bit-identical"
    mism = (host != dev).mean()
    assert mism < 1e-3, f"noisy voxels differ on {mism:.2e} of the volume (libm vs CUDA rounding)"
    assert np.abs(host.astype(int) - dev.astype(int)).max() <= 1


def dataclass_replace(ph, **kw):
    import dataclasses
    return dataclasses.replace(ph, **kw)


def test_host_buffer_call_and_logic_mirror(cuda_lib):
    """The drop-in call: host volume in, markers + host body mask out, through MamriLogic's method names."""
    from mamri_pose_estimation_b200.logic import MamriLogic, MamriParameterNode, MarkupsFiducialNode, ScalarVolumeNode
    ph = phantom.small_phantom(dims=(96, 80, 48), n_fiducials=6, seed=12, spacing=(1.2, 1.2, 2.4))
    vol = phantom.generate(ph)
    ora = seg.detect_fiducials(vol, _geom(ph))
    logic = MamriLogic()
    pNode = MamriParameterNode(inputVolume=ScalarVolumeNode(vol, ph.spacing, ph.origin, ph.direction))
    logic.volume_threshold_segmentation(pNode)
    node = logic.scene.get("DetectedFiducials")
    if ora.fiducials:
        assert node.GetNumberOfControlPoints() == len(ora.fiducials)
        assert np.allclose(node.points(), ora.ras_points, atol=1e-9)
        assert [node.GetNthControlPointLabel(i) for i in range(len(ora.fiducials))] == ora.marker_labels
    else:
        assert node is None
    assert pNode.segmentationNode is not None and pNode.segmentationNode.body_label == ora.body_label
    assert np.array_equal(pNode.segmentationNode.body_mask, ora.body_mask)
    # a broken input volume is logged and ignored, like the reference's guarded pull (Mamri.py:1306-1307)
    logic.volume_threshold_segmentation(MamriParameterNode(inputVolume=None))
    # entry point through the mirror
    pts, nrm, tgt = phantom.surface_candidates(20000, seed=3)
    pNode.segmentationNode.surface_points, pNode.segmentationNode.surface_normals = pts, nrm
    t = MarkupsFiducialNode("target")
    t.AddControlPoint(tgt)
    pNode.targetFiducialNode = t
    logic.findAndSetEntryPoint(pNode)
    wi, wd = kin.find_entry_point(pts, nrm, tgt)
    assert wi >= 0
    assert np.allclose(pNode.entryPointFiducialNode.GetNthControlPointPositionWorld(0), pts[wi].astype(np.float64))


def test_entry_search_parity_and_path_sampling(cuda_lib):
    from mamri_pose_estimation_b200.detector import FiducialDetector
    det = FiducialDetector((64, 64, 64))
    for n, seed in ((1, 1), (1000, 2), (262144, 3)):
        pts, nrm, tgt = phantom.surface_candidates(n, seed=seed)
        r = det.entry_search(torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda(), tgt)
        wi, wd = kin.find_entry_point(pts, nrm, tgt)
        assert r["index"] == wi
        if wi >= 0:
            assert r["distance"] == wd, "float64 distance is bit-identical"
            assert np.array_equal(r["point"], pts[wi].astype(np.float64))
    # nothing suitable: all normals along y
    pts, nrm, tgt = phantom.surface_candidates(512, seed=4)
    bad = np.zeros_like(nrm); bad[:, 1] = 1.0
    r = det.entry_search(torch.from_numpy(pts).cuda(), torch.from_numpy(bad).cuda(), tgt)
    assert r["index"] == -1 and r["n_suitable"] == 0
    # needle-path sampling against a voxel mask (extension; numpy restatement of the same rule)
    pts, nrm, tgt = phantom.surface_candidates(4096, seed=5)
    dims = (64, 64, 64)
    mask = np.ones(dims[::-1], dtype=np.uint8)
    mask[20:44, 10:30, 30:64] = 0                               # an obstacle the needle may not cross
    m = np.array([[0.2, 0, 0, 32], [0, 0.2, 0, 32], [0, 0, 0.2, 32]], dtype=np.float64)   # RAS mm -> index
    S = 16
    r = det.entry_search(torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda(), tgt, n_path_samples=S,
                         path_mask=torch.from_numpy(mask).cuda(), ras_to_index=m, path_free_value=1)
    p64 = pts.astype(np.float64)
    d = p64 - tgt
    d2 = d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2
    ok = (d2 <= 80.0 ** 2) & ((np.abs(nrm[:, 0].astype(np.float64)) - 2 * np.abs(nrm[:, 1].astype(np.float64))) > -0.5)
    for k in range(1, S + 1):
        q = p64 + (k / (S + 1)) * (tgt - p64)
        idx = np.rint(q @ m[:, :3].T + m[:, 3]).astype(np.int64)
        inside = ((idx >= 0) & (idx < np.array(dims))).all(axis=1)
        val = np.zeros(len(pts), dtype=np.int64)
        ii = idx[inside]
        val[inside] = mask[ii[:, 2], ii[:, 1], ii[:, 0]]
        ok &= val == 1
    want = int(np.flatnonzero(ok)[np.argmin(np.sqrt(d2)[ok])]) if ok.any() else -1
    assert r["index"] == want and r["n_suitable"] == int(ok.sum())
    det.close()


def test_batch_detector_matches_single(cuda_lib):
    from mamri_pose_estimation_b200.detector import BatchDetector, FiducialDetector, generate_phantom_cuda
    specs = [phantom.small_phantom(dims=(128, 96, 64), seed=40, scan_index=i) for i in range(7)]
    vols = [generate_phantom_cuda(p) for p in specs]
    bd = BatchDetector(specs[0].dims, n_contexts=3)
    batch = bd.run(vols, specs[0].spacing, specs[0].origin, specs[0].direction)
    hosts = [v.cpu().pin_memory() for v in vols]
    bodies = [torch.empty(v.shape, dtype=torch.uint8).pin_memory() for v in vols]
    batch_h = bd.run_host(hosts, specs[0].spacing, specs[0].origin, specs[0].direction, body_out=bodies)
    one = FiducialDetector(specs[0].dims)
    for v, rb, rh, body in zip(vols, batch, batch_h, bodies):
        r1 = one.detect(v, spacing=specs[0].spacing, origin=specs[0].origin, direction=specs[0].direction, want_body=True)
        for r in (rb, rh):
            assert r.n_labels == r1.n_labels and r.body_label == r1.body_label
            assert [(m.label, m.count, m.sum_idx, m.sum_mom) for m in r.markers] == \
                   [(m.label, m.count, m.sum_idx, m.sum_mom) for m in r1.markers]
        assert np.array_equal(body.numpy(), r1.body_mask.cpu().numpy())
    # the pool's ring still holds the mask / label volumes of the last n_contexts scans (7 = waves of 3 + 3 + 1)
    r_last = one.detect(vols[6], spacing=specs[0].spacing, origin=specs[0].origin, direction=specs[0].direction,
                        want_mask=True, want_labels=True)
    assert torch.equal(batch[6].mask, r_last.mask) and torch.equal(batch[6].labels, r_last.labels)
    # the vectorised table (what the NCCL gather ships) equals the per-object packing
    from mamri_pose_estimation_b200 import distributed as mdist
    assert np.array_equal(batch.table(mdist.TABLE_SLOTS), mdist.pack_table(list(batch)))
    # the per-scan pipelined path (no wave graph) gives the same tables
    import os
    os.environ["MAMRI_NO_WAVE_GRAPH"] = "1"
    try:
        bd2 = BatchDetector(specs[0].dims, n_contexts=2)
        assert np.array_equal(bd2.run(vols, specs[0].spacing, specs[0].origin, specs[0].direction).table(), batch.table())
        bd2.close()
    finally:
        del os.environ["MAMRI_NO_WAVE_GRAPH"]
    bd.close(); one.close()


def test_capacity_errors_are_reported_not_fatal(cuda_lib):
    from mamri_pose_estimation_b200 import _capi
    from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector
    rng = np.random.default_rng(0)
    vol = torch.from_numpy((rng.random((16, 64, 64)) < 0.3).astype(np.uint8) * 200).cuda()
    small = FiducialDetector((64, 64, 16), max_runs=1 << 20, max_markers=4)
    with pytest.raises(_capi.MamriError) as e:
        small.detect(vol, params=DetectParams(lower=1, upper=255, close_radius=0, min_volume=1, max_volume=5))
    assert e.value.code == _capi.MAMRI_ERR_CAPACITY
    # the context stays usable
    r = small.detect(vol, params=DetectParams(lower=1, upper=255, close_radius=0, min_volume=1e9, max_volume=2e9))
    assert r.n_labels > 0 and not r.markers
    with pytest.raises(_capi.MamriError):
        small.detect(torch.zeros((17, 64, 64), dtype=torch.uint8, device="cuda"))   # larger than the context
    with pytest.raises(_capi.MamriError):
        small.detect(vol, params=DetectParams(connectivity=18))
    small.close()


def test_size_independent_properties(cuda_lib):
    """Closing is extensive and idempotent; 26-components are unions of 6-components; counts sum to the mask."""
    from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda
    ph = phantom.config_c4(dims=(256, 256, 128), n_blobs=400)
    d_vol = generate_phantom_cuda(ph)
    det = FiducialDetector(ph.dims, max_runs=ph.n_voxels // 4, max_markers=16384)
    r6 = det.detect(d_vol, spacing=ph.spacing, params=DetectParams(connectivity=6), want_mask=True, want_labels=True)
    raw = det.detect(d_vol, spacing=ph.spacing, params=DetectParams(close_radius=0), want_mask=True).mask
    assert bool((r6.mask >= raw).all()), "closing is extensive"
    again = det.detect(r6.mask * 200, spacing=ph.spacing, params=DetectParams(lower=1, upper=255), want_mask=True).mask
    assert torch.equal(again, r6.mask), "closing is idempotent"
    l6 = r6.labels.clone()
    r26 = det.detect(d_vol, spacing=ph.spacing, params=DetectParams(connectivity=26), want_labels=True)
    assert r26.n_labels <= r6.n_labels
    pair = torch.unique(torch.stack([l6.flatten().long(), r26.labels.flatten().long()]), dim=1)
    assert pair.shape[1] == r6.n_labels + 1, "every 6-component lies in exactly one 26-component"
    assert int(det.label_counts(r26.n_labels).astype(np.int64).sum()) == r26.n_foreground == int(r6.mask.sum())
    det.close()


def test_begin_end_and_device_tables(cuda_lib):
    """mamri_pool_detect_begin/_end: same results as the one-call form, and the fixed-size marker tables the scans'
    last kernels write on the device equal the host-packed ones bit for bit (what the NCCL gather ships)."""
    from mamri_pose_estimation_b200.detector import BatchDetector
    from mamri_pose_estimation_b200.distributed import TABLE_SLOTS, pack_table
    from mamri_pose_estimation_b200 import _capi
    specs = [phantom.small_phantom(dims=(96, 80, 48), n_fiducials=5 + i, n_blobs=3, seed=60 + i, spacing=(1.2, 1.2, 2.4)) for i in range(3)]
    vols = [torch.from_numpy(phantom.generate(p)).cuda() for p in specs]
    sp, org, dr = specs[0].spacing, specs[0].origin, specs[0].direction
    for env_graph in (True, False):
        bd = BatchDetector(specs[0].dims, n_contexts=4)
        if not env_graph:
            bd.context(0).set_profiling(True)               # forces the per-context (no wave graph) path
        want = bd.run(vols, sp, org, dr) if env_graph else None
        tables = torch.full((3, TABLE_SLOTS, 8), -1.0, dtype=torch.float64, device="cuda")
        bd.begin(vols, sp, org, dr, tables=tables)
        with pytest.raises(_capi.MamriError):
            bd.begin(vols, sp, org, dr)                     # one batch at a time
        got = bd.end()
        assert np.array_equal(tables.cpu().numpy(), pack_table(got))
        for i, (ph, v) in enumerate(zip(specs, vols)):
            ora = seg.detect_fiducials(v.cpu().numpy(), _geom(ph))
            _assert_equal_detection(got[i], ora)
            if want is not None:
                assert [m.label for m in got[i].markers] == [m.label for m in want[i].markers]
        with pytest.raises(RuntimeError):
            bd.end()
        with pytest.raises(ValueError):
            bd.begin(vols * 2, sp, org, dr)                 # more scans than contexts
        bd.close()


def test_batch_pipeline_matches_single_batches(cuda_lib):
    """BatchPipeline: batches enqueued ahead of the collection of the previous one give the results of plain runs."""
    from mamri_pose_estimation_b200.detector import BatchPipeline
    specs = [[phantom.small_phantom(dims=(96, 80, 48), n_fiducials=4 + i, n_blobs=2, seed=80 + 10 * b + i, spacing=(1.2, 1.2, 2.4))
              for i in range(3)] for b in range(4)]
    vols = [[torch.from_numpy(phantom.generate(p)).cuda() for p in batch] for batch in specs]
    sp, org, dr = specs[0][0].spacing, specs[0][0].origin, specs[0][0].direction
    bp = BatchPipeline(specs[0][0].dims, n_contexts=3, depth=2)
    got = []
    bp.submit(vols[0], sp, org, dr)
    for b in range(4):
        if b + 1 < 4:
            bp.submit(vols[b + 1], sp, org, dr)
        got.append(bp.result())
    with pytest.raises(RuntimeError):
        bp.result()
    for b in range(4):
        for i in range(3):
            ora = seg.detect_fiducials(vols[b][i].cpu().numpy(), _geom(specs[b][i]))
            _assert_equal_detection(got[b][i], ora)
    bp.submit(vols[0], sp, org, dr)
    bp.submit(vols[1], sp, org, dr)
    with pytest.raises(RuntimeError):
        bp.submit(vols[2], sp, org, dr)                    # both pools busy
    bp.result(); bp.result()
    bp.close()


def test_c3_batch_full_size(cuda_lib):
    """Config C3 at full size (512x512x256, Rician sigma 15 -> thousands of noise voxels above the threshold before
    closing), 12 of the 64 scans through the pipelined pools: every scan against the C oracle (labels, markers, body),
    mask + label volumes bit-exact for the scans whose ring buffers are still live, and the pose stage on the batch."""
    from mamri_pose_estimation_b200.detector import BatchPipeline, generate_phantom_cuda
    specs = [phantom.config_c3(scan_index=i) for i in range(12)]
    vols = [generate_phantom_cuda(p) for p in specs]
    sp, org, dr = specs[0].spacing, specs[0].origin, specs[0].direction
    bp = BatchPipeline(specs[0].dims, n_contexts=4, depth=2)
    got = []
    bp.submit(vols[0:4], sp, org, dr)
    for b in range(3):
        if b + 1 < 3:
            bp.submit(vols[4 * (b + 1):4 * (b + 2)], sp, org, dr)
        res = bp.result()
        if b == 2:                                            # its ring buffers are not overwritten any more
            last_masks = [res[i].mask.cpu().numpy() for i in range(4)]
            last_labels = [res[i].labels.cpu().numpy().view(np.uint32) for i in range(4)]
        got.extend(res[i] for i in range(4))
    for i, (ph, v) in enumerate(zip(specs, vols)):
        ora = c_oracle.detect_fiducials(v.cpu().numpy(), _geom(ph), want_body_mask=False)
        if i >= 8:
            _assert_equal_detection(got[i], ora, last_masks[i - 8], last_labels[i - 8])
        else:
            _assert_equal_detection(got[i], ora)
        assert len(got[i].markers) == 6 and got[i].n_labels == ora.n_labels
    poses = bp.pools[0].context(0).pose_estimate([g.ras_points for g in got])
    for pose, g in zip(poses, got):
        ident = kin.joint_detection(g.ras_points)
        assert pose.identified == {jn: [m["id"] for m in ms] for jn, ms in ident.items()}
    bp.close()


def test_host_pipeline_matches_run_host(cuda_lib):
    """BatchPipeline.submit_host / result (mamri_pool_detect_host_begin / _end): host volumes in, host body masks +
    markers out, batches enqueued ahead of the previous batch's collection."""
    from mamri_pose_estimation_b200.detector import BatchPipeline
    specs = [[phantom.small_phantom(dims=(96, 80, 48), n_fiducials=4 + i, n_blobs=2, seed=120 + 10 * b + i, spacing=(1.2, 1.2, 2.4))
              for i in range(3)] for b in range(3)]
    vols = [[phantom.generate(p) for p in batch] for batch in specs]
    bodies = [[np.zeros(v.shape, np.uint8) for v in batch] for batch in vols]
    sp, org, dr = specs[0][0].spacing, specs[0][0].origin, specs[0][0].direction
    bp = BatchPipeline(specs[0][0].dims, n_contexts=3, depth=2)
    got = []
    bp.submit_host(vols[0], sp, org, dr, body_out=bodies[0])
    for b in range(3):
        if b + 1 < 3:
            bp.submit_host(vols[b + 1], sp, org, dr, body_out=bodies[b + 1])
        got.append(bp.result())
    torch.cuda.synchronize()
    for b in range(3):
        for i in range(3):
            ora = seg.detect_fiducials(vols[b][i], _geom(specs[b][i]))
            _assert_equal_detection(got[b][i], ora)
            assert np.array_equal(bodies[b][i], ora.body_mask)
            assert got[b][i].body_mask is bodies[b][i]
    bp.submit_host(vols[0], sp, org, dr)                      # no body masks asked for
    r = bp.result()
    assert r[0].body_mask is None and r[0].mask is None
    bp.close()


@pytest.mark.gpu
def test_bit_packed_body_mask_equals_uint8_body_mask(cuda_lib):
    """The *_bits calls return the body labelmap at 1 bit per voxel (8x fewer bytes over PCIe): unpacked, it is the
    uint8 labelmap of the plain calls and of the oracle -- single context, device output, pool and pipeline; ragged
    rows (nx % 32 != 0) included; an empty scan gives an all-zero mask."""
    import ctypes as C
    import torch
    from mamri_pose_estimation_b200.detector import (BatchPipeline, DetectParams, FiducialDetector, body_bits_shape,
                                                     unpack_body_bits, _desc)
    for dims, seed in (((96, 80, 48), 12), ((70, 45, 33), 13)):
        ph = phantom.small_phantom(dims=dims, n_fiducials=5, seed=seed)
        vol = phantom.generate(ph)
        ora = seg.detect_fiducials(vol, _geom(ph))
        nx, ny, nz = dims
        det = FiducialDetector(dims)
        det.reserve_staging(vol.nbytes, int(np.prod(body_bits_shape(vol.shape))) * 4)
        bits = np.full(body_bits_shape(vol.shape), -1, dtype=np.int32)
        res = det.detect_host(vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction, body_bits_out=bits)
        assert res.body_label == ora.body_label
        assert np.array_equal(unpack_body_bits(bits, vol.shape), ora.body_mask)
        # device-resident form through the C ABI
        d_bits = torch.full(body_bits_shape(vol.shape), -1, dtype=torch.int32, device="cuda")
        d = _desc(vol.shape, "uint16", ph.spacing, ph.origin, ph.direction)
        p = DetectParams().to_c()
        dv = torch.from_numpy(vol).cuda()
        rc = cuda_lib.mamri_detect_bits_async(det._ctx, C.byref(d), dv.data_ptr(), C.byref(p), None, None, d_bits.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        det._pending = (None, None, None, None)
        det.collect()
        assert np.array_equal(unpack_body_bits(d_bits.cpu(), vol.shape), ora.body_mask)
        # empty scan
        bits[:] = -1
        det.detect_host(np.zeros_like(vol), spacing=ph.spacing, origin=ph.origin, direction=ph.direction, body_bits_out=bits)
        assert not bits.any()
        det.close()
    # pool / pipeline
    ph = phantom.small_phantom(dims=(96, 80, 48), n_fiducials=5, seed=14)
    vols = [phantom.generate(phantom.small_phantom(dims=(96, 80, 48), n_fiducials=5, seed=14 + i)) for i in range(3)]
    bp = BatchPipeline((96, 80, 48), n_contexts=4, depth=2)
    hv = [torch.from_numpy(v).pin_memory() for v in vols]
    hb = [torch.empty(body_bits_shape(vols[0].shape), dtype=torch.int32).pin_memory() for _ in vols]
    bp.submit_host(hv, ph.spacing, ph.origin, ph.direction, body_bits_out=hb)
    r = bp.result()
    for i, v in enumerate(vols):
        ora = seg.detect_fiducials(v, _geom(ph))
        assert r[i].body_label == ora.body_label
        assert np.array_equal(unpack_body_bits(hb[i], v.shape), ora.body_mask if ora.body_mask is not None else np.zeros_like(v, dtype=np.uint8))
    bp.close()


@pytest.mark.gpu
def test_entry_search_at_config_c5_size(cuda_lib):
    """BASELINE config C5 at full size on one GPU: 1,048,576 skin-surface candidates; the winner (index and float64
    distance) is bit-identical to the reference loop restated in NumPy (Mamri.py:1008-1023), with and without sharding
    the candidates into blocks and reducing the block winners the way distributed.gather_entry_results does."""
    import torch
    from mamri_pose_estimation_b200.detector import FiducialDetector
    n = 1 << 20
    pts, nrm, tgt = phantom.surface_candidates(n)
    wi, wd = kin.find_entry_point(pts, nrm, tgt)
    det = FiducialDetector((64, 64, 64))
    p, q = torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda()
    r = det.entry_search(p, q, tgt)
    assert r["index"] == wi and r["distance"] == wd and wi >= 0
    best = (float("inf"), -1)
    for lo in range(0, n, n // 8):                            # 8 blocks, as 8 ranks would hold them
        b = det.entry_search(p[lo:lo + n // 8].contiguous(), q[lo:lo + n // 8].contiguous(), tgt)
        if b["index"] >= 0 and (b["distance"], b["index"] + lo) < best:
            best = (b["distance"], b["index"] + lo)
    assert best == (wd, wi)
    # argument errors are reported, not guessed around
    with pytest.raises(ValueError):
        det.entry_search(p, q, tgt, n_path_samples=8)
    det.close()


@pytest.mark.gpu
def test_one_process_two_devices(cuda_lib):
    """One process driving two GPUs (mamri_create takes any device): the opt-in to more than 48 KB of dynamic shared
    memory is per device, so a context created on the second device must run the tile kernels too."""
    import torch
    from mamri_pose_estimation_b200.detector import FiducialDetector
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    ph = phantom.config_c2()
    vol = phantom.generate(phantom.small_phantom(dims=(512, 64, 48), seed=5))        # rows of 512 voxels: > 48 KB tiles
    sm = phantom.small_phantom(dims=(512, 64, 48), seed=5)
    ora = seg.detect_fiducials(vol, _geom(sm))
    for dev in (0, 1):
        det = FiducialDetector((512, 64, 48), device=dev)
        res = det.detect(torch.from_numpy(vol).to(f"cuda:{dev}"), spacing=sm.spacing, origin=sm.origin, direction=sm.direction,
                         want_mask=True, want_labels=True)
        assert np.array_equal(res.mask.cpu().numpy(), ora.closed) and np.array_equal(res.labels.cpu().numpy().view(np.uint32), ora.labels)
        det.close()
