"""GPU tests of the batched marker-table consumers (mamri_pose_estimate, pose.cu) against oracle/kinematics.py:
identical L-shape assignments (joint_detection + _sort_l_shaped_markers, Mamri.py:1343-1363, 1782-1792), the
baseplate registration (vtkLandmarkTransform restatement) to 1e-9, and joint angles within 1e-5 rad of the
reference's SciPy TRF solve -- the device solver is the same algorithm restated (Trust Region Reflective, 2-point
Jacobian, ftol = xtol = 1e-6), so it must end in SciPy's minimum, not merely in some minimum."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from mamri_pose_estimation_b200 import robot as rb
from oracle import kinematics as kin

ANGLE_TOL = 1e-5      # rad, vs SciPy TRF stopped at 1e-6
MATRIX_TOL = 1e-9


def _base(rng):
    """Baseplate lying on the scanner table (local z anterior, as phantom.robot_base_matrix: the reference forces the
    three baseplate markers to one RAS y, Mamri.py:1371-1373), turned about the table normal and shifted."""
    a = rng.uniform(-0.4, 0.4)
    turn = np.eye(4)
    turn[:3, :3] = [[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]]
    turn[:3, 3] = rng.uniform(-40, 40, 3)
    return turn @ rb._rot("X", -90.0)


def _scene(rng, links=("Baseplate", "Joint4", "Joint6"), noise=0.15, extra=2, shuffle=True, pose_deg=35.0):
    """Marker centroids of a posed robot (+ noise, + spurious points); baseplate first so that the reference's
    first-match rule recovers the true assignment."""
    theta = np.radians(rng.uniform(-pose_deg, pose_deg, 6))
    base = _base(rng)
    pos = rb.marker_positions_ras(theta, base, links)
    pts = [pos[l] + rng.normal(0, noise, (3, 3)) for l in links]
    pts = np.concatenate(pts)
    if shuffle:                                   # within-link order is arbitrary in a scan (label order)
        for i in range(len(links)):
            pts[3 * i:3 * i + 3] = pts[3 * i:3 * i + 3][rng.permutation(3)]
    if extra:
        pts = np.concatenate([pts, rng.uniform(-300, 300, (extra, 3)) + np.array([0, 0, 900.0])])
    return pts, theta, base


same_basin = []        # per IK solved in this module: did the device solver and SciPy end in the same minimum?


def _compare(pose, pts, check_angles=True):
    ang, ident, base = kin.pose_from_markers(pts)
    assert pose.identified == {jn: [m["id"] for m in ms] for jn, ms in ident.items()}
    if base is None:
        assert pose.base_matrix is None
    else:
        assert np.abs(pose.base_matrix - base).max() < MATRIX_TOL
    if ang is None:
        assert pose.joint_angles is None
    elif check_angles:
        # The device solver is the reference's algorithm restated, so (1) it reports success like SciPy does, with a
        # cost and an rms error that are the reference's residual function at its angles, inside the joint limits,
        # (2) its cost is SciPy's cost (both stop at ftol = 1e-6, i.e. short of the exact minimum by design), and
        # (3) the angles are SciPy's.
        assert pose.ik_converged and pose.ik_termination in (1, 2, 3, 4)
        j6 = [pts[i] for i in pose.identified["Joint6"]]
        j4 = [pts[i] for i in pose.identified["Joint4"]] if "Joint4" in pose.identified else None
        f = lambda x: np.array(kin.ik_error(x, j6, base, False, j4))
        err = f(pose.joint_angles)
        assert abs(0.5 * float(err @ err) - pose.ik_cost) < 1e-9 * max(1.0, pose.ik_cost)
        assert abs(math.sqrt(float(np.mean(err[:9] ** 2))) - pose.ik_rms_error) < 1e-9
        lim = np.radians([kin.ROBOT_BY_NAME[n]["joint_limits"] for n in kin.ARTICULATED_CHAIN])
        assert np.all(pose.joint_angles >= lim[:, 0]) and np.all(pose.joint_angles <= lim[:, 1])
        ref_err = f(ang)
        assert abs(pose.ik_cost - 0.5 * float(ref_err @ ref_err)) <= 1e-5 * max(1.0, pose.ik_cost)
        # Same algorithm, same iterates: where the markers fit the model the two stop within ANGLE_TOL of each other.
        # In a scene whose markers do not fit (mis-sorted L: residuals of several mm, flat valleys) both stop at
        # ftol somewhere along the valley floor, where rounding decides the last step: 1e-2 rad there.
        tol = ANGLE_TOL if pose.ik_cost < 10.0 else 1e-2
        same_basin.append(bool(np.abs(pose.joint_angles - ang).max() < tol))
        if not same_basin[-1]:
            assert np.abs(pose.joint_angles - ang).max() < 0.05 or pose.ik_cost >= 10.0, (pose.joint_angles, ang, pose.ik_cost)
    return ang


def test_batch_of_posed_robots_matches_the_reference_chain(cuda_lib):
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(2024)
    det = FiducialDetector((32, 32, 32))
    scenes = [_scene(rng) for _ in range(24)]
    scenes += [_scene(rng, links=("Baseplate", "Joint6"), extra=0) for _ in range(8)]       # no secondary markers
    poses = det.pose_estimate([s[0] for s in scenes])
    n_ik = 0
    for pose, (pts, theta, base) in zip(poses, scenes):
        ang = _compare(pose, pts)
        n_ik += ang is not None
    assert n_ik >= 28
    assert sum(same_basin) >= math.ceil(0.95 * len(same_basin)), f"{sum(same_basin)} of {len(same_basin)} solves end where SciPy's does"
    det.close()


def test_effector_correction_saved_baseplate_and_initial_guess(cuda_lib):
    """The three inputs of _solve_full_chain_ik / _get_baseplate_transform besides the scan (Mamri.py:1376-1447):
    apply_correction (effector markers turned 180 deg about z, :1511-1514), the saved baseplate transform (preferred
    when the parameter node says so, fall-back when the scan shows no baseplate, :1376-1408), and the current joint
    angles as the first initial guess (:1425)."""
    from mamri_pose_estimation_b200.detector import FiducialDetector
    import scipy.optimize
    rng = np.random.default_rng(321)
    det = FiducialDetector((32, 32, 32))
    lim = np.radians([kin.ROBOT_BY_NAME[n]["joint_limits"] for n in kin.ARTICULATED_CHAIN])
    # --- apply_correction: markers of a robot whose effector L is mounted turned by 180 deg
    theta = np.radians(rng.uniform(-30, 30, 6))
    base = _base(rng)
    world = rb.link_world_transforms(theta, base)
    j6_local = np.asarray(rb.LINK_BY_NAME["Joint6"]["markers"]) * np.array([-1.0, -1.0, 1.0])
    j6 = j6_local @ world["Joint6"][:3, :3].T + world["Joint6"][:3, 3]
    bp = rb.marker_positions_ras(theta, base, ("Baseplate",))["Baseplate"]
    pts = np.concatenate([bp, j6]) + rng.normal(0, 0.05, (6, 3))
    ang, ident, bm = kin.pose_from_markers(pts, apply_correction=True)
    pose = det.pose_estimate([pts], apply_correction=True)[0]
    assert ang is not None and pose.joint_angles is not None
    assert pose.identified == {jn: [m["id"] for m in ms] for jn, ms in ident.items()}
    assert np.abs(pose.joint_angles - ang).max() < ANGLE_TOL
    plain = det.pose_estimate([pts])[0]                      # without the correction the same markers give another pose
    assert np.abs(plain.joint_angles - pose.joint_angles).max() > 1e-3
    # --- saved baseplate: a scan with the effector markers only
    pts6, theta6, base6 = _scene(rng, links=("Baseplate", "Joint6"), extra=0, shuffle=False)
    full = det.pose_estimate([pts6])[0]
    assert full.base_source == "scan"
    only6 = pts6[3:6]
    none = det.pose_estimate([only6])[0]
    assert none.base_matrix is None and none.joint_angles is None and none.base_source == ""
    saved = det.pose_estimate([only6], saved_base=full.base_matrix)[0]
    assert saved.base_source == "saved" and np.array_equal(saved.base_matrix, full.base_matrix)
    j6t = [only6[i] for i in saved.identified["Joint6"]]
    ref = scipy.optimize.least_squares(kin.ik_error, [0.0] * 6, bounds=(lim[:, 0], lim[:, 1]), args=(j6t, full.base_matrix, False, None),
                                       method="trf", ftol=1e-6, xtol=1e-6)
    assert np.abs(saved.joint_angles - ref.x).max() < ANGLE_TOL
    # preferred over the scan's own baseplate when asked for; ignored otherwise
    shifted = full.base_matrix.copy()
    shifted[:3, 3] += [3.0, 0.0, -2.0]
    pref = det.pose_estimate([pts6], saved_base=shifted, prefer_saved_base=True)[0]
    assert pref.base_source == "saved" and np.array_equal(pref.base_matrix, shifted)
    assert det.pose_estimate([pts6], saved_base=shifted)[0].base_source == "scan"
    # --- initial guess: the run from the current angles and the run from zeros, lower cost of the successful ones
    j6t = [pts6[i] for i in full.identified["Joint6"]]
    guess = np.clip(theta6 + 0.05, lim[:, 0], lim[:, 1])
    best = None
    for g0 in (guess, np.zeros(6)):
        r = scipy.optimize.least_squares(kin.ik_error, g0, bounds=(lim[:, 0], lim[:, 1]), args=(j6t, full.base_matrix, False, None),
                                         method="trf", ftol=1e-6, xtol=1e-6)
        if r.success and (best is None or r.cost < best.cost):
            best = r
    warm = det.pose_estimate([pts6], initial_angles=[guess])[0]
    assert np.abs(warm.joint_angles - best.x).max() < ANGLE_TOL
    assert warm.ik_iterations > full.ik_iterations           # two runs instead of one
    det.close()


def test_exact_markers_recover_the_pose(cuda_lib):
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(7)
    det = FiducialDetector((32, 32, 32))
    pts, theta, base = _scene(rng, noise=0.0, extra=0, shuffle=False)
    pose = det.pose_estimate([pts])[0]
    _compare(pose, pts)
    # exact markers: every minimum the solver can stop in with a tiny residual is an exact IK solution
    # (float32 landmarks and the y-flatten leave ~1e-5 mm in the registration)
    if pose.ik_rms_error < 1e-3:
        got = rb.marker_positions_ras(pose.joint_angles, pose.base_matrix, ("Joint6",))["Joint6"]
        assert np.abs(got - pts[[i for i in pose.identified["Joint6"]]]).max() < 1e-2
    det.close()


def test_degenerate_scans(cuda_lib):
    from mamri_pose_estimation_b200 import _capi
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(3)
    det = FiducialDetector((32, 32, 32))
    # no baseplate: matching only.  (Joint6's 45/20 L would itself pass for the 40/20 baseplate within the 5 mm
    # tolerance -- the reference's first-match rule -- so only Joint4's markers are in this scan.)
    pts, _, _ = _scene(rng, links=("Joint4",), extra=1)
    few = np.array([[0.0, 0, 0], [40, 0, 0]])
    empty = np.zeros((0, 3))
    far = rng.uniform(-500, 500, (20, 3))                                    # nothing forms an L
    poses = det.pose_estimate([pts, few, empty, far])
    for pose, p in zip(poses, (pts, few, empty, far)):
        _compare(pose, p)
    assert poses[0].base_matrix is None and poses[0].joint_angles is None and poses[0].identified
    assert poses[1].identified == {} and poses[2].identified == {} and poses[2].n_points == 0
    with pytest.raises(_capi.MamriError):
        det.pose_estimate([rng.uniform(-500, 500, (65, 3))])                 # beyond the matcher's limit
    det.close()


def test_first_match_rule_and_used_ids(cuda_lib):
    """Two candidate triplets for one link: the first 3-combination in node order wins and consumes its points."""
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(11)
    det = FiducialDetector((32, 32, 32))
    pos = rb.marker_positions_ras(np.zeros(6), np.eye(4), ("Baseplate",))["Baseplate"]
    second = pos + np.array([200.0, 0, 0]) + rng.normal(0, 0.1, (3, 3))
    pts = np.concatenate([second[[2, 0, 1]], pos + rng.normal(0, 0.1, (3, 3))])
    pose = det.pose_estimate([pts])[0]
    _compare(pose, pts, check_angles=False)
    assert sorted(pose.identified["Baseplate"]) == [0, 1, 2]
    det.close()


def test_c1_scan_to_joint_angles_on_device(cuda_lib):
    """Config C1 end to end on the device: segmentation -> marker table -> matching -> registration -> IK;
    angles within 1e-5 rad of the reference chain (SciPy) fed with the oracle's centroids."""
    from mamri_pose_estimation_b200 import phantom
    from mamri_pose_estimation_b200.logic import MamriLogic, MamriParameterNode, ScalarVolumeNode
    from oracle import segmentation as seg
    ph = phantom.config_c1()
    vol = phantom.generate(ph)
    ora = seg.detect_fiducials(vol, seg.Geometry(ph.spacing, ph.origin, ph.direction))
    ang, ident, base = kin.pose_from_markers(ora.ras_points)
    assert ang is not None
    logic = MamriLogic()
    logic.volume_threshold_segmentation(MamriParameterNode(inputVolume=ScalarVolumeNode(vol, ph.spacing, ph.origin, ph.direction)))
    got = logic.joint_detection()
    assert {k: [m["id"] for m in v] for k, v in got.items()} == {k: [m["id"] for m in v] for k, v in ident.items()}
    pose = logic.estimate_pose()
    assert np.abs(pose.base_matrix - base).max() < MATRIX_TOL
    assert np.abs(pose.joint_angles - ang).max() < ANGLE_TOL
    # the whole of process() (Mamri.py:850-881) in one call, from a device-resident volume this time
    angles = MamriLogic().process(MamriParameterNode(inputVolume=ScalarVolumeNode(torch.from_numpy(vol).cuda(), ph.spacing, ph.origin,
                                                                                   ph.direction)))
    assert np.abs(angles - ang).max() < ANGLE_TOL
    assert MamriLogic().process(MamriParameterNode(inputVolume=ScalarVolumeNode(np.zeros((16, 32, 64), np.uint16)))) is None


def test_batch_detector_poses(cuda_lib):
    """A batch of scans -> marker tables -> one pose call for the whole batch; equals the per-scan oracle chain."""
    from mamri_pose_estimation_b200 import phantom
    from mamri_pose_estimation_b200.detector import BatchDetector
    from oracle import segmentation as seg
    specs = [phantom.small_phantom(dims=(96, 80, 48), n_fiducials=6, n_blobs=2, seed=40 + i, spacing=(1.2, 1.2, 2.4)) for i in range(3)]
    vols = [phantom.generate(p) for p in specs]
    bd = BatchDetector(specs[0].dims, n_contexts=4)
    res = bd.run([torch.from_numpy(v).cuda() for v in vols], specs[0].spacing, specs[0].origin, specs[0].direction)
    poses = bd.estimate_poses(res)
    assert len(poses) == 3
    # the same stage queued on the device right behind the scans, fed from the device-written tables
    tables = torch.zeros((3, 32, 8), dtype=torch.float64, device="cuda")
    bd.begin([torch.from_numpy(v).cuda() for v in vols], specs[0].spacing, specs[0].origin, specs[0].direction, tables=tables)
    dev_poses = bd.context(0).pose_from_tables(tables)
    bd.end()
    for a, b in zip(dev_poses, poses):
        assert a.identified == b.identified and a.n_points == b.n_points
        assert (a.base_matrix is None) == (b.base_matrix is None)
        if a.base_matrix is not None:
            assert np.array_equal(a.base_matrix, b.base_matrix)
        assert (a.joint_angles is None) == (b.joint_angles is None)
        if a.joint_angles is not None:
            assert np.array_equal(a.joint_angles, b.joint_angles)
    for pose, ph, vol in zip(poses, specs, vols):
        ora = seg.detect_fiducials(vol, seg.Geometry(ph.spacing, ph.origin, ph.direction))
        assert pose.n_points == len(ora.fiducials)
        _compare(pose, ora.ras_points, check_angles=False)
    bd.close()


def test_collision_sampling_matches_the_oracle(cuda_lib):
    """mamri_collision_check (stand-in for _check_collision, Mamri.py:1555-1575): per configuration the set of links
    with a sample point inside the body labelmap and the number of such points equal the NumPy restatement."""
    from mamri_pose_estimation_b200 import phantom
    from mamri_pose_estimation_b200.detector import FiducialDetector
    rng = np.random.default_rng(99)
    dims = (160, 144, 96)                                          # x, y, z
    sp = np.array([2.5, 2.5, 4.0])
    org = -0.5 * sp * (np.array(dims) - 1)                         # volume centred on the LPS origin
    z, y, x = np.meshgrid(np.arange(dims[2]), np.arange(dims[1]), np.arange(dims[0]), indexing="ij")
    body = ((((x - 79.5) / 44.0) ** 2 + ((y - 40.0) / 26.0) ** 2 + ((z - 47.5) / 36.0) ** 2) <= 1.0).astype(np.uint8)
    m = np.array([[-1 / sp[0], 0, 0, -org[0] / sp[0]], [0, -1 / sp[1], 0, -org[1] / sp[1]], [0, 0, 1 / sp[2], -org[2] / sp[2]]])
    base = phantom.robot_base_matrix()
    base[:3, 3] = [20.0, -170.0, -60.0]                            # baseplate on the table under the body
    parts = {}                                                     # boxes around each link's axis as stand-ins for the STL vertices
    for name, (lo, hi) in {"Joint1": (0, 30), "Joint2": (0, 150), "Joint3": (0, 10), "Joint4": (0, 155), "Joint5": (0, 13),
                           "Joint6": (0, 60)}.items():
        n = 400
        parts[name] = np.stack([rng.uniform(-15, 15, n), rng.uniform(-15, 15, n), rng.uniform(lo, hi, n)], axis=1).astype(np.float32)
    configs = np.radians(rng.uniform(-100, 100, (40, 6)))
    configs[0] = 0.0                                               # arm straight up: through the body
    det = FiducialDetector((32, 32, 32))
    got = det.collision_check(parts, configs, base, torch.from_numpy(body).cuda(), m)
    n_hit = 0
    for cfg, g in zip(configs, got):
        links, inside = kin.check_collision_voxel(cfg, base, parts, body, m)
        assert g["links"] == links and g["n_points_inside"] == inside
        assert g["collision"] == bool(links) and g["first_link"] == (links[0] if links else None)
        n_hit += bool(links)
    assert 0 < n_hit < len(configs), "the test poses should include clear and colliding ones"
    # no points / no configurations
    assert det.collision_check({}, configs[:2], base, torch.from_numpy(body).cuda(), m) == [
        {"collision": False, "links": [], "first_link": None, "n_points_inside": 0}] * 2
    assert det.collision_check(parts, np.zeros((0, 6)), base, torch.from_numpy(body).cuda(), m) == []
    det.close()
