"""Compares the device IK with the SciPy chain on the scenes of tests/test_gpu_pose.py (debug aid)."""
import os, sys, importlib.util
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
spec = importlib.util.spec_from_file_location("tp", os.path.join(os.path.dirname(__file__), "..", "tests", "test_gpu_pose.py"))
tp = importlib.util.module_from_spec(spec); spec.loader.exec_module(tp)
from oracle import kinematics as kin
from mamri_pose_estimation_b200.detector import FiducialDetector
rng = np.random.default_rng(2024)
det = FiducialDetector((32, 32, 32))
scenes = [tp._scene(rng) for _ in range(24)] + [tp._scene(rng, links=("Baseplate", "Joint6"), extra=0) for _ in range(8)]
poses = det.pose_estimate([s[0] for s in scenes])
for i, (pose, (pts, theta, base)) in enumerate(zip(poses, scenes)):
    ang, ident, b = kin.pose_from_markers(pts)
    j6 = [pts[k] for k in pose.identified["Joint6"]]
    j4 = [pts[k] for k in pose.identified["Joint4"]] if "Joint4" in pose.identified else None
    e = np.array(kin.ik_error(ang, j6, b, False, j4))
    print(i, "dev cost %.6g it %d rms %.4g | scipy cost %.6g | max dtheta %.3g | dtruth dev %.3g scipy %.3g" % (
        pose.ik_cost, pose.ik_iterations, pose.ik_rms_error, 0.5 * float(e @ e), np.abs(pose.joint_angles - ang).max(),
        np.abs(pose.joint_angles - theta).max(), np.abs(ang - theta).max()))
    f = lambda x: np.array(kin.ik_error(x, j6, b, False, j4))
    h = 1e-6
    grad = np.array([(0.5 * np.sum(f(pose.joint_angles + h * e_) ** 2) - 0.5 * np.sum(f(pose.joint_angles - h * e_) ** 2)) / (2 * h) for e_ in np.eye(6)])
    lim = np.radians([kin.ROBOT_BY_NAME[n]["joint_limits"] for n in kin.ARTICULATED_CHAIN])
    interior = (pose.joint_angles > lim[:, 0] + 1e-9) & (pose.joint_angles < lim[:, 1] - 1e-9)
    print("    conv", pose.ik_converged, "grad interior max %.3g" % (np.abs(grad[interior]).max() if interior.any() else 0), "active", int((~interior).sum()))
from mamri_pose_estimation_b200 import phantom
from oracle import segmentation as seg
ph = phantom.config_c1()
pts = np.concatenate([ph.truth["marker_ras"][l] for l in ("Baseplate", "Joint4", "Joint6")])
ang, ident, b = kin.pose_from_markers(pts)
pose = det.pose_estimate([pts])[0]
print("C1 truth markers: scipy-vs-dev", np.abs(pose.joint_angles - ang).max(), "truth", np.abs(np.array(ph.truth["pose_rad"]) - ang).max(), pose.identified, {k: [m["id"] for m in v] for k, v in ident.items()})
