"""The C-ABI library loads without a GPU and exports every symbol include/mamri_b200.h declares;
argument validation that needs no device works; there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mamri_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MAMRI_API\s+[\w\s\*]+?\b(mamri_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("mamri_create", "mamri_destroy", "mamri_detect_async", "mamri_detect_host_async",
                 "mamri_detect_collect", "mamri_entry_search", "mamri_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(cuda_lib):
    from mamri_pose_estimation_b200 import _capi
    for name in declared_symbols():
        assert hasattr(cuda_lib, name), f"{name} declared in the header but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes signature"
    assert cuda_lib.mamri_version().decode().startswith("mamri_b200")


def test_integration_doc_names_every_entry_point():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_symbols() if n not in doc]
    assert not missing, f"INTEGRATION.md does not say what {missing} replace"


def test_ctypes_structs_match_header_sizes(cuda_lib, tmp_path):
    """sizeof of every struct as gcc sees the header == sizeof of the ctypes mirror."""
    import subprocess
    from mamri_pose_estimation_b200 import _capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mamri_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(mamri_volume_desc),sizeof(mamri_params),sizeof(mamri_marker),sizeof(mamri_summary),'
                   'sizeof(mamri_entry_result),sizeof(mamri_link),sizeof(mamri_robot),sizeof(mamri_pose),sizeof(mamri_collision_result));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    mirrors = [_capi.VolumeDesc, _capi.Params, _capi.Marker, _capi.Summary, _capi.EntryResult, _capi.Link, _capi.Robot,
               _capi.Pose, _capi.CollisionResult]
    assert sizes == [C.sizeof(m) for m in mirrors]
    p = _capi.Params()
    cuda_lib.mamri_default_params(C.byref(p))
    assert (p.lower, p.upper, p.close_radius, p.connectivity, p.min_volume, p.max_volume) == (65.0, 65535.0, 2, 6, 50.0, 1500.0)


def test_default_robot_matches_the_oracle_restatement_of_robot_config(cuda_lib):
    """mamri_default_robot == oracle/kinematics.ROBOT (itself compared with robot_config.json when the
    reference tree is mounted, tests/test_oracle_kinematics.py)."""
    from mamri_pose_estimation_b200 import _capi
    from oracle import kinematics as kin
    r = _capi.Robot()
    cuda_lib.mamri_default_robot(C.byref(r))
    assert r.n_links == len(kin.ROBOT) and r.distance_tolerance == kin.DISTANCE_TOLERANCE and r.secondary_weight == 0.05
    names = [j["name"] for j in kin.ROBOT]
    assert (names[r.base_link], names[r.effector_link], names[r.secondary_link]) == ("Baseplate", "Joint6", "Joint4")
    for i, j in enumerate(kin.ROBOT):
        l = r.links[i]
        assert l.parent == (names.index(j["parent"]) if j["parent"] else -1)
        assert l.axis == _capi.AXIS_CODES[j.get("articulation_axis")]
        assert bool(l.has_markers) == bool(j.get("has_markers"))
        assert l.chain_index == (kin.ARTICULATED_CHAIN.index(j["name"]) if j["name"] in kin.ARTICULATED_CHAIN else -1)
        assert list(l.translate) == [float(v) for v in (j.get("translate") or [0, 0, 0])]
        if j.get("has_markers"):
            assert list(l.marker_coords) == [float(v) for p_ in j["local_marker_coords"] for v in p_]
            assert list(l.arm_lengths) == [float(v) for v in j["arm_lengths"]]
        if j["name"] in kin.ARTICULATED_CHAIN:
            assert list(l.limits_deg) == [float(v) for v in j["joint_limits"]]


def test_no_cpu_fallback_without_gpu(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mamri_pose_estimation_b200 import _capi
    ctx = C.c_void_p()
    rc = cuda_lib.mamri_create(C.byref(ctx), 0, 64, 64, 64, 0, 0)
    assert rc == _capi.MAMRI_ERR_NO_DEVICE and not ctx.value
    assert b"no CPU fallback" in cuda_lib.mamri_last_error(None)
    pool = C.c_void_p()
    rc = cuda_lib.mamri_pool_create(C.byref(pool), 0, 2, 64, 64, 64, 0, 0)
    assert rc == _capi.MAMRI_ERR_NO_DEVICE and not pool.value
    assert b"no CPU fallback" in cuda_lib.mamri_pool_last_error(None)
    assert cuda_lib.mamri_pool_create(C.byref(pool), 0, 0, 64, 64, 64, 0, 0) == _capi.MAMRI_ERR_INVALID_ARG
    from mamri_pose_estimation_b200.detector import BatchDetector, FiducialDetector
    with pytest.raises(RuntimeError):
        FiducialDetector((64, 64, 64))
    with pytest.raises(RuntimeError):
        BatchDetector((64, 64, 64))
    from mamri_pose_estimation_b200.logic import MamriLogic
    with pytest.raises(RuntimeError):
        MamriLogic()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "mamri_pose_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "libmamri_oracle" not in text, f"{f} links the oracle"


def _build_c_example(tmp_path):
    import subprocess
    exe = tmp_path / "detect_host"
    lib_dir = os.path.join(ROOT, "mamri_pose_estimation_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "detect_host.c"), "-o", str(exe), "-L", lib_dir, "-lmamri_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    return exe


def test_plain_c_caller_compiles_and_links(cuda_lib, tmp_path):
    """examples/detect_host.c: the header is valid C99 and the library links from a plain C program."""
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_plain_c_caller_matches_the_oracle(cuda_lib, tmp_path):
    import subprocess
    import numpy as np
    from mamri_pose_estimation_b200 import phantom
    from oracle import segmentation as seg
    ph = phantom.small_phantom(dims=(96, 80, 48), n_fiducials=6, seed=12, spacing=(1.2, 1.2, 2.4))
    vol = phantom.generate(ph)
    ora = seg.detect_fiducials(vol, seg.Geometry(ph.spacing, (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1)))
    path = tmp_path / "vol.u16"
    vol.tofile(path)
    exe = _build_c_example(tmp_path)
    r = subprocess.run([str(exe), str(path), "96", "80", "48", "1.2", "1.2", "2.4"], capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert lines[0] == f"labels {ora.n_labels} markers {len(ora.fiducials)} body {ora.body_label} body_voxels {int(ora.body_mask.sum())}"
    for line, label, f, ras in zip(lines[1:], ora.marker_labels, ora.fiducials, ora.ras_points):
        tok = line.split()
        assert tok[0] == label.replace("mm³", "mm3")
        assert np.allclose([float(v) for v in tok[2:5]], ras, atol=1e-5)
    assert lines[-1] == f"body_mask_voxels {int(ora.body_mask.sum())}"
