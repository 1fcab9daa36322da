"""Runs a few device-resident scans of one BASELINE config through a single context: the command
ncu wraps for the launch list / full captures (see profiles/README.md)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mamri_pose_estimation_b200 import phantom
from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2", choices=["c1", "c2", "c4"])
ap.add_argument("--scans", type=int, default=3)
ap.add_argument("--conn", type=int, default=6)
a = ap.parse_args()
ph = {"c1": phantom.config_c1, "c2": phantom.config_c2, "c4": phantom.config_c4}[a.config]()
vol = generate_phantom_cuda(ph)
nx, ny, nz = ph.dims
det = FiducialDetector(ph.dims, max_runs=(nx * ny * nz) // 8)
mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device="cuda")
lab = torch.empty((nz, ny, nx), dtype=torch.int32, device="cuda")
det.set_profiling(True)
for i in range(a.scans):
    r = det.detect(vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction,
                   params=DetectParams(connectivity=a.conn), out_mask=mask, out_labels=lab)
    print(i, r.n_labels, r.n_runs, len(r.markers), r.body_label, det.stage_times_ms())
kt = det.kernel_times_ms()
for n, ms in kt:
    print(f'{ms*1e3:9.2f} us  {n}')
print(f'sum {sum(m for _, m in kt)*1e3:.1f} us')
det.close()
