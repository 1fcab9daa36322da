"""Single-context, single-stream latency of one device-resident scan through the captured graph
(no profiling events): the per-scan time a caller sees when scans arrive one at a time."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mamri_pose_estimation_b200 import phantom
from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2", choices=["c1", "c2", "c4"])
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
ph = {"c1": phantom.config_c1, "c2": phantom.config_c2, "c4": phantom.config_c4}[a.config]()
nx, ny, nz = ph.dims
vols = [generate_phantom_cuda({"c1": phantom.config_c1, "c2": phantom.config_c2, "c4": phantom.config_c4}[a.config](**({"scan_index": i} if a.config == "c2" else {})))
        for i in range(2 if a.config == "c4" else 4)]
det = FiducialDetector(ph.dims, max_runs=(nx * ny * nz) // 8)
mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device="cuda")
lab = torch.empty((nz, ny, nx), dtype=torch.int32, device="cuda")
kw = dict(spacing=ph.spacing, origin=ph.origin, direction=ph.direction, params=DetectParams(), out_mask=mask, out_labels=lab)
for i in range(5):
    det.detect(vols[i % len(vols)], **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.reps):
    det.detect_async(vols[i % len(vols)], **kw)
    r = det.collect()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"{a.config}: {ms*1e3:.1f} us/scan serial (graph, incl. collect sync) = {nx*ny*nz/ms/1e6:.1f} Gvox/s; labels {r.n_labels} runs {r.n_runs}")
det.close()
