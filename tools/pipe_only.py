"""Times BatchPipeline (two alternating pools) against BatchDetector.run on a device-resident C2 batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mamri_pose_estimation_b200 import phantom
from mamri_pose_estimation_b200.detector import BatchDetector, BatchPipeline, generate_phantom_cuda
S = int(os.environ.get("SCANS", "8")); K = int(os.environ.get("MAMRI_BENCH_CONTEXTS", "8")); steps = int(os.environ.get("STEPS", "40"))
depth = int(os.environ.get("DEPTH", "2"))
specs = [phantom.config_c2(scan_index=i) for i in range(S)]
vols = [generate_phantom_cuda(p) for p in specs]
sp, org, dr = specs[0].spacing, specs[0].origin, specs[0].direction
bp = BatchPipeline(specs[0].dims, n_contexts=K, depth=depth)
def run(n):
    bp.submit(vols, sp, org, dr)
    r = None
    for k in range(n):
        if k + 1 < n:
            bp.submit(vols, sp, org, dr)
        r = bp.result()
    return r
run(6)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = run(steps)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"pipeline depth {depth} S={S} K={K}: {ms*1e3/S:.1f} us/scan, {S*512*512*256/ms/1e6:.1f} Gvox/s  (labels {r[0].n_labels}, markers {len(r[0].markers)})")
bp.close()
