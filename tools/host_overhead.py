"""Where does the host spend its time in BatchDetector.run? (cProfile over a few steps)"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mamri_pose_estimation_b200 import phantom
from mamri_pose_estimation_b200.detector import BatchDetector, generate_phantom_cuda
S = 8
specs = [phantom.config_c2(scan_index=i) for i in range(S)]
vols = [generate_phantom_cuda(p) for p in specs]
bd = BatchDetector(specs[0].dims, n_contexts=int(os.environ.get("MAMRI_BENCH_CONTEXTS", "3")))
sp, org, dr = specs[0].spacing, specs[0].origin, specs[0].direction
for _ in range(3):
    bd.run(vols, sp, org, dr)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    bd.run(vols, sp, org, dr)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host loop {1e3*(t1-t0)/20:.3f} ms/step, with final sync {1e3*(t2-t0)/20:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    bd.run(vols, sp, org, dr)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
