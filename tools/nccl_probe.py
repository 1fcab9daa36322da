"""Times the per-step pieces of the multi-GPU bench separately (debug aid): all-gather alone, wave alone, both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from mamri_pose_estimation_b200 import phantom
from mamri_pose_estimation_b200.detector import BatchDetector, generate_phantom_cuda
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
S = 8
gi = torch.zeros((S, 32, 8), dtype=torch.float64, device=dev); go = torch.zeros((world * S, 32, 8), dtype=torch.float64, device=dev)
def timeit(fn, n=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3
def ag():
    dist.all_gather_into_tensor(go, gi)
def ag_sync():
    dist.all_gather_into_tensor(go, gi); torch.cuda.synchronize()
specs = [phantom.config_c2(scan_index=rank * S + i) for i in range(S)]
vols = [generate_phantom_cuda(p, device=local) for p in specs]
bd = BatchDetector(specs[0].dims, device=local, n_contexts=8)
sp, org, dr = specs[0].spacing, specs[0].origin, specs[0].direction
def wave(): bd.run(vols, sp, org, dr)
def wave_ag():
    bd.begin(vols, sp, org, dr, tables=gi); w = dist.all_gather_into_tensor(go, gi, async_op=True); r = bd.end(); w.wait()
def wave_ag_sync():
    bd.begin(vols, sp, org, dr, tables=gi); r = bd.end(); dist.all_gather_into_tensor(go, gi); torch.cuda.synchronize()
def wave_barrier():
    bd.run(vols, sp, org, dr); dist.barrier()
res = {"allgather_async_ms": timeit(ag), "allgather_sync_ms": timeit(ag_sync), "wave_ms": timeit(wave, 20),
       "wave_then_gather_overlapped_ms": timeit(wave_ag, 20), "wave_then_gather_serial_ms": timeit(wave_ag_sync, 20)}
if rank == 0: print(res, flush=True)
bd.close(); dist.destroy_process_group()
