"""Host <-> device copy ceilings of the box at N concurrent ranks, one process per GPU: pinned H2D alone, D2H alone, and
the two mixes the end-to-end path moves per scan (128 MiB in with the 64 MiB uint8 body labelmap out, or with the
8 MiB bit-packed one).  bench.py measures the mix it uses itself (`e2e.copy_ceiling`); this is the stand-alone form.

    python tools/pcie_probe.py
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every figure is the aggregate over the N ranks, timed with CUDA events between barriers, max over the ranks."""
import json
import os

import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
n = 128 << 20
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(n // 2, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
d_out = [torch.empty(n // 2, dtype=torch.uint8, device=dev) for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(h2d, out_bytes, reps=10):
    """GB/s (aggregate) host-to-device and device-to-host when both run at once; out_bytes per 128 MiB in (0 = none)."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        for i in range(4):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[i].copy_(h_in[i], non_blocking=True)
            if out_bytes:
                with torch.cuda.stream(s2):
                    h_out[i][:out_bytes].copy_(d_out[i][:out_bytes], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return (world * reps * 4 * n / ms / 1e6 if h2d else 0.0), world * reps * 4 * out_bytes / ms / 1e6


for _ in range(2):
    run(True, n // 2, 2)
res = {"n_gpus": world, "h2d_alone_gb_s": run(True, 0)[0], "d2h_alone_gb_s": run(False, n // 2)[1]}
a, b = run(True, n // 2)
res["mix_uint8_body"] = {"h2d_gb_s": a, "d2h_gb_s": b, "scans_per_s": a * 1e9 / n, "gvoxel_per_s": a / 2}
a, b = run(True, n // 16)
res["mix_bit_packed_body"] = {"h2d_gb_s": a, "d2h_gb_s": b, "scans_per_s": a * 1e9 / n, "gvoxel_per_s": a / 2}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
