"""The N>1 path on CPU: world_size-2 gloo.  Each rank runs the (CPU oracle as stand-in for the) per-scan
detection on its shard and the marker tables are all-gathered exactly as bench.py does over NCCL; the
gathered table must equal the single-process result, and the sharded entry search must find the global
arg-min."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mamri_pose_estimation_b200 import distributed as mdist
from mamri_pose_estimation_b200 import phantom
from oracle import kinematics as kin
from oracle import segmentation as seg

N_SCANS = 4


def _scan_result(i):
    ph = phantom.small_phantom(dims=(40, 32, 24), seed=30, scan_index=i)
    det = seg.detect_fiducials(phantom.generate(ph), seg.Geometry(ph.spacing, ph.origin, ph.direction), min_vol=20, max_vol=600)
    by = {s.label: s for s in det.stats}
    markers = [types.SimpleNamespace(label=f["id"], count=by[f["id"]].count, volume_mm3=f["vol"], centroid_ras=r)
               for f, r in zip(det.fiducials, det.ras_points)]
    return types.SimpleNamespace(markers=markers, n_labels=det.n_labels, body_label=det.body_label)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = mdist.shard_indices(N_SCANS, rank, world)
    local = torch.from_numpy(mdist.pack_table([_scan_result(i) for i in mine]))
    full = mdist.unshard(mdist.gather_tables(local), N_SCANS, world)
    # sharded entry search: contiguous candidate blocks, global arg-min
    pts, nrm, tgt = phantom.surface_candidates(4096, seed=9)
    per = 4096 // world
    li, ld = kin.find_entry_point(pts[rank * per:(rank + 1) * per], nrm[rank * per:(rank + 1) * per], tgt)
    gi, gd = mdist.gather_entry_results(li, ld, rank * per, torch.device("cpu"))
    if rank == 0:
        q.put((full.numpy(), gi, gd))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gather_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, gi, gd = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = mdist.pack_table([_scan_result(i) for i in range(N_SCANS)])
    assert np.array_equal(full, want)
    pts, nrm, tgt = phantom.surface_candidates(4096, seed=9)
    wi, wd = kin.find_entry_point(pts, nrm, tgt)
    assert (gi, gd) == (wi, wd)
