"""Timings of the BASELINE configs that bench.py's contract line does not carry (it reports C2/C3): C1, C4 (6- and
26-connectivity), C5 (entry search, with and without needle-path sampling; sharded when launched under torchrun),
the skin-surface stage and the batched pose stage.  One JSON line per measurement on rank 0; CUDA-event timed,
3 warm-ups, median of `--reps`.

    python tools/bench_configs.py [--reps 10] [--skip-c4]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py --only c5
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mamri_pose_estimation_b200 import phantom
from mamri_pose_estimation_b200 import robot as rb
from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--skip-c4", action="store_true")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
PEAK = 6544.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=args.reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms), min(ms)


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def want(name):
    return not args.only or name in args.only.split(",")


def scan_config(name, ph, conn):
    nx, ny, nz = ph.dims
    n = nx * ny * nz
    vol = generate_phantom_cuda(ph, device=local)
    det = FiducialDetector(ph.dims, device=local, max_runs=max(n // 8, 1 << 20), max_markers=4096)
    mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device=dev)
    lab = torch.empty((nz, ny, nx), dtype=torch.int32, device=dev)
    prm = DetectParams(connectivity=conn)
    res = [None]

    def run():
        res[0] = det.detect(vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction, params=prm, out_mask=mask, out_labels=lab)
    med, best = timed(run)
    det.set_profiling(True)
    run(); run()
    stages = {k: round(v, 4) for k, v in det.stage_times_ms().items()}
    det.set_profiling(False)
    r = res[0]
    emit(config=name, dims=[nx, ny, nz], connectivity=conn, ms_per_scan=med, ms_best=best, gvoxel_per_s=n / med / 1e6,
         algorithmic_gb_per_s=7.0 * n / med / 1e6, frac_of_hbm_peak=7.0 * n / med / 1e6 / PEAK, n_labels=r.n_labels, n_runs=r.n_runs,
         n_markers=len(r.markers), stages_ms=stages, note="one scan alone through one context (captured graph), collect included; "
         "u8 mask + u32 labels materialised; 7 algorithmic B/voxel")
    det.close()
    return vol


if rank == 0 and want("c1"):
    scan_config("C1", phantom.config_c1(), 6)
if rank == 0 and want("c2"):
    scan_config("C2", phantom.config_c2(), 6)
if rank == 0 and want("c4") and not args.skip_c4:
    for conn in (6, 26):
        scan_config("C4", phantom.config_c4(), conn)
        torch.cuda.empty_cache()

# ---- skin-surface stage on the C2 body
if rank == 0 and want("surface"):
    ph = phantom.config_c2()
    vol = generate_phantom_cuda(ph, device=local)
    det = FiducialDetector(ph.dims, device=local)
    r = det.detect(vol, spacing=ph.spacing, origin=ph.origin, direction=ph.direction, want_body=True)
    out = [None]

    def from_runs():
        out[0] = det.body_surface(None, shape_zyx=tuple(vol.shape), spacing=ph.spacing, origin=ph.origin, direction=ph.direction)

    def from_u8():
        out[0] = det.body_surface(r.body_mask, spacing=ph.spacing, origin=ph.origin, direction=ph.direction)
    for nm, fn in (("body of the last scan (run table)", from_runs), ("uint8 body labelmap", from_u8)):
        med, best = timed(fn)
        emit(config="surface/C2", source=nm, ms=med, ms_best=best, n_candidates=int(out[0][0].shape[0]), body_voxels=det.last_body_voxels,
             note="count pass + emit pass (two library calls, each with its own host sync); points + normals float32")
    surf_pts, surf_nrm = out[0]
    # the self-contained chain: candidates of the real body -> closest suitable entry point
    centre = surf_pts.double().mean(dim=0).cpu().numpy()
    tgt = centre + np.array([60.0, 5.0, 10.0])
    med, best = timed(lambda: det.entry_search(surf_pts, surf_nrm, tgt))
    emit(config="entry/C2-surface", n_candidates=int(surf_pts.shape[0]), ms=med, ms_best=best,
         mcandidates_per_s=surf_pts.shape[0] / med / 1e3)
    det.close()

# ---- C5: 1M analytic candidates, sharded in contiguous blocks over the ranks
if want("c5"):
    from mamri_pose_estimation_b200.distributed import gather_entry_results
    N = 1 << 20
    pts, nrm, tgt = phantom.surface_candidates(N)
    per = N // world
    lo = rank * per
    p_d = torch.from_numpy(pts[lo:lo + per]).to(dev)
    n_d = torch.from_numpy(nrm[lo:lo + per]).to(dev)
    det = FiducialDetector((64, 64, 64), device=local)
    dims = (512, 512, 256)
    ph = phantom.config_c2()
    sp, org = np.array(ph.spacing), np.array(ph.origin)
    body = torch.zeros(dims[::-1], dtype=torch.uint8, device=dev)       # path mask: 1 = free (inside the body), replicated per rank
    zz, yy, xx = torch.meshgrid(torch.arange(dims[2], device=dev), torch.arange(dims[1], device=dev), torch.arange(dims[0], device=dev), indexing="ij")
    e = ph.ellipsoids[0]
    body[((xx - e[0]) / e[3]) ** 2 + ((yy - e[1]) / e[4]) ** 2 + ((zz - e[2]) / e[5]) ** 2 <= 1.0] = 1
    del zz, yy, xx
    # RAS mm -> voxel index for identity direction: lps = -ras(x,y), index = (lps - origin) / spacing
    m = np.array([[-1 / sp[0], 0, 0, -org[0] / sp[0]], [0, -1 / sp[1], 0, -org[1] / sp[1]], [0, 0, 1 / sp[2], -org[2] / sp[2]]])
    for S in (0, 64):
        win = [None]

        def run():
            r = det.entry_search(p_d, n_d, tgt, n_path_samples=S, path_mask=body if S else None, ras_to_index=m if S else None, path_free_value=1)
            win[0] = gather_entry_results(r["index"], r["distance"], lo, dev)
        med, best = timed(run)
        t = torch.tensor([med], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        med = float(t.item())
        emit(config="C5", n_gpus=world, n_candidates=N, n_path_samples=S, ms=med, mcandidates_per_s=N / med / 1e3,
             candidate_gb_per_s=24.0 * N / med / 1e6, winner=win[0][0], distance=win[0][1],
             note="contiguous candidate blocks per rank, one all-gather of the (distance, index) winners; result copy + host sync included")
    det.close()

# ---- pose stage: 64 scans' marker tables -> matching + registration + IK
if rank == 0 and want("pose"):
    rng = np.random.default_rng(1)
    base = phantom.robot_base_matrix()
    scenes = []
    for i in range(64):
        th = np.radians(rng.uniform(-35, 35, 6))
        pos = rb.marker_positions_ras(th, base, ("Baseplate", "Joint4", "Joint6"))
        scenes.append(np.concatenate([pos[l] + rng.normal(0, 0.15, (3, 3)) for l in ("Baseplate", "Joint4", "Joint6")] + [rng.uniform(-300, 300, (3, 3)) + [0, 0, 900.0]]))
    det = FiducialDetector((32, 32, 32), device=local)
    out = [None]

    def run():
        out[0] = det.pose_estimate(scenes)
    med, best = timed(run)
    emit(config="pose", n_scans=64, points_per_scan=12, ms=med, ms_best=best, scans_per_s=64 / med * 1e3,
         ik_iterations_mean=float(np.mean([p.ik_iterations for p in out[0]])), note="host tables in, poses out (H2D + kernel + D2H + sync + Python unpacking)")
    det.close()

# ---- CPU baselines of SURVEY 8d on the box's host cores: (B1) the NumPy/SciPy oracle, single-threaded, and
# (B2) the C/OpenMP restatement on all host threads -- both on config C1 (B2 also on C2); restatements, not SimpleITK
if rank == 0 and want("cpu"):
    import time
    from oracle import c_oracle
    from oracle import segmentation as seg
    ph = phantom.config_c1()
    vol = phantom.generate(ph)
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    t0 = time.perf_counter()
    ora = seg.detect_fiducials(vol, geom)
    dt = time.perf_counter() - t0
    emit(config="cpu/B1", what="oracle/segmentation.py (NumPy + scipy.ndimage), one thread", dims=list(ph.dims), seconds=dt,
         gvoxel_per_s=vol.size / dt / 1e9, n_labels=ora.n_labels, n_markers=len(ora.fiducials))
    c_oracle.use_all_cores()
    for nm, p in (("C1", ph), ("C2", phantom.config_c2())):
        v = vol if nm == "C1" else phantom.generate(p)
        c_oracle.run_pipeline(v)
        t0 = time.perf_counter()
        for _ in range(3):
            c_oracle.run_pipeline(v)
        dt = (time.perf_counter() - t0) / 3
        emit(config="cpu/B2", what="oracle/c (C + OpenMP), all host threads", scan=nm, dims=list(p.dims), threads=c_oracle.num_threads(),
             cpu_count=os.cpu_count(), seconds=dt, gvoxel_per_s=v.size / dt / 1e9)

if world > 1:
    dist.destroy_process_group()
