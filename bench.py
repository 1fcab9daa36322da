#!/usr/bin/env python
"""Benchmark of the fiducial seg + CCL + stats hot path (BASELINE.json metric: Gvoxel/s and scans/s,
% of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # CUDA path (one process per GPU)
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU arm (C oracle, all host threads)

A step = one pass of the hot path (threshold -> ball closing -> connected components -> label statistics
-> marker filter, with the closed mask and the label volume materialised) over one batch of
`--scans-per-gpu` synthetic 512x512x256 uint16 phantoms per GPU (BASELINE config C2, the same workload at
every N: weak scaling).  Scans shard across ranks with no data-path collective; the only exchange is one NCCL
all-gather of the per-scan marker tables per step.  Prints ONE JSON line on rank 0, which also carries

  e2e        the same metric through the host-buffer call (pinned host volumes in, marker table + body labelmap
             out, copies inside the timed region) and the plain-copy ceiling of that path measured at the same N
  roofline   the dominant kernel against the measured HBM peak, and the whole pipeline at 7 algorithmic B/voxel
             for the batch and for one scan alone
  configs    every other BASELINE config as BASELINE.json words it: C1 and C4 (6 / 26) as lone scans, C3 (64
             sigma-15 scans sharded over the N ranks, gathered tables compared with a one-GPU run of the same
             batch), C5 (1 M entry-point candidates sharded over the N ranks, 0 and 64 path samples), each with
             its own fraction of the roofline and parity booleans against the oracle
  cpu_baseline, clocks, rank_ms_per_step (spread over the ranks), gpu_launches
MAMRI_BENCH_NO_GATHER=1 drops the all-gather (to attribute its cost), MAMRI_BENCH_BODY_U8=1 returns the body
labelmap of the e2e path as uint8 instead of 1 bit per voxel.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gvoxel/s fiducial seg+CCL+stats"
ALGO_BYTES_PER_VOXEL = {"threshold_pack": 2.0 + 1.0 / 8.0,    # read u16 once, write 1 bit
                        "materialise": 1.0 + 4.0 + 1.0 / 8.0}   # write u8 mask + u32 label, read 1 bit
DIMS = (512, 512, 256)
MAX_TABLE = 32          # marker slots per scan in the gathered table


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, sampled every 2 ms through NVML on a
    background thread (nvidia-smi -lms needs longer to start than a short run lasts)."""

    def __init__(self, gpu_index: int):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.recording = False               # samples are kept only while the timed region runs
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:          # no NVML: report that instead of inventing numbers
            self._err = str(e)

    def _run(self):
        nv = self._nv
        flags = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            if not self.recording:
                self._stop.wait(0.001)
                continue
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in flags.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def begin(self) -> None:
        self.recording = True

    def stop(self) -> dict:
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable: " + getattr(self, "_err", "?")]}
        self._stop.set()
        self._thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


def run_reference(args):
    """CPU arm: the path restated in C (oracle/c, OpenMP over all host threads) -- SimpleITK, which the
    reference calls at Mamri.py:1308-1310, is not installable here.  One C2 scan per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mamri_pose_estimation_b200 import phantom
    from oracle import c_oracle
    c_oracle.use_all_cores()
    ph = phantom.config_c2()
    vol = phantom.generate(ph)
    n = vol.size
    for _ in range(args.warmup):
        c_oracle.run_pipeline(vol)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.run_pipeline(vol)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e9
    cores = c_oracle.num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gvoxel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "scans_per_s": args.steps / dt,
            "config": {"workload": "C2: 512x512x256 uint16 phantom, 6 fiducials, Rician sigma 10; 1 scan per step"},
            "cpu_baseline": {"value": val, "unit": "Gvoxel/s", "cores": cores, "kind": "port",
                             "sample": "one 512x512x256 scan per step through oracle/c (threshold, closing, CCL, "
                                       "label sums); SimpleITK itself is not installable offline"},
            "e2e": {"value": val, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# helpers of the CUDA arm
# ----------------------------------------------------------------------------------------------------------------
class LoneScan:
    """One device-resident scan through one context: the literal BASELINE configs C1 / C2 / C4 ("one scan, 1 B200").
    Timed through the bare C ABI (mamri_detect_async + mamri_detect_collect with preallocated structs, closed mask and
    label volume materialised) with CUDA events around `reps` back-to-back scans: what a caller that hands over one
    scan at a time waits for, graph launch and result synchronisation included."""

    def __init__(self, ph, device, conn=6, max_markers=4096):
        import ctypes as C
        import torch
        from mamri_pose_estimation_b200 import _capi
        from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, _desc, generate_phantom_cuda
        self.C, self.torch, self.ph = C, torch, ph
        nx, ny, nz = ph.dims
        self.n_vox = nx * ny * nz
        dev = torch.device(f"cuda:{device}")
        self.vol = generate_phantom_cuda(ph, device=device)
        self.det = FiducialDetector(ph.dims, device=device, max_runs=max(self.n_vox // 8, 1 << 20), max_markers=max_markers)
        self.mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device=dev)
        self.labels = torch.empty((nz, ny, nx), dtype=torch.int32, device=dev)
        self.params = DetectParams(connectivity=conn)
        self.desc = _desc((nz, ny, nx), "uint16", ph.spacing, ph.origin, ph.direction)
        self.summary = _capi.Summary()
        self.markers = (_capi.Marker * max_markers)()
        self.max_markers = max_markers

    def result(self):
        return self.det.detect(self.vol, spacing=self.ph.spacing, origin=self.ph.origin, direction=self.ph.direction,
                               params=self.params, out_mask=self.mask, out_labels=self.labels)

    def time_bare(self, reps, warm=4):
        C, torch, lib = self.C, self.torch, self.det._lib
        p = self.params.to_c()
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        vp, mp, lp = C.c_void_p(self.vol.data_ptr()), C.c_void_p(self.mask.data_ptr()), C.c_void_p(self.labels.data_ptr())

        def one():
            rc = lib.mamri_detect_async(self.det._ctx, C.byref(self.desc), vp, C.byref(p), mp, lp, None, s)
            rc |= lib.mamri_detect_collect(self.det._ctx, C.byref(self.summary), self.markers, self.max_markers)
            if rc:
                raise RuntimeError(f"lone scan failed: {rc}")
        for _ in range(warm):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            one()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def stages(self, reps=6):
        self.det.set_profiling(True)
        acc = {}
        for i in range(reps + 2):
            self.result()
            if i >= 2:
                for k, v in self.det.stage_times_ms().items():
                    acc[k] = acc.get(k, 0.0) + v / reps
        self.det.set_profiling(False)
        return acc

    def close(self):
        self.det.close()


def oracle_parity(lone, res):
    """Full-size comparison of one scan with the C oracle: closed mask, label volume, kept labels, centroids."""
    import numpy as np
    from oracle import c_oracle
    from oracle import segmentation as seg
    ph = lone.ph
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    host = lone.vol.cpu().numpy()
    ora = c_oracle.detect_fiducials(host, geom, connectivity=lone.params.connectivity, want_body_mask=False)
    return {"mask_bit_exact": bool(np.array_equal(lone.mask.cpu().numpy(), ora.closed)),
            "labels_bit_exact": bool(np.array_equal(lone.labels.cpu().numpy().view(np.uint32), ora.labels)),
            "markers_equal": [m.label for m in res.markers] == [f["id"] for f in ora.fiducials],
            "max_centroid_err_mm": float(max([np.abs(np.array(m.centroid_lps) - np.array(f["centroid"])).max()
                                              for m, f in zip(res.markers, ora.fiducials)] or [0.0])),
            "n_markers": len(res.markers), "n_labels": res.n_labels, "body_label_equal": res.body_label == ora.body_label}


def run_native(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from mamri_pose_estimation_b200 import _capi, phantom
    from mamri_pose_estimation_b200.detector import (BatchPipeline, DetectParams, FiducialDetector, body_bits_shape,
                                                     generate_phantom_cuda)
    from mamri_pose_estimation_b200.distributed import gather_entry_results, pack_table, unshard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa = None
    if world > 1 and os.environ.get("MAMRI_BENCH_NUMA", "1") != "0":
        # one process per GPU: run on the CPUs next to this GPU *before* any pinned host buffer is allocated, so that
        # the staging memory of the e2e path is local to the GPU's PCIe root (first touch)
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[local]) if visible and visible.split(",")[local].isdigit() else local
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1} & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus)"
        except Exception as e:                      # no NVML / not permitted: run unpinned
            numa = f"unpinned: {e}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    no_gather = os.environ.get("MAMRI_BENCH_NO_GATHER", "0") == "1"       # experiment: the step without its one collective
    peak, peak_src = measured_peak_gbs()

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_values(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.tolist()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S = args.scans_per_gpu
    nx, ny, nz = DIMS
    n_vox = nx * ny * nz
    specs = [phantom.config_c2(scan_index=rank * S + i) for i in range(S)]
    vols = [generate_phantom_cuda(p, device=local) for p in specs]
    torch.cuda.synchronize()
    sp, org, dr = specs[0].spacing, specs[0].origin, specs[0].direction
    params = DetectParams()
    n_ctx = int(os.environ.get("MAMRI_BENCH_CONTEXTS", "8"))
    depth = int(os.environ.get("MAMRI_BENCH_DEPTH", "2")) if S <= n_ctx else 1      # pools alternating (batches in flight)
    bp = BatchPipeline(DIMS, device=local, n_contexts=n_ctx, depth=depth)
    bd = bp.pools[0]
    # the single exchange of the path: the scans' last kernels write their fixed-size marker tables into gather_in,
    # and the all-gather is queued behind them (on the batch's stream) before the host waits for the results
    gather_in = [torch.zeros((S, MAX_TABLE, 8), dtype=torch.float64, device=dev) for _ in range(depth)]
    gather_out = [torch.zeros((world * S, MAX_TABLE, 8), dtype=torch.float64, device=dev) for _ in range(depth)]
    use_gather = world > 1 and not no_gather
    gstream = torch.cuda.Stream(device=dev)
    gather_done = [torch.cuda.Event() for _ in range(depth)]

    def run_steps(n):
        """n steps, software-pipelined over the two pools: batch k+1 is enqueued before batch k is collected."""
        if S > n_ctx:                                   # more scans than contexts: several waves per step, no pipelining
            for _ in range(n):
                res = bd.run(vols, sp, org, dr, params)
                if use_gather:
                    gather_in[0].copy_(torch.from_numpy(pack_table(res)), non_blocking=True)
                    dist.all_gather_into_tensor(gather_out[0], gather_in[0])
            return res, 0
        def submit(k):
            slot = k % depth
            if use_gather:
                bp.streams[slot].wait_event(gather_done[slot])       # the tables of this slot's previous batch have been gathered
            st = bp.submit(vols, sp, org, dr, params, tables=gather_in[slot] if use_gather else None)
            if use_gather:
                # the all-gather runs on a stream of its own behind the batch: the pool's stream is free for its next
                # batch, and a rank that is a little late does not hold up the others' scans, only their gather
                gstream.wait_stream(st)
                with torch.cuda.stream(gstream):
                    dist.all_gather_into_tensor(gather_out[slot], gather_in[slot])
                    gather_done[slot].record(gstream)
        res, submitted = None, 0
        for k in range(n):
            while submitted < n and bp.pending() < depth:
                submit(submitted)
                submitted += 1
            res = bp.result()
        if use_gather:
            torch.cuda.current_stream().wait_stream(gstream)            # the timed region ends when the last gather has
        return res, (n - 1) % depth

    # ---------------- device-resident throughput (the contract's `value`)
    # at least two batches per pool: the first sizes the run-table grids from its run count, the second re-captures the graph
    n_warm = max(args.warmup, 3, 2 * depth)
    res, _ = run_steps(n_warm)
    sampler = ClockSampler(local) if rank == 0 else None       # NVML thread (set up before the barrier: nvmlInit takes ms)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.begin()                                        # 2 ms period, timed region only
    e0.record()
    res, last_slot = run_steps(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # NVML can be slow to answer on some hosts: if the timed region saw fewer than 5 clock samples, keep the same load
    # running (untimed) until the sampler has them -- every rank takes part, rank 0 decides
    extra_steps = 0
    t_stop = time.perf_counter() + 2.0
    while True:
        more = torch.tensor([1 if (sampler is not None and sampler._thread is not None and len(sampler.samples) < 5
                                   and time.perf_counter() < t_stop) else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(more, op=dist.ReduceOp.MAX)
        if not int(more.item()):
            break
        run_steps(10)
        extra_steps += 10
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["untimed_load_steps_for_sampling"] = extra_steps
    rank_ms = all_values(ms)
    ms_max = max(rank_ms)
    value = world * S * n_vox * args.steps / (ms_max * 1e-3) / 1e9

    gathered_ok = None
    if use_gather:                                      # the device-written tables equal the host-packed ones, on every rank's slot
        mine = gather_out[last_slot][rank * S:(rank + 1) * S].cpu().numpy()
        gathered_ok = bool(np.array_equal(mine, pack_table(res)))

    # ---------------- end to end through the host-buffer call: pinned host u16 in, marker table + body labelmap out.
    # The body labelmap goes back at 1 bit per voxel (mamri_pool_detect_host_bits_begin: 8x fewer bytes on the link that
    # bounds this path); MAMRI_BENCH_BODY_U8=1 measures the uint8 form instead.
    body_u8 = os.environ.get("MAMRI_BENCH_BODY_U8", "0") == "1"
    h_vols = [torch.empty((nz, ny, nx), dtype=torch.uint16).pin_memory() for _ in range(S)]
    for h, v in zip(h_vols, vols):
        h.copy_(v)

    def body_buffers():
        if body_u8:
            return [torch.empty((nz, ny, nx), dtype=torch.uint8).pin_memory() for _ in range(S)]
        return [torch.empty(body_bits_shape((nz, ny, nx)), dtype=torch.int32).pin_memory() for _ in range(S)]
    # two sets of host body-mask buffers: batch k+1 is enqueued while batch k's results are still being written
    bodies = [body_buffers() for _ in range(depth)]
    body_bytes = bodies[0][0].numel() * bodies[0][0].element_size()
    body_kw = "body_out" if body_u8 else "body_bits_out"
    torch.cuda.synchronize()

    def run_host_steps(n):
        """n end-to-end steps, software-pipelined over the two pools like run_steps: the next batch's H2D copies start
        while the previous batch's last scans still compute and drain (the PCIe link never idles between steps)."""
        if depth == 1:
            for _ in range(n):                          # more scans than contexts: chunks of n_ctx, one after the other
                for first in range(0, S, n_ctx):
                    bd.begin_host(h_vols[first:first + n_ctx], sp, org, dr, params, **{body_kw: bodies[0][first:first + n_ctx]})
                    r = bd.end()
            return r
        r, submitted = None, 0
        for k in range(n):
            while submitted < n and bp.pending() < depth:
                bp.submit_host(h_vols, sp, org, dr, params, **{body_kw: bodies[submitted % depth]})
                submitted += 1
            r = bp.result()
            if use_gather:                              # host-packed tables on this path (the marker tables are on the host anyway)
                gather_in[k % depth].copy_(torch.from_numpy(pack_table(r)), non_blocking=True)
                dist.all_gather_into_tensor(gather_out[k % depth], gather_in[k % depth])
        return r

    e2e_steps = max(4, min(args.steps, 30))
    res_h = run_host_steps(2 * depth)
    barrier()
    e0.record()
    res_h = run_host_steps(e2e_steps)
    e1.record()
    barrier()
    e2e_ms = all_max(e0.elapsed_time(e1))
    e2e_value = world * S * n_vox * e2e_steps / (e2e_ms * 1e-3) / 1e9

    # the ceiling of that path on this box, measured the same way at the same N: the same bytes in the same directions
    # with plain copies (one process per GPU, pinned buffers, both directions at once), no kernels at all
    d_in = [torch.empty((nz, ny, nx), dtype=torch.uint16, device=dev) for _ in range(2)]
    d_body = [torch.empty_like(bodies[0][0], device=dev) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copy_steps(n):
        for _ in range(n):
            for i in range(S):
                with torch.cuda.stream(s_in):
                    d_in[i % 2].copy_(h_vols[i], non_blocking=True)
                with torch.cuda.stream(s_out):
                    bodies[0][i].copy_(d_body[i % 2], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in)
        torch.cuda.current_stream().wait_stream(s_out)
    copy_steps(2)
    barrier()
    e0.record()
    s_in.wait_stream(torch.cuda.current_stream()); s_out.wait_stream(torch.cuda.current_stream())
    copy_steps(e2e_steps)
    e1.record()
    barrier()
    ceil_ms = all_max(e0.elapsed_time(e1))
    ceiling = world * S * n_vox * e2e_steps / (ceil_ms * 1e-3) / 1e9
    n_mk = sum(len(r.markers) for r in res_h)
    table_bytes = S * ctypes.sizeof(_capi.Summary) + n_mk * ctypes.sizeof(_capi.Marker)      # written by the device into pinned memory
    e2e = {"value": e2e_value, "unit": "Gvoxel/s", "scans_per_s": e2e_value * 1e9 / n_vox, "steps": e2e_steps,
           "h2d_bytes_per_step": world * S * n_vox * 2, "d2h_bytes_per_step": world * (S * body_bytes + table_bytes),
           "body_labelmap": "uint8" if body_u8 else "1 bit per voxel",
           "copy_ceiling": {"value": ceiling, "unit": "Gvoxel/s",
                            "h2d_gb_per_s": world * S * n_vox * 2 * e2e_steps / (ceil_ms * 1e-3) / 1e9,
                            "d2h_gb_per_s": world * S * body_bytes * e2e_steps / (ceil_ms * 1e-3) / 1e9,
                            "how": f"the same {S} x 128 MiB host-to-device and {S} x {body_bytes >> 20} MiB device-to-host per step and "
                                   f"rank as plain concurrent copies from / to the same pinned buffers on {world} rank(s), no kernels"},
           "frac_of_copy_ceiling": e2e_value / ceiling,
           "api": "BatchPipeline.submit_host/result -> mamri_pool_detect_host_bits_begin / mamri_pool_detect_end (pinned host u16 "
                  "volumes in; marker table + body labelmap out), two pools alternating"}
    del d_in, d_body

    # ---------------- the BASELINE configs, each as BASELINE.json words it
    configs = {}
    stages, roof, cpu, parity = None, None, None, None
    c2_ms = None
    if rank == 0:
        # C2: one clinical-size scan alone
        lone = LoneScan(phantom.config_c2(), local)
        c2_ms = lone.time_bare(30)
        acc = lone.stages(8)
        stages = {k: round(v, 4) for k, v in acc.items()}
        r0 = lone.result()
        configs["C2"] = {"what": "one 512x512x256 scan alone on one B200 (bare C ABI, collect included)", "ms_per_scan": c2_ms,
                         "gvoxel_per_s": n_vox / c2_ms / 1e6, "frac_of_hbm_peak": 7.0 * n_vox / c2_ms / 1e6 / peak,
                         "kernels_per_scan": lone.det.kernel_launches, "n_runs": r0.n_runs, "n_labels": r0.n_labels,
                         "n_markers": len(r0.markers)}
        dom = max(ALGO_BYTES_PER_VOXEL, key=lambda k: acc[k])
        achieved = ALGO_BYTES_PER_VOXEL[dom] * n_vox / (acc[dom] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_voxel": ALGO_BYTES_PER_VOXEL[dom],
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_VOXEL[dom] * n_vox,
                "launch_ms": acc[dom], "timing": "CUDA events on the launching stream, mean of 8 scans through one "
                                                 "context after 2 warm-ups (kernel timed alone: burst peak applies)",
                "kernels": {k: {"launch_ms": acc[k], "achieved": ALGO_BYTES_PER_VOXEL[k] * n_vox / (acc[k] * 1e-3) / 1e9,
                                "frac": ALGO_BYTES_PER_VOXEL[k] * n_vox / (acc[k] * 1e-3) / 1e9 / peak} for k in ALGO_BYTES_PER_VOXEL},
                # whole pipeline at 7 algorithmic B/voxel (read u16 once, write u8 mask + u32 label once):
                # `batch` = the timed region above (all contexts in flight), `serial` = one scan alone (config C2)
                "pipeline": {"algorithmic_bytes_per_voxel": 7.0,
                             "batch": {"achieved": 7.0 * value / world, "frac": 7.0 * value / world / peak},   # per GPU
                             "serial": {"achieved": 7.0 * n_vox / (c2_ms * 1e-3) / 1e9,
                                        "frac": 7.0 * n_vox / (c2_ms * 1e-3) / 1e9 / peak, "ms_per_scan": c2_ms,
                                        "how": "30 scans one after the other through the bare C ABI (graph launch, "
                                               "result synchronisation and collect included), CUDA events around the loop",
                                        "stage_sum_ms": sum(acc.values())}}}
        # the stages between the two streaming kernels: what they have to move (1 bit per voxel volumes, 4 bytes per mask
        # word of run bases, the run table) against the time they take alone -- latency-bound by design
        n_words = n_vox // 32
        mid_bytes = {"closing": 2 * n_vox // 8,                                           # raw mask in, closed mask out
                     "ccl": n_vox // 8 + 4 * n_words + 5 * 4 * r0.n_runs,                 # mask in; word_base + 5 run-table columns out
                     "stats_filter": 7 * 4 * r0.n_runs}                                   # run-table columns in
        roof["stages"] = {k: {"ms": acc[k], "algorithmic_bytes": b, "achieved": b / (acc[k] * 1e-3) / 1e9,
                              "frac": b / (acc[k] * 1e-3) / 1e9 / peak,
                              "bound": "latency: dependent L2 loads / atomics on 1 bit/voxel data and the run table, not HBM"}
                          for k, b in mid_bytes.items() if acc.get(k)}
        if not args.no_cpu_baseline:
            parity = oracle_parity(lone, r0)
            configs["C2"]["parity"] = parity
        lone.close()
        # C1: the reference's own CPU-runnable case
        ph1 = phantom.config_c1()
        lone = LoneScan(ph1, local)
        ms1 = lone.time_bare(30)
        r1 = lone.result()
        v1 = ph1.dims[0] * ph1.dims[1] * ph1.dims[2]
        configs["C1"] = {"what": "one 256x256x128 scan with 9 fiducials alone", "ms_per_scan": ms1, "gvoxel_per_s": v1 / ms1 / 1e6,
                         "frac_of_hbm_peak": 7.0 * v1 / ms1 / 1e6 / peak, "kernels_per_scan": lone.det.kernel_launches,
                         "n_runs": r1.n_runs, "n_labels": r1.n_labels, "n_markers": len(r1.markers)}
        if not args.no_cpu_baseline:
            configs["C1"]["parity"] = oracle_parity(lone, r1)
        lone.close()
    # C4: merge stress, 6- and 26-connectivity (rank 0; the 1 GiB volume + 2 GiB label volume fit one GPU)
    if rank == 0 and not args.skip_c4:
        ph4 = phantom.config_c4()
        v4 = ph4.dims[0] * ph4.dims[1] * ph4.dims[2]
        for conn in (6, 26):
            lone = LoneScan(ph4, local, conn=conn, max_markers=16384)
            ms4 = lone.time_bare(8, warm=3)
            r4 = lone.result()
            c = {"what": f"one 1024x1024x512 scan (32 fiducials + 2000 blobs, sigma 20) alone, {conn}-connectivity",
                 "ms_per_scan": ms4, "gvoxel_per_s": v4 / ms4 / 1e6, "frac_of_hbm_peak": 7.0 * v4 / ms4 / 1e6 / peak,
                 "kernels_per_scan": lone.det.kernel_launches, "n_runs": r4.n_runs, "n_labels": r4.n_labels, "n_markers": len(r4.markers)}
            if not args.no_cpu_baseline:
                c["parity"] = oracle_parity(lone, r4)
            configs[f"C4_conn{conn}"] = c
            lone.close()
            del lone
            torch.cuda.empty_cache()

    # C3: the 64-scan noisy batch (sigma 15) sharded round-robin over the ranks: scan i -> rank i % world
    n3 = args.c3_scans
    mine3 = list(range(rank, n3, world))
    per3 = (n3 + world - 1) // world
    vols3 = [generate_phantom_cuda(phantom.config_c3(i), device=local) for i in mine3]
    t_in = torch.zeros((per3, MAX_TABLE, 8), dtype=torch.float64, device=dev)
    t_out = torch.zeros((world * per3, MAX_TABLE, 8), dtype=torch.float64, device=dev)

    def run_c3(volumes, tables):
        """All of a rank's scans in chunks of n_ctx through the two alternating pools; device-written tables."""
        chunks = [volumes[i:i + n_ctx] for i in range(0, len(volumes), n_ctx)]
        outs = []
        for k, ch in enumerate(chunks):
            if bp.pending() >= depth:
                outs.append(bp.result())
            bp.submit(ch, sp, org, dr, params, tables=tables[k * n_ctx:k * n_ctx + len(ch)])
        while bp.pending():
            outs.append(bp.result())
        return outs

    def step_c3():
        run_c3(vols3, t_in)
        if world > 1:
            dist.all_gather_into_tensor(t_out, t_in)
        else:
            t_out.copy_(t_in)
    step_c3(); step_c3()
    barrier()
    reps3 = 3
    e0.record()
    for _ in range(reps3):
        step_c3()
    e1.record()
    barrier()
    ms3 = all_max(e0.elapsed_time(e1)) / reps3
    gathered = unshard(t_out, n3, world).cpu().numpy()
    same = None
    if rank == 0:                                       # SURVEY 4 tier 5: the same batch on one GPU gives the same tables, byte for byte
        single = torch.zeros((n3, MAX_TABLE, 8), dtype=torch.float64, device=dev)
        for first in range(0, n3, n_ctx):
            ids = list(range(first, min(first + n_ctx, n3)))
            vs = [vols3[mine3.index(i)] if i in mine3 else generate_phantom_cuda(phantom.config_c3(i), device=local) for i in ids]
            run_c3(vs, single[first:first + len(ids)])
            del vs
        torch.cuda.synchronize()
        same = bool(np.array_equal(single.cpu().numpy(), gathered))
        configs["C3"] = {"what": f"{n3} noisy 512x512x256 scans (sigma 15), scan i on rank i % {world}; one all-gather of the device-written tables",
                         "n_scans": n3, "ms_per_batch": ms3, "scans_per_s": n3 / (ms3 * 1e-3), "gvoxel_per_s": n3 * n_vox / ms3 / 1e6,
                         "frac_of_hbm_peak_per_gpu": 7.0 * n3 * n_vox / ms3 / 1e6 / peak / world,
                         "gathered_equals_single_gpu": same,
                         "markers_per_scan": [int((gathered[i, :, 0] > 0).sum()) for i in range(min(n3, 8))]}
    del vols3
    torch.cuda.empty_cache()

    # C5: 1,048,576 skin-surface candidates in contiguous blocks over the ranks, global arg-min of the gathered winners
    N5 = 1 << 20
    pts, nrm, tgt = phantom.surface_candidates(N5)
    per5 = N5 // world
    lo5 = rank * per5
    hi5 = N5 if rank == world - 1 else lo5 + per5
    p_d, n_d = torch.from_numpy(pts[lo5:hi5]).to(dev), torch.from_numpy(nrm[lo5:hi5]).to(dev)
    det5 = FiducialDetector((64, 64, 64), device=local)
    ph2 = phantom.config_c2()
    sp2, org2 = np.array(ph2.spacing), np.array(ph2.origin)
    body5 = torch.zeros((nz, ny, nx), dtype=torch.uint8, device=dev)      # path mask: 1 = free (inside the body), replicated per rank
    e = ph2.ellipsoids[0]
    zz = torch.arange(nz, device=dev, dtype=torch.float32).view(-1, 1, 1)
    yy = torch.arange(ny, device=dev, dtype=torch.float32).view(1, -1, 1)
    xx = torch.arange(nx, device=dev, dtype=torch.float32).view(1, 1, -1)
    body5[((xx - float(e[0])) / float(e[3])) ** 2 + ((yy - float(e[1])) / float(e[4])) ** 2 + ((zz - float(e[2])) / float(e[5])) ** 2 <= 1.0] = 1
    # RAS mm -> voxel index for identity direction: lps = -ras(x,y), index = (lps - origin) / spacing
    m5 = np.array([[-1 / sp2[0], 0, 0, -org2[0] / sp2[0]], [0, -1 / sp2[1], 0, -org2[1] / sp2[1]], [0, 0, 1 / sp2[2], -org2[2] / sp2[2]]])
    c5 = {}
    for samples in (0, 64):
        win = [None]

        def search():
            r = det5.entry_search(p_d, n_d, tgt, n_path_samples=samples, path_mask=body5 if samples else None,
                                  ras_to_index=m5 if samples else None, path_free_value=1)
            win[0] = gather_entry_results(r["index"], r["distance"], lo5, dev)
        for _ in range(3):
            search()
        barrier()
        e0.record()
        for _ in range(10):
            search()
        e1.record()
        barrier()
        ms5 = all_max(e0.elapsed_time(e1)) / 10
        entry = {"ms": ms5, "mcandidates_per_s": N5 / ms5 / 1e3, "winner": win[0][0], "distance": win[0][1]}
        if rank == 0:
            if world > 1:                               # the same search on one GPU
                pa, na = torch.from_numpy(pts).to(dev), torch.from_numpy(nrm).to(dev)
                ra = det5.entry_search(pa, na, tgt, n_path_samples=samples, path_mask=body5 if samples else None,
                                       ras_to_index=m5 if samples else None, path_free_value=1)
                entry["sharded_equals_single_gpu"] = bool(ra["index"] == win[0][0] and ra["distance"] == win[0][1])
                del pa, na
            if samples == 0 and not args.no_cpu_baseline:
                from oracle import kinematics as kin
                wi, wd = kin.find_entry_point(pts, nrm, tgt)
                entry["equals_oracle"] = bool(wi == win[0][0] and wd == win[0][1])
            c5[f"path_samples_{samples}"] = entry
    if rank == 0:
        configs["C5"] = {"what": f"closest suitable entry point among {N5} skin-surface candidates, contiguous blocks over {world} rank(s), "
                                 "one all-gather of the (distance, index) winners; result copy + host sync included", **c5}
    det5.close()
    del body5, p_d, n_d

    # ---------------- CPU baseline (oracle port) on a bounded sample of the batch's scans
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import c_oracle
        c_oracle.use_all_cores()
        n_cpu = min(S, 4)                        # bounded sample: a few of the batch's scans, ~2-10 s of CPU work
        hosts = [v.cpu().numpy() for v in vols[:n_cpu]]
        c_oracle.run_pipeline(hosts[0])          # warm-up (page faults, OpenMP team start)
        dt_total = 0.0
        for h in hosts:
            dt_total += c_oracle.run_pipeline(h)[4]
        cpu = {"value": n_cpu * n_vox / dt_total / 1e9, "unit": "Gvoxel/s", "cores": c_oracle.num_threads(),
               "kind": "port", "sample": f"{n_cpu} of the batch's {nx}x{ny}x{nz} scans, once each after one warm-up, "
               "through oracle/c (threshold, closing, CCL, label sums; OpenMP over all host threads); SimpleITK "
               "itself is not installable offline", "seconds": dt_total}

    if rank == 0:
        rank_ms_step = [m / args.steps for m in rank_ms]
        line = {"metric": METRIC, "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                "scans_per_s": world * S * args.steps / (ms_max * 1e-3),
                "config": {"workload": f"C2: {nx}x{ny}x{nz} uint16 phantom (baseplate + end-effector fiducials, Rician "
                                       f"sigma 10), {S} distinct scans per GPU per step at every N (weak scaling of the C2 workload; "
                                       "BASELINE's other configs, C3's sigma-15 64-scan batch included, are in `configs`)",
                           "scans_per_gpu": S, "outputs": "u8 closed mask + u32 label volume materialised per scan; "
                                                          "marker table to host",
                           "l2": f"no flush: each step streams {S} x 128 MiB of distinct inputs per GPU (> 126 MB L2)",
                           "parallelism": f"scan-sharded x{world}, " + ("no collective (MAMRI_BENCH_NO_GATHER=1)" if no_gather else
                                          "one NCCL all-gather of the device-written marker tables per step, queued behind the scans"),
                           "pipelining": "steps are software-pipelined over two pools of contexts: batch k+1 is enqueued before the "
                                         "results of batch k are collected; all K batches complete inside the timed region"},
                "rank_ms_per_step": {"min": min(rank_ms_step), "median": statistics.median(rank_ms_step), "max": max(rank_ms_step)},
                "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "stages_ms": stages,
                "gpu_launches": args.steps * S * bd.kernel_launches_per_scan, "kernels_per_scan": bd.kernel_launches_per_scan,
                "parity": parity, "configs": configs}
        if gathered_ok is not None:
            line["gathered_tables_equal_host_packed"] = gathered_ok
        if numa is not None:
            line["config"]["cpu_affinity_rank0"] = numa
        print(json.dumps(line), flush=True)
    bp.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--scans-per-gpu", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-c4", action="store_true", help="leave the 1024x1024x512 config out of the `configs` block")
    ap.add_argument("--c3-scans", type=int, default=64, help="scans of the sharded sigma-15 batch (config C3)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 5
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 100     # ~0.1 s timed: enough NVML clock samples
        args.warmup = args.warmup if args.warmup is not None else 3
        run_native(args)


if __name__ == "__main__":
    main()
