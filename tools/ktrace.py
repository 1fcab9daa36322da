"""In-pipeline kernel timeline of one lone scan (trace build of the library).

    python -m mamri_pose_estimation_b200.build --trace
    MAMRI_LIB=mamri_pose_estimation_b200/libmamri_b200_trace.so python tools/ktrace.py --config c2

Every kernel of the trace build stamps the GPU's nanosecond timer when its first CTA gets past the dependency on the
kernel before it (csrc/common.cuh: ktrace), and at a few points inside the long ones.  The scan runs through the
captured graph exactly as in production (no events between the kernels), so the differences between consecutive
stamps are what each kernel really costs in the pipeline, launch gap included -- which ncu cannot show.
"""
import argparse
import ctypes as C
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MAMRI_LIB", os.path.join(ROOT, "mamri_pose_estimation_b200", "libmamri_b200_trace.so"))

import torch  # noqa: E402
from mamri_pose_estimation_b200 import _capi, phantom  # noqa: E402
from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda  # noqa: E402

# slot order of csrc/common.cuh: enum KId
NAMES = ["threshold", "close", "erode", "runs_scan", "union_slices", "union_z1", "union_z2", "flatten_rank", "select",
         "label_cluster", "stats", "final", "materialise", "end", "label.U1", "label.U2", "label.F", "label.FIX", "label.S",
         "label.end", "runs.lookback", "runs.write", "close.loaded", "close.dilated", "close.eroded", "stats.finalise",
         "close.lastCTA", "runs.lastCTA", "threshold.lastCTA",
         "uslice.n", "uslice.n.last", "uslice.join", "uslice.join.last", "uslice.flat", "uslice.flat.last", "uslice.end", "uslice.end.last",
         "uz1.end", "uz1.end.last", "uz2.end", "uz2.end.last", "rank.end", "rank.end.last", "select.end", "select.end.last",
         "runs.lookback.last"]
LAST = {11, 13, 26, 27, 28, 30, 32, 34, 36, 38, 40, 42, 44, 45}      # slots written by ktrace_last (complemented)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--conn", type=int, default=6)
    a = ap.parse_args()
    mk = {"c1": phantom.config_c1, "c2": phantom.config_c2, "c3": lambda: phantom.config_c3(0), "c4": phantom.config_c4}[a.config]
    ph = mk()
    nx, ny, nz = ph.dims
    vol = generate_phantom_cuda(ph)
    lib = _capi.load()
    det = FiducialDetector(ph.dims, max_runs=(nx * ny * nz) // 8)
    mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device="cuda")
    lab = torch.empty((nz, ny, nx), dtype=torch.int32, device="cuda")
    kw = dict(spacing=ph.spacing, origin=ph.origin, direction=ph.direction, params=DetectParams(connectivity=a.conn),
              out_mask=mask, out_labels=lab)
    for _ in range(4):                                   # warm-up: sizes the grids, captures the graph
        r = det.detect(vol, **kw)
    buf = (C.c_uint64 * 48)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows, totals = [], []
    for _ in range(a.reps):
        if lib.mamri_ktrace_reset() != 0:
            raise SystemExit("this is not the trace build: python -m mamri_pose_estimation_b200.build --trace")
        torch.cuda.synchronize()
        e0.record()
        det.detect_async(vol, **kw)
        r = det.collect()
        e1.record()
        torch.cuda.synchronize()
        totals.append(e0.elapsed_time(e1) * 1e3)
        lib.mamri_ktrace_read(buf, 48)
        row = [int(v) for v in buf]
        for j in LAST:                               # KT_FINAL / KT_END hold the complement of the LAST stamp (ktrace_last)
            if row[j] != (1 << 64) - 1:
                row[j] = ~row[j] & ((1 << 64) - 1)
        rows.append(row)
    print(f"{a.config} conn {a.conn}: {r.n_runs} runs, {r.n_labels} labels, {len(r.markers)} markers, "
          f"{det.kernel_launches} kernels per scan; event-timed scan (enqueue .. collect): median {statistics.median(totals):.1f} us, "
          f"best {min(totals):.1f} us")
    none = (1 << 64) - 1
    stamped = [i for i in range(len(NAMES)) if all(row[i] != none for row in rows)]
    t0 = [row[0] for row in rows]
    order = sorted(stamped, key=lambda i: statistics.median(row[i] - b for row, b in zip(rows, t0)))
    prev = None
    print(f"{'stamp':>16} {'at us (median)':>15} {'since previous':>15}")
    for i in order:
        at = statistics.median((row[i] - b) / 1e3 for row, b in zip(rows, t0))
        print(f"{NAMES[i]:>16} {at:15.2f} {'' if prev is None else format(at - prev, '15.2f')}")
        prev = at
    det.close()


if __name__ == "__main__":
    main()
