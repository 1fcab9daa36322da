"""Single-context, single-stream latency of one device-resident scan through the captured graph
(no profiling events): the per-scan time a caller sees when scans arrive one at a time.

Two loops: through the Python mirror (FiducialDetector.detect_async + collect, result objects built), and
through the bare C ABI (mamri_detect_async + mamri_detect_collect with preallocated structs), which is
what a C caller of include/mamri_b200.h pays."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mamri_pose_estimation_b200 import _capi, phantom
from mamri_pose_estimation_b200.detector import DetectParams, FiducialDetector, generate_phantom_cuda, _desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--conn", type=int, default=6)
    a = ap.parse_args()
    mk = {"c1": lambda i: phantom.config_c1(), "c2": lambda i: phantom.config_c2(scan_index=i),
          "c3": lambda i: phantom.config_c3(i), "c4": lambda i: phantom.config_c4()}[a.config]
    ph = mk(0)
    nx, ny, nz = ph.dims
    vols = [generate_phantom_cuda(mk(i)) for i in range(1 if a.config in ("c1", "c4") else 4)]
    det = FiducialDetector(ph.dims, max_runs=(nx * ny * nz) // 8)
    mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device="cuda")
    lab = torch.empty((nz, ny, nx), dtype=torch.int32, device="cuda")
    prm = DetectParams(connectivity=a.conn)
    kw = dict(spacing=ph.spacing, origin=ph.origin, direction=ph.direction, params=prm, out_mask=mask, out_labels=lab)
    for i in range(5):
        r = det.detect(vols[i % len(vols)], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.reps):
        det.detect_async(vols[i % len(vols)], **kw)
        r = det.collect()
    e1.record()
    torch.cuda.synchronize()
    ms_py = e0.elapsed_time(e1) / a.reps
    # bare C ABI
    lib = det._lib
    d = _desc((nz, ny, nx), "uint16", ph.spacing, ph.origin, ph.direction)
    p = prm.to_c()
    summ = _capi.Summary()
    markers = (_capi.Marker * det.max_markers)()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptrs = [C.c_void_p(v.data_ptr()) for v in vols]
    pm, pl = C.c_void_p(mask.data_ptr()), C.c_void_p(lab.data_ptr())
    e0.record()
    for i in range(a.reps):
        rc = lib.mamri_detect_async(det._ctx, C.byref(d), ptrs[i % len(ptrs)], C.byref(p), pm, pl, None, stream)
        rc |= lib.mamri_detect_collect(det._ctx, C.byref(summ), markers, det.max_markers)
        assert rc == 0
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / a.reps
    nv = nx * ny * nz
    print(f"{a.config} conn {a.conn}: {det.kernel_launches} kernels/scan; labels {r.n_labels} runs {r.n_runs} markers {summ.n_markers}; "
          f"serial scan (graph, incl. collect sync): python mirror {ms_py*1e3:.1f} us = {nv/ms_py/1e6:.1f} Gvox/s; "
          f"bare C ABI {ms_c*1e3:.1f} us = {nv/ms_c/1e6:.1f} Gvox/s = {7*nv/ms_c/1e6:.0f} GB/s of 7 B/voxel")
    det.close()


if __name__ == "__main__":
    main()
