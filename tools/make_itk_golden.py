#!/usr/bin/env python
"""ITK pinning kit: golden vectors from the REAL SimpleITK calls of the reference.

The oracle (oracle/segmentation.py) restates what MamriLogic.volume_threshold_segmentation asks SimpleITK to do
(Mamri/Mamri.py:1308-1310, 1318-1323).  SimpleITK cannot be installed in the build container, so the restatement has
never been compared with ITK itself ("parity unpinned").  This script closes that gap wherever SimpleITK exists (any
3D Slicer Python console, or `pip install SimpleITK numpy`):

    python tools/make_itk_golden.py                 # writes tests/golden/itk/*.npz   (needs SimpleITK)
    python -m pytest tests/test_itk_golden.py       # oracle vs ITK on CPU; with a B200 also the CUDA path vs ITK

It runs, on seeded phantoms that need nothing but NumPy to regenerate, exactly the calls of the reference

    binary   = sitk.BinaryMorphologicalClosing(sitk.BinaryThreshold(img, lo, hi), [r]*3, sitk.sitkBall)   # :1308
    labeled  = sitk.ConnectedComponent(binary)                                                            # :1309
    stats    = sitk.LabelShapeStatisticsImageFilter(); stats.Execute(labeled)                             # :1309
    kept     = [lbl for lbl in stats.GetLabels() if lo_v <= stats.GetPhysicalSize(lbl) <= hi_v]           # :1310
    body     = max(other labels, key=stats.GetPhysicalSize)                                               # :1320-1322

and stores every intermediate (threshold mask, closed mask, label volume, per-label pixel count / physical size /
centroid / principal moments / principal axes, kept ids, body id) plus the structuring element ITK really used
(dilation of a single voxel) and the SimpleITK version.  The cases cover what the parity tests cover: face and full
connectivity, radii 1-3, objects touching the border (safe-border rule), ragged row lengths, LPS-flipped direction,
int16 / float32 voxels, and binary opening (the north_star's extension; sitk.BinaryMorphologicalOpening).

`--backend oracle` writes the same files from the oracle instead (used by the test-suite to exercise this kit's
plumbing where SimpleITK is missing; such files are marked backend="oracle" and prove nothing about ITK).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mamri_pose_estimation_b200 import phantom      # noqa: E402  (NumPy only)

DEFAULT_OUT = os.path.join(ROOT, "tests", "golden", "itk")

# name, dims (x, y, z), seed, connectivity, close radius, open radius, LPS flip, border-touching object, dtype, lower
CASES = [
    dict(name="i1_c6_r2", dims=(48, 40, 32), seed=21, conn=6, radius=2, open_radius=0, flip=False, border=False, dtype="uint16", lo=65.0),
    dict(name="i2_c26_r2_flip_border_ragged", dims=(37, 29, 23), seed=22, conn=26, radius=2, open_radius=0, flip=True, border=True, dtype="uint16", lo=65.0),
    dict(name="i3_c6_r1_border", dims=(64, 24, 20), seed=23, conn=6, radius=1, open_radius=0, flip=False, border=True, dtype="uint16", lo=65.0),
    dict(name="i4_c6_r3_border_ragged", dims=(45, 33, 27), seed=24, conn=6, radius=3, open_radius=0, flip=False, border=True, dtype="uint16", lo=65.0),
    dict(name="i5_c26_r1_int16", dims=(40, 36, 28), seed=25, conn=26, radius=1, open_radius=0, flip=False, border=True, dtype="int16", lo=65.0),
    dict(name="i6_c6_r2_float32_fractional", dims=(33, 31, 29), seed=26, conn=6, radius=2, open_radius=0, flip=True, border=False, dtype="float32", lo=64.5),
    dict(name="i7_c6_r0", dims=(32, 32, 24), seed=27, conn=6, radius=0, open_radius=0, flip=False, border=True, dtype="uint16", lo=65.0),
    dict(name="i8_open1_close2_border", dims=(48, 40, 32), seed=28, conn=6, radius=2, open_radius=1, flip=False, border=True, dtype="uint16", lo=65.0),
    dict(name="i9_open2_close0_ragged", dims=(37, 29, 23), seed=29, conn=26, radius=0, open_radius=2, flip=False, border=True, dtype="uint16", lo=65.0),
    dict(name="i10_noisy_c6_r2", dims=(64, 48, 40), seed=30, conn=6, radius=2, open_radius=0, flip=False, border=True, dtype="uint16", lo=65.0, sigma=22.0),
]
MIN_VOL, MAX_VOL, UPPER = 20.0, 600.0, 65535.0


def make_volume(c):
    ph = phantom.small_phantom(dims=c["dims"], seed=c["seed"], flip_lps=c["flip"], touch_border=c["border"],
                               sigma=c.get("sigma", 12.0))
    vol = phantom.generate(ph)
    if c["dtype"] == "int16":
        vol = np.minimum(vol, 32767).astype(np.int16)
    elif c["dtype"] == "float32":
        vol = vol.astype(np.float32) + np.float32(0.25)
    return ph, np.ascontiguousarray(vol)


def upper_for(dtype):
    # the reference passes 65535 (Mamri.py:1308); for int16 volumes that is not representable and ITK rejects
    # lower > upper after the cast, so the kit (like the library) uses the type's maximum there -- stated in DESIGN.md
    return 32767.0 if np.dtype(dtype) == np.int16 else UPPER


def run_sitk(c, ph, vol):
    import SimpleITK as sitk
    img = sitk.GetImageFromArray(vol)                       # [z, y, x] -> ITK x-fastest
    img.SetSpacing(tuple(float(v) for v in ph.spacing))
    img.SetOrigin(tuple(float(v) for v in ph.origin))
    img.SetDirection(tuple(float(v) for v in ph.direction))
    binary0 = sitk.BinaryThreshold(img, c["lo"], upper_for(vol.dtype))
    binary = binary0
    if c["open_radius"] > 0:
        binary = sitk.BinaryMorphologicalOpening(binary, [c["open_radius"]] * 3, sitk.sitkBall)
    opened = binary
    if c["radius"] > 0:
        binary = sitk.BinaryMorphologicalClosing(binary, [c["radius"]] * 3, sitk.sitkBall)      # Mamri.py:1308
    labeled = sitk.ConnectedComponent(binary, c["conn"] == 26)                                   # :1309 (default False)
    stats = sitk.LabelShapeStatisticsImageFilter()
    stats.Execute(labeled)
    labels = list(stats.GetLabels())
    rec = dict(threshold=sitk.GetArrayFromImage(binary0).astype(np.uint8), opened=sitk.GetArrayFromImage(opened).astype(np.uint8),
               closed=sitk.GetArrayFromImage(binary).astype(np.uint8), labels=sitk.GetArrayFromImage(labeled).astype(np.uint32),
               label_ids=np.array(labels, dtype=np.int64),
               counts=np.array([stats.GetNumberOfPixels(l) for l in labels], dtype=np.int64),
               physical_size=np.array([stats.GetPhysicalSize(l) for l in labels], dtype=np.float64),
               centroid=np.array([stats.GetCentroid(l) for l in labels], dtype=np.float64).reshape(-1, 3),
               principal_moments=np.array([stats.GetPrincipalMoments(l) for l in labels], dtype=np.float64).reshape(-1, 3),
               principal_axes=np.array([stats.GetPrincipalAxes(l) for l in labels], dtype=np.float64).reshape(-1, 9))
    # the structuring element ITK really used: dilation of one voxel
    for r in (1, 2, 3):
        one = sitk.Image(4 * r + 1, 4 * r + 1, 4 * r + 1, sitk.sitkUInt8)
        one[2 * r, 2 * r, 2 * r] = 1
        rec[f"ball{r}"] = sitk.GetArrayFromImage(sitk.BinaryDilate(one, [r] * 3, sitk.sitkBall)).astype(np.uint8)
    rec["backend"] = np.array("SimpleITK " + sitk.Version().VersionString())
    return rec


def run_oracle(c, ph, vol):
    from oracle import segmentation as seg
    geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
    binary0 = seg.binary_threshold(vol, c["lo"], upper_for(vol.dtype))
    opened = seg.binary_opening(binary0, c["open_radius"])
    closed = seg.binary_closing_safe_border(opened, c["radius"])
    labels, k = seg.connected_components(closed, c["conn"])
    st = seg.label_shape_statistics(labels, k, geom)
    rec = dict(threshold=binary0, opened=opened, closed=closed, labels=labels, label_ids=np.arange(1, k + 1, dtype=np.int64),
               counts=np.array([s.count for s in st], dtype=np.int64),
               physical_size=np.array([s.physical_size for s in st], dtype=np.float64),
               centroid=np.array([s.centroid for s in st], dtype=np.float64).reshape(-1, 3),
               principal_moments=np.array([s.principal_moments for s in st], dtype=np.float64).reshape(-1, 3),
               principal_axes=np.array([s.principal_axes.ravel() for s in st], dtype=np.float64).reshape(-1, 9))
    for r in (1, 2, 3):
        b = np.zeros((4 * r + 1,) * 3, dtype=np.uint8)
        b[r:3 * r + 1, r:3 * r + 1, r:3 * r + 1] = seg.ball_structure(r)
        rec[f"ball{r}"] = b
    rec["backend"] = np.array("oracle")
    return rec


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--out", default=DEFAULT_OUT)
    ap.add_argument("--backend", choices=["sitk", "oracle"], default="sitk")
    ap.add_argument("--only", default=None, help="comma-separated case names")
    a = ap.parse_args(argv)
    if a.backend == "sitk":
        try:
            import SimpleITK  # noqa: F401
        except ImportError:
            raise SystemExit("SimpleITK is not importable here.  Run this script where it is (3D Slicer's Python console: "
                             "exec(open('tools/make_itk_golden.py').read()), or `pip install SimpleITK`), then commit "
                             "tests/golden/itk/*.npz.")
    os.makedirs(a.out, exist_ok=True)
    only = set(a.only.split(",")) if a.only else None
    for c in CASES:
        if only and c["name"] not in only:
            continue
        ph, vol = make_volume(c)
        rec = run_sitk(c, ph, vol) if a.backend == "sitk" else run_oracle(c, ph, vol)
        # the filter of Mamri.py:1310 and the body choice of :1320-1322, on ITK's own numbers
        size = rec["physical_size"]
        keep = (size >= MIN_VOL) & (size <= MAX_VOL)
        kept = rec["label_ids"][keep]
        rest = rec["label_ids"][~keep]
        body = int(rest[np.argmax(size[~keep])]) if rest.size else 0          # max() returns the first maximum
        path = os.path.join(a.out, c["name"] + ".npz")
        np.savez_compressed(path, volume=vol, spacing=np.array(ph.spacing, dtype=np.float64), origin=np.array(ph.origin, dtype=np.float64),
                            direction=np.array(ph.direction, dtype=np.float64), lo=np.float64(c["lo"]), hi=np.float64(upper_for(vol.dtype)),
                            conn=np.int64(c["conn"]), radius=np.int64(c["radius"]), open_radius=np.int64(c["open_radius"]),
                            min_vol=np.float64(MIN_VOL), max_vol=np.float64(MAX_VOL), kept=kept, body_label=np.int64(body),
                            threshold_bits=np.packbits(rec.pop("threshold")), opened_bits=np.packbits(rec.pop("opened")),
                            closed_bits=np.packbits(rec.pop("closed")), **rec)
        print(f"{c['name']}: {len(rec['label_ids'])} labels, {len(kept)} kept, body {body}, {os.path.getsize(path)} bytes "
              f"[{str(rec['backend'])}]")


if __name__ == "__main__":
    main()
