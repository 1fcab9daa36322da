"""Generates the golden fixtures under tests/golden/ from the NumPy oracle (cross-checked against the
brute-force and C restatements while generating).  The reference ships no golden vectors
 and SimpleITK cannot run here, so these pin OUR oracle against regressions and give
the GPU tests fixed inputs; they are not outputs of the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mamri_pose_estimation_b200 import phantom      # noqa: E402
from oracle import bruteforce as bf                 # noqa: E402
from oracle import c_oracle                         # noqa: E402
from oracle import segmentation as seg              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    dict(name="g1_phantom_48x40x32_c6", dims=(48, 40, 32), seed=21, conn=6, radius=2, flip=False, border=False),
    dict(name="g2_phantom_37x29x23_c26_flip", dims=(37, 29, 23), seed=22, conn=26, radius=2, flip=True, border=True),
    dict(name="g3_phantom_64x24x20_r1", dims=(64, 24, 20), seed=23, conn=6, radius=1, flip=False, border=True),
]


def main():
    for c in CASES:
        ph = phantom.small_phantom(dims=c["dims"], seed=c["seed"], flip_lps=c["flip"], touch_border=c["border"])
        vol = phantom.generate(ph)
        geom = seg.Geometry(ph.spacing, ph.origin, ph.direction)
        det = seg.detect_fiducials(vol, geom, close_radius=c["radius"], connectivity=c["conn"], min_vol=20.0, max_vol=600.0)
        # cross-checks: definitional restatement and C restatement
        assert np.array_equal(det.closed, bf.closing_itk_pipeline(seg.binary_threshold(vol), c["radius"]))
        lab_bf, k_bf = bf.flood_fill_labels(det.closed, c["conn"])
        assert k_bf == det.n_labels and np.array_equal(lab_bf, det.labels)
        cd = c_oracle.detect_fiducials(vol, geom, close_radius=c["radius"], connectivity=c["conn"], min_vol=20.0, max_vol=600.0)
        assert np.array_equal(cd.closed, det.closed) and np.array_equal(cd.labels, det.labels) and cd.fiducials == det.fiducials
        assert det.n_labels < 65536
        by = {s.label: s for s in det.stats}
        markers = np.array([[f["id"], by[f["id"]].count, f["vol"], *f["centroid"], *by[f["id"]].sum_idx, *by[f["id"]].sum_mom]
                            for f in det.fiducials], dtype=np.float64).reshape(-1, 15)
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), volume=vol, closed_bits=np.packbits(det.closed),
                            labels=det.labels.astype(np.uint16), counts=det.counts.astype(np.int64), markers=markers,
                            body_label=np.int64(det.body_label), spacing=np.array(ph.spacing), origin=np.array(ph.origin),
                            direction=np.array(ph.direction), conn=np.int64(c["conn"]), radius=np.int64(c["radius"]),
                            min_vol=np.float64(20.0), max_vol=np.float64(600.0))
        print(c["name"], det.n_labels, len(det.fiducials), det.body_label, os.path.getsize(os.path.join(HERE, c["name"] + ".npz")))


if __name__ == "__main__":
    main()
