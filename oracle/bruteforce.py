"""Second, independent restatement from the set-theoretic definitions, for small
volumes only (<= ~48^3).  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

It shares no code with ``oracle.segmentation`` beyond the ball definition being
re-derived here, and follows ITK's own pipeline literally (pad by r, dilate with
outside = background, erode with outside = foreground, crop by r;
itk::BinaryMorphologicalClosingImageFilter, called at Mamri/Mamri.py:1308), a
flood-fill labelling (itk::ConnectedComponentImageFilter semantics, :1309) and
direct per-voxel moment sums in physical space (ShapeLabelMapFilter's
commented-out "basic implementation", :1309).
"""
from __future__ import annotations

from collections import deque

import numpy as np


def ball(radius: int):
    """Voxel (dz,dy,dx) is in ITK's ball iff the centre of that voxel, taken at
    index+0.5, lies in the ellipsoid centred at r+0.5 with semi-axes r+0.5."""
    r = int(radius)
    offs = []
    for iz in range(2 * r + 1):
        for iy in range(2 * r + 1):
            for ix in range(2 * r + 1):
                q = sum(((i + 0.5) - (r + 0.5)) ** 2 / (0.5 * (2 * r + 1)) ** 2 for i in (ix, iy, iz))
                if q <= 1.0:
                    offs.append((iz - r, iy - r, ix - r))
    return offs


def _shifted(a: np.ndarray, dz: int, dy: int, dx: int, fill: bool) -> np.ndarray:
    """out[p] = a[p + d], with ``fill`` where p + d is outside."""
    out = np.full(a.shape, fill, dtype=bool)
    nz, ny, nx = a.shape
    zs = slice(max(0, -dz), min(nz, nz - dz)); zt = slice(max(0, dz), min(nz, nz + dz))
    ys = slice(max(0, -dy), min(ny, ny - dy)); yt = slice(max(0, dy), min(ny, ny + dy))
    xs = slice(max(0, -dx), min(nx, nx - dx)); xt = slice(max(0, dx), min(nx, nx + dx))
    out[zs, ys, xs] = a[zt, yt, xt]
    return out


def dilate(a: np.ndarray, offs, outside: bool = False) -> np.ndarray:
    out = np.zeros(a.shape, dtype=bool)
    for d in offs:
        out |= _shifted(a, *d, fill=outside)
    return out


def erode(a: np.ndarray, offs, outside: bool = True) -> np.ndarray:
    out = np.ones(a.shape, dtype=bool)
    for d in offs:
        out &= _shifted(a, *d, fill=outside)
    return out


def closing_itk_pipeline(mask: np.ndarray, radius: int) -> np.ndarray:
    r = int(radius)
    if r == 0:
        return mask.astype(np.uint8).copy()
    b = ball(r)
    p = np.pad(mask.astype(bool), r, mode="constant", constant_values=False)
    d = dilate(p, b, outside=False)
    e = erode(d, b, outside=True)
    return e[r:-r, r:-r, r:-r].astype(np.uint8)


def opening_itk_pipeline(mask: np.ndarray, radius: int) -> np.ndarray:
    """itk::BinaryMorphologicalOpeningImageFilter literally: erode with the outside
    of the image counted as foreground (BinaryErodeImageFilter's BoundaryToForeground
    = true), then dilate with the outside as background; no padding."""
    r = int(radius)
    if r == 0:
        return mask.astype(np.uint8).copy()
    b = ball(r)
    e = erode(mask.astype(bool), b, outside=True)
    return dilate(e, b, outside=False).astype(np.uint8)


def flood_fill_labels(mask: np.ndarray, connectivity: int = 6):
    """Raster scan; every unlabelled foreground voxel starts the next label and
    is flooded -- labels are consecutive in order of first voxel by construction."""
    nz, ny, nx = mask.shape
    if connectivity == 6:
        nb = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    elif connectivity == 26:
        nb = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1) if (a, b, c) != (0, 0, 0)]
    else:
        raise ValueError(connectivity)
    lab = np.zeros(mask.shape, dtype=np.uint32)
    k = 0
    fg = mask != 0
    for z, y, x in zip(*np.nonzero(fg)):      # np.nonzero yields raster (C) order
        if lab[z, y, x]:
            continue
        k += 1
        lab[z, y, x] = k
        q = deque([(z, y, x)])
        while q:
            cz, cy, cx = q.popleft()
            for dz, dy, dx in nb:
                pz, py, px = cz + dz, cy + dy, cx + dx
                if 0 <= pz < nz and 0 <= py < ny and 0 <= px < nx and fg[pz, py, px] and not lab[pz, py, px]:
                    lab[pz, py, px] = k
                    q.append((pz, py, px))
    return lab, k


def shape_stats_direct(labels: np.ndarray, k: int, spacing, origin, direction):
    """count, physical size, centroid and second central moments straight from
    per-voxel physical points (no index-space shortcut)."""
    d = np.asarray(direction, dtype=np.float64).reshape(3, 3)
    s = np.asarray(spacing, dtype=np.float64)
    o = np.asarray(origin, dtype=np.float64)
    out = []
    for l in range(1, k + 1):
        z, y, x = np.nonzero(labels == l)
        idx = np.stack([x, y, z], axis=1).astype(np.float64)
        n = idx.shape[0]
        cidx = idx.mean(axis=0)
        centroid = o + d @ (s * cidx)
        pts = o[None, :] + (idx * s[None, :]) @ d.T
        m = pts.T @ pts / n - np.outer(centroid, centroid)
        m[np.diag_indices(3)] += s * s / 12.0
        w, v = np.linalg.eigh(m)
        axes = v.T.copy()
        axes[2] *= np.linalg.det(axes)
        out.append(dict(label=l, count=n, physical_size=n * float(np.prod(s)), centroid_index=cidx,
                        centroid=centroid, principal_moments=w, principal_axes=axes))
    return out
