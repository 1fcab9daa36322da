"""Oracle for the skin-surface candidate stage.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference derives its
entry-point candidates from Slicer's closed-surface representation of
"AutoBodySegmentation" (``Mamri/Mamri.py:1338-1339``, ``_get_body_polydata``
at ``:994``) and ``vtkPolyDataNormals`` (``:997-1003``).  Flying edges,
smoothing and decimation are Slicer internals that exist nowhere outside it,
so NO PARITY with the reference is claimed for this stage: the candidate set is
*defined* here on the voxel grid (the same definition ``surface.cu`` implements)
and validated geometrically against analytic shapes in the tests.

Definition
  candidate  body voxel with at least one of its 6 face neighbours outside the
             body (outside the volume counts as outside); ascending linear index
  point      physical centre of the voxel, LPS -> RAS ``[-x, -y, z]``
             (``Mamri.py:1317`` convention), rounded to float32 (vtkPoints)
  normal     g = sum over the ITK radius-2 ball (d.d <= 6, 81 voxels) of
             d * body(p + d) points into the body; n_index = -g;
             n_lps = Direction @ (n_index / spacing); RAS flip; unit length in
             float64; rounded to float32; zero vector where g vanishes.
"""
from __future__ import annotations

import numpy as np

from .segmentation import Geometry, ball_offsets


def surface_voxels(body: np.ndarray) -> np.ndarray:
    """bool [nz,ny,nx]: body voxels with a face neighbour that is not body."""
    b = np.asarray(body) != 0
    p = np.pad(b, 1, mode="constant", constant_values=False)
    interior = (p[1:-1, 1:-1, :-2] & p[1:-1, 1:-1, 2:] & p[1:-1, :-2, 1:-1] & p[1:-1, 2:, 1:-1]
                & p[:-2, 1:-1, 1:-1] & p[2:, 1:-1, 1:-1])
    return b & ~interior


def ball_moments(body: np.ndarray, zyx: np.ndarray) -> np.ndarray:
    """Integer first moment (gx, gy, gz) of the body inside the radius-2 ball around each voxel of `zyx` [n,3]."""
    b = np.asarray(body) != 0
    p = np.pad(b, 2, mode="constant", constant_values=False)
    g = np.zeros((len(zyx), 3), dtype=np.int64)
    z, y, x = zyx[:, 0] + 2, zyx[:, 1] + 2, zyx[:, 2] + 2
    for dz, dy, dx in ball_offsets(2):
        m = p[z + dz, y + dy, x + dx].astype(np.int64)
        g[:, 0] += dx * m
        g[:, 1] += dy * m
        g[:, 2] += dz * m
    return g


def body_surface(body: np.ndarray, geom: Geometry):
    """(points float32 [n,3] RAS, normals float32 [n,3] RAS, linear_index int64 [n]) of the body labelmap
    `body[z,y,x]`; operation order is the one `surface.cu` uses (separately rounded float64 operations)."""
    s = surface_voxels(body)
    zyx = np.argwhere(s)                                  # C order = ascending linear index
    nz, ny, nx = s.shape
    lin = zyx[:, 2] + nx * (zyx[:, 1] + ny * zyx[:, 0])
    sp = np.asarray(geom.spacing, dtype=np.float64)
    org = np.asarray(geom.origin, dtype=np.float64)
    D = np.asarray(geom.direction, dtype=np.float64).reshape(3, 3)
    ix, iy, iz = sp[0] * zyx[:, 2].astype(np.float64), sp[1] * zyx[:, 1].astype(np.float64), sp[2] * zyx[:, 0].astype(np.float64)
    lps = np.stack([org[k] + ((D[k, 0] * ix + D[k, 1] * iy) + D[k, 2] * iz) for k in range(3)], axis=1)
    pts = np.stack([-lps[:, 0], -lps[:, 1], lps[:, 2]], axis=1).astype(np.float32)
    g = ball_moments(body, zyx).astype(np.float64)
    qx, qy, qz = -g[:, 0] / sp[0], -g[:, 1] / sp[1], -g[:, 2] / sp[2]
    v = np.stack([(D[k, 0] * qx + D[k, 1] * qy) + D[k, 2] * qz for k in range(3)], axis=1)
    ln = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    ok = ln > 0.0
    n = np.zeros_like(v)
    n[ok] = v[ok] / ln[ok, None]
    nrm = np.stack([-n[:, 0], -n[:, 1], n[:, 2]], axis=1)
    nrm[~ok] = 0.0
    return pts, nrm.astype(np.float32), lin.astype(np.int64)
