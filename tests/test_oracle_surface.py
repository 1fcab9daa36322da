"""CPU tests of the skin-surface candidate definition (oracle/surface.py): set-theoretic properties and
geometric validation on analytic shapes -- the reference's own surface (Slicer closed-surface conversion +
vtkPolyDataNormals, Mamri.py:994-1003) cannot be produced outside Slicer, so there is no parity target."""
import numpy as np
import pytest

from oracle import segmentation as seg
from oracle import surface as srf


def _ellipsoid(dims_xyz, centre, semi):
    nx, ny, nz = dims_xyz
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    q = ((x - centre[0]) / semi[0]) ** 2 + ((y - centre[1]) / semi[1]) ** 2 + ((z - centre[2]) / semi[2]) ** 2
    return (q <= 1.0).astype(np.uint8)


def test_surface_is_body_minus_6_erosion():
    rng = np.random.default_rng(5)
    body = (rng.random((12, 17, 37)) < 0.7).astype(np.uint8)
    s = srf.surface_voxels(body)
    from scipy import ndimage
    er = ndimage.binary_erosion(body != 0, structure=ndimage.generate_binary_structure(3, 1), border_value=0)
    assert np.array_equal(s, (body != 0) & ~er)
    assert not s[body == 0].any()


def test_ball_moment_matches_direct_sum():
    rng = np.random.default_rng(6)
    body = (rng.random((9, 10, 40)) < 0.5).astype(np.uint8)
    zyx = np.argwhere(srf.surface_voxels(body))[::7]
    g = srf.ball_moments(body, zyx)
    nz, ny, nx = body.shape
    for (z, y, x), gi in zip(zyx, g):
        acc = np.zeros(3, dtype=np.int64)
        for dz in range(-2, 3):
            for dy in range(-2, 3):
                for dx in range(-2, 3):
                    if dx * dx + dy * dy + dz * dz <= 6 and 0 <= z + dz < nz and 0 <= y + dy < ny and 0 <= x + dx < nx:
                        if body[z + dz, y + dy, x + dx]:
                            acc += (dx, dy, dz)
        assert tuple(acc) == tuple(gi)


@pytest.mark.parametrize("spacing,flip", [((1.0, 1.0, 1.0), False), ((0.8, 0.8, 1.6), True)])
def test_normals_follow_the_analytic_ellipsoid(spacing, flip):
    dims = (72, 64, 48)
    centre, semi = (35.5, 31.0, 23.5), (28.0, 22.0, 17.0)
    body = _ellipsoid(dims, centre, semi)
    direction = (-1, 0, 0, 0, -1, 0, 0, 0, 1) if flip else (1, 0, 0, 0, 1, 0, 0, 0, 1)
    geom = seg.Geometry(spacing, (3.0, -7.0, 11.0), direction)
    pts, nrm, lin = srf.body_surface(body, geom)
    assert len(pts) == int(srf.surface_voxels(body).sum()) and np.all(np.diff(lin) > 0)
    # analytic outward normal of the ellipsoid at each candidate, in RAS
    nx_, ny_ = dims[0], dims[1]
    x, y, z = lin % nx_, (lin // nx_) % ny_, lin // (nx_ * ny_)
    sp = np.array(spacing)
    D = np.array(direction, dtype=np.float64).reshape(3, 3)
    gidx = np.stack([(x - centre[0]) / semi[0] ** 2, (y - centre[1]) / semi[1] ** 2, (z - centre[2]) / semi[2] ** 2], axis=1)
    n_lps = (gidx / sp) @ D.T
    n_ras = n_lps * np.array([-1.0, -1.0, 1.0])
    n_ras /= np.linalg.norm(n_ras, axis=1, keepdims=True)
    cosang = np.sum(n_ras * nrm.astype(np.float64), axis=1)
    assert np.all(np.abs(np.linalg.norm(nrm.astype(np.float64), axis=1) - 1.0) < 1e-6)
    ang = np.degrees(np.arccos(np.clip(cosang, -1, 1)))
    assert np.median(ang) < 4.0 and ang.max() < 20.0, (np.median(ang), ang.max())
    # points are voxel centres in RAS
    lps = np.array(geom.origin) + (np.stack([x, y, z], axis=1) * sp) @ D.T
    assert np.allclose(pts, lps * np.array([-1.0, -1.0, 1.0]), atol=1e-4)


def test_degenerate_bodies():
    geom = seg.Geometry((1, 1, 1), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    pts, nrm, lin = srf.body_surface(np.zeros((5, 6, 7), np.uint8), geom)
    assert len(pts) == 0 and len(nrm) == 0
    one = np.zeros((5, 6, 7), np.uint8); one[2, 3, 4] = 1
    pts, nrm, lin = srf.body_surface(one, geom)
    assert len(pts) == 1 and np.array_equal(nrm, np.zeros((1, 3), np.float32))       # symmetric -> zero moment
    assert np.array_equal(pts[0], np.array([-4.0, -3.0, 2.0], np.float32))
    full = np.ones((5, 6, 7), np.uint8)                                                # the volume border is surface
    pts, nrm, lin = srf.body_surface(full, geom)
    assert len(pts) == 5 * 6 * 7 - 3 * 4 * 5
