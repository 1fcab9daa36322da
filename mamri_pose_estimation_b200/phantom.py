"""Synthetic MRI phantoms for tests and benchmarks (BASELINE.json ``configs``;
recipe in SURVEY.md §8d): uint16 volume, air 0, one "body" ellipsoid (140),
petroleum-jelly fiducial ellipsoids (400) where a posed MAMRI arm carries its
markers (robot.py), optional spurious bright blobs (300), Rician noise
``v = round(sqrt((A+n1)^2 + n2^2))`` driven by counter-based Philox4x32-10.

Two generators share one description (`Phantom`): the NumPy one here (tests,
CPU) and the CUDA one in ``csrc/phantom.cu`` (benchmarks; keeps batches of scans
off PCIe).  They consume identical Philox counters; the noise-free signal is
bit-identical, noisy voxels agree except where libm/CUDA logf/sincosf rounding
flips a final integer rounding (tests bound the mismatch fraction).

This is data generation, not part of the detection path.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import robot

BASE_SEED = 0x4D414D52          # "MAMR"
AIR, BODY, BLOB, FIDUCIAL = 0.0, 140.0, 300.0, 400.0
FOV_MM = 410.0

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_STREAM_TAG = 0x50484E54        # "PHNT": 4th counter word of the noise stream


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32 with 10 rounds (Salmon et al. 2011) on uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32).copy(); c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy(); c3 = np.asarray(c3, dtype=np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = _M0 * c0.astype(np.uint64)
        p1 = _M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
        kk0 = np.uint32((k0 + r * _W0) & 0xFFFFFFFF)
        kk1 = np.uint32((k1 + r * _W1) & 0xFFFFFFFF)
        c0, c1, c2, c3 = hi1 ^ c1 ^ kk0, lo1, hi0 ^ c3 ^ kk1, lo0
    return c0, c1, c2, c3


@dataclasses.dataclass
class Phantom:
    name: str
    dims: Tuple[int, int, int]                   # (nx, ny, nz)
    spacing: Tuple[float, float, float]
    origin: Tuple[float, float, float]
    direction: Tuple[float, ...]
    ellipsoids: np.ndarray                       # [E,7] f32: cx,cy,cz (index), ax,ay,az (voxels), intensity; painted in order
    sigma: float
    seed: int
    scan_index: int = 0
    truth: dict = dataclasses.field(default_factory=dict)

    @property
    def n_voxels(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]


# ----------------------------------------------------------------------------- numpy generator
def paint_signal(ph: Phantom) -> np.ndarray:
    """Noise-free uint16 signal; later ellipsoids overwrite earlier ones.  The
    inside test is evaluated in float32 with separately rounded operations
    (sub, div, mul, add in x,y,z order) so the CUDA painter reproduces it bit for bit."""
    nx, ny, nz = ph.dims
    vol = np.zeros((nz, ny, nx), dtype=np.uint16)
    f = np.float32
    for cx, cy, cz, ax, ay, az, inten in np.asarray(ph.ellipsoids, dtype=np.float32):
        x0, x1 = max(0, int(math.floor(cx - ax))), min(nx - 1, int(math.ceil(cx + ax)))
        y0, y1 = max(0, int(math.floor(cy - ay))), min(ny - 1, int(math.ceil(cy + ay)))
        z0, z1 = max(0, int(math.floor(cz - az))), min(nz - 1, int(math.ceil(cz + az)))
        if x0 > x1 or y0 > y1 or z0 > z1:
            continue
        qx = ((np.arange(x0, x1 + 1, dtype=np.float32) - f(cx)) / f(ax)) ** 2
        qy = ((np.arange(y0, y1 + 1, dtype=np.float32) - f(cy)) / f(ay)) ** 2
        qz = ((np.arange(z0, z1 + 1, dtype=np.float32) - f(cz)) / f(az)) ** 2
        q = (qx[None, None, :] + qy[None, :, None]) + qz[:, None, None]
        sub = vol[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]
        sub[q <= f(1.0)] = np.uint16(inten)
    return vol


def add_rician_noise(signal: np.ndarray, sigma: float, seed: int, scan_index: int = 0,
                     chunk: int = 1 << 22) -> np.ndarray:
    """One Philox call per voxel PAIR p = linear_index // 2: counter (p_lo, p_hi,
    scan_index, TAG), key (seed_lo, seed_hi); words 0,1 drive the even voxel,
    2,3 the odd one.  u = ((w >> 8) + 0.5) * 2^-24; Box-Muller in float32."""
    flat = signal.reshape(-1)
    n = flat.size
    out = np.empty(n, dtype=np.uint16)
    if sigma <= 0:
        out[:] = flat
        return out.reshape(signal.shape)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    f = np.float32
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        p = np.arange(s // 2, (e + 1) // 2, dtype=np.uint64)
        r0, r1, r2, r3 = philox4x32_10((p & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                                       (p >> np.uint64(32)).astype(np.uint32),
                                       np.full(p.shape, scan_index, dtype=np.uint32),
                                       np.full(p.shape, _STREAM_TAG, dtype=np.uint32), k0, k1)
        ua = np.stack([r0, r2], axis=1).reshape(-1)
        ub = np.stack([r1, r3], axis=1).reshape(-1)
        off = s - 2 * (s // 2)
        ua, ub = ua[off:off + (e - s)], ub[off:off + (e - s)]
        u1 = ((ua >> np.uint32(8)).astype(np.float32) + f(0.5)) * f(2.0 ** -24)
        u2 = ((ub >> np.uint32(8)).astype(np.float32) + f(0.5)) * f(2.0 ** -24)
        rad = np.sqrt(f(-2.0) * np.log(u1)) * f(sigma)
        ang = f(2.0) * u2                                  # in units of pi
        n1 = rad * np.cos(np.float32(math.pi) * ang)
        n2 = rad * np.sin(np.float32(math.pi) * ang)
        a = flat[s:e].astype(np.float32) + n1
        v = np.rint(np.sqrt(a * a + n2 * n2))
        out[s:e] = np.clip(v, 0, 65535).astype(np.uint16)
    return out.reshape(signal.shape)


def generate(ph: Phantom) -> np.ndarray:
    """uint16 volume [nz, ny, nx]."""
    return add_rician_noise(paint_signal(ph), ph.sigma, ph.seed, ph.scan_index)


# ----------------------------------------------------------------------------- geometry helpers
def _centered_geometry(dims, fov=FOV_MM):
    spacing = tuple(fov / d for d in dims)
    # voxel centres symmetric about 0 (LPS): index (n-1)/2 maps to 0
    origin = tuple(-0.5 * (d - 1) * s for d, s in zip(dims, spacing))
    return spacing, origin, (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)


def ras_to_index(p_ras: np.ndarray, spacing, origin) -> np.ndarray:
    """RAS -> LPS (flip x,y; Mamri.py:1317 inverted) -> continuous index (identity direction)."""
    p = np.asarray(p_ras, dtype=np.float64)
    lps = p * np.array([-1.0, -1.0, 1.0])
    return (lps - np.asarray(origin)) / np.asarray(spacing)


def _body_ellipsoid(dims) -> List[float]:
    nx, ny, nz = dims
    return [(nx - 1) / 2.0, (ny - 1) / 2.0, (nz - 1) / 2.0, 0.30 * nx, 0.22 * ny, 0.40 * nz, BODY]


def _inside_margin(c, body, margin_vox: float) -> bool:
    """True if index point c is within `margin_vox` voxels (conservatively) of the body ellipsoid."""
    q = sum(((c[i] - body[i]) / (body[3 + i] + margin_vox)) ** 2 for i in range(3))
    return q <= 1.0


# Pose used for the robot-carrying phantoms (degrees).  Found by a random search so that all four
# marker triplets fit the 410 mm field of view beside and above the supine body, with every pair of fiducials
# at least 5 voxels apart even at C1's 1.6 x 1.6 x 3.2 mm voxels (the 20 mm short arms of the
# L-shapes must stay out of the coarse z direction or the radius-2 closing would fuse them).
ROBOT_POSE_DEG = (-25.0, -35.0, 60.0, 25.0, 70.0, -20.0)
ROBOT_BASE_RAS = (75.0, -174.0, -95.0)       # translation of the baseplate frame in RAS, mm


def robot_base_matrix() -> np.ndarray:
    """Baseplate frame in RAS: lying on the scanner table, i.e. rotated -90 deg about x so that its
    local z (the arm's axis) points anterior (+y).  The reference assumes exactly this: it forces the
    three baseplate markers to one common RAS y before registering them (Mamri.py:1371-1373)."""
    m = np.eye(4)
    m[:3, 3] = ROBOT_BASE_RAS
    return m @ robot._rot("X", -90.0)


def _robot_fiducials(dims, spacing, origin, links: Sequence[str], rng: np.random.Generator):
    pose = [math.radians(a) for a in ROBOT_POSE_DEG]
    pos = robot.marker_positions_ras(pose, robot_base_matrix(), links)
    ells, centres = [], {}
    for name in links:
        centres[name] = pos[name]
        for p in pos[name]:
            c = ras_to_index(p, spacing, origin)
            semi_mm = rng.uniform(3.0, 5.5, size=3)
            ells.append([c[0], c[1], c[2], semi_mm[0] / spacing[0], semi_mm[1] / spacing[1],
                         semi_mm[2] / spacing[2], FIDUCIAL])
    return ells, centres


def _extent_along(u, semi) -> float:
    """Half-width of an axis-aligned ellipsoid along unit direction u (support function)."""
    return math.sqrt(sum((u[a] * semi[a]) ** 2 for a in range(3)))


def _check_layout(dims, ells: np.ndarray, body, min_body_gap=8.0, min_border=6.0, min_gap=5.0):
    """Asserts the recipe's clearances for fiducials: >= 8 voxels from the body, >= 6 from the
    border and >= 5 from each other (a radius-2 closing bridges gaps of up to 4 voxels)."""
    fid = [e for e in ells if e[6] == FIDUCIAL]
    for i, e in enumerate(fid):
        r = max(e[3:6])
        assert not _inside_margin(e[:3], body, min_body_gap + r), f"fiducial {i} too close to the body"
        for a in range(3):
            assert e[a] - e[3 + a] >= min_border and e[a] + e[3 + a] <= dims[a] - 1 - min_border, \
                f"fiducial {i} too close to the border on axis {a}: {e[:6]}"
        for j in range(i):
            o = fid[j]
            d = [e[a] - o[a] for a in range(3)]
            n = math.sqrt(sum(v * v for v in d))
            u = [v / n for v in d]
            gap = n - _extent_along(u, e[3:6]) - _extent_along(u, o[3:6])
            assert gap >= min_gap, f"fiducials {i},{j} only {gap:.1f} voxels apart"


def _make(name, dims, sigma, links, seed, scan_index=0, n_extra_fid=0, n_blobs=0, check=True) -> Phantom:
    spacing, origin, direction = _centered_geometry(dims)
    rng = np.random.Generator(np.random.Philox(seed))
    body = _body_ellipsoid(dims)
    fid, centres = _robot_fiducials(dims, spacing, origin, links, rng)
    blobs: List[List[float]] = []
    nx, ny, nz = dims
    # extra free-floating fiducials (C4): rejection-sample clear positions
    tries = 0
    while n_extra_fid > 0 and tries < 100000:
        tries += 1
        c = [rng.uniform(16, nx - 17), rng.uniform(16, ny - 17), rng.uniform(16, nz - 17)]
        semi_mm = rng.uniform(3.0, 5.5, size=3)
        e = [c[0], c[1], c[2], semi_mm[0] / spacing[0], semi_mm[1] / spacing[1], semi_mm[2] / spacing[2], FIDUCIAL]
        r = max(e[3:6])
        if _inside_margin(c, body, 10.0 + r):
            continue
        if any(math.dist(c, o[:3]) - r - max(o[3:6]) < 8.0 for o in fid):
            continue
        fid.append(e)
        n_extra_fid -= 1
    assert n_extra_fid == 0, "could not place the extra fiducials"
    if n_blobs:
        # spurious bright blobs: random spheres r in [1,12] voxels anywhere (may overlap each other and
        # the body); every 10th starts a thin diagonal chain (1-voxel spheres stepping by (1,1,1)) that is
        # 26- but not 6-connected; every 7th is centred on a multiple-of-32/8 plane to straddle CCL tiles.
        for b in range(n_blobs):
            r = float(rng.uniform(1.0, 12.0))
            c = [rng.uniform(14, nx - 15), rng.uniform(14, ny - 15), rng.uniform(14, nz - 15)]
            if b % 7 == 0:
                c = [32.0 * round(c[0] / 32.0) - 0.5, 8.0 * round(c[1] / 8.0) - 0.5, 8.0 * round(c[2] / 8.0) - 0.5]
            if any(math.dist(c, o[:3]) - r - max(o[3:6]) < 8.0 for o in fid):
                continue
            if b % 10 == 0:
                for s in range(int(rng.integers(6, 40))):
                    cc = [math.floor(c[0]) + 3 * s, math.floor(c[1]) + 3 * s, math.floor(c[2]) + s]
                    if cc[0] > nx - 15 or cc[1] > ny - 15 or cc[2] > nz - 15:
                        break
                    if any(math.dist(cc, o[:3]) - 1.0 - max(o[3:6]) < 8.0 for o in fid):
                        break
                    blobs.append([cc[0], cc[1], cc[2], 0.6, 0.6, 0.6, BLOB])
            else:
                blobs.append([c[0], c[1], c[2], r, r, r, BLOB])
    ells = np.asarray([body] + blobs + fid, dtype=np.float32)
    if check:
        _check_layout(dims, ells, body)
    truth = dict(pose_rad=[math.radians(a) for a in ROBOT_POSE_DEG], base=robot_base_matrix(),
                 marker_ras=centres, n_fiducials=len(fid), n_blobs=len(blobs))
    return Phantom(name=name, dims=tuple(dims), spacing=spacing, origin=origin, direction=direction,
                   ellipsoids=ells, sigma=float(sigma), seed=int(seed), scan_index=int(scan_index), truth=truth)


# ----------------------------------------------------------------------------- BASELINE.json configs
def config_c1(seed: int = BASE_SEED + 1) -> Phantom:
    """256x256x128, sigma 10, 9 fiducials (Baseplate, Joint4, Joint6 triplets)."""
    return _make("C1", (256, 256, 128), 10.0, ("Baseplate", "Joint4", "Joint6"), seed)


def config_c2(seed: int = BASE_SEED + 2, scan_index: int = 0, sigma: float = 10.0, name: str = "C2") -> Phantom:
    """512x512x256, sigma 10, baseplate + end-effector (Joint6) fiducials."""
    return _make(name, (512, 512, 256), sigma, ("Baseplate", "Joint6"), seed, scan_index)


def config_c3(scan_index: int, seed: int = BASE_SEED + 3) -> Phantom:
    """One of the 64 scans of the batch config: C2 geometry, sigma 15, per-scan noise stream."""
    return config_c2(seed=seed, scan_index=scan_index, sigma=15.0, name="C3")


def config_c4(seed: int = BASE_SEED + 4, dims=(1024, 1024, 512), n_blobs: int = 2000) -> Phantom:
    """1024x1024x512, sigma 20, 32 fiducials (12 on the robot + 20 free) + 2000 spurious blobs."""
    return _make("C4", dims, 20.0, robot.MARKER_LINKS, seed, n_extra_fid=20, n_blobs=n_blobs)


def small_phantom(dims=(64, 48, 40), n_fiducials=5, n_blobs=6, sigma=12.0, seed=1, scan_index=0,
                  spacing=(1.5, 1.5, 3.0), flip_lps=False, touch_border=False) -> Phantom:
    """Voxel-unit phantom for CPU-sized parity tests: same ingredients, arbitrary (ragged) dims."""
    nx, ny, nz = dims
    rng = np.random.Generator(np.random.Philox(seed))
    body = [(nx - 1) / 2.0, (ny - 1) / 2.0, (nz - 1) / 2.0, 0.22 * nx, 0.18 * ny, 0.30 * nz, BODY]
    ells = [body]
    for _ in range(n_blobs):
        r = float(rng.uniform(0.8, 4.0))
        ells.append([rng.uniform(0, nx - 1), rng.uniform(0, ny - 1), rng.uniform(0, nz - 1), r, r, r, BLOB])
    for i in range(n_fiducials):
        a = rng.uniform(1.2, 3.2, size=3)
        c = [rng.uniform(min(4, nx / 2), max(nx - 5, nx / 2)), rng.uniform(min(4, ny / 2), max(ny - 5, ny / 2)),
             rng.uniform(min(3, nz / 2), max(nz - 4, nz / 2))]
        if touch_border and i == 0:
            c[0] = 0.5
        ells.append([c[0], c[1], c[2], a[0], a[1], a[2], FIDUCIAL])
    direction = (-1.0, 0, 0, 0, -1.0, 0, 0, 0, 1.0) if flip_lps else (1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0)
    origin = (12.5, -7.25, 3.0)
    return Phantom(name="small", dims=tuple(dims), spacing=tuple(spacing), origin=origin, direction=direction,
                   ellipsoids=np.asarray(ells, dtype=np.float32), sigma=float(sigma), seed=int(seed),
                   scan_index=int(scan_index))


# ----------------------------------------------------------------------------- entry-search candidates (C5)
def surface_candidates(n: int, seed: int = BASE_SEED + 5, dims=(512, 512, 256)):
    """C5: `n` points uniformly (in parameter space) on the C2 body ellipsoid surface with analytic
    outward unit normals, float32, RAS mm; target 40 mm under the surface on the +x (left-right) side."""
    spacing, origin, _ = _centered_geometry(dims)
    body = _body_ellipsoid(dims)
    semi = np.array([body[3] * spacing[0], body[4] * spacing[1], body[5] * spacing[2]])   # mm
    rng = np.random.Generator(np.random.Philox(seed))
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pts = v * semi[None, :]
    nrm = pts / (semi[None, :] ** 2)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    target = np.array([semi[0] - 40.0, 10.0, 15.0])
    return pts.astype(np.float32), nrm.astype(np.float32), target
