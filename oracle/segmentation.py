"""Oracle for stages 1-4a: threshold, ball closing, connected components,
label-shape statistics, candidate filter, body selection.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): CPU restatement, on
numpy/scipy.ndimage, of what ``MamriLogic.volume_threshold_segmentation``
(``Mamri/Mamri.py:1304-1323``) asks SimpleITK to do.  PARITY UNPINNED: SimpleITK
is not available here; semantics follow ITK 5.x as documented in SURVEY.md §8c.

Array convention everywhere: ``vol[z, y, x]`` C-contiguous (x fastest), the
layout of an ITK/SimpleITK buffer; linear index = x + nx*(y + ny*z).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import ndimage

# Constants of the reference (Mamri/Mamri.py:810-812, :1308).
INTENSITY_THRESHOLD = 65.0
UPPER_THRESHOLD = 65535.0
MIN_VOLUME_THRESHOLD = 50.0
MAX_VOLUME_THRESHOLD = 1500.0
CLOSE_RADIUS = 2


# --------------------------------------------------------------------------- #
# (c-1) sitk.BinaryThreshold(img, lo, hi)            Mamri/Mamri.py:1308
# --------------------------------------------------------------------------- #
def cast_threshold(value: float, dtype: np.dtype) -> object:
    """``static_cast<InputPixelType>(double)`` as itk::BinaryThresholdImageFilter
    applies to its bounds.  Integer types truncate toward zero; an out-of-range
    bound is clamped to the type's range (documented deviation for int16 input,
    where the reference's 65535 is not representable and ITK would reject
    lower > upper)."""
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return dtype.type(value)
    info = np.iinfo(dtype)
    if np.isnan(value):
        return dtype.type(0)
    v = np.trunc(value)
    return dtype.type(int(min(max(v, info.min), info.max)))


def binary_threshold(vol: np.ndarray, lo: float = INTENSITY_THRESHOLD,
                     hi: float = UPPER_THRESHOLD) -> np.ndarray:
    """uint8 {0,1}: 1 where lo <= v <= hi (both inclusive), compared in the
    input pixel type.  NaN compares false.

    Restates itk::BinaryThresholdImageFilter (ITK 5.x,
    Modules/Filtering/Thresholding/include/itkBinaryThresholdImageFilter.h:
    Functor::BinaryThreshold::operator() ``m_LowerThreshold <= A && A <=
    m_UpperThreshold ? m_InsideValue : m_OutsideValue``; the bounds reach the functor
    through ``static_cast<InputPixelType>`` of SimpleITK's double arguments)."""
    lo_c = cast_threshold(lo, vol.dtype)
    hi_c = cast_threshold(hi, vol.dtype)
    return ((vol >= lo_c) & (vol <= hi_c)).astype(np.uint8)


# --------------------------------------------------------------------------- #
# (c-2) itk::FlatStructuringElement<3>::Ball(r, radiusIsParametric=false)
# --------------------------------------------------------------------------- #
def ball_offsets(radius: int) -> np.ndarray:
    """Offsets (dz, dy, dx) of ITK's ball: voxel centres inside the ellipsoid of
    axes 2r+1, i.e. dx^2+dy^2+dz^2 <= (r+0.5)^2  <=>  <= r*r + r for integers.
    r=1 -> 19 voxels, r=2 -> 81, r=3 -> 179.

    Restates itk::FlatStructuringElement<3>::Ball(radius, radiusIsParametric=false)
    (Modules/Filtering/MathematicalMorphology/include/itkFlatStructuringElement.hxx):
    an EllipsoidInteriorExteriorSpatialFunction with axes = size = 2r+1 centred at
    r + 0.5, flood-filled from the centre pixel with the centre-inclusion strategy
    (a pixel belongs iff its centre index + 0.5 is inside).  SimpleITK's sitkBall
    maps to this constructor.  tools/make_itk_golden.py records the element ITK
    really used (dilation of one voxel) as ``ball{r}``."""
    r = int(radius)
    g = np.arange(-r, r + 1)
    dz, dy, dx = np.meshgrid(g, g, g, indexing="ij")
    keep = dx * dx + dy * dy + dz * dz <= r * r + r
    return np.stack([dz[keep], dy[keep], dx[keep]], axis=1)


def ball_structure(radius: int) -> np.ndarray:
    r = int(radius)
    s = np.zeros((2 * r + 1,) * 3, dtype=bool)
    off = ball_offsets(r)
    s[off[:, 0] + r, off[:, 1] + r, off[:, 2] + r] = True
    return s


# --------------------------------------------------------------------------- #
# (c-3) sitk.BinaryMorphologicalClosing(binary, [r]*3, sitkBall)  Mamri.py:1308
# --------------------------------------------------------------------------- #
def binary_opening(mask: np.ndarray, radius: int) -> np.ndarray:
    """sitk.BinaryMorphologicalOpening(mask, [r]*3, sitkBall): the north_star's
    "open/close" -- the reference itself only closes (Mamri.py:1308), so this is
    an extension (``mamri_params.open_radius``, 0 = off = the reference).

    Restates itk::BinaryMorphologicalOpeningImageFilter::GenerateData
    (Modules/Filtering/BinaryMathematicalMorphology/include/
    itkBinaryMorphologicalOpeningImageFilter.hxx): BinaryErodeImageFilter, whose
    constructor sets BoundaryToForeground = true (outside the image counts as
    foreground: an object is not eroded from the image border), then
    BinaryDilateImageFilter (outside = background).  There is no safe-border
    padding in the opening filter."""
    r = int(radius)
    if r == 0:
        return mask.astype(np.uint8).copy()
    st = ball_structure(r)
    e = ndimage.binary_erosion(mask.astype(bool), structure=st, border_value=1)
    d = ndimage.binary_dilation(e, structure=st, border_value=0)
    return d.astype(np.uint8)


def binary_closing_safe_border(mask: np.ndarray, radius: int = CLOSE_RADIUS) -> np.ndarray:
    """itk::BinaryMorphologicalClosingImageFilter with SafeBorder=true
    (Modules/Filtering/BinaryMathematicalMorphology/include/
    itkBinaryMorphologicalClosingImageFilter.hxx, GenerateData: ConstantPadImageFilter
    by the kernel radius -> BinaryDilateImageFilter -> BinaryErodeImageFilter ->
    CropImageFilter; SimpleITK's procedural call leaves safeBorder at its default true):
    ConstantPad(r, 0) -> BinaryDilate (outside = background) -> BinaryErode
    (outside = foreground) -> Crop(r).  For voxels of the original domain this is
    the closing of the zero-extended mask on an unbounded grid.  Restated as:
    zero-pad by 2r, plain dilate and erode (outside = 0), crop 2r -- the erosion
    of an original-domain voxel only samples the r-apron, where the padded
    dilation is exact."""
    r = int(radius)
    if r == 0:
        return mask.astype(np.uint8).copy()
    st = ball_structure(r)
    p = np.pad(mask.astype(bool), 2 * r, mode="constant", constant_values=False)
    d = ndimage.binary_dilation(p, structure=st, border_value=0)
    e = ndimage.binary_erosion(d, structure=st, border_value=0)
    c = (slice(2 * r, -2 * r),) * 3
    return e[c].astype(np.uint8)


# --------------------------------------------------------------------------- #
# (c-4) sitk.ConnectedComponent(closed)              Mamri/Mamri.py:1309
# --------------------------------------------------------------------------- #
def canonical_relabel(labels: np.ndarray) -> Tuple[np.ndarray, int]:
    """Relabel so that label k (1-based) is the component whose minimum linear
    voxel index is the k-th smallest -- ITK's consecutive numbering and the
    north_star's canonical form."""
    flat = labels.ravel()
    fg = np.flatnonzero(flat)
    if fg.size == 0:
        return np.zeros(labels.shape, dtype=np.uint32), 0
    vals = flat[fg]
    uniq, first = np.unique(vals, return_index=True)   # first occurrence in raster order
    order = np.argsort(first, kind="stable")
    lut = np.zeros(int(uniq.max()) + 1, dtype=np.uint32)
    lut[uniq[order]] = np.arange(1, uniq.size + 1, dtype=np.uint32)
    out = np.zeros(flat.shape, dtype=np.uint32)
    out[fg] = lut[vals]
    return out.reshape(labels.shape), int(uniq.size)


def connected_components(mask: np.ndarray, connectivity: int = 6) -> Tuple[np.ndarray, int]:
    """itk::ConnectedComponentImageFilter; SimpleITK default fullyConnected=False
    is face connectivity (6), True is 26.  uint32 labels, background 0.

    Restates Modules/Segmentation/ConnectedComponents/include/
    itkConnectedComponentImageFilter.hxx + itkScanlineFilterCommon.h: provisional
    labels per x-run in raster order, LinkLabels keeps the smaller label as the
    root of a union, CreateConsecutive numbers the roots 1..K in increasing order
    of provisional label -- i.e. by each object's first voxel in raster order,
    which is what ``canonical_relabel`` produces."""
    if connectivity == 6:
        st = ndimage.generate_binary_structure(3, 1)
    elif connectivity == 26:
        st = ndimage.generate_binary_structure(3, 3)
    else:
        raise ValueError("connectivity must be 6 or 26")
    lab, _ = ndimage.label(mask != 0, structure=st)
    return canonical_relabel(lab)


# --------------------------------------------------------------------------- #
# (c-5) sitk.LabelShapeStatisticsImageFilter().Execute(labeled)  Mamri.py:1309
# --------------------------------------------------------------------------- #
@dataclasses.dataclass
class Geometry:
    """Image geometry as SimpleITK holds it (LPS): physical = origin +
    direction @ (spacing * index)."""
    spacing: Tuple[float, float, float] = (1.0, 1.0, 1.0)      # (sx, sy, sz)
    origin: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    direction: Tuple[float, ...] = (1, 0, 0, 0, 1, 0, 0, 0, 1)  # row-major 3x3

    def matrix(self) -> np.ndarray:
        d = np.asarray(self.direction, dtype=np.float64).reshape(3, 3)
        return d * np.asarray(self.spacing, dtype=np.float64)[None, :]

    def index_to_physical(self, cidx: Sequence[float]) -> np.ndarray:
        return np.asarray(self.origin, dtype=np.float64) + self.matrix() @ np.asarray(cidx, dtype=np.float64)

    def voxel_volume(self) -> float:
        s = self.spacing
        return float(s[0]) * float(s[1]) * float(s[2])


@dataclasses.dataclass
class LabelStats:
    label: int
    count: int
    sum_idx: Tuple[int, int, int]            # (sum x, sum y, sum z), exact integers
    sum_mom: Tuple[int, int, int, int, int, int]  # xx yy zz xy xz yz, exact integers
    physical_size: float
    centroid_index: np.ndarray
    centroid: np.ndarray                     # physical (LPS)
    principal_moments: np.ndarray
    principal_axes: np.ndarray               # rows = axes (ITK: V^T, last row * det)


def integer_sums(labels: np.ndarray, n_labels: int):
    """Exact per-label integer sums via float64 bincount (all partial sums are
    integers < 2^53 for volumes up to 1024x1024x512)."""
    nz, ny, nx = labels.shape
    flat = labels.ravel()
    fg = np.flatnonzero(flat)
    lab = flat[fg].astype(np.int64)
    x = (fg % nx).astype(np.float64)
    y = ((fg // nx) % ny).astype(np.float64)
    z = (fg // (nx * ny)).astype(np.float64)
    m = n_labels + 1
    bc = lambda w=None: np.bincount(lab, weights=w, minlength=m)[1:]
    cnt = bc().astype(np.int64)
    sums = [np.rint(bc(w)).astype(np.int64) for w in (x, y, z)]
    moms = [np.rint(bc(w)).astype(np.int64) for w in (x * x, y * y, z * z, x * y, x * z, y * z)]
    return cnt, np.stack(sums, axis=1), np.stack(moms, axis=1)


def moments_from_sums(count: int, sum_idx, sum_mom, geom: Geometry):
    """ShapeLabelMapFilter's second central moments from exact index sums
    (itk::ShapeLabelMapFilter::ThreadedProcessLabelObject,
    Modules/Filtering/LabelMap/include/itkShapeLabelMapFilter.hxx: sums over the
    label object's lines, "Normalize using the total mass", "Center the second
    order moments", then ``centralMoments[i][i] += spacing[i]^2 / 12`` -- the
    second moment of one pixel --, vnl_symmetric_eigensystem (ascending), principal
    axes = V^T with the last row multiplied by the determinant).
    With p = o + A i (A = direction*spacing):  E[pp^T]-cc^T = A Cov(i) A^T, plus
    spacing_i^2/12 on the diagonal (second moment of one voxel box).  Principal
    moments ascending; principal axes = rows of V^T with the last row multiplied
    by det to make a proper rotation."""
    n = float(count)
    sx, sy, sz = (float(v) for v in sum_idx)
    xx, yy, zz, xy, xz, yz = (float(v) for v in sum_mom)
    cx, cy, cz = sx / n, sy / n, sz / n
    cov = np.array([[xx / n - cx * cx, xy / n - cx * cy, xz / n - cx * cz],
                    [xy / n - cx * cy, yy / n - cy * cy, yz / n - cy * cz],
                    [xz / n - cx * cz, yz / n - cy * cz, zz / n - cz * cz]], dtype=np.float64)
    a = geom.matrix()
    m = a @ cov @ a.T
    m[np.diag_indices(3)] += np.asarray(geom.spacing, dtype=np.float64) ** 2 / 12.0
    w, v = np.linalg.eigh(m)
    axes = v.T.copy()
    axes[2] *= np.linalg.det(axes)
    return w, axes


def label_shape_statistics(labels: np.ndarray, n_labels: int, geom: Geometry) -> List[LabelStats]:
    cnt, sums, moms = integer_sums(labels, n_labels)
    vv = geom.voxel_volume()
    out = []
    for k in range(n_labels):
        n = int(cnt[k])
        cidx = sums[k].astype(np.float64) / n
        pm, pa = moments_from_sums(n, sums[k], moms[k], geom)
        out.append(LabelStats(label=k + 1, count=n,
                              sum_idx=tuple(int(v) for v in sums[k]),
                              sum_mom=tuple(int(v) for v in moms[k]),
                              physical_size=n * vv, centroid_index=cidx,
                              centroid=geom.index_to_physical(cidx),
                              principal_moments=pm, principal_axes=pa))
    return out


# --------------------------------------------------------------------------- #
# (c-6) candidate filter, RAS flip, body label        Mamri/Mamri.py:1310-1323
# --------------------------------------------------------------------------- #
@dataclasses.dataclass
class Detection:
    binary: np.ndarray              # after threshold
    closed: np.ndarray              # after closing (uint8)
    labels: np.ndarray              # uint32, ITK-consecutive
    n_labels: int
    counts: np.ndarray              # per label voxel count, label order
    fiducials: List[dict]           # [{"vol","centroid"(LPS),"id"}] ascending label (Mamri.py:1310)
    ras_points: np.ndarray          # M x 3, [-x,-y,z] (Mamri.py:1317)
    marker_labels: List[str]        # f"M_{id}_{vol:.0f}mm3" (Mamri.py:1317)
    body_label: int                 # 0 = none (Mamri.py:1319,1321)
    body_mask: Optional[np.ndarray]
    stats: List[LabelStats]         # only for the kept labels (+ body) unless full_stats


def select_candidates(counts: np.ndarray, voxel_volume: float,
                      min_vol: float = MIN_VOLUME_THRESHOLD,
                      max_vol: float = MAX_VOLUME_THRESHOLD):
    """Labels kept by the list comprehension at Mamri.py:1310 (inclusive bounds,
    ascending label order) and the body label of :1320-1322 (largest physical
    size among the others; first maximum = lowest label on ties)."""
    vols = counts.astype(np.float64) * voxel_volume
    keep = (vols >= min_vol) & (vols <= max_vol)
    kept = (np.flatnonzero(keep) + 1).tolist()
    rest = np.flatnonzero(~keep)
    body = 0
    if rest.size:
        body = int(rest[np.argmax(vols[rest])]) + 1     # argmax returns the first maximum
    return kept, body


def detect_fiducials(vol: np.ndarray, geom: Geometry, lo: float = INTENSITY_THRESHOLD,
                     hi: float = UPPER_THRESHOLD, close_radius: int = CLOSE_RADIUS,
                     connectivity: int = 6, min_vol: float = MIN_VOLUME_THRESHOLD,
                     max_vol: float = MAX_VOLUME_THRESHOLD, full_stats: bool = False,
                     open_radius: int = 0) -> Detection:
    """The whole of Mamri.py:1308-1323 on one volume (``open_radius`` > 0: the
    thresholded mask is opened first -- north_star extension, not in the reference)."""
    binary = binary_threshold(vol, lo, hi)
    closed = binary_closing_safe_border(binary_opening(binary, open_radius), close_radius)
    labels, k = connected_components(closed, connectivity)
    counts = np.bincount(labels.ravel(), minlength=k + 1)[1:].astype(np.int64)
    kept, body = select_candidates(counts, geom.voxel_volume(), min_vol, max_vol)
    want = set(kept) | ({body} if body else set())
    if full_stats:
        stats = label_shape_statistics(labels, k, geom)
    else:
        # statistics only for the labels the reference reads beyond PhysicalSize
        sel = np.zeros(k + 1, dtype=np.uint32)
        order = sorted(want)
        sel[order] = np.arange(1, len(order) + 1, dtype=np.uint32)
        sub = sel[labels]
        stats = label_shape_statistics(sub, len(order), geom)
        for s, lbl in zip(stats, order):
            s.label = lbl
    by_label = {s.label: s for s in stats}
    fiducials = [{"vol": by_label[l].physical_size,
                  "centroid": tuple(float(c) for c in by_label[l].centroid),
                  "id": int(l)} for l in kept]
    ras = np.array([[-f["centroid"][0], -f["centroid"][1], f["centroid"][2]] for f in fiducials],
                   dtype=np.float64).reshape(-1, 3)
    names = [f"M_{f['id']}_{f['vol']:.0f}mm³" for f in fiducials]
    body_mask = (labels == body).astype(np.uint8) if body else None
    return Detection(binary=binary, closed=closed, labels=labels, n_labels=k, counts=counts,
                     fiducials=fiducials, ras_points=ras, marker_labels=names,
                     body_label=body, body_mask=body_mask, stats=stats)
