"""CPU oracle for the per-object, per-channel hand-crafted feature extraction.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, call or
link this module.  The only permitted users are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

What it restates (``NB:<n>`` = raw JSON line n of
``/root/reference/channel_importance_hand_crafted_features.ipynb``):

* ``basic_statistical_features``  -> NB:220-264  (17 scalars per channel)
* ``glcm_features``               -> NB:269-308  (6 scalars per channel)
* the extraction loop / table     -> NB:327-334, NB:358-364

The notebook calls three scikit-image functions.  scikit-image is not installed in
this image (and cannot be fetched), so their published algorithms are restated here:

* ``shannon_entropy``  (skimage.measure, used at NB:262)
* ``greycomatrix``     (skimage.feature, used at NB:298; later renamed graycomatrix)
* ``greycoprops``      (skimage.feature, used at NB:301-306)

The notebook pins no versions (no requirements file); the era evidence points at
scikit-image ~0.17.  The numpy / scipy calls are the very same library calls the
notebook makes (numpy 2.3.5 / scipy 1.18.1 in this image).

Pinning status
--------------
* basic features: PINNED.  ``tests/golden/make_golden.py`` executes the notebook's own
  cell 13 (loaded from /root/reference at generation time) on seeded inputs and the
  committed fixtures are compared with this module in ``tests/test_oracle_cpu.py``.
* GLCM features: pinned against scikit-image's published docstring example for
  ``greycomatrix`` and its ``test_texture.py`` property values, and
  against cell 13 executed with the restated skimage functions injected.  The real
  scikit-image binary was never run here, so this half is "parity pinned to published
  known answers only".
* masked statistics, 4-direction GLCM, shape descriptors, spatial moments: the
  reference has no code for them, so there is no
  reference output to compare with.  What pins the functions below instead:
  - spatial moments (x3) and the region's area, centroid, bounding box, extent and
    second-order central moments (x2): OpenCV's ``cv2.moments`` / ``cv2.boundingRect``, a
    third-party implementation of the same published definitions, through the committed
    fixture ``tests/golden/cv2_moments_golden.npz`` (``tests/golden/make_cv2_golden.py``);
    major / minor axis and eccentricity follow from those moments by scikit-image's
    published regionprops formulas;
  - perimeter (x2): scikit-image's ``measure.perimeter`` restated line by line over the real
    ``scipy.ndimage`` primitives it calls (``tests/test_oracle_cpu.py``);
  - 4-direction GLCM (x4): scikit-image's four-angle docstring example;
  - masked statistics (x1): numpy / scipy on ``x[mask > 0]`` -- the reference's own calls on
    the selected pixels; a definition, not a comparison: "parity unpinned" for this row.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import stats as _stats

# ----------------------------------------------------------------------------------
# Schema (NB:241-262 for the basic block, NB:301-306 for the GLCM block)
# ----------------------------------------------------------------------------------
BASIC_NAMES = (
    ["min_intensity"]
    + ["percentile%d0_intensity" % k for k in range(1, 10)]
    + ["max_intensity", "total_intensity", "mean_intensity", "std_intensity",
       "kurtosis_intensity", "skew_intensity", "shannon_entropy"]
)
GLCM_PROPS = ["contrast", "dissimilarity", "homogeneity", "ASM", "energy", "correlation"]
# NB:242-250: the arguments really are 0.1 ... 0.9 *percent*.
NOTEBOOK_PERCENTILES = tuple(k / 10.0 for k in range(1, 10))
# Extension blocks (no reference counterpart)
ANGLE_TAGS = ["", "_a45", "_a90", "_a135"]
ANGLES4 = (0.0, np.pi / 4, np.pi / 2, 3 * np.pi / 4)
SHAPE_NAMES = [
    "area", "perimeter", "bbox_area", "extent", "centroid_row", "centroid_col",
    "major_axis_length", "minor_axis_length", "eccentricity", "circularity",
]
MOMENT_NAMES = [
    "weighted_centroid_row", "weighted_centroid_col",
    "nu20", "nu11", "nu02", "nu30", "nu21", "nu12", "nu03",
]


def column_names(n_channels, glcm=True, four_directions=False, shape=False, moments=False):
    """Column order of the feature table (NB:330-331, NB:334: dict insertion order)."""
    cols = []
    for ch in range(n_channels):
        cols += ["%s_Ch%d" % (nm, ch + 1) for nm in BASIC_NAMES]
    if glcm:
        tags = ANGLE_TAGS if four_directions else ANGLE_TAGS[:1]
        for ch in range(n_channels):
            for tag in tags:
                cols += ["%s%s_Ch%d" % (p, tag, ch + 1) for p in GLCM_PROPS]
    if shape:
        for ch in range(n_channels):
            cols += ["%s_Ch%d" % (nm, ch + 1) for nm in SHAPE_NAMES]
    if moments:
        for ch in range(n_channels):
            cols += ["%s_Ch%d" % (nm, ch + 1) for nm in MOMENT_NAMES]
    return cols


# ----------------------------------------------------------------------------------
# scikit-image restatements
# ----------------------------------------------------------------------------------
def shannon_entropy(image, base=2):
    """skimage.measure.shannon_entropy: entropy of the distinct raw values."""
    _, counts = np.unique(image, return_counts=True)
    return _stats.entropy(counts, base=base)


def glcm_offset(distance, angle):
    """(d_row, d_col) exactly as skimage's ``_glcm_loop`` derives it (C ``round``)."""
    def c_round(v):
        return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    return c_round(math.sin(angle) * distance), c_round(math.cos(angle) * distance)


def greycomatrix(image, distances, angles, levels=256, symmetric=False, normed=False,
                 pair_mask=None):
    """skimage.feature.greycomatrix restated.

    ``pair_mask`` is an extension (boolean (h,w)): a pair is counted only when both of
    its pixels are inside the mask.
    """
    image = np.ascontiguousarray(image)
    if image.ndim != 2:
        raise ValueError("greycomatrix wants a 2-D image")
    if not np.issubdtype(image.dtype, np.integer):
        raise ValueError("greycomatrix wants an integer image")
    if image.size and int(image.max()) >= levels:
        raise ValueError("image maximum must be smaller than levels")
    h, w = image.shape
    out = np.zeros((levels, levels, len(distances), len(angles)), dtype=np.uint32)
    for ai, angle in enumerate(angles):
        for di, dist in enumerate(distances):
            dr, dc = glcm_offset(dist, angle)
            r0, r1 = max(0, -dr), min(h, h - dr)
            c0, c1 = max(0, -dc), min(w, w - dc)
            if r1 <= r0 or c1 <= c0:
                continue
            a = image[r0:r1, c0:c1]
            b = image[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
            if pair_mask is not None:
                keep = pair_mask[r0:r1, c0:c1] & pair_mask[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
                a, b = a[keep], b[keep]
            flat = a.astype(np.int64).ravel() * levels + b.astype(np.int64).ravel()
            counts = np.bincount(flat, minlength=levels * levels)
            out[:, :, di, ai] = counts.reshape(levels, levels).astype(np.uint32)
    if symmetric:
        out = out + out.transpose(1, 0, 2, 3)
    if normed:
        out = out.astype(np.float64)
        sums = out.sum(axis=(0, 1), keepdims=True)
        sums[sums == 0] = 1
        out = out / sums
    return out


def greycoprops(P, prop="contrast"):
    """skimage.feature.greycoprops restated.  Returns [n_dist, n_angle]."""
    P = np.asarray(P)
    if P.ndim != 4 or P.shape[0] != P.shape[1]:
        raise ValueError("P must be [levels, levels, n_dist, n_angle]")
    levels = P.shape[0]
    P = P.astype(np.float64)
    sums = P.sum(axis=(0, 1), keepdims=True)
    sums[sums == 0] = 1
    P = P / sums
    I, J = np.ogrid[0:levels, 0:levels]
    if prop in ("contrast", "dissimilarity", "homogeneity"):
        if prop == "contrast":
            weights = (I - J) ** 2
        elif prop == "dissimilarity":
            weights = np.abs(I - J)
        else:
            weights = 1.0 / (1.0 + (I - J) ** 2)
        weights = weights.reshape(levels, levels, 1, 1)
        return (P * weights).sum(axis=(0, 1))
    if prop == "ASM":
        return (P ** 2).sum(axis=(0, 1))
    if prop == "energy":
        return np.sqrt((P ** 2).sum(axis=(0, 1)))
    if prop == "correlation":
        I = I.reshape(levels, 1, 1, 1).astype(np.float64)
        J = J.reshape(1, levels, 1, 1).astype(np.float64)
        mu_i = (I * P).sum(axis=(0, 1))
        mu_j = (J * P).sum(axis=(0, 1))
        di = I - mu_i
        dj = J - mu_j
        std_i = np.sqrt((P * di ** 2).sum(axis=(0, 1)))
        std_j = np.sqrt((P * dj ** 2).sum(axis=(0, 1)))
        cov = (P * (di * dj)).sum(axis=(0, 1))
        res = np.ones_like(cov)
        ok = ~((std_i < 1e-15) | (std_j < 1e-15))
        res[ok] = cov[ok] / (std_i[ok] * std_j[ok])
        return res
    raise ValueError("unknown GLCM property %r" % (prop,))


# ----------------------------------------------------------------------------------
# Notebook feature functions (cell 13)
# ----------------------------------------------------------------------------------
def basic_values(values):
    """The 17 basic scalars of one channel (NB:241-262), for any array of pixel values."""
    v = np.asarray(values)
    out = [v.min()]
    out += [np.percentile(v, q) for q in NOTEBOOK_PERCENTILES]
    out += [v.max(), v.sum(), v.mean(), v.std(),
            _stats.kurtosis(v.ravel()), _stats.skew(v.ravel()), shannon_entropy(v)]
    return [float(x) for x in out]


def basic_statistical_features(image):
    """Restatement of NB:220-264: dict of 17 features per channel, notebook key order."""
    feats = {}
    for ch in range(image.shape[2]):
        vals = basic_values(image[:, :, ch])
        for name, val in zip(BASIC_NAMES, vals):
            feats["%s_Ch%d" % (name, ch + 1)] = val
    return feats


def quantise(plane, vmax=None):
    """NB:293-295: (x / max) * 255 -> uint8 (truncation); 0/0 -> NaN -> 0."""
    t = np.array(plane, copy=True)
    with np.errstate(all="ignore"):
        t = (t / (t.max() if vmax is None else vmax)) * 255
        t = np.nan_to_num(t, nan=0.0).astype("uint8")
    return t


def glcm_values(plane, angles=(0.0,), distance=5, levels=256, pair_mask=None, vmax=None):
    """6 GLCM properties per angle for one channel (NB:293-306)."""
    q = quantise(plane, vmax)
    P = greycomatrix(q, distances=[distance], angles=list(angles), levels=levels,
                     pair_mask=pair_mask)
    out = []
    for ai in range(len(angles)):
        Pa = P[:, :, :, ai:ai + 1]
        out += [float(greycoprops(Pa, prop=p)[0, 0]) for p in GLCM_PROPS]
    return out


def glcm_features(image):
    """Restatement of NB:269-308: dict of 6 features per channel."""
    feats = {}
    for ch in range(image.shape[2]):
        vals = glcm_values(image[:, :, ch])
        for name, val in zip(GLCM_PROPS, vals):
            feats["%s_Ch%d" % (name, ch + 1)] = val
    return feats


def glcm_counts(plane, angles=(0.0,), distance=5, levels=256, pair_mask=None, vmax=None):
    """uint32[levels, levels, n_angles] bins of one channel (the 'GLCM bins' parity target)."""
    q = quantise(plane, vmax)
    return greycomatrix(q, [distance], list(angles), levels=levels, pair_mask=pair_mask)[:, :, 0, :]


# ----------------------------------------------------------------------------------
# Extension specifications (no reference counterpart -- SURVEY.md A.6)
# ----------------------------------------------------------------------------------
def masked_basic_values(plane, mask):
    """x1: NB:241-262 applied to the multiset plane[mask > 0]; empty mask -> all NaN."""
    sel = np.asarray(plane)[np.asarray(mask) > 0]
    if sel.size == 0:
        return [float("nan")] * len(BASIC_NAMES)
    return basic_values(sel)


def masked_glcm_values(plane, mask, angles=(0.0,), distance=5, levels=256):
    """x1/x4: quantise with the max over masked pixels; count a pair only if both pixels
    are inside the mask.  Empty mask -> GLCM of zero pairs (0,0,0,0,0,1)."""
    m = np.asarray(mask) > 0
    plane = np.asarray(plane)
    if not m.any():
        vmax = 0
    else:
        vmax = plane[m].max()
    # pixels outside the mask never enter a pair, so their quantised value is irrelevant;
    # zero them so that the uint8 cast is defined.
    work = np.where(m, plane, 0)
    return glcm_values(work, angles=angles, distance=distance, levels=levels,
                       pair_mask=m, vmax=vmax)


_PERIM_W = np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]], dtype=np.int64)


def perimeter_classes(mask):
    """Counts (n_straight, n_diagonal, n_corner) of skimage.measure.perimeter(neighbourhood=4).

    border = mask minus its 4-connected erosion (pixels outside the image count as
    background); each border pixel is classified by the weighted sum of its 3x3 border
    neighbourhood: {5,7,15,17,25,27} -> 1, {21,33} -> sqrt(2), {13,23} -> (1+sqrt(2))/2.
    """
    m = (np.asarray(mask) > 0).astype(np.int64)
    h, w = m.shape
    p = np.zeros((h + 2, w + 2), dtype=np.int64)
    p[1:-1, 1:-1] = m
    er = p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
    border = m - er
    b = np.zeros((h + 2, w + 2), dtype=np.int64)
    b[1:-1, 1:-1] = border
    conv = np.zeros((h, w), dtype=np.int64)
    for dr in range(3):
        for dc in range(3):
            conv += _PERIM_W[dr, dc] * b[dr:dr + h, dc:dc + w]
    hist = np.bincount(conv.ravel(), minlength=50)
    n1 = int(hist[[5, 7, 15, 17, 25, 27]].sum())
    n2 = int(hist[[21, 33]].sum())
    n3 = int(hist[[13, 23]].sum())
    return n1, n2, n3


def shape_values(mask):
    """x2: shape descriptors of one mask plane (all mask pixels form one region, as in
    skimage.measure.regionprops with a single label).  Order = SHAPE_NAMES.

    Everything is derived from exact integer sums (area, sum r, sum c, sum r^2, sum c^2,
    sum r*c, bounding box, perimeter class counts) with cancellation-free formulas, so a
    second implementation can agree to ~1e-15:
      inertia tensor [[a, b], [b, c]] = [[mu02, -mu11], [-mu11, mu20]] / area
      l1 = (a+c)/2 + D,  D = sqrt(((a-c)/2)^2 + b^2),  l2 = (a*c - b^2) / l1
      major/minor axis = 4*sqrt(l1), 4*sqrt(l2);  eccentricity = sqrt(2*D/l1)
    Empty mask -> area 0, perimeter 0, the rest NaN."""
    m = np.asarray(mask) > 0
    area = int(m.sum())
    n1, n2, n3 = perimeter_classes(m)
    perim = n1 + n2 * math.sqrt(2.0) + n3 * ((1.0 + math.sqrt(2.0)) / 2.0)
    if area == 0:
        return [0.0, 0.0] + [float("nan")] * (len(SHAPE_NAMES) - 2)
    rr, cc = np.nonzero(m)
    rr, cc = rr.astype(np.int64), cc.astype(np.int64)      # exact integer sums (< 2^63 for any plane)
    bbox_area = float((int(rr.max()) - int(rr.min()) + 1) * (int(cc.max()) - int(cc.min()) + 1))
    extent = area / bbox_area
    sr, sc = int(rr.sum()), int(cc.sum())
    srr, scc, src = int((rr * rr).sum()), int((cc * cc).sum()), int((rr * cc).sum())
    a2 = float(area) * float(area)
    a = float(area * scc - sc * sc) / a2            # mu02 / area  (column variance)
    c = float(area * srr - sr * sr) / a2            # mu20 / area  (row variance)
    b = -float(area * src - sr * sc) / a2           # -mu11 / area
    half_diff = float((area * scc - sc * sc) - (area * srr - sr * sr)) / a2 * 0.5
    D = math.sqrt(half_diff * half_diff + b * b)
    l1 = (a + c) * 0.5 + D
    if l1 > 0:
        l2 = max((a * c - b * b) / l1, 0.0)
        ecc = math.sqrt(min(max(2.0 * D / l1, 0.0), 1.0))
    else:
        l2, ecc = 0.0, 0.0
    major = 4.0 * math.sqrt(l1)
    minor = 4.0 * math.sqrt(l2)
    circ = 4.0 * math.pi * area / (perim * perim) if perim > 0 else float("nan")
    return [float(area), perim, bbox_area, extent, sr / float(area), sc / float(area),
            major, minor, ecc, circ]


def moment_values(plane, mask=None):
    """x3: intensity-weighted spatial moments (order <= 3) of one channel.
    Order = MOMENT_NAMES.  Zero total weight -> all NaN."""
    img = np.asarray(plane).astype(np.float64)
    if mask is not None:
        img = np.where(np.asarray(mask) > 0, img, 0.0)
    h, w = img.shape
    r = np.arange(h, dtype=np.float64)[:, None]
    c = np.arange(w, dtype=np.float64)[None, :]
    m00 = img.sum()
    if m00 == 0:
        return [float("nan")] * len(MOMENT_NAMES)
    cr = (r * img).sum() / m00
    cc = (c * img).sum() / m00
    dr, dc = r - cr, c - cc

    def nu(p, q):
        mu = ((dr ** p) * (dc ** q) * img).sum()
        return mu / m00 ** ((p + q) / 2.0 + 1.0)

    return [cr, cc, nu(2, 0), nu(1, 1), nu(0, 2), nu(3, 0), nu(2, 1), nu(1, 2), nu(0, 3)]


# ----------------------------------------------------------------------------------
# Table assembly (NB:358-364)
# ----------------------------------------------------------------------------------
def extract_object(image, mask=None, glcm=True, four_directions=False, shape=False,
                   moments=False):
    """One table row (float64[F_total]) for one (h,w,C) object."""
    C = image.shape[2]
    row = []
    for ch in range(C):
        if mask is None:
            row += basic_values(image[:, :, ch])
        else:
            row += masked_basic_values(image[:, :, ch], mask[:, :, ch])
    if glcm:
        angles = ANGLES4 if four_directions else ANGLES4[:1]
        for ch in range(C):
            if mask is None:
                row += glcm_values(image[:, :, ch], angles=angles)
            else:
                row += masked_glcm_values(image[:, :, ch], mask[:, :, ch], angles=angles)
    if shape:
        for ch in range(C):
            m = np.ones(image.shape[:2], bool) if mask is None else mask[:, :, ch]
            row += shape_values(m)
    if moments:
        for ch in range(C):
            row += moment_values(image[:, :, ch], None if mask is None else mask[:, :, ch])
    return np.asarray(row, dtype=np.float64)


def oracle_extract(images, masks=None, **opts):
    """float64[N, F_total] table + column names for a sequence of (h,w,C) objects."""
    rows = []
    for i, img in enumerate(images):
        m = None if masks is None else masks[i]
        with np.errstate(all="ignore"):
            rows.append(extract_object(np.asarray(img), None if m is None else np.asarray(m),
                                       **opts))
    table = np.vstack(rows) if rows else np.zeros((0, 0))
    C = np.asarray(images[0]).shape[2] if len(images) else 0
    return table, column_names(C, **{k: v for k, v in opts.items()})
