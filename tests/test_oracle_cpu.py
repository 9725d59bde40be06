"""CPU tests: the oracle (numpy + C restatements) against the reference-generated golden
fixtures and against scikit-image's published known answers."""
import numpy as np
import pytest

from conftest import compare_tables, parity_distributions
from oracle import c_oracle
from oracle import notebook_oracle as orc


def test_greycomatrix_known_answer():
    # scikit-image docstring example
    img = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [0, 2, 2, 2], [2, 2, 3, 3]], dtype=np.uint8)
    P = orc.greycomatrix(img, [1], [0, np.pi / 4, np.pi / 2, 3 * np.pi / 4], levels=4)
    want = [
        [[2, 2, 1, 0], [0, 2, 0, 0], [0, 0, 3, 1], [0, 0, 0, 1]],
        [[1, 1, 3, 0], [0, 1, 1, 0], [0, 0, 0, 2], [0, 0, 0, 0]],
        [[3, 0, 2, 0], [0, 2, 2, 0], [0, 0, 1, 2], [0, 0, 0, 0]],
        [[2, 0, 0, 0], [1, 1, 2, 0], [0, 0, 2, 1], [0, 0, 0, 0]],
    ]
    for a in range(4):
        assert (P[:, :, 0, a] == np.array(want[a])).all()


def test_greycoprops_known_answer():
    # scikit-image test_texture.py values (symmetric, normed, distances [1, 2], angle 0)
    img = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [0, 2, 2, 2], [2, 2, 3, 3]], dtype=np.uint8)
    P = orc.greycomatrix(img, [1, 2], [0], levels=4, symmetric=True, normed=True)
    want = {"contrast": (0.58333333, 1.25), "dissimilarity": (0.41666667, 1.0),
            "homogeneity": (0.80833333, 0.525), "ASM": (0.14583333, 0.1796875),
            "energy": (0.38188131, 0.42389562), "correlation": (0.71953255, 0.41176470)}
    for prop, (a, b) in want.items():
        got = orc.greycoprops(P, prop)
        np.testing.assert_allclose(got[:, 0], [a, b], rtol=2e-7)


def test_glcm_offsets():
    pi = np.pi
    assert [orc.glcm_offset(5, a) for a in (0, pi / 4, pi / 2, 3 * pi / 4)] == [(0, 5), (4, 4), (5, 0), (4, -4)]
    assert [orc.glcm_offset(1, a) for a in (0, pi / 4, pi / 2, 3 * pi / 4)] == [(0, 1), (1, 1), (1, 0), (1, -1)]


def test_glcm_edge_cases():
    # zero pairs -> 0,0,0,0,0,1 ; constant / all-zero -> 0,0,1,1,1,1
    assert orc.glcm_values(np.full((9, 5), 7, np.uint16)) == [0, 0, 0, 0, 0, 1]
    assert orc.glcm_values(np.full((16, 16), 9, np.uint16)) == [0, 0, 1, 1, 1, 1]
    assert orc.glcm_values(np.zeros((16, 16), np.uint16)) == [0, 0, 1, 1, 1, 1]


def test_numpy_oracle_matches_reference_golden(golden):
    """Basic columns: bit-identical to the reference's own cell 13 (same numpy/scipy calls)."""
    for name, (img, want, cols) in golden.items():
        table, ocols = orc.oracle_extract([img])
        assert ocols == cols, name
        g, w = table[0], want
        same = (g == w) | (np.isnan(g) & np.isnan(w))
        assert same.all(), (name, [cols[i] for i in np.flatnonzero(~same)][:5])


def test_feature_count_matches_notebook(golden):
    # NB:317 prints 69 features for 3 channels
    img, want, cols = golden["blob_64x64x3"]
    assert len(cols) == 69 and cols[0] == "min_intensity_Ch1" and cols[51] == "contrast_Ch1"


def _planar(img):
    return np.ascontiguousarray(img.transpose(2, 0, 1))[None]


def test_c_oracle_matches_reference_golden(golden):
    for name, (img, want, cols) in golden.items():
        got = c_oracle.table(_planar(img))
        compare_tables(got, want[None], cols, rtol=1e-10, atol=1e-10, label=name)


def test_c_oracle_matches_numpy_oracle_extensions():
    rng = np.random.default_rng(5)
    for (h, w) in [(64, 64), (37, 91), (20, 9)]:
        d = parity_distributions(rng, h, w)
        names = sorted(d)
        img = np.stack([d[k] for k in names], axis=2)
        yy, xx = np.mgrid[0:h, 0:w]
        masks = []
        for k in range(len(names)):
            m = ((yy - h / 2) ** 2 / (h / (2.2 + 0.3 * k)) ** 2 + (xx - w / 2) ** 2 / (w / 2.5) ** 2) < 1
            if k == 3:
                m[:] = False                      # empty mask
            if k == 4:
                m[:] = True
            masks.append(m)
        mask = np.stack(masks, axis=2).astype(np.uint8)
        want, cols = orc.oracle_extract([img], [mask], glcm=True, four_directions=True, shape=True,
                                        moments=True)
        got = c_oracle.table(_planar(img), _planar(mask), glcm=True, n_angles=4, shape=True,
                             moments=True)
        compare_tables(got, want, cols, rtol=1e-9, atol=1e-9, label="%dx%d" % (h, w))
        # unmasked, 4 directions
        want, cols = orc.oracle_extract([img], glcm=True, four_directions=True, shape=True, moments=True)
        got = c_oracle.table(_planar(img), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(got, want, cols, rtol=1e-9, atol=1e-9, label="nomask %dx%d" % (h, w))


def test_c_oracle_glcm_counts_bit_exact():
    rng = np.random.default_rng(9)
    plane = rng.integers(0, 4096, (33, 47)).astype(np.uint16)
    want = orc.glcm_counts(plane, angles=orc.ANGLES4)
    got = c_oracle.glcm_counts(plane, n_angles=4)
    for a in range(4):
        assert (got[a] == want[:, :, a]).all()
    mask = (rng.random((33, 47)) < 0.6)
    want = orc.glcm_counts(np.where(mask, plane, 0), angles=orc.ANGLES4, pair_mask=mask,
                           vmax=plane[mask].max())
    got = c_oracle.glcm_counts(plane, mask.astype(np.uint8), n_angles=4)
    for a in range(4):
        assert (got[a] == want[:, :, a]).all()


def test_quantiser_multiply_shift_is_exact():
    """k3_quant (csrc/k3_glcm.cuh): (255*x*ceil(2^(24+l)/max)) >> (24+l) == the notebook's float64
    expression (x/max)*255 -> uint8 (NB:294-295), checked for all x on a spread of maxima."""
    rng = np.random.default_rng(1)
    maxima = np.unique(np.concatenate([np.arange(1, 600), 2 ** np.arange(1, 17) - 1, 2 ** np.arange(1, 16),
                                       2 ** np.arange(1, 16) + 1, rng.integers(600, 65536, 900), [65535, 65534, 4095]]))
    for vmax in maxima:
        vmax = int(vmax)
        x = np.arange(0, vmax + 1, dtype=np.uint64)
        l = 0 if vmax <= 1 else int(vmax - 1).bit_length()
        sh = 24 + l
        mul = -(-(1 << sh) // vmax)
        assert mul == int(np.ceil(np.ldexp(1.0, sh) / float(vmax)))   # the device computes it in double
        assert mul < 2 ** 32
        q = (x * np.uint64(255) * np.uint64(mul)) >> np.uint64(sh)
        ref = ((x.astype(np.float64) / float(vmax)) * 255).astype(np.uint8)
        assert (q == ref).all(), vmax


def test_quantiser_single_multiply_is_exact():
    """k3_magic_fast (csrc/k3_glcm.cuh): for 256 < max <= 4103 the quantiser is one multiply-high,
    umulhi(x, 255 * (floor((2^32 - 1) / max) + 1)); checked against the notebook's float64 expression
    (NB:294-295) for every x <= max and every maximum of the range."""
    for vmax in range(257, 4104):
        mul = 255 * (0xFFFFFFFF // vmax + 1)
        assert mul < 2 ** 32
        x = np.arange(0, vmax + 1, dtype=np.uint64)
        q = (x * np.uint64(mul)) >> np.uint64(32)
        ref = ((x.astype(np.float64) / float(vmax)) * 255).astype(np.uint8)
        assert (q == ref).all(), vmax


# ---- x2 / x3: the extension blocks against independent third-party implementations ---------------------------
def _cv2_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_moments_golden.npz"))


def test_moments_match_opencv_golden():
    """x3 (spatial moments): centroid and the seven normalised central moments of the masked plane against
    OpenCV's cv2.moments (fixture: tests/golden/make_cv2_golden.py).  OpenCV's x is the column, y the row."""
    from imfeat_b200 import synth
    g = _cv2_golden()
    K = {k: i for i, k in enumerate(g["keys"].tolist())}
    for case, wm in zip(g["cases"].tolist(), g["weighted"]):
        seed, obj, ch, h, w, shrink = case
        px, mk = synth.synth_plane(seed, obj, ch, h, w, shrink)
        got = orc.moment_values(px, mk)
        want = [wm[K["m01"]] / wm[K["m00"]], wm[K["m10"]] / wm[K["m00"]], wm[K["nu02"]], wm[K["nu11"]], wm[K["nu20"]],
                wm[K["nu03"]], wm[K["nu12"]], wm[K["nu21"]], wm[K["nu30"]]]
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-13, err_msg=str(case))
        # the C restatement follows the same definition
        tab = c_oracle.table(px[None, None].astype(np.uint16), mk[None, None], glcm=False, shape=True, moments=True)
        np.testing.assert_allclose(tab[0, -9:], want, rtol=1e-9, atol=1e-13, err_msg=str(case))
        np.testing.assert_allclose(tab[0, -19:-9], orc.shape_values(mk), rtol=1e-9, atol=1e-12, err_msg=str(case))


def test_shape_matches_opencv_golden():
    """x2 (region descriptors): area, centroid, bounding box and extent against OpenCV; major / minor axis and
    eccentricity from OpenCV's central moments of the mask through scikit-image's published regionprops formulas
    (inertia tensor [[mu02, -mu11], [-mu11, mu20]] / area, axes 4 sqrt(eigenvalue), ecc = sqrt(1 - l2 / l1))."""
    from imfeat_b200 import synth
    g = _cv2_golden()
    K = {k: i for i, k in enumerate(g["keys"].tolist())}
    for case, bm, rect in zip(g["cases"].tolist(), g["binary"], g["rect"].tolist()):
        seed, obj, ch, h, w, shrink = case
        _, mk = synth.synth_plane(seed, obj, ch, h, w, shrink)
        got = dict(zip(orc.SHAPE_NAMES, orc.shape_values(mk)))
        area = bm[K["m00"]]
        assert got[orc.SHAPE_NAMES[0]] == area == float((mk > 0).sum())
        vals = orc.shape_values(mk)
        # order: area, perimeter, bbox_area, extent, centroid_r, centroid_c, major, minor, eccentricity, circularity
        assert vals[2] == float(rect[2] * rect[3])
        np.testing.assert_allclose(vals[3], area / float(rect[2] * rect[3]), rtol=1e-15)
        np.testing.assert_allclose(vals[4], bm[K["m01"]] / area, rtol=1e-12)
        np.testing.assert_allclose(vals[5], bm[K["m10"]] / area, rtol=1e-12)
        a, b, c = bm[K["mu20"]] / area, -bm[K["mu11"]] / area, bm[K["mu02"]] / area      # x = column
        ev = np.linalg.eigvalsh(np.array([[a, b], [b, c]]))
        l1, l2 = float(ev[1]), max(float(ev[0]), 0.0)
        np.testing.assert_allclose(vals[6], 4.0 * np.sqrt(l1), rtol=1e-9, err_msg=str(case))
        np.testing.assert_allclose(vals[7], 4.0 * np.sqrt(l2), rtol=1e-7, atol=1e-9, err_msg=str(case))
        np.testing.assert_allclose(vals[8], np.sqrt(max(1.0 - l2 / l1, 0.0)), rtol=1e-7, atol=1e-9, err_msg=str(case))


def test_perimeter_matches_scipy_ndimage_restatement():
    """x2 perimeter: scikit-image's measure.perimeter(neighborhood=4) is a few lines over scipy.ndimage
    (binary_erosion with the 4-connected cross and border_value=0, a 3x3 convolution with weights
    [[10, 2, 10], [2, 1, 2], [10, 2, 10]], a histogram and a weight table).  Those lines, on the real
    scipy.ndimage primitives, against the oracle's own erosion / convolution, on masks that touch the border,
    single pixels, lines, full planes and synthetic ellipses."""
    from scipy import ndimage as ndi
    from imfeat_b200 import synth

    def skimage_perimeter(image):
        strel = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)
        image = image.astype(np.uint8)
        eroded = ndi.binary_erosion(image, strel, border_value=0)
        border = image - eroded
        weights = np.zeros(50, dtype=np.float64)
        weights[[5, 7, 15, 17, 25, 27]] = 1
        weights[[21, 33]] = np.sqrt(2)
        weights[[13, 23]] = (1 + np.sqrt(2)) / 2
        conv = ndi.convolve(border, np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]]), mode="constant", cval=0)
        return float(np.bincount(conv.ravel(), minlength=50) @ weights)

    rng = np.random.default_rng(5)
    masks = [np.ones((7, 9), np.uint8), np.zeros((5, 5), np.uint8), np.eye(6, dtype=np.uint8)]
    one = np.zeros((5, 6), np.uint8); one[2, 3] = 1
    line = np.zeros((6, 8), np.uint8); line[3, 1:7] = 1
    masks += [one, line, (rng.random((33, 47)) < 0.6).astype(np.uint8), (rng.random((64, 64)) < 0.9).astype(np.uint8)]
    masks += [synth.synth_plane(3, k, 0, 64, 64, 256 - 40 * k)[1] for k in range(4)]
    # a square, exactly: 4 straight sides of the border ring
    sq = np.zeros((10, 10), np.uint8); sq[2:8, 2:8] = 1
    masks.append(sq)
    for m in masks:
        n1, n2, n3 = orc.perimeter_classes(m)
        got = n1 + n2 * np.sqrt(2.0) + n3 * (1.0 + np.sqrt(2.0)) / 2.0
        np.testing.assert_allclose(got, skimage_perimeter(m), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(orc.shape_values(m)[1], skimage_perimeter(m), rtol=1e-13, atol=1e-13)
