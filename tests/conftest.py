import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


INT_EXACT = ("min_intensity", "max_intensity", "total_intensity", "area", "bbox_area")
BIT_EXACT = tuple("percentile%d0_intensity" % k for k in range(1, 10))


_ANGLE_OF_TAG = {"": 0, "_a45": 1, "_a90": 2, "_a135": 3}
_OFFSETS = {5: [(0, 5), (4, 4), (5, 0), (4, -4)]}


def corr_is_degenerate(image, mask, ch, angle, distance=5):
    """Exact test for greycoprops' special case (NB:306): is one marginal of the GLCM of channel `ch`, direction
    `angle`, constant -- i.e. do all pairs share their left or their right gray level?  Integer arithmetic only
    (quantiser floor(255*x/max), NB:294-295), so there is no rounding to argue about."""
    plane = np.asarray(image)[:, :, ch].astype(np.int64)
    m = None if mask is None else (np.asarray(mask)[:, :, ch] != 0)
    vmax = int(plane[m].max()) if (m is not None and m.any()) else (int(plane.max()) if m is None else 0)
    q = plane * 255 // vmax if vmax > 0 else np.zeros_like(plane)
    if distance in _OFFSETS:
        dr, dc = _OFFSETS[distance][angle]
    else:
        import math
        a = [0.0, math.pi / 4, math.pi / 2, 3 * math.pi / 4][angle]
        rnd = lambda v: int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))
        dr, dc = rnd(math.sin(a) * distance), rnd(math.cos(a) * distance)
    h, w = q.shape
    r1, c0, c1 = h - dr, max(0, -dc), w - max(0, dc)
    if r1 <= 0 or c1 <= c0:
        return False
    I, J = q[0:r1, c0:c1], q[dr:dr + r1, c0 + dc:c1 + dc]
    if m is not None:
        ok = m[0:r1, c0:c1] & m[dr:dr + r1, c0 + dc:c1 + dc]
        I, J = I[ok], J[ok]
    if I.size == 0:
        return False
    return bool(I.min() == I.max() or J.min() == J.max())


def compare_tables(got, want, cols, rtol=1e-9, atol=1e-9, label="", images=None, masks=None, distance=5):
    """Integer-valued columns and percentiles must match bit for bit; floating columns within
    rtol (north_star allows 1e-5; the kernels are held to 1e-9 here).

    images / masks (sequences of (h, w, C) objects, row i = object i) are only needed by batches that
    contain planes with a degenerate GLCM correlation, see below."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.shape[1] == len(cols)
    bad = []
    for j, name in enumerate(cols):
        g, w = got[:, j], want[:, j]
        base = name.rsplit("_Ch", 1)[0]
        if base in INT_EXACT or base in BIT_EXACT:
            ok = (g == w) | (np.isnan(g) & np.isnan(w))
        else:
            ok = np.isclose(g, w, rtol=rtol, atol=atol, equal_nan=True)
            if base.startswith("correlation") and not ok.all() and images is not None:
                # greycoprops returns 1 when a marginal std is < 1e-15.  When exactly one marginal
                # is constant the CPU float path sometimes misses that test by rounding (std ~1e-14)
                # and returns rounding noise ~0 instead; the kernel decides from exact integer
                # variances and returns 1 ("degenerate GLCM").  That signature is accepted
                # for a cell only if the marginal of THAT plane and direction really is constant,
                # established here in integer arithmetic -- a kernel returning 1 for a genuinely
                # tiny correlation still fails.
                ch = int(name.rsplit("_Ch", 1)[1]) - 1
                angle = _ANGLE_OF_TAG[base[len("correlation"):]]
                for i in np.flatnonzero(~ok):
                    if g[i] == 1.0 and abs(w[i]) < 1e-6 and corr_is_degenerate(
                            images[i], None if masks is None else masks[i], ch, angle, distance):
                        ok[i] = True
        if not ok.all():
            i = int(np.flatnonzero(~ok)[0])
            bad.append("%s row %d: got %r want %r" % (name, i, g[i], w[i]))
    assert not bad, "%s %d bad columns, first: %s" % (label, len(bad), "; ".join(bad[:6]))


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "notebook_golden.npz")
    z = np.load(path)
    names = sorted(k[3:] for k in z.files if k.startswith("in_"))
    return {n: (z["in_" + n], z["out_" + n], [str(c) for c in z["cols_" + n]]) for n in names}


def parity_distributions(rng, h, w):
    """The parity distributions of SURVEY.md 8(d), as a dict name -> uint16 [h, w]."""
    yy, xx = np.mgrid[0:h, 0:w]
    out = {
        "uniform12": rng.integers(0, 4096, (h, w)),
        "full16": rng.integers(0, 65536, (h, w)),
        "ties8": rng.choice([3, 17, 17, 250, 900, 901, 4000, 65535], (h, w)),
        "halfzero": np.clip(rng.normal(0, 300, (h, w)), 0, 65535),
        "constant": np.full((h, w), 1234),
        "allzero": np.zeros((h, w)),
        "saturated": np.full((h, w), 65535),
        "gradient": (yy * 37 + xx * 11) % 65536,
        "poisson_blob": np.clip(rng.poisson(400, (h, w)) + 3000 * np.clip(
            1 - ((yy - h / 2.1) / (h / 3.0)) ** 2 - ((xx - w / 1.9) / (w / 4.0)) ** 2, 0, None), 0, 65535),
        "two_level": np.where(rng.random((h, w)) < 0.999, 100, 65535),
        "low_outlier": np.where(rng.random((h, w)) < 0.002, 0, rng.integers(30000, 30100, (h, w))),
    }
    return {k: v.astype(np.uint16) for k, v in out.items()}
