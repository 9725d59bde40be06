"""Generate tests/golden/notebook_golden.npz by running the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

What it does: loads cell 13 of
/root/reference/channel_importance_hand_crafted_features.ipynb (the cell that defines
``basic_statistical_features`` NB:220-264 and ``glcm_features`` NB:269-308) as text,
``exec``s it unmodified, and calls the two functions on seeded uint16 (h,w,C) objects.
The numpy / scipy calls inside are the real libraries.  scikit-image is not installed
here, so ``skimage.feature`` / ``skimage.measure`` are provided as stub modules backed
by the restatements in ``oracle/notebook_oracle.py`` (validated against scikit-image's
published known answers in tests/test_oracle_cpu.py).  Consequently the
17 basic columns of the fixture are reference-executed end to end; the 6 GLCM columns
are reference control flow (quantiser, argument literals, key order) over restated
skimage kernels.
"""
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
NB = "/root/reference/channel_importance_hand_crafted_features.ipynb"


def load_reference_functions():
    from oracle import notebook_oracle as orc
    feat = types.ModuleType("skimage.feature")
    feat.greycomatrix = orc.greycomatrix
    feat.greycoprops = orc.greycoprops
    meas = types.ModuleType("skimage.measure")
    meas.shannon_entropy = orc.shannon_entropy
    pkg = types.ModuleType("skimage")
    pkg.feature, pkg.measure = feat, meas
    sys.modules.update({"skimage": pkg, "skimage.feature": feat, "skimage.measure": meas})
    nb = json.load(open(NB))
    src = "".join(nb["cells"][13]["source"])
    assert "def basic_statistical_features" in src and "def glcm_features" in src
    ns = {}
    warnings.filterwarnings("ignore")          # the notebook does this globally (NB:23)
    exec(compile(src, NB + ":cell13", "exec"), ns)
    return ns["basic_statistical_features"], ns["glcm_features"]


def make_inputs():
    """Seeded objects covering SURVEY.md 8(d)'s test distributions."""
    rng = np.random.default_rng(20261018)
    objs = {}

    def planes(h, w, gens):
        return np.stack([g(h, w) for g in gens], axis=2).astype(np.uint16)

    uni12 = lambda h, w: rng.integers(0, 4096, (h, w))
    full16 = lambda h, w: rng.integers(0, 65536, (h, w))
    ties8 = lambda h, w: rng.choice([3, 17, 17, 250, 900, 901, 4000, 65535], (h, w))
    halfzero = lambda h, w: np.clip(rng.normal(0, 300, (h, w)), 0, 65535)
    const = lambda h, w: np.full((h, w), 1234)
    zero = lambda h, w: np.zeros((h, w))
    sat = lambda h, w: np.full((h, w), 65535)

    def outlier(h, w):
        a = np.full((h, w), 100)
        a[h // 2, w // 3] = 65535
        return a

    def blob(h, w):
        yy, xx = np.mgrid[0:h, 0:w]
        d2 = ((yy - h / 2.1) / (h / 3.0)) ** 2 + ((xx - w / 1.9) / (w / 4.0)) ** 2
        return np.clip(rng.poisson(400, (h, w)) + 3000 * np.clip(1 - d2, 0, None), 0, 65535)

    def gradient(h, w):
        yy, xx = np.mgrid[0:h, 0:w]
        return (yy * 37 + xx * 11) % 65536

    objs["mix_64x64x4"] = planes(64, 64, [uni12, full16, ties8, halfzero])
    objs["degenerate_64x64x4"] = planes(64, 64, [const, zero, sat, outlier])
    objs["blob_64x64x3"] = planes(64, 64, [blob, gradient, blob])
    objs["rect_37x91x3"] = planes(37, 91, [uni12, blob, halfzero])
    objs["big_128x128x2"] = planes(128, 128, [full16, blob])
    objs["narrow_9x5x2"] = planes(9, 5, [uni12, ties8])       # w <= 5: zero GLCM pairs
    objs["narrow_7x6x2"] = planes(7, 6, [uni12, full16])      # exactly one pair per row
    objs["single_1x1x2"] = planes(1, 1, [uni12, const])
    objs["odd_33x17x3"] = planes(33, 17, [full16, halfzero, gradient])
    return objs


def main():
    basic, glcm = load_reference_functions()
    objs = make_inputs()
    out = {}
    for name, img in objs.items():
        feats = {}
        feats.update(basic(img))          # NB:362
        feats.update(glcm(img))           # NB:363
        out["in_" + name] = img
        out["out_" + name] = np.array([float(v) for v in feats.values()], dtype=np.float64)
        out["cols_" + name] = np.array(list(feats.keys()))
        print(name, img.shape, len(feats))
    path = os.path.join(HERE, "notebook_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
