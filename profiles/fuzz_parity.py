"""Randomised differential run of the CUDA path against the C oracle (development tool, outside the test-suite).

    python profiles/fuzz_parity.py [--rounds 40] [--seed 1]

Every round draws a batch of variable-size objects (sizes, channel counts, value ranges and mask styles at random:
none, sparse boxes, blobs, stripes, single rows / columns, scattered pixels, full, empty), extracts it with all
feature blocks through the host entry point and compares every row with oracle/imfeat_ref.c the way the tests do
(integers, percentiles bit for bit, the rest to 1e-9).  Prints one line per round and a total.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import imfeat_b200 as imf  # noqa: E402
from conftest import compare_tables  # noqa: E402
from oracle import c_oracle  # noqa: E402


def planar(a):
    return np.ascontiguousarray(np.asarray(a).transpose(0, 3, 1, 2))


def draw_plane(rng, h, w):
    kind = int(rng.integers(0, 7))
    if kind == 0:
        return rng.integers(0, 65536, (h, w))
    if kind == 1:
        return rng.integers(100, 3000, (h, w))
    if kind == 2:                                           # smooth blob + noise
        r, c = np.mgrid[0:h, 0:w]
        z = 2000.0 * np.exp(-(((r - h / 2) / (h / 3 + 1)) ** 2 + ((c - w / 2) / (w / 3 + 1)) ** 2))
        return np.clip(300 + z + rng.normal(0, 20, (h, w)), 0, 65535).astype(np.int64)
    if kind == 3:
        return np.full((h, w), int(rng.integers(0, 65536)))  # constant
    if kind == 4:
        return rng.integers(0, 6, (h, w)) * int(rng.integers(1, 13000))   # few levels, wide
    if kind == 5:
        return np.clip(rng.normal(30000, 9000, (h, w)), 0, 65535).astype(np.int64)
    base = rng.integers(0, 256, (h, w))
    base[rng.random((h, w)) < 0.01] = 65535                 # outliers
    return base


def draw_mask(rng, h, w):
    kind = int(rng.integers(0, 9))
    m = np.zeros((h, w), np.uint8)
    if kind == 0:
        m[:] = 1
    elif kind == 1:
        pass                                                # empty
    elif kind == 2:
        r0, c0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        m[r0:r0 + int(rng.integers(1, 16)), c0:c0 + int(rng.integers(1, 16))] = 1
    elif kind == 3:
        r, c = np.mgrid[0:h, 0:w]
        m[((r - h * rng.random()) / (h * 0.4 + 1)) ** 2 + ((c - w * rng.random()) / (w * 0.4 + 1)) ** 2 < 1] = 1
    elif kind == 4:
        m[::int(rng.integers(2, 7)), :] = 1                 # stripes
    elif kind == 5:
        m[int(rng.integers(0, h)), :] = 1                   # one row
    elif kind == 6:
        m[:, int(rng.integers(0, w))] = 1                   # one column
    elif kind == 7:
        m[rng.random((h, w)) < rng.random() * 0.2] = 1      # scattered
    else:
        m[h // 2:, :] = 1
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    ex = imf.FeatureExtractor(glcm=True, four_directions=True, shape=True, moments=True)
    total, t0 = 0, time.time()
    for rd in range(a.rounds):
        C = int(rng.integers(1, 5))
        n = int(rng.integers(4, 40))
        big = rng.random() < 0.5
        hs, ws = (int(rng.integers(60, 129)), int(rng.integers(60, 129))) if big else (int(rng.integers(2, 65)), int(rng.integers(2, 65)))
        use_mask = rng.random() < 0.8
        img, msk = np.zeros((n, hs, ws, C), np.uint16), np.zeros((n, hs, ws, C), np.uint8)
        sizes = np.zeros((n, 2), np.int32)
        objs, masks = [], []
        for i in range(n):
            h, w = int(rng.integers(1, hs + 1)), int(rng.integers(1, ws + 1))
            if i == 0:
                h, w = hs, ws
            o = np.stack([draw_plane(rng, h, w) for _ in range(C)], axis=2).astype(np.uint16)
            m = np.stack([draw_mask(rng, h, w) for _ in range(C)], axis=2)
            img[i, :h, :w], msk[i, :h, :w], sizes[i] = o, m, (h, w)
            objs.append(o)
            masks.append(m)
        got = ex.extract_host_hwc(img, msk if use_mask else None, sizes=sizes)
        cols = imf.feature_columns(C, n_angles=4, shape=True, moments=True)
        for i in range(n):
            want = c_oracle.table(planar(objs[i][None]), planar(masks[i][None]) if use_mask else None, glcm=True, n_angles=4,
                                  shape=True, moments=True)
            compare_tables(got[i:i + 1], want, cols, label="round %d object %d %s" % (rd, i, objs[i].shape),
                           images=[objs[i]], masks=[masks[i]] if use_mask else None)
        total += n
        print("round %3d: %2d objects, C=%d, stride %dx%d, masks=%s  ok" % (rd, n, C, hs, ws, use_mask), flush=True)
    print("fuzz ok: %d objects in %d rounds, %.0f s" % (total, a.rounds, time.time() - t0))


if __name__ == "__main__":
    main()
