"""Small run of every kernel path for compute-sanitizer (racecheck / memcheck), not a benchmark.

    compute-sanitizer --tool racecheck python profiles/sanitize_run.py
Covers: K12 (window hit and miss), K2 full-range ring, K1 FP64 fallback, K4w, K3 front + bins kernels with
4-bit and 8-bit tables and their wrapped-counter fallbacks, the K3 ring kernel (unmasked), finalize, pack.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import imfeat_b200 as imf

rng = np.random.default_rng(0)
n, h, w, c = 40, 64, 64, 4
img = rng.integers(100, 3000, (n, h, w, c)).astype(np.uint16)
img[:, :, :, 1] = rng.integers(0, 65536, (n, h, w))                       # wide range: K2 ring, FP64 moments
img[:, :, :, 2] = rng.choice([7, 7, 7, 900], (n, h, w))                     # few levels: counters wrap -> fallbacks
mask = (rng.random((n, h, w, c)) < 0.6).astype(np.uint8)
mask[:, :, :, 2] = 1
t1 = imf.extract_features(img, mask, four_directions=True, shape=True, moments=True)
t2 = imf.extract_features(img)                                              # notebook mode: ring K3
t3 = imf.extract_features([img[0, :37, :53], img[1, :20, :9]], [mask[0, :37, :53], mask[1, :20, :9]],
                          four_directions=True, shape=True, moments=True)  # odd sizes, size table
print("ok", t1.shape, t2.shape, t3.shape, float(np.nansum(t1)) > 0)
