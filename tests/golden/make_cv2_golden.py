"""Generate tests/golden/cv2_moments_golden.npz with OpenCV -- an independent, third-party implementation
of the image moments (cv2.moments: raw, central and normalised central moments up to order 3) and of the
bounding rectangle -- on seeded synthetic planes.  The fixture pins the SPECIFICATION of the extension blocks
x2 / x3 (oracle/notebook_oracle.py: shape_values, moment_values), which the reference notebook does not have:

    python tests/golden/make_cv2_golden.py        (needs cv2; the committed .npz is what the tests read)

OpenCV's x is the column and y the row: cv2 nu_pq = sum (c - cc)^p (r - cr)^q I / m00^((p+q)/2 + 1), the same
normalisation scikit-image's moments_normalized uses.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from imfeat_b200 import synth  # noqa: E402

KEYS = ["m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03", "mu20", "mu11", "mu02", "mu30", "mu21",
        "mu12", "mu03", "nu20", "nu11", "nu02", "nu30", "nu21", "nu12", "nu03"]
# (seed, object, channel, h, w, mask_shrink): the planes are regenerated from these by the tests
CASES = [(7, 0, 0, 64, 64, 256), (7, 1, 3, 64, 64, 256), (7, 2, 5, 64, 64, 128), (7, 3, 1, 40, 56, 256),
         (7, 4, 2, 33, 21, 200), (7, 5, 0, 128, 96, 32), (7, 6, 7, 17, 128, 256), (7, 7, 4, 5, 9, 256)]


def main():
    out = {"cases": np.array(CASES, dtype=np.int64), "keys": np.array(KEYS)}
    weighted, binary, rects = [], [], []
    for seed, obj, ch, h, w, shrink in CASES:
        px, mk = synth.synth_plane(seed, obj, ch, h, w, shrink)
        img = np.where(mk > 0, px, 0).astype(np.float64)
        m = cv2.moments(img, binaryImage=False)                 # intensity-weighted, masked
        b = cv2.moments(mk.astype(np.uint8), binaryImage=True)  # the mask as a region
        weighted.append([m[k] for k in KEYS])
        binary.append([b[k] for k in KEYS])
        rects.append(cv2.boundingRect(mk.astype(np.uint8)))     # x, y, w, h
    out["weighted"] = np.array(weighted, dtype=np.float64)
    out["binary"] = np.array(binary, dtype=np.float64)
    out["rect"] = np.array(rects, dtype=np.int64)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "cv2_moments_golden.npz"), **out)
    print("wrote", len(CASES), "cases, OpenCV", cv2.__version__)


if __name__ == "__main__":
    main()
