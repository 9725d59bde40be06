"""Small device-resident run of the hot path for ncu captures (not a benchmark).

    python profiles/prof_run.py [--objects 2000] [--mode full|notebook] [--reps 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imfeat_b200 as imf

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=2000)
ap.add_argument("--mode", default="full")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
full = a.mode == "full"
ex = imf.FeatureExtractor(glcm=True, four_directions=full, shape=full, moments=full)
planes, masks, _ = ex.synth(0, 0, a.objects, 12, 64, 64, with_masks=True)
for _ in range(a.reps):
    out = ex.extract_planar(planes, masks if full else None, hs=64, ws=64)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out[0, 0]))
