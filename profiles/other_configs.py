"""Device-resident timings of BASELINE.json configs[3] (channel-ablation sweep) and configs[4]
(variable-size objects up to 128x128x18 with sparse masks).  Not bench lines: evidence that the other
configurations run on the same kernels at comparable rates.

    python profiles/other_configs.py [--abl-objects 20000] [--var-objects 4000]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import imfeat_b200 as imf
from imfeat_b200 import ablation


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_split(ex, fn):
    ex.enable_timing(True)
    fn()
    torch.cuda.synchronize()
    t = ex.kernel_times()
    ex.enable_timing(False)
    return t


ap = argparse.ArgumentParser()
ap.add_argument("--abl-objects", type=int, default=20000)
ap.add_argument("--var-objects", type=int, default=4000)
a = ap.parse_args()
res = {}

# cfg4: leave-one-channel-out and per-channel permutation re-extraction, 12 ablations each
ex = imf.FeatureExtractor(glcm=True, four_directions=True, shape=True, moments=True)
N = a.abl_objects
planes, masks, _ = ex.synth(0, 0, N, 12, 64, 64, with_masks=True)
base = timed(lambda: ex.extract_planar(planes, masks, hs=64, ws=64))
out_l = torch.empty((12, N, ex.row_width(11)), dtype=torch.float64, device="cuda")
out_p = torch.empty((12, N, ex.row_width(12)), dtype=torch.float64, device="cuda")
t_l = timed(lambda: ablation.channel_ablation_sweep(ex, planes, masks, hs=64, ws=64, mode="loco", out=out_l), 1)
t_p = timed(lambda: ablation.channel_ablation_sweep(ex, planes, masks, hs=64, ws=64, mode="permute", out=out_p), 1)
res["cfg4_ablation"] = {"objects": N, "base_ms": base, "base_objects_per_s": N / base * 1e3,
                        "loco_12_ms": t_l, "loco_objects_per_s": 12 * N / t_l * 1e3,
                        "permute_12_ms": t_p, "permute_objects_per_s": 12 * N / t_p * 1e3}
del planes, masks, out_l, out_p
torch.cuda.empty_cache()

# cfg5: h, w ~ U{16..128}, C = 18, sparse masks (mask_shrink 25/256 of the blob ellipse)
N = a.var_objects
planes, masks, sizes = ex.synth(1, 0, N, 18, 128, 128, with_masks=True, variable=True, hmin=16, wmin=16,
                                mask_shrink=25)
t_v = timed(lambda: ex.extract_planar(planes, masks, sizes, hs=128, ws=128))
px = int((sizes[:, 0].long() * sizes[:, 1].long()).sum().item()) * 18
split = kernel_split(ex, lambda: ex.extract_planar(planes, masks, sizes, hs=128, ws=128))
res["cfg5_variable_sparse"] = {"objects": N, "ms": t_v, "objects_per_s": N / t_v * 1e3, "valid_pixels": px,
                               "gpixels_per_s": px / t_v / 1e6, "kernel_ms": split}
# same objects without masks, notebook defaults
exn = imf.FeatureExtractor(glcm=True)
t_n = timed(lambda: exn.extract_planar(planes, None, sizes, hs=128, ws=128))
res["cfg5_variable_notebook_mode"] = {"objects": N, "ms": t_n, "objects_per_s": N / t_n * 1e3,
                                      "gpixels_per_s": px / t_n / 1e6}
print(json.dumps(res, indent=1, default=str))
