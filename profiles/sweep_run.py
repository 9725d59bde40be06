"""Device-resident timing of the headline step under different IMFEAT_* switches (development sweeps, not a benchmark).

    python profiles/sweep_run.py --env IMFEAT_OVERLAP=0 --env "IMFEAT_OVERLAP=1 IMFEAT_K4_FILL=4" ...

Every --env runs in a fresh process (the switches are read once at context creation) and prints one line:
ms per step (CUDA events, 10,000 objects 64x64x12 + masks, all blocks), the kernel-group times, and whether the
table is bit-identical to the one of the first configuration.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(a):
    sys.path.insert(0, ROOT)
    import torch
    import imfeat_b200 as imf
    full = a.mode == "full"
    ex = imf.FeatureExtractor(glcm=True, four_directions=full, shape=full, moments=full)
    if a.mode == "cfg5":
        ex = imf.FeatureExtractor(glcm=True, four_directions=True, shape=True, moments=True)
        planes, masks, sizes = ex.synth(5, 0, 8192, 18, 128, 128, with_masks=True, variable=True, hmin=16, wmin=16, mask_shrink=32)
        hs = ws = 128
    else:
        planes, masks, sizes = ex.synth(0, 0, a.objects, 12, 64, 64, with_masks=True)
        hs = ws = 64
        if a.mode == "wide":
            ex = imf.FeatureExtractor(glcm=True, four_directions=True, shape=True, moments=True)
            g = torch.Generator(device="cuda")
            g.manual_seed(16)
            planes = torch.randint(0, 65536, tuple(planes.shape), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
        if a.mode == "notebook":
            masks = None
    out = ex.extract_planar(planes, masks, sizes, hs=hs, ws=ws)
    for _ in range(3):
        ex.extract_planar(planes, masks, sizes, hs=hs, ws=ws, out=out)
    torch.cuda.synchronize()
    ex.enable_timing(True)
    ex.kernel_times(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ex.extract_planar(planes, masks, sizes, hs=hs, ws=ws, out=out)
    e1.record()
    torch.cuda.synchronize()
    kms, kc = ex.kernel_times(reset=True)
    digest = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"ms": e0.elapsed_time(e1) / a.steps, "k12": kms[0] / max(kc[0], 1), "k2": kms[1] / max(kc[1], 1),
                      "k3": kms[2] / max(kc[2], 1), "k4": kms[3] / max(kc[3], 1), "sha": digest}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--objects", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mode", default="full", choices=["full", "notebook", "cfg5", "wide"])
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        return child(a)
    first = None
    for spec in a.env or [""]:
        env = dict(os.environ)
        lib = os.path.join(ROOT, "profiles", "_variants", "default.so")
        for kv in spec.split():
            k, v = kv.split("=", 1)
            if k == "LIB":                                 # a build variant (profiles/build_variant.sh) takes the library's place
                lib = os.path.join(ROOT, v)
            else:
                env[k] = v
        if os.path.exists(lib):
            shutil.copyfile(lib, os.path.join(ROOT, "interpretable-multichannel-image-analysis_b200", "libimfeat.so"))
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--objects", str(a.objects), "--steps", str(a.steps),
                              "--mode", a.mode], env=env, capture_output=True, text=True)
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
        try:
            d = json.loads(line)
        except Exception:
            print("%-50s FAILED rc=%d %s" % (spec, res.returncode, (res.stderr or "")[-400:]))
            continue
        first = first or d["sha"]
        print("%-50s %7.3f ms  k12 %.3f k2 %.3f k3 %.3f k4 %.3f  %s" % (spec or "(default)", d["ms"], d["k12"], d["k2"], d["k3"], d["k4"],
                                                                      "same" if d["sha"] == first else "DIFFERENT " + d["sha"]), flush=True)


if __name__ == "__main__":
    main()
