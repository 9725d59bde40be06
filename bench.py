#!/usr/bin/env python
"""Benchmark of the hot path: per-object, per-channel feature extraction (objects/second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--objects M]

One "step" = one pass of the hot path over one batch of synthetic objects.  Workload at N=1 is
BASELINE.json configs[1]: 10,000 synthetic 64x64x12 uint16 objects + uint8 masks, every feature
block (17 masked intensity/percentile/entropy statistics, GLCM x 4 directions, 10 shape, 9
spatial-moment features per channel).  For N>1 the workload is configs[2]: 1,000,000 such objects in
total, sharded over the N ranks (strong scaling), extracted with the product call
``imfeat_b200.distributed.extract_sharded`` -- slab pipeline, every rank ends up with the full
float64[1e6, 720] table -- all inside the timed region.  Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

C, HS, WS = 12, 64, 64
SEED = 0
FULL = dict(glcm=True, four_directions=True, shape=True, moments=True)
F_FULL = 17 + 24 + 10 + 9
UNIT = "objects/s"
METRIC = "objects/sec (12-ch 64x64 uint16+mask), all feature blocks"


def peak_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def algorithmic_bytes_per_object(h, w, c, mask_bytes, feats_per_channel):
    """B = 2*h*w*C + m*h*w*C + 8*F*C (valid pixels only, one pass)."""
    return 2 * h * w * c + mask_bytes * h * w * c + 8 * feats_per_channel * c


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the notebook's CPU code path (numpy/scipy restatement, oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    first, count, notebook = args
    from imfeat_b200 import synth
    from oracle import notebook_oracle as orc
    objs = synth.synth_objects(SEED, first, count, C, HS, WS)
    t0 = time.perf_counter()
    if notebook:
        orc.oracle_extract([o[0] for o in objs])                              # NB:358-364 as it stands: no masks, 23 features
    else:
        orc.oracle_extract([o[0] for o in objs], [o[1] for o in objs], **FULL)
    return time.perf_counter() - t0


def cpu_reference_rate(n_objects, procs, notebook=False, start="fork"):
    """objects/s of the notebook-style CPU path on `procs` host processes (generation untimed)."""
    if procs <= 1:
        dt = _cpu_worker((0, n_objects, notebook))
        return n_objects / dt
    import multiprocessing as mp
    per = max(1, n_objects // procs)
    jobs = [(k * per, per, notebook) for k in range(procs)]
    with mp.get_context(start).Pool(procs) as pool:             # "spawn" when this process holds a CUDA context
        pool.map(_cpu_worker, [(0, 1, notebook)] * procs)     # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    return per * procs / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 2 * cores                                      # objects per step (bounded sample: ~0.2 s per core)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_rate(cores, cores)
    rates = [cpu_reference_rate(per_step, cores) for _ in range(args.steps)]
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "sample of the same synthetic 64x64x12 uint16+mask objects: %d per step, all feature blocks" % per_step,
                   "code": "oracle/notebook_oracle.py (numpy/scipy restatement of notebook cell 13; the reference is a Jupyter notebook, nothing to compile)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d objects per step x %d steps on %d processes" % (per_step, args.steps, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
KERNEL_GROUPS = ["k12_basic (min..entropy, 17 columns)", "k2_full_range (worklist)", "k3_glcm (front + bins + finalize)",
                 "k4_shape_moments"]


def _timed(torch, fn, steps, warmup=2):
    """Device time per call of fn() (CUDA events on the current stream)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def side_records(torch, imf, dev, local, planes, masks, n_obj, steps, peak):
    """Measurements outside the main timed region: the notebook's own configuration, BASELINE.json configs[3]
    and configs[4], and a full-16-bit batch that takes the wide-range paths."""
    rec = {}
    # ---- notebook-parity mode (no masks, 1 direction, 23 features/channel), device resident ----
    exn = imf.FeatureExtractor(device=local)
    outn = torch.empty((n_obj, exn.row_width(C)), dtype=torch.float64, device=dev)
    exn.extract_planar(planes, None, hs=HS, ws=WS, out=outn)
    exn.enable_timing(True)
    exn.kernel_times(reset=True)
    nb_ms = _timed(torch, lambda: exn.extract_planar(planes, None, hs=HS, ws=WS, out=outn), steps, warmup=2)
    nkms, nkcalls = exn.kernel_times(reset=True)
    exn.enable_timing(False)
    b_nb = n_obj * algorithmic_bytes_per_object(HS, WS, C, 0, 23)
    rec["notebook_mode"] = {
        "objects_per_s": n_obj / (nb_ms * 1e-3), "ms_per_step": nb_ms,
        "workload": "same %d objects, masks=None, 23 features/channel (the notebook's literals, NB:242-250, NB:298)" % n_obj,
        "roofline_path_frac": b_nb / (nb_ms * 1e-3) / 1e9 / peak,
        "kernels": [{"kernel": KERNEL_GROUPS[k], "ms_per_launch": nkms[k] / nkcalls[k]} for k in range(4) if nkcalls[k]]}
    del outn
    # ---- configs[3]: channel-importance sweep, 12 leave-one-channel-out + 12 permutation re-extractions ----
    n4 = 100000
    ex4 = imf.FeatureExtractor(device=local, **FULL)
    p4, m4, _ = ex4.synth(SEED, 0, n4, C, HS, WS, with_masks=True)
    out_l = torch.empty((C, n4, ex4.row_width(C - 1)), dtype=torch.float64, device=dev)
    idx_l = imf.ablation.sweep_index(p4, "loco")
    ms_l = _timed(torch, lambda: imf.ablation.channel_ablation_sweep(ex4, p4, m4, hs=HS, ws=WS, mode="loco", out=out_l, index=idx_l), 2, warmup=1)
    del out_l
    out_p = torch.empty((C, n4, ex4.row_width(C)), dtype=torch.float64, device=dev)
    idx_p = imf.ablation.sweep_index(p4, "permute", seed=42)
    ms_p = _timed(torch, lambda: imf.ablation.channel_ablation_sweep(ex4, p4, m4, hs=HS, ws=WS, mode="permute", out=out_p, index=idx_p), 2, warmup=1)
    del out_p, p4, m4
    rec["cfg4_ablation_sweep"] = {
        "workload": "configs[3]: %d objects, 12 leave-one-channel-out + 12 per-channel permutation re-extractions on the device (index tensors only, no pixel moves)" % n4,
        "loco_ms": ms_l, "loco_object_extractions_per_s": C * n4 / (ms_l * 1e-3),
        "permute_ms": ms_p, "permute_object_extractions_per_s": C * n4 / (ms_p * 1e-3)}
    torch.cuda.empty_cache()
    # ---- configs[4]: variable-size objects up to 128x128x18, sparse masks ----
    n5, c5, s5 = 8192, 18, 128
    p5, m5, z5 = ex4.synth(SEED + 5, 0, n5, c5, s5, s5, with_masks=True, variable=True, hmin=16, wmin=16, mask_shrink=32)
    out5 = torch.empty((n5, ex4.row_width(c5)), dtype=torch.float64, device=dev)
    ex4.enable_timing(True)
    ex4.kernel_times(reset=True)
    # (two warm-up calls: the context hands out two K3 scratch slots in turn, and each grows on its first use here)
    ms5 = _timed(torch, lambda: ex4.extract_planar(p5, m5, z5, hs=s5, ws=s5, out=out5), 4, warmup=2)
    k5, c5n = ex4.kernel_times(reset=True)
    px5 = float((z5[:, 0].double() * z5[:, 1].double()).sum().item()) * c5
    b5 = px5 * 3 + n5 * 8 * F_FULL * c5
    rec["cfg5_variable_sparse"] = {
        "workload": "configs[4]: %d objects, h,w ~ U{16..128}, 18 channels, masks 1-10%% of the tile, fixed stride 128x128 + size table, all feature blocks" % n5,
        "ms": ms5, "objects_per_s": n5 / (ms5 * 1e-3), "gpixel_per_s": px5 / (ms5 * 1e-3) / 1e9,
        "roofline_path_frac": b5 / (ms5 * 1e-3) / 1e9 / peak,
        "kernels": [{"kernel": KERNEL_GROUPS[k], "ms_per_launch": k5[k] / c5n[k]} for k in range(4) if c5n[k]]}
    del p5, m5, z5, out5
    # ---- full 16-bit range: K12's window fails for every tile, K2 full-range + K1 FP64 + K3 general quantiser ----
    g = torch.Generator(device=dev)
    g.manual_seed(16)
    p16 = torch.randint(0, 65536, tuple(planes.shape), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
    out16 = torch.empty((n_obj, ex4.row_width(C)), dtype=torch.float64, device=dev)
    ms16 = _timed(torch, lambda: ex4.extract_planar(p16, masks, hs=HS, ws=WS, out=out16), 4, warmup=2)
    k16, c16 = ex4.kernel_times(reset=True)
    ex4.enable_timing(False)
    rec["full_16bit_range"] = {
        "workload": "same masks, pixels uniform over 0..65535: every tile leaves the 4,096-value histogram window and the integer moment range",
        "ms": ms16, "objects_per_s": n_obj / (ms16 * 1e-3),
        "kernels": [{"kernel": KERNEL_GROUPS[k], "ms_per_launch": k16[k] / c16[k]} for k in range(4) if c16[k]]}
    return rec


def run_b200(args):
    import torch
    import imfeat_b200 as imf
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ex = imf.FeatureExtractor(device=local, **FULL)
    width = ex.row_width(C)
    if world == 1:
        n_total = n_obj = args.objects
        first = 0
    else:
        n_total = args.total_objects
        first, stop, _ = imf.distributed.shard_range(n_total, world, rank)
        n_obj = stop - first
    planes, masks, _ = ex.synth(SEED, first, n_obj, C, HS, WS, with_masks=True)
    table = None
    if world > 1:
        table = imf.distributed.ShardedTable(n_total, width, device=dev, transport=args.transport)
        out = None
    else:
        out = torch.empty((n_obj, width), dtype=torch.float64, device=dev)

    def step():
        if world == 1:
            ex.extract_planar(planes, masks, hs=HS, ws=WS, out=out)
        else:
            # the product call: slab pipeline, rows delivered to every rank while the next slab is computed;
            # it returns when the full table is complete on every rank
            imf.distributed.extract_sharded(ex, planes, masks, hs=HS, ws=WS, n_objects=n_total, table=table,
                                            slab_objects=args.slab)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ex.enable_timing(True)
    ex.kernel_times(reset=True)
    launches0 = ex.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.result()
    ms_total = e0.elapsed_time(e1)
    kms, kcalls = ex.kernel_times(reset=True)
    ex.enable_timing(False)
    launches = ex.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_total / (ms_step * 1e-3)

    # ---- N > 1: a foreign shard's rows in MY table against a local recomputation (outside the timed region) ----
    identity = None
    if world > 1:
        other = (rank + 1) % world
        o_first, o_stop, _ = imf.distributed.shard_range(n_total, world, other)
        n_chk = min(512, o_stop - o_first)
        cp, cm, _ = ex.synth(SEED, o_first, n_chk, C, HS, WS, with_masks=True)
        mine = ex.extract_planar(cp, cm, hs=HS, ws=WS)
        theirs = table.table[o_first:o_first + n_chk]
        same = bool(((mine == theirs) | (torch.isnan(mine) & torch.isnan(theirs))).all().item())
        flag = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        identity = {"bit_identical": bool(flag.item() == 1.0),
                    "check": "every rank recomputed %d objects of its neighbour's shard locally and compared them with the rows delivered into its own table" % n_chk}
        del cp, cm, mine

    # ---- end to end through the host-buffer entry point (pinned inputs, H2D + D2H inside) ----
    n_e2e = min(n_obj, args.objects)
    hwc = torch.empty((n_e2e, HS, WS, C), dtype=torch.uint16).pin_memory()
    mhwc = torch.empty((n_e2e, HS, WS, C), dtype=torch.uint8).pin_memory()
    hwc.copy_(planes[:n_e2e, :, :HS * WS].reshape(n_e2e, C, HS, WS).permute(0, 2, 3, 1))
    mhwc.copy_(masks[:n_e2e, :, :HS * WS].reshape(n_e2e, C, HS, WS).permute(0, 2, 3, 1))
    h_img, h_mask = hwc.numpy(), mhwc.numpy()
    h_out = torch.empty((n_e2e, width), dtype=torch.float64).pin_memory().numpy()
    e2e_steps = max(2, min(args.steps, 5))
    ex.extract_host_hwc(h_img, h_mask, out=h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ex.extract_host_hwc(h_img, h_mask, out=h_out)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n_e2e / e2e_s
    # the same call with the masks in the bit-packed host format (packed once, outside the timed region)
    h_bits = torch.from_numpy(imf.pack_mask_bits(h_mask)).pin_memory().numpy()
    ex.extract_host_hwc(h_img, h_bits, out=h_out, masks_packed=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ex.extract_host_hwc(h_img, h_bits, out=h_out, masks_packed=True)
    torch.cuda.synchronize()
    e2e_bits_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_bits_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_bits_s = float(t.item())
    dev_rows = out[:n_e2e] if world == 1 else table.table[first:first + n_e2e]
    assert np.array_equal(h_out, dev_rows.cpu().numpy(), equal_nan=True), "host and device paths disagree"

    peak, peak_src = peak_hbm_gbs()
    side = {}
    if world == 1 and not args.no_side:
        side = side_records(torch, imf, dev, local, planes, masks, n_obj, args.steps, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    per_kernel = []
    tot_k = sum(kms) or 1.0
    for k in range(4):
        if kcalls[k]:
            per_kernel.append({"kernel": KERNEL_GROUPS[k], "ms_per_step": kms[k] / args.steps, "launch_groups": int(kcalls[k]),
                               "share": kms[k] / tot_k})
    dom = max(per_kernel, key=lambda d: d["share"])
    b_path = n_total * algorithmic_bytes_per_object(HS, WS, C, 1, F_FULL)
    b_basic = n_obj * algorithmic_bytes_per_object(HS, WS, C, 1, 17)      # kernel times are per rank
    achieved = b_path / (ms_step * 1e-3) / 1e9
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if world == 1 and n_obj == 10000 and os.path.exists(tpath):          # measured for exactly this workload
        try:
            tj = json.load(open(tpath))
            traffic, tsrc = tj["step_total_bytes"], "profiles/r2_traffic.json (ncu dram__bytes_read+write, summed over the kernels of one step)"
        except Exception:
            traffic = None
    # what bounds each kernel group, from the committed ncu capture of this workload (profiles/r2_traffic.json:
    # shared-memory wavefronts per SM clock; 1.0 is the pipe's peak) -- the HBM fraction alone would not say
    try:
        tk = json.load(open(tpath))["kernels"] if (world == 1 and n_obj == 10000) else {}
    except Exception:
        tk = {}
    if tk:
        smem = lambda *keys: [round(v["shared_wavefronts_per_sm_clock"], 3) for k, v in tk.items() if any(q in k for q in keys)]
        for d in per_kernel:
            if d["kernel"].startswith("k3_"):
                d["bound"] = "shared-memory pipe"
                d["smem_wavefronts_per_sm_clock"] = smem("k3a_front", "k3_glcm_kernel")
            elif d["kernel"].startswith("k12"):
                d["bound"] = "instruction issue (62 % of peak, ncu)"
                d["smem_wavefronts_per_sm_clock"] = smem("k12_basic")
            elif d["kernel"].startswith("k4_"):
                d["bound"] = "instruction issue (64 % of peak, ncu)"
                d["smem_wavefronts_per_sm_clock"] = smem("k4w_shape")
    basic_ms = sum(d["ms_per_step"] for d in per_kernel[:2] if d["kernel"].startswith(("k12", "k2_")))
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        rate = cpu_reference_rate(args.cpu_objects, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "first %d of the %d objects, one process (the notebook's mode), all feature blocks, oracle/notebook_oracle.py" % (args.cpu_objects, n_obj)}
        n_nb = max(512, 4 * cores)
        nb_rate = cpu_reference_rate(n_nb, cores, notebook=True, start="spawn")
        side.setdefault("notebook_mode", {})["cpu_baseline"] = {
            "value": nb_rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d objects on %d processes, the notebook's own path (NB:358-364: no masks, 23 features per channel)" % (n_nb, cores)}
    workload = ("cfg2: %d synthetic 64x64x12 uint16 objects + uint8 masks; per channel 17 masked intensity + 24 GLCM "
                "(4 directions, 256 levels) + 10 shape + 9 moment features" % n_obj) if world == 1 else (
        "cfg3: %d synthetic 64x64x12 uint16 objects + uint8 masks in total, %d per GPU resident in HBM (generated on the device), "
        "all feature blocks (720 columns); every rank ends with the full float64 table" % (n_total, n_obj))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong",
        "vs_baseline": None, "dtype": "u16 pixels, integer + f64 accumulation, f64 table",
        "data": "synthetic",
        "config": {"workload": workload,
                   "l2": "inputs (%.0f MB per GPU and step) exceed the 126 MB L2" % ((planes.numel() * 2 + masks.numel()) / 1e6),
                   "collective": "none" if world == 1 else
                   "imfeat_b200.distributed.extract_sharded: slabs of %d objects; transport %s (%s); all deliveries complete inside the timed region" % (
                       args.slab, table.transport,
                       "per-peer device-to-device copies through CUDA IPC mappings on a side stream: copy engines over NVLink, no SM-resident collective kernel"
                       if table.transport == "p2p" else "all_gather_into_tensor per slab on a side stream")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_img.nbytes + h_mask.nbytes),
                "d2h_bytes_per_step": int(h_out.nbytes), "objects_per_gpu": int(n_e2e),
                "call": "FeatureExtractor.extract_host_hwc -> imfeat_extract_host_hwc (pinned host buffers, README (h,w,c) layout)"},
        "e2e_bitmask": {"value": world * n_e2e / e2e_bits_s, "unit": UNIT, "h2d_bytes_per_step": int(h_img.nbytes + h_bits.nbytes),
                        "d2h_bytes_per_step": int(h_out.nbytes),
                        "call": "the same call with the masks in the bit-packed host format (opts.host_mask_bits; imfeat_b200.pack_mask_bits): "
                                "a byte mask is a third of the PCIe traffic of an object, a packed one 4%; expanded to bytes on the device"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "whole step (every kernel of one extraction); dominant: " + dom["kernel"],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "traffic_source": tsrc, "algorithmic_bytes": b_path,
                     "note": "SURVEY 8(d): algorithmic bytes = N*(2hwC + 1*hwC + 8*60*C), counted ONCE for the whole step however often a kernel re-reads the tile; "
                             "divided by the CUDA-event time of the step.  ncu: K3's two kernels run at 77-82% of the shared-memory pipe (their roofline), "
                             "K12 / K4 at 62-64% instruction issue; none is HBM-bound (profiles/)"},
        "roofline_kernels": per_kernel,
        "roofline_basic_block": {"ms_per_step": basic_ms, "frac": b_basic / (basic_ms * 1e-3) / 1e9 / peak if basic_ms else None,
                                 "note": "the 17 intensity/percentile/entropy columns: N*(2hwC + hwC + 8*17*C) bytes over the time of K12 + the full-range worklist kernel"},
    }
    if identity:
        line["gather_check"] = identity
    line.update(side)
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--objects", type=int, default=10000, help="objects per step at N=1 (and of the end-to-end leg per GPU)")
    ap.add_argument("--total-objects", type=int, default=1000000, help="objects in total at N>1 (configs[2])")
    ap.add_argument("--slab", type=int, default=16384, help="objects per slab of the sharded pipeline")
    ap.add_argument("--transport", default=None, choices=[None, "p2p", "collective"])
    ap.add_argument("--cpu-objects", type=int, default=96, help="CPU-baseline sample size (headline configuration, one process)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the notebook-mode / cfg4 / cfg5 / 16-bit side records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
