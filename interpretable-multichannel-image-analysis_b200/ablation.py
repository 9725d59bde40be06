"""Device-side channel-ablation loop (channel importance by re-extraction).

The reference derives channel importance downstream, from a classifier's importances grouped by
column name (NB:458-462) and from sklearn's permutation importance on feature columns
(NB:493-497); it never re-extracts.  north_star asks for the re-extraction variant: for each
channel, (a) leave it out, or (b) permute it across objects, and re-run the kernels without host
round trips.  All index tensors are built once on the device; the loop only enqueues kernels.

Because every feature is a function of one channel plane (NB:239, NB:291), both ablations are
exactly predictable from the base table -- tests/test_gpu_parity.py checks that equality.
"""
import numpy as np


def loco_channel_lists(n_channels):
    """[C][C-1] int32: row k is the channel list with channel k left out."""
    return np.array([[c for c in range(n_channels) if c != k] for k in range(n_channels)], dtype=np.int32)


def permutation_sources(n_objects, n_channels, seed=42):
    """[C][N][C] int32 source-object tables: table k permutes channel k across objects with
    ``np.random.default_rng(seed).permutation(N)`` (echoing random_state=42 at NB:496); one
    independent permutation per channel, drawn in channel order."""
    rng = np.random.default_rng(seed)
    ident = np.tile(np.arange(n_objects, dtype=np.int32)[:, None], (1, n_channels))
    out = np.empty((n_channels, n_objects, n_channels), dtype=np.int32)
    for k in range(n_channels):
        out[k] = ident
        out[k][:, k] = rng.permutation(n_objects).astype(np.int32)
    return out


def sweep_index(planes, mode="loco", seed=42):
    """The device index tensor of a sweep: channel lists [C, C-1] for "loco", source-object tables
    [C, N, C] for "permute" (built on the host once, copied once)."""
    import torch
    N, C = int(planes.shape[0]), int(planes.shape[1])
    if mode == "loco":
        return torch.from_numpy(loco_channel_lists(C)).to(planes.device)
    if mode == "permute":
        return torch.from_numpy(permutation_sources(N, C, seed)).to(planes.device)
    raise ValueError("mode must be 'loco' or 'permute'")


def channel_ablation_sweep(extractor, planes, masks=None, sizes=None, hs=None, ws=None,
                           mode="loco", seed=42, out=None, index=None):
    """Re-extract once per ablated channel, entirely on the device.

    mode "loco"    -> tensor [C, N, row_width(C-1)]  (column suffixes are positional, NB:241;
                      ``FeatureExtractor.columns(C-1, channel_ids=...)`` restores original ids)
    mode "permute" -> tensor [C, N, row_width(C)]; permuting objects is only meaningful for
                      equal-size objects, so ``sizes`` must be None.
    ``index`` is the tensor ``sweep_index`` returns (built here when not given).
    """
    import torch
    N, C = int(planes.shape[0]), int(planes.shape[1])
    dev = planes.device
    if mode == "permute" and sizes is not None:
        raise ValueError("channel permutation across objects needs equal-size objects")
    if index is None:
        index = sweep_index(planes, mode, seed)
    if mode == "loco":
        width = extractor.row_width(C - 1)
        res = out if out is not None else torch.empty((C, N, width), dtype=torch.float64, device=dev)
        for k in range(C):
            extractor.extract_planar(planes, masks, sizes, hs=hs, ws=ws, chan=index[k], out=res[k])
        return res
    if mode == "permute":
        width = extractor.row_width(C)
        res = out if out is not None else torch.empty((C, N, width), dtype=torch.float64, device=dev)
        for k in range(C):
            extractor.extract_planar(planes, masks, None, hs=hs, ws=ws, src_obj=index[k], out=res[k])
        return res
    raise ValueError("mode must be 'loco' or 'permute'")


class CapturedSweep:
    """A whole channel-ablation sweep captured once into a CUDA graph: ``replay()`` re-runs the C
    re-extractions (5 kernels each) on the current contents of ``planes`` / ``masks`` with a single
    launch and no host work in between -- e.g. once per bootstrap replicate or per augmented copy
    written into the same buffers.  The index tensors, the output block and the extractor's internal
    buffers are fixed at capture time, so the sweep works on a private clone of the extractor (its own
    context: later, larger batches through the original extractor can neither move nor share the work
    buffers and scheduler counters whose addresses the graph holds)."""

    def __init__(self, extractor, planes, masks=None, sizes=None, hs=None, ws=None, mode="loco", seed=42):
        import torch
        extractor = extractor.clone()
        self.extractor, self.planes, self.masks = extractor, planes, masks
        self.index = sweep_index(planes, mode, seed)
        kw = dict(sizes=sizes, hs=hs, ws=ws, mode=mode, seed=seed, index=self.index)
        # one eager run sizes the extractor's internal buffers (nothing may be allocated while capturing)
        self.out = channel_ablation_sweep(extractor, planes, masks, **kw)
        torch.cuda.synchronize(planes.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            channel_ablation_sweep(extractor, planes, masks, out=self.out, **kw)

    def replay(self):
        self.graph.replay()
        return self.out


def channel_importance_from_sweep(base_score, ablated_scores):
    """Importance of channel k = score drop when it is ablated (host-side helper; the classifier
    that produces the scores stays on the host, as in NB:404-417)."""
    return [float(base_score - s) for s in ablated_scores]


def group_importances_by_channel(columns, importances, threshold=0.0):
    """Channel grouping of per-feature importances (NB:458-462) with an exact channel-id match
    (the notebook's substring test ``ch in x`` lets 'Ch1' also match Ch10..Ch12)."""
    groups = {}
    for name, imp in zip(columns, importances):
        ch = int(name.rsplit("_Ch", 1)[1])
        if imp > threshold:
            groups.setdefault(ch, []).append(float(imp))
    return groups
