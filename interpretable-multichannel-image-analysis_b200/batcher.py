"""Pinned-host batcher: the step before the hot path.

The reference documents one ``.h5`` file per object with keys ``image`` (16-bit (h,w,c)), ``mask``
((h,w,c), optional), ``label``, ``donor``, ``experiment``, ``channels`` (README.md:5-14) but ships
no loader.  This batcher is the B200-side replacement for that loader: objects are packed, as
they arrive, into fixed-stride page-locked slabs ``uint16[B, Hs, Ws, C]`` (+ ``uint8`` or bit-packed masks)
with a padded-size table ``int32[B, 2]``, which ``imfeat_extract_host_hwc`` DMA-copies without an extra
staging copy and transposes on the device; two slabs alternate so that host packing overlaps the GPU.

h5py / libhdf5 are not part of this image, so ``read_h5_object`` is import-guarded and has never run
(UNTESTED); any source of (h,w,c) uint16 arrays works.
"""
import numpy as np


def _pinned_empty(shape, dtype):
    """Page-locked numpy array when CUDA is available, ordinary memory otherwise (CPU tests)."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
            return t.numpy(), t
    except Exception:
        pass
    return np.empty(shape, dtype=dtype), None


class PinnedBatcher:
    """Collects (h_i, w_i, C) objects into fixed-stride page-locked slabs and extracts full slabs.

    Two slabs are used in turn: ``flush`` hands a full slab to a worker thread (the C call releases the GIL) and
    returns at once, so packing slab k+1 on the host overlaps the copies and kernels of slab k.  With
    ``packed_masks=True`` the masks are stored bit-packed (the C ABI's host_mask_bits format): a third less PCIe
    traffic per object.

    >>> b = PinnedBatcher(extractor, capacity=4096, hs=128, ws=128, channels=12, with_masks=True)
    >>> for img, msk, label in objects: b.add(img, msk, label)
    >>> table, labels = b.finish()
    """

    def __init__(self, extractor, capacity, hs, ws, channels, with_masks=False, packed_masks=False, asynchronous=True):
        self.ex = extractor
        self.capacity, self.hs, self.ws, self.c = int(capacity), int(hs), int(ws), int(channels)
        self.with_masks, self.packed = bool(with_masks), bool(with_masks and packed_masks)
        self.mask_bytes = ((hs * ws * channels + 63) // 64) * 8
        self._slabs = []
        for _ in range(2 if asynchronous else 1):
            images, img_t = _pinned_empty((self.capacity, hs, ws, channels), np.uint16)
            if not with_masks:
                masks, msk_t = None, None
            elif self.packed:
                masks, msk_t = _pinned_empty((self.capacity, self.mask_bytes), np.uint8)
            else:
                masks, msk_t = _pinned_empty((self.capacity, hs, ws, channels), np.uint8)
            self._slabs.append(dict(images=images, masks=masks, sizes=np.zeros((self.capacity, 2), dtype=np.int32),
                                    keep=(img_t, msk_t), pending=None))
        self._cur = 0
        self._pad = np.zeros((hs, ws, channels), dtype=np.uint8) if self.packed else None
        self.count = 0
        self.labels = []
        self._tables = []
        self._all_labels = []
        self.variable = False
        self._pool = None
        if asynchronous:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=1)       # one extraction in flight: the context is not re-entrant

    # the slab being filled (kept as attributes for callers that look at them)
    @property
    def images(self):
        return self._slabs[self._cur]["images"]

    @property
    def masks(self):
        return self._slabs[self._cur]["masks"]

    @property
    def sizes(self):
        return self._slabs[self._cur]["sizes"]

    def add(self, image, mask=None, label=None):
        image = np.asarray(image)
        if image.dtype != np.uint16 or image.ndim != 3 or image.shape[2] != self.c:
            raise ValueError("objects must be uint16 (h, w, %d) arrays (README.md:8)" % self.c)
        h, w = image.shape[:2]
        if h > self.hs or w > self.ws or h < 1 or w < 1:
            raise ValueError("object %dx%d does not fit the %dx%d stride" % (h, w, self.hs, self.ws))
        if self.with_masks != (mask is not None):
            raise ValueError("masks must be given for every object or for none")
        slab = self._slabs[self._cur]
        i = self.count
        slab["images"][i, :h, :w] = image
        if mask is not None:
            if self.packed:
                self._pad[...] = 0
                self._pad[:h, :w] = np.asarray(mask) != 0
                bits = np.packbits(self._pad.reshape(-1), bitorder="little")
                slab["masks"][i, :bits.size] = bits
                slab["masks"][i, bits.size:] = 0
            else:
                slab["masks"][i, :h, :w] = np.asarray(mask) != 0
        slab["sizes"][i] = (h, w)
        self.variable |= (h != self.hs or w != self.ws)
        self.labels.append(label)
        self.count += 1
        if self.count == self.capacity:
            self.flush()

    def _collect(self, slab):
        if slab["pending"] is not None:
            self._tables.append(slab["pending"].result())
            slab["pending"] = None

    def flush(self):
        """Hand the objects collected so far to the extractor (one hot-path call) and start a new slab.  Returns the
        slab's table when the batcher is synchronous, None when the call runs in the background."""
        if self.count == 0:
            return None
        n, slab = self.count, self._slabs[self._cur]
        kw = dict(masks_packed=True) if self.packed else {}
        args = (slab["images"][:n], None if slab["masks"] is None else slab["masks"][:n],
                slab["sizes"][:n].copy() if self.variable else None)
        self._all_labels += self.labels
        self.count, self.labels, self.variable = 0, [], False
        if self._pool is None:
            table = self.ex.extract_host_hwc(*args, **kw)
            self._tables.append(table)
            return table
        slab["pending"] = self._pool.submit(self.ex.extract_host_hwc, *args, **kw)
        self._cur ^= 1
        self._collect(self._slabs[self._cur])                # the other slab must be free before it is refilled
        return None

    def finish(self):
        """Flush the tail, wait for everything in flight and return (float64 [N, row_width], labels)."""
        self.flush()
        if self._pool is not None:
            # results in submission order: the slab that was submitted first is the one about to be refilled
            for k in (self._cur, self._cur ^ 1):
                self._collect(self._slabs[k])
        width = self.ex.row_width(self.c)
        table = np.vstack(self._tables) if self._tables else np.zeros((0, width))
        labels = self._all_labels
        self._tables, self._all_labels = [], []
        return table, labels


def read_h5_object(path):
    """One object file as README.md:5-14 describes it -> dict(image, mask, label, channels, ...).
    Needs h5py, which this image does not have; raises ImportError otherwise."""
    import h5py                                    # noqa: F401  (optional dependency)
    out = {}
    with h5py.File(path, "r") as f:
        out["image"] = np.asarray(f["image"], dtype=np.uint16)
        for key in ("mask", "label", "donor", "experiment", "channels"):
            if key in f:
                v = f[key][()]
                out[key] = np.asarray(v) if key == "mask" else v
    return out
