"""Pinned-host batcher: the step before the hot path.

The reference documents one ``.h5`` file per object with keys ``image`` (16-bit (h,w,c)), ``mask``
((h,w,c), optional), ``label``, ``donor``, ``experiment``, ``channels`` (README.md:5-14) but ships
no loader.  This batcher is the B200-side replacement for that loader: objects are packed, as
they arrive, into fixed-stride page-locked slabs ``uint16[B, Hs, Ws, C]`` (+ ``uint8`` masks) with a
padded-size table ``int32[B, 2]``, which ``imfeat_extract_host_hwc`` DMA-copies without an extra
staging copy and transposes on the device.

h5py / libhdf5 are not part of this image, so ``read_h5_object`` is import-guarded; any source of
(h,w,c) uint16 arrays works.
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
    """Collects (h_i, w_i, C) objects into a fixed-stride slab and extracts full slabs.

    >>> b = PinnedBatcher(extractor, capacity=4096, hs=128, ws=128, channels=12, with_masks=True)
    >>> for img, msk, label in objects: b.add(img, msk, label)
    >>> table, labels = b.finish()
    """

    def __init__(self, extractor, capacity, hs, ws, channels, with_masks=False):
        self.ex = extractor
        self.capacity, self.hs, self.ws, self.c = int(capacity), int(hs), int(ws), int(channels)
        self.images, self._img_t = _pinned_empty((self.capacity, hs, ws, channels), np.uint16)
        self.masks, self._msk_t = (_pinned_empty((self.capacity, hs, ws, channels), np.uint8)
                                   if with_masks else (None, None))
        self.sizes = np.zeros((self.capacity, 2), dtype=np.int32)
        self.count = 0
        self.labels = []
        self._tables = []
        self._all_labels = []
        self.variable = False

    def add(self, image, mask=None, label=None):
        image = np.asarray(image)
        if image.dtype != np.uint16 or image.ndim != 3 or image.shape[2] != self.c:
            raise ValueError("objects must be uint16 (h, w, %d) arrays (README.md:8)" % self.c)
        h, w = image.shape[:2]
        if h > self.hs or w > self.ws or h < 1 or w < 1:
            raise ValueError("object %dx%d does not fit the %dx%d stride" % (h, w, self.hs, self.ws))
        if (self.masks is None) != (mask is None):
            raise ValueError("masks must be given for every object or for none")
        i = self.count
        self.images[i, :h, :w] = image
        if mask is not None:
            self.masks[i, :h, :w] = np.asarray(mask) != 0
        self.sizes[i] = (h, w)
        self.variable |= (h != self.hs or w != self.ws)
        self.labels.append(label)
        self.count += 1
        if self.count == self.capacity:
            self.flush()

    def flush(self):
        """Extract the objects collected so far (one hot-path call) and start a new slab."""
        if self.count == 0:
            return None
        n = self.count
        table = self.ex.extract_host_hwc(
            self.images[:n], None if self.masks is None else self.masks[:n],
            self.sizes[:n] if self.variable else None)
        self._tables.append(table)
        self._all_labels += self.labels
        self.count, self.labels, self.variable = 0, [], False
        return table

    def finish(self):
        """Flush the tail and return (float64 [N, row_width], labels)."""
        self.flush()
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
