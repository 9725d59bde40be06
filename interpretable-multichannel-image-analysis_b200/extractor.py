"""Host-side mirror of the reference's feature-extraction interface over the CUDA library.

Reference interface being mirrored (NB = raw line of the reference notebook):
  * ``basic_statistical_features(image) -> dict``   NB:220-264
  * ``glcm_features(image) -> dict``                NB:269-308
  * the extraction loop that fills ``df_features``  NB:327-334, NB:358-364
``extract_features(images, masks, channels)`` replaces that loop for a whole batch.

PyTorch is used only for device memory and streams; all arithmetic is in libimfeat.so.
"""
import ctypes

import numpy as np

from . import _lib, schema


def _torch():
    import torch
    return torch


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.ImfeatError(
            "no CUDA device visible: this package has no CPU fallback (B200 / sm_100a only)")
    return torch


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _np_ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def plane_stride_for(hs, ws):
    return (hs * ws + 7) & ~7


def pack_mask_bits(masks):
    """uint8/bool masks [N, ...] -> the bit-packed host format of the C ABI (include/imfeat.h, host_mask_bits):
    element k of an object is bit k & 7 of byte k >> 3, every object padded to a multiple of 8 bytes.  A mask is a
    third of the bytes that cross PCIe per object; packed it is 4 %."""
    m = np.asarray(masks)
    n = m.shape[0]
    bits = np.packbits(m.reshape(n, -1) != 0, axis=1, bitorder="little")
    want = ((m[0].size + 63) // 64) * 8 if n else 0
    if bits.shape[1] != want:
        bits = np.concatenate([bits, np.zeros((n, want - bits.shape[1]), np.uint8)], axis=1)
    return np.ascontiguousarray(bits)


def _planes3(t):
    """[N, C, Hs, Ws] -> ([N, C, plane_stride], hs, ws); pads the plane to a multiple of 8
    elements when Hs*Ws is not one (the C ABI wants 16-byte aligned planes)."""
    torch = _torch()
    N, C, hs, ws = (int(v) for v in t.shape)
    stride = plane_stride_for(hs, ws)
    if stride == hs * ws:
        return t.reshape(N, C, stride), hs, ws
    p = torch.zeros((N, C, stride), dtype=t.dtype, device=t.device)
    p[:, :, :hs * ws] = t.reshape(N, C, hs * ws)
    return p, hs, ws


class FeatureExtractor:
    """One context (lookup tables, staging buffers) on one GPU + a fixed option set.

    Defaults are the notebook's literals: 17 basic features + GLCM(distance 5, angle 0, 256
    levels), percentile arguments 0.1..0.9 (NB:242-250, NB:298).
    """

    def __init__(self, device=None, basic=True, glcm=True, four_directions=False, shape=False,
                 moments=False, percentiles=schema.NOTEBOOK_PERCENTILES, distance=5):
        torch = _require_cuda()
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else torch.device(device).index or 0
        if len(percentiles) != 9:
            raise ValueError("exactly nine percentile arguments are supported (NB:242-250)")
        self.opts = _lib.ImfeatOpts()
        self.lib.imfeat_default_opts(ctypes.byref(self.opts))
        self.opts.want_basic = int(bool(basic))
        self.opts.want_glcm = int(bool(glcm))
        self.opts.n_angles = 4 if four_directions else 1
        self.opts.glcm_distance = int(distance)
        self.opts.want_shape = int(bool(shape))
        self.opts.want_moments = int(bool(moments))
        for k, q in enumerate(percentiles):
            self.opts.percentiles[k] = float(q)
        self._ctx = ctypes.c_void_p()
        _lib.check(self.lib.imfeat_create(int(self.device), ctypes.byref(self._ctx)))

    # -- bookkeeping -----------------------------------------------------------------------
    def clone(self):
        """A second extractor with the same options and its OWN context (work buffers, scheduler counters):
        what a captured CUDA graph must use, because a graph bakes the context's buffer addresses in."""
        o = self.opts
        return FeatureExtractor(device=self.device, basic=bool(o.want_basic), glcm=bool(o.want_glcm),
                                four_directions=o.n_angles == 4, shape=bool(o.want_shape),
                                moments=bool(o.want_moments), percentiles=tuple(o.percentiles),
                                distance=int(o.glcm_distance))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.imfeat_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_angles(self):
        return int(self.opts.n_angles)

    def row_width(self, c_out):
        return int(self.lib.imfeat_row_width(int(c_out), ctypes.byref(self.opts)))

    def columns(self, n_channels, channel_ids=None):
        o = self.opts
        return schema.feature_columns(n_channels, glcm=bool(o.want_glcm), n_angles=o.n_angles,
                                      shape=bool(o.want_shape), moments=bool(o.want_moments),
                                      basic=bool(o.want_basic), channel_ids=channel_ids)

    def launch_count(self):
        return int(self.lib.imfeat_launch_count(self._ctx))

    def enable_timing(self, on=True):
        _lib.check(self.lib.imfeat_enable_timing(self._ctx, int(bool(on))), self._ctx)

    def kernel_times(self, reset=True):
        """Accumulated (ms, launches) per kernel group K1..K4 since the last reset."""
        ms = (ctypes.c_double * 4)()
        calls = (ctypes.c_int64 * 4)()
        _lib.check(self.lib.imfeat_kernel_times(self._ctx, ms, calls, int(bool(reset))), self._ctx)
        return list(ms), list(calls)

    def _stream(self, stream=None):
        torch = _torch()
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    # -- device-resident hot path ------------------------------------------------------------
    def extract_planar(self, planes, masks=None, sizes=None, hs=None, ws=None, src_obj=None,
                       chan=None, out=None, status=None, stream=None):
        """planes: cuda uint16 [N, C, Hs, Ws] or [N, C, plane_stride] (then give hs, ws).
        Returns a cuda float64 [N, row_width] table (``out`` is reused when given)."""
        torch = _torch()
        assert planes.is_cuda and planes.is_contiguous() and planes.element_size() == 2
        if planes.dim() == 4:
            planes, hs, ws = _planes3(planes)
            if masks is not None and masks.dim() == 4:
                masks = _planes3(masks)[0]
        assert hs is not None and ws is not None
        N, C, stride = (int(v) for v in planes.shape)
        if masks is not None:
            assert masks.is_cuda and masks.is_contiguous() and masks.element_size() == 1
            assert masks.numel() == planes.numel()
        c_out = C if chan is None else int(chan.numel())
        width = self.row_width(c_out)
        if out is None:
            out = torch.empty((N, width), dtype=torch.float64, device=planes.device)
        assert out.is_cuda and out.dtype == torch.float64 and out.stride(1) == 1
        for idx in (sizes, src_obj, chan):
            assert idx is None or (idx.is_cuda and idx.dtype == torch.int32 and idx.is_contiguous())
        # through the torch op (csrc_torch/imfeat_torch.cpp): tensors in, current stream taken from torch
        ops = _lib.load_torch_ops()
        o = self.opts
        args = (planes.view(torch.int16) if planes.dtype != torch.int16 else planes,
                None if masks is None else masks.view(torch.uint8).reshape(N, C, stride), sizes, src_obj, chan, int(hs), int(ws),
                bool(o.want_basic), bool(o.want_glcm), int(o.n_angles), int(o.glcm_distance), bool(o.want_shape),
                bool(o.want_moments), [float(q) for q in o.percentiles], out,
                None if status is None else status.view(torch.int32), int(self._ctx.value))
        try:
            if stream is None:
                ops.extract(*args)
            else:
                with torch.cuda.stream(stream):
                    ops.extract(*args)
        except RuntimeError as exc:                        # TORCH_CHECK (shape / dtype / device misuse) or a C-ABI error
            raise _lib.ImfeatError(str(exc).split("\n")[0]) from None
        return out

    def pack_hwc(self, images, masks=None, sizes=None, stream=None):
        """cuda uint16 [N, hs, ws, C] (README.md:8 layout) -> planar [N, C, plane_stride]."""
        torch = _torch()
        assert images.is_cuda and images.is_contiguous() and images.element_size() == 2
        N, hs, ws, C = (int(v) for v in images.shape)
        stride = plane_stride_for(hs, ws)
        planes = torch.empty((N, C, stride), dtype=torch.uint16, device=images.device)
        pmasks = None
        if masks is not None:
            assert masks.is_cuda and masks.is_contiguous() and masks.element_size() == 1
            pmasks = torch.empty((N, C, stride), dtype=torch.uint8, device=images.device)
        _lib.check(self.lib.imfeat_pack_hwc_device(
            self._ctx, _ptr(images), _ptr(masks), _ptr(sizes), N, C, hs, ws, stride, _ptr(planes),
            _ptr(pmasks), self._stream(stream)), self._ctx)
        return planes, pmasks, hs, ws

    def glcm_counts(self, planes, masks=None, sizes=None, hs=None, ws=None, stream=None):
        """Raw GLCM bins (NB:298): cuda int32 [N, C, n_angles, 256, 256] (values are uint32)."""
        torch = _torch()
        if planes.dim() == 4:
            planes, hs, ws = _planes3(planes)
            if masks is not None and masks.dim() == 4:
                masks = _planes3(masks)[0]
        N, C, stride = (int(v) for v in planes.shape)
        ops = _lib.load_torch_ops()
        args = (planes.view(torch.int16) if planes.dtype != torch.int16 else planes,
                None if masks is None else masks.view(torch.uint8).reshape(N, C, stride), sizes, int(hs), int(ws),
                int(self.n_angles), int(self.opts.glcm_distance), int(self._ctx.value))
        if stream is None:
            return ops.glcm_counts(*args)
        with torch.cuda.stream(stream):
            return ops.glcm_counts(*args)

    def synth(self, seed, first, count, c, hs, ws, with_masks=True, variable=False, hmin=1, wmin=1,
              mask_shrink=256, stream=None):
        """Generate synthetic objects on the device (numpy mirror: synth.py).
        Returns (planes [N,C,stride] uint16, masks or None, sizes or None)."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        stride = plane_stride_for(hs, ws)
        planes = torch.empty((count, c, stride), dtype=torch.uint16, device=dev)
        masks = torch.empty((count, c, stride), dtype=torch.uint8, device=dev) if with_masks else None
        sizes = torch.empty((count, 2), dtype=torch.int32, device=dev) if variable else None
        _lib.check(self.lib.imfeat_synth_device(
            self._ctx, ctypes.c_uint64(seed), int(first), int(count), c, hs, ws, stride,
            int(variable), hmin, wmin, mask_shrink, _ptr(planes), _ptr(masks), _ptr(sizes),
            self._stream(stream)), self._ctx)
        return planes, masks, sizes

    # -- host buffers (the reference-facing call) ----------------------------------------------
    def extract_host_hwc(self, images, masks=None, sizes=None, out=None, return_status=False, masks_packed=False):
        """images: numpy uint16 [N, hs, ws, C] (pinned or pageable), masks uint8 same shape -- or, with
        masks_packed=True, the bit-packed form ``pack_mask_bits`` returns --, sizes int32 [N, 2] or None.
        Returns numpy float64 [N, row_width]."""
        images = np.ascontiguousarray(images)
        assert images.dtype == np.uint16 and images.ndim == 4
        N, hs, ws, C = images.shape
        opts = self.opts
        if masks is not None and masks_packed:
            masks = np.ascontiguousarray(masks)
            assert masks.dtype == np.uint8 and masks.shape == (N, ((hs * ws * C + 63) // 64) * 8)
            opts = _lib.ImfeatOpts.from_buffer_copy(self.opts)
            opts.host_mask_bits = 1
        elif masks is not None:
            masks = np.ascontiguousarray(masks)
            if masks.dtype == np.bool_:
                masks = masks.view(np.uint8)
            assert masks.dtype == np.uint8 and masks.shape == images.shape
        if sizes is not None:
            sizes = np.ascontiguousarray(sizes, dtype=np.int32)
            assert sizes.shape == (N, 2)
        width = self.row_width(C)
        if out is None:
            out = np.empty((N, width), dtype=np.float64)
        status = np.zeros(N, dtype=np.uint32) if return_status else None
        _lib.check(self.lib.imfeat_extract_host_hwc(
            self._ctx, _np_ptr(images), _np_ptr(masks), _np_ptr(sizes), N, C, hs, ws,
            ctypes.byref(opts), _np_ptr(out), int(out.strides[0] // 8), _np_ptr(status)),
            self._ctx)
        return (out, status) if return_status else out

    def extract_host_planar(self, planes, masks=None, sizes=None, hs=None, ws=None,
                            return_status=False, masks_packed=False):
        """planes: numpy uint16 [N, C, Hs, Ws] or [N, C, plane_stride] (give hs, ws); masks the same shape, or
        with masks_packed=True ``pack_mask_bits(masks.reshape(N * C, -1))`` (every plane packed on its own)."""
        planes = np.ascontiguousarray(planes)
        assert planes.dtype == np.uint16
        N, C = planes.shape[:2]
        if planes.ndim == 4:
            hs, ws = planes.shape[2:]
            stride = hs * ws
        else:
            stride = planes.shape[2]
        opts = self.opts
        if masks is not None and masks_packed:
            masks = np.ascontiguousarray(masks)
            assert masks.dtype == np.uint8 and masks.size == N * C * ((stride + 63) // 64) * 8
            opts = _lib.ImfeatOpts.from_buffer_copy(self.opts)
            opts.host_mask_bits = 1
        elif masks is not None:
            masks = np.ascontiguousarray(masks).view(np.uint8)
        if sizes is not None:
            sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        width = self.row_width(C)
        out = np.empty((N, width), dtype=np.float64)
        status = np.zeros(N, dtype=np.uint32) if return_status else None
        _lib.check(self.lib.imfeat_extract_host(
            self._ctx, _np_ptr(planes), _np_ptr(masks), _np_ptr(sizes), N, C, hs, ws, stride,
            ctypes.byref(opts), _np_ptr(out), width, _np_ptr(status)), self._ctx)
        return (out, status) if return_status else out


# ---------------------------------------------------------------------------------------------
# Reference-shaped functional API
# ---------------------------------------------------------------------------------------------
_CACHE = {}


def get_extractor(device=None, **opts):
    torch = _require_cuda()
    dev = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    key = (dev,) + tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in opts.items()))
    if key not in _CACHE:
        _CACHE[key] = FeatureExtractor(device=dev, **opts)
    return _CACHE[key]


def _as_batch(images, masks):
    """Normalise the accepted input forms to (uint16 [N,hs,ws,C], uint8 masks or None, sizes or None)."""
    if isinstance(images, np.ndarray) and images.ndim == 4:
        return images, masks, None
    if isinstance(images, np.ndarray) and images.ndim == 3:
        return images[None], None if masks is None else np.asarray(masks)[None], None
    objs = [np.asarray(im) for im in images]
    if not objs:
        raise ValueError("empty image list")
    C = objs[0].shape[2]
    hs = max(o.shape[0] for o in objs)
    ws = max(o.shape[1] for o in objs)
    same = all(o.shape == objs[0].shape for o in objs)
    if same:
        batch = np.stack(objs)
        mb = None if masks is None else np.stack([np.asarray(m) for m in masks])
        return batch, mb, None
    batch = np.zeros((len(objs), hs, ws, C), dtype=np.uint16)
    mb = None if masks is None else np.zeros((len(objs), hs, ws, C), dtype=np.uint8)
    sizes = np.zeros((len(objs), 2), dtype=np.int32)
    for i, o in enumerate(objs):
        if o.shape[2] != C:
            raise ValueError("all objects must have the same channel count")
        h, w = o.shape[:2]
        batch[i, :h, :w] = o
        if mb is not None:
            mb[i, :h, :w] = np.asarray(masks[i]) != 0
        sizes[i] = (h, w)
    return batch, mb, sizes


def _as_uint16(batch):
    """The reference's objects are 16-bit unsigned images (README.md:8).  Other integer types are accepted
    when every value fits; floating-point input (the notebook's own demo feeds MedNIST JPEGs scaled to
    [0, 1], NB:328) is refused rather than truncated: rescale it first, e.g. ``from_unit_float``."""
    if batch.dtype == np.uint16:
        return batch
    if np.issubdtype(batch.dtype, np.integer) and (batch.size == 0 or (batch.min() >= 0 and batch.max() <= 65535)):
        return batch.astype(np.uint16)
    raise TypeError("images must be 16-bit unsigned integers (README.md:8), got %s" % batch.dtype)


def from_unit_float(images, scale=255):
    """Adapter for the notebook's demo data (NB:328, NB:360: ``add_two_noise_channels(im) / 255.``, float64
    in [0, 1]): back to the integers they came from, ``rint(images * scale)`` as uint16.  Scale-free columns
    (kurtosis, skew, entropy, all GLCM properties) are unchanged by the rescaling; intensity columns come
    out in units of 1/scale."""
    a = np.rint(np.asarray(images, dtype=np.float64) * scale)
    if a.size and (a.min() < 0 or a.max() > 65535):
        raise ValueError("scaled values leave the 16-bit range")
    return a.astype(np.uint16)


def extract_features(images, masks=None, channels=None, *, basic=True, glcm=True, four_directions=False,
                     shape=False, moments=False, percentiles=schema.NOTEBOOK_PERCENTILES,
                     distance=5, as_frame=False, device=None, return_status=False):
    """Drop-in for the notebook's extraction loop (NB:358-364).

    images   : uint16 ndarray [N, h, w, C], a single (h, w, C) object, or a sequence of
               (h_i, w_i, C) objects (README.md:8 layout)
    masks    : same shape(s), non-zero = inside; None = the notebook's semantics (no mask)
    channels : optional list of channel indices to extract (column suffixes stay positional,
               ``Ch1..`` as in NB:241); a list of names is accepted and ignored (metadata)
    Returns the float64 [N, 23*C] table in the notebook's column order (a pandas DataFrame with
    the notebook's column names when ``as_frame`` is true).
    """
    batch, mb, sizes = _as_batch(images, masks)
    batch = _as_uint16(batch)
    if channels is not None and len(channels) and not isinstance(channels[0], str):
        sel = [int(c) for c in channels]
        batch = np.ascontiguousarray(batch[..., sel])
        if mb is not None:
            mb = np.ascontiguousarray(mb[..., sel])
    if mb is not None and mb.dtype != np.uint8:
        mb = (np.asarray(mb) != 0).astype(np.uint8)
    ex = get_extractor(device, basic=basic, glcm=glcm, four_directions=four_directions, shape=shape,
                       moments=moments, percentiles=tuple(percentiles), distance=distance)
    res = ex.extract_host_hwc(batch, mb, sizes, return_status=return_status)
    table, status = res if return_status else (res, None)
    if as_frame:
        import pandas as pd
        table = pd.DataFrame(table, columns=ex.columns(batch.shape[3]))
    return (table, status) if return_status else table


def basic_statistical_features(image, device=None):
    """Same call and return shape as the reference function (NB:220-264): dict of 17 scalars
    per channel keyed ``<name>_Ch<k>``, for one (M, N, C) uint16 image."""
    image = np.asarray(image)
    row = extract_features(image[None], glcm=False, device=device)[0]
    return dict(zip(schema.feature_columns(image.shape[2], glcm=False), (float(v) for v in row)))


def glcm_features(image, device=None):
    """Same call and return shape as the reference function (NB:269-308)."""
    image = np.asarray(image)
    row = extract_features(image[None], basic=False, glcm=True, device=device)[0]    # same dtype / range checks
    return dict(zip(schema.feature_columns(image.shape[2], basic=False, glcm=True), (float(v) for v in row)))
