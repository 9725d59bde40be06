"""ctypes binding of libimfeat.so (the C ABI in include/imfeat.h).

There is no CPU fallback: if the library is missing or no B200 is visible the import of the
compute entry points fails loudly.
"""
import ctypes
import os

from . import build as _build

_LIB = None


class ImfeatOpts(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("want_basic", ctypes.c_int32),
        ("want_glcm", ctypes.c_int32),
        ("n_angles", ctypes.c_int32),
        ("glcm_distance", ctypes.c_int32),
        ("want_shape", ctypes.c_int32),
        ("want_moments", ctypes.c_int32),
        ("host_mask_bits", ctypes.c_int32),
        ("percentiles", ctypes.c_double * 9),
    ]


EXPORTS = [
    "imfeat_default_opts", "imfeat_abi_version", "imfeat_create", "imfeat_destroy",
    "imfeat_last_error", "imfeat_row_width", "imfeat_extract_device", "imfeat_extract_host",
    "imfeat_extract_host_hwc", "imfeat_glcm_counts_device", "imfeat_pack_hwc_device",
    "imfeat_synth_device", "imfeat_launch_count", "imfeat_enable_timing", "imfeat_kernel_times",
    "imfeat_minmax_fit_device", "imfeat_minmax_transform_device",
]


class ImfeatError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """Load the shared library; raises if it cannot be found/built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise ImfeatError("CUDA library %s is missing (run __graft_entry__.build())" % path)
        _build.build_library()
    L = ctypes.CDLL(path)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
    po = ctypes.POINTER(ImfeatOpts)
    L.imfeat_default_opts.argtypes = [po]
    L.imfeat_default_opts.restype = None
    L.imfeat_abi_version.argtypes = []
    L.imfeat_abi_version.restype = ctypes.c_int
    L.imfeat_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.imfeat_create.restype = ctypes.c_int
    L.imfeat_destroy.argtypes = [vp]
    L.imfeat_destroy.restype = ctypes.c_int
    L.imfeat_last_error.argtypes = [vp]
    L.imfeat_last_error.restype = ctypes.c_char_p
    L.imfeat_launch_count.argtypes = [vp]
    L.imfeat_launch_count.restype = i64
    L.imfeat_row_width.argtypes = [i32, po]
    L.imfeat_row_width.restype = i64
    L.imfeat_extract_device.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i64, po,
                                        vp, i64, vp, vp]
    L.imfeat_extract_device.restype = ctypes.c_int
    L.imfeat_extract_host.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i64, po, vp, i64, vp]
    L.imfeat_extract_host.restype = ctypes.c_int
    L.imfeat_extract_host_hwc.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, po, vp, i64, vp]
    L.imfeat_extract_host_hwc.restype = ctypes.c_int
    L.imfeat_glcm_counts_device.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i64, po, vp, vp]
    L.imfeat_glcm_counts_device.restype = ctypes.c_int
    L.imfeat_pack_hwc_device.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i64, vp, vp, vp]
    L.imfeat_pack_hwc_device.restype = ctypes.c_int
    L.imfeat_synth_device.argtypes = [vp, u64, i64, i64, i32, i32, i32, i64, i32, i32, i32, i32,
                                      vp, vp, vp, vp]
    L.imfeat_synth_device.restype = ctypes.c_int
    L.imfeat_enable_timing.argtypes = [vp, i32]
    L.imfeat_enable_timing.restype = ctypes.c_int
    L.imfeat_kernel_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64), i32]
    L.imfeat_kernel_times.restype = ctypes.c_int
    L.imfeat_minmax_fit_device.argtypes = [vp, vp, i64, i32, i64, vp, vp]
    L.imfeat_minmax_fit_device.restype = ctypes.c_int
    L.imfeat_minmax_transform_device.argtypes = [vp, vp, i64, i32, i64, vp, vp, i64, vp]
    L.imfeat_minmax_transform_device.restype = ctypes.c_int
    _LIB = L
    return L


_TORCH_OPS = None


def load_torch_ops(build_if_missing=True):
    """Load the thin PyTorch C++ extension (csrc_torch/imfeat_torch.cpp -> libimfeat_torch.so) that registers
    ``torch.ops.imfeat.extract`` / ``glcm_counts`` / ``row_width`` over the same C ABI.  Raises if it is missing."""
    global _TORCH_OPS
    if _TORCH_OPS is not None:
        return _TORCH_OPS
    import torch
    load(build_if_missing)
    path = _build.TORCH_EXT_PATH
    if not os.path.exists(path):
        if not build_if_missing:
            raise ImfeatError("torch extension %s is missing (run __graft_entry__.build())" % path)
        _build.build_torch_extension()
    torch.ops.load_library(path)
    _TORCH_OPS = torch.ops.imfeat
    return _TORCH_OPS


def check(rc, ctx=None):
    if rc != 0:
        msg = load().imfeat_last_error(ctx)
        raise ImfeatError("imfeat error %d: %s" % (rc, msg.decode() if msg else "?"))
