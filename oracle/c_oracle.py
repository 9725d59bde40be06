"""ctypes front end of the plain-C oracle (oracle/imfeat_ref.c).  TEST INFRASTRUCTURE ONLY
(see the header of imfeat_ref.c): used by tests/, smoke() and bench.py's cpu_baseline."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libimfeat_ref.so")
_lib = None

N_BASIC, N_GLCM, N_SHAPE, N_MOM = 17, 6, 10, 9


def build(force=False):
    src = os.path.join(_HERE, "imfeat_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.imfeat_ref_table.restype = None
        _lib.imfeat_ref_plane.restype = None
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def table(planes, masks=None, sizes=None, glcm=True, n_angles=1, shape=False, moments=False):
    """planes: uint16[N, C, Hs, Ws] (plane-compact when sizes is given) -> float64[N, F]."""
    planes = np.ascontiguousarray(planes, dtype=np.uint16)
    N, C, Hs, Ws = planes.shape
    if masks is not None:
        masks = np.ascontiguousarray(masks, dtype=np.uint8)
        assert masks.shape == planes.shape
    if sizes is not None:
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        assert sizes.shape == (N, 2)
    na = n_angles if glcm else 0
    F = C * (N_BASIC + N_GLCM * na + (N_SHAPE if shape else 0) + (N_MOM if moments else 0))
    out = np.empty((N, F), dtype=np.float64)
    lib().imfeat_ref_table(
        _ptr(planes), _ptr(masks), _ptr(sizes), ctypes.c_long(N), ctypes.c_int(C),
        ctypes.c_int(Hs), ctypes.c_int(Ws), ctypes.c_long(Hs * Ws), ctypes.c_int(int(glcm)),
        ctypes.c_int(n_angles), ctypes.c_int(int(shape)), ctypes.c_int(int(moments)), _ptr(out))
    return out


def glcm_counts(plane, mask=None, n_angles=1):
    """uint32[n_angles, 256, 256] raw GLCM bins of one (h, w) uint16 plane."""
    plane = np.ascontiguousarray(plane, dtype=np.uint16)
    h, w = plane.shape
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
    counts = np.zeros((n_angles, 256, 256), dtype=np.uint32)
    lib().imfeat_ref_plane(_ptr(plane), _ptr(mask), ctypes.c_int(h), ctypes.c_int(w),
                           ctypes.c_int(n_angles), None, None, None, None, _ptr(counts))
    return counts
