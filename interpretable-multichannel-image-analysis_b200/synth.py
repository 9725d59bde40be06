"""Synthetic imaging-flow-cytometry-like objects from a counter-based hash.

numpy mirror of ``synth_kernel`` in csrc/aux_kernels.cuh (keep the two in sync): every pixel is a
pure function of (seed, object, channel, pixel index), so any object of a device-generated data
set can be regenerated on the host for sampled parity checks.

Per plane: integer background ``offset`` (100..1000) + ~Gaussian integer noise (sum of four bytes
of the hash, sigma ~ 9..37) + a paraboloid blob over an ellipse with peak 500..3095; the mask is
the (optionally shrunk) ellipse.  Values are clamped to uint16.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _sm64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def object_key(seed, obj):
    return _sm64(np.uint64(seed) ^ _sm64(np.asarray(obj, dtype=np.uint64)))


def object_size(seed, obj, hs, ws, variable=False, hmin=1, wmin=1):
    if not variable:
        return hs, ws
    k = int(object_key(seed, obj))
    return hmin + (k & 0xFFFF) % (hs - hmin + 1), wmin + ((k >> 16) & 0xFFFF) % (ws - wmin + 1)


def synth_plane(seed, obj, ch, h, w, mask_shrink=256):
    """(uint16[h,w], uint8[h,w]) of one plane."""
    okey = object_key(seed, obj)
    with np.errstate(over="ignore"):
        k = _sm64(okey + np.uint64(ch + 1) * np.uint64(0xD1B54A32D192ED03))
    k1, k2 = int(_sm64(k ^ np.uint64(1))), int(_sm64(k ^ np.uint64(2)))
    offset = 100 + (k1 & 0xFFFF) % 901
    sig = 8 + ((k1 >> 16) & 0xFF) % 25
    amp = 500 + ((k1 >> 24) & 0xFFFF) % 2596
    fx = 56 + (k2 & 0xFF) % 56
    fy = 56 + ((k2 >> 8) & 0xFF) % 56
    RX = max(2, (2 * w * fx) >> 8)
    RY = max(2, (2 * h * fy) >> 8)
    cx2 = (w - 1) + ((k2 >> 16) & 0xFF) % (w // 4 + 1) - w // 8
    cy2 = (h - 1) + ((k2 >> 24) & 0xFF) % (h // 4 + 1) - h // 8
    D = RX * RX * RY * RY
    idx = np.arange(h * w, dtype=np.int64)
    r, c = idx // w, idx % w
    dx, dy = 2 * c - cx2, 2 * r - cy2
    E = dx * dx * (RY * RY) + dy * dy * (RX * RX)
    blob = np.where(E < D, amp * (D - E) // D, 0)
    with np.errstate(over="ignore"):
        u = _sm64(k + np.uint64(0x1000) + idx.astype(np.uint64))
    s4 = ((u & np.uint64(0xFF)) + ((u >> np.uint64(8)) & np.uint64(0xFF))
          + ((u >> np.uint64(16)) & np.uint64(0xFF)) + ((u >> np.uint64(24)) & np.uint64(0xFF))).astype(np.int64)
    noise = ((s4 - 510) * sig) >> 7
    val = np.clip(offset + noise + blob, 0, 65535).astype(np.uint16)
    mask = (E * 256 < D * mask_shrink).astype(np.uint8)
    return val.reshape(h, w), mask.reshape(h, w)


def synth_objects(seed, first, count, c, hs, ws, variable=False, hmin=1, wmin=1, mask_shrink=256):
    """List of (image uint16 (h,w,c), mask uint8 (h,w,c)) in the reference's (h,w,c) layout."""
    out = []
    for obj in range(first, first + count):
        h, w = object_size(seed, obj, hs, ws, variable, hmin, wmin)
        planes = [synth_plane(seed, obj, ch, h, w, mask_shrink) for ch in range(c)]
        out.append((np.stack([p[0] for p in planes], axis=2), np.stack([p[1] for p in planes], axis=2)))
    return out


def synth_batch_hwc(seed, first, count, c, hs, ws, mask_shrink=256):
    """Fixed-size batch: (uint16[N,hs,ws,c], uint8[N,hs,ws,c])."""
    objs = synth_objects(seed, first, count, c, hs, ws, mask_shrink=mask_shrink)
    return np.stack([o[0] for o in objs]), np.stack([o[1] for o in objs])
