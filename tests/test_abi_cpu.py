"""CPU tests of the boundary: the library builds, loads and exports every symbol the header
declares (no compute calls without a GPU), the schema matches the notebook, and the product
refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

import imfeat_b200 as imf
from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "imfeat.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(imfeat_[a-z_0-9]+)\s*\(", hdr)))


def test_library_builds_and_exports_header_symbols():
    from imfeat_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), "header declares %s but the library does not export it" % n
    assert sorted(_lib.EXPORTS) == names
    assert L.imfeat_abi_version() == 1


def test_torch_extension_builds_and_registers_the_ops():
    """The thin PyTorch C++ extension over the C ABI (north_star: "calls CUDA through a thin PyTorch C++/CUDA
    extension"): builds with g++, loads, and registers imfeat::extract / glcm_counts / row_width with the
    schema SURVEY 8(b) sketches.  No compute call without a GPU."""
    import torch
    from imfeat_b200 import _lib, build
    path = build.build_torch_extension()
    assert os.path.exists(path)
    ops = _lib.load_torch_ops()
    schema = str(torch.ops.imfeat.extract.default._schema)
    for piece in ("Tensor planes", "Tensor? masks", "Tensor? sizes", "Tensor? src_obj", "Tensor? chan", "float[] percentiles",
                  "int ctx=0", "-> Tensor"):
        assert piece in schema, schema
    assert "Tensor planes" in str(torch.ops.imfeat.glcm_counts.default._schema)
    assert ops.row_width(3, True, True, 1, False, False) == 69                          # NB:317
    assert ops.row_width(12, True, True, 4, True, True) == 720
    if not torch.cuda.is_available():
        with pytest.raises((RuntimeError, NotImplementedError)):                        # no CPU kernel is registered
            ops.extract(torch.zeros((1, 1, 64), dtype=torch.int16), None, None, None, None, 8, 8, True, True, 1, 5,
                        False, False, [], None, None, 0)


def test_default_opts_are_the_notebook_literals():
    from imfeat_b200 import _lib
    L = _lib.load()
    o = _lib.ImfeatOpts()
    L.imfeat_default_opts(ctypes.byref(o))
    assert o.struct_size == ctypes.sizeof(_lib.ImfeatOpts)
    assert (o.want_basic, o.want_glcm, o.n_angles, o.glcm_distance) == (1, 1, 1, 5)     # NB:298
    assert list(o.percentiles) == [k / 10.0 for k in range(1, 10)]                      # NB:242-250
    assert L.imfeat_row_width(3, ctypes.byref(o)) == 69                                 # NB:317
    assert L.imfeat_row_width(12, ctypes.byref(o)) == 276
    o.n_angles, o.want_shape, o.want_moments = 4, 1, 1
    assert L.imfeat_row_width(12, ctypes.byref(o)) == 12 * (17 + 24 + 10 + 9)


def test_schema_matches_notebook_order(golden):
    img, want, cols = golden["blob_64x64x3"]
    assert imf.feature_columns(3) == cols
    from oracle import notebook_oracle as orc
    for kw in [dict(glcm=True, n_angles=4, shape=True, moments=True), dict(glcm=False)]:
        okw = dict(glcm=kw["glcm"], four_directions=kw.get("n_angles", 1) == 4,
                   shape=kw.get("shape", False), moments=kw.get("moments", False))
        assert imf.feature_columns(5, **kw) == orc.column_names(5, **okw)


def test_channel_column_index():
    from imfeat_b200 import schema
    cols = imf.feature_columns(4, glcm=True, n_angles=4, shape=True)
    for slot in range(4):
        idx = schema.channel_column_index(4, slot, glcm=True, n_angles=4, shape=True)
        assert len(idx) == 17 + 24 + 10
        assert all(cols[i].endswith("_Ch%d" % (slot + 1)) for i in idx)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(imf.ImfeatError):
        imf.extract_features(np.zeros((1, 8, 8, 2), np.uint16))
    with pytest.raises(imf.ImfeatError):
        imf.MinMaxScaler().fit(np.zeros((4, 3)))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "interpretable-multichannel-image-analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "imfeat_ref" not in txt, f


def test_synth_is_deterministic_and_in_range():
    from imfeat_b200 import synth
    a, m = synth.synth_plane(7, 11, 3, 64, 64)
    b, m2 = synth.synth_plane(7, 11, 3, 64, 64)
    assert (a == b).all() and (m == m2).all() and a.dtype == np.uint16
    assert 0.10 < m.mean() < 0.65 and a.max() <= 4500 and a.min() >= 0
    c, _ = synth.synth_plane(7, 12, 3, 64, 64)
    assert (a != c).any()
    fr = [synth.synth_plane(0, i, 0, 64, 64)[1].mean() for i in range(40)]
    assert 0.12 < min(fr) and max(fr) < 0.65
    h, w = synth.object_size(3, 5, 128, 128, True, 16, 16)
    assert 16 <= h <= 128 and 16 <= w <= 128


def test_pinned_batcher_packing_cpu():
    """Packing logic of the batcher (no GPU needed: a stub extractor records what it is given)."""
    from imfeat_b200 import PinnedBatcher

    class Stub:
        calls = []

        def row_width(self, c):
            return 23 * c

        def extract_host_hwc(self, images, masks=None, sizes=None):
            self.calls.append((images.copy(), None if masks is None else masks.copy(),
                               None if sizes is None else sizes.copy()))
            return np.full((images.shape[0], 23 * images.shape[3]), float(len(self.calls)))

    rng = np.random.default_rng(0)
    stub = Stub()
    b = PinnedBatcher(stub, capacity=3, hs=16, ws=12, channels=2, with_masks=True)
    objs = [(rng.integers(0, 4096, (h, w, 2)).astype(np.uint16), rng.random((h, w, 2)) < 0.5)
            for h, w in [(16, 12), (9, 7), (16, 12), (5, 12), (16, 12)]]
    for k, (img, m) in enumerate(objs):
        b.add(img, m, label="o%d" % k)
    table, labels = b.finish()
    assert table.shape == (5, 46) and labels == ["o%d" % k for k in range(5)]
    assert len(stub.calls) == 2                            # a full slab of 3, then the tail of 2
    imgs, masks, sizes = stub.calls[0]
    assert sizes.tolist() == [[16, 12], [9, 7], [16, 12]]
    assert (imgs[1, :9, :7] == objs[1][0]).all() and (masks[1, :9, :7] == objs[1][1]).all()
    imgs, masks, sizes = stub.calls[1]
    assert sizes.tolist() == [[5, 12], [16, 12]]
    with pytest.raises(ValueError):
        b.add(np.zeros((17, 12, 2), np.uint16), np.zeros((17, 12, 2), bool))
    with pytest.raises(ValueError):
        b.add(np.zeros((4, 4, 2), np.float32), np.zeros((4, 4, 2), bool))
