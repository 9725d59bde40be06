"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.
Bit-exact for integer outputs, percentiles and GLCM bins; 1e-9 relative (north_star allows 1e-5)
for floating-point statistics and Haralick features."""
import numpy as np
import pytest

import imfeat_b200 as imf
from conftest import compare_tables, parity_distributions
from oracle import c_oracle
from oracle import notebook_oracle as orc

pytestmark = pytest.mark.gpu


def _planar(img):
    return np.ascontiguousarray(np.asarray(img).transpose(0, 3, 1, 2))


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch


def test_golden_fixtures(golden, torch_mod):
    """Reference-generated vectors (tests/golden/make_golden.py) through extract_features."""
    for name, (img, want, cols) in golden.items():
        got = imf.extract_features(img[None])
        assert imf.feature_columns(img.shape[2]) == cols
        compare_tables(got, want[None], cols, label=name)


def test_reference_shaped_functions(golden, torch_mod):
    img, want, cols = golden["mix_64x64x4"]
    d = imf.basic_statistical_features(img)            # NB:362
    g = imf.glcm_features(img)                         # NB:363
    feats = {}
    feats.update(d)
    feats.update(g)
    assert list(feats.keys()) == cols
    compare_tables(np.array([list(feats.values())]), want[None], cols, label="dicts")
    frame = imf.extract_features(img[None], as_frame=True)
    assert list(frame.columns) == cols


@pytest.mark.parametrize("shape", [(64, 64), (37, 91), (128, 128), (9, 5), (7, 6), (1, 1), (33, 17),
                                   (3, 300), (181, 181), (2, 16384), (20, 8), (50, 16), (41, 32), (30, 256)])
def test_distributions_all_blocks(shape, torch_mod):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    d = parity_distributions(rng, h, w)
    names = sorted(d)
    img = np.stack([d[k] for k in names], axis=2)[None]
    want = c_oracle.table(_planar(img), glcm=True, n_angles=4, shape=True, moments=True)
    got, status = imf.extract_features(img, four_directions=True, shape=True, moments=True,
                                       return_status=True)
    cols = imf.feature_columns(len(names), n_angles=4, shape=True, moments=True)
    compare_tables(got, want, cols, label=str(shape), images=img)
    if h * w > 1:
        assert status[0] & 4          # constant channels present
    # notebook-default call
    want = c_oracle.table(_planar(img))
    got = imf.extract_features(img)
    compare_tables(got, want, imf.feature_columns(len(names)), label="default " + str(shape), images=img)


def test_against_numpy_oracle_directly(torch_mod):
    """Small case checked against the numpy restatement itself (not the C port)."""
    rng = np.random.default_rng(3)
    d = parity_distributions(rng, 48, 40)
    img = np.stack([d[k] for k in sorted(d)], axis=2)
    want, cols = orc.oracle_extract([img], glcm=True, four_directions=True, shape=True, moments=True)
    got = imf.extract_features(img[None], four_directions=True, shape=True, moments=True)
    compare_tables(got, want, cols, label="numpy oracle", images=img[None])


def test_glcm_bins_bit_exact(torch_mod):
    torch = torch_mod
    rng = np.random.default_rng(11)
    for (h, w) in [(64, 64), (33, 47), (128, 128), (9, 5)]:
        d = parity_distributions(rng, h, w)
        names = sorted(d)
        img = np.stack([d[k] for k in names], axis=2)[None]
        ex = imf.get_extractor(four_directions=True)
        planes = torch.from_numpy(_planar(img)).cuda()
        counts = ex.glcm_counts(planes).cpu().numpy().view(np.uint32)
        for c, k in enumerate(names):
            want = orc.glcm_counts(d[k], angles=orc.ANGLES4)
            for a in range(4):
                assert (counts[0, c, a] == want[:, :, a]).all(), (h, w, k, a)
        # masked bins
        mask = (rng.random((1, h, w, len(names))) < 0.7).astype(np.uint8)
        counts = ex.glcm_counts(planes, torch.from_numpy(_planar(mask)).cuda()).cpu().numpy().view(np.uint32)
        for c, k in enumerate(names):
            want = c_oracle.glcm_counts(d[k], mask[0, :, :, c], n_angles=4)
            assert (counts[0, c] == want).all(), ("masked", h, w, k)


@pytest.mark.parametrize("shape", [(64, 64), (37, 91), (128, 128), (20, 9), (20, 8), (50, 16), (41, 32), (30, 256)])
def test_masked_all_blocks(shape, torch_mod):
    h, w = shape
    rng = np.random.default_rng(77 + h)
    d = parity_distributions(rng, h, w)
    names = sorted(d)
    img = np.stack([d[k] for k in names], axis=2)[None]
    yy, xx = np.mgrid[0:h, 0:w]
    masks = []
    for k in range(len(names)):
        m = ((yy - h / 2) ** 2 / (h / (2.2 + 0.3 * k)) ** 2 + (xx - w / 2) ** 2 / (w / 2.5) ** 2) < 1
        if k == 3:
            m = np.zeros_like(m)
        if k == 4:
            m = np.ones_like(m)
        if k == 5:
            m = rng.random((h, w)) < 0.05         # sparse, disconnected
        masks.append(m)
    mask = np.stack(masks, axis=2).astype(np.uint8)[None]
    want = c_oracle.table(_planar(img), _planar(mask), glcm=True, n_angles=4, shape=True, moments=True)
    got, status = imf.extract_features(img, mask, four_directions=True, shape=True, moments=True,
                                       return_status=True)
    cols = imf.feature_columns(len(names), n_angles=4, shape=True, moments=True)
    compare_tables(got, want, cols, label="masked " + str(shape), images=img, masks=mask)
    assert status[0] & 1              # the empty mask was flagged
    # mask values other than 0/1 count as inside
    got2 = imf.extract_features(img, mask * 200, four_directions=True, shape=True, moments=True)
    assert np.array_equal(got, got2, equal_nan=True)


def test_variable_sizes(torch_mod):
    rng = np.random.default_rng(5)
    objs, masks = [], []
    for i in range(24):
        h, w = int(rng.integers(6, 129)), int(rng.integers(6, 129))
        objs.append(rng.integers(0, 4096, (h, w, 5)).astype(np.uint16))
        masks.append((rng.random((h, w, 5)) < 0.5).astype(np.uint8))
    got = imf.extract_features(objs, four_directions=True, shape=True, moments=True)
    gotm = imf.extract_features(objs, masks, four_directions=True, shape=True, moments=True)
    cols = imf.feature_columns(5, n_angles=4, shape=True, moments=True)
    for i, (o, m) in enumerate(zip(objs, masks)):
        want = c_oracle.table(_planar(o[None]), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(got[i:i + 1], want, cols, label="var %d %s" % (i, o.shape), images=[o])
        want = c_oracle.table(_planar(o[None]), _planar(m[None]), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(gotm[i:i + 1], want, cols, label="var masked %d %s" % (i, o.shape), images=[o], masks=[m])


def test_odd_size_batch_uses_every_group(torch_mod):
    """More tiles than SMs with a slot size that is not a multiple of 16 pixels: every thread group of the
    GLCM kernel stages tiles, so the per-group prefetch buffers must stay 16-byte aligned."""
    rng = np.random.default_rng(17)
    n, h, w, c = 64, 37, 91, 5
    img = rng.integers(0, 3000, (n, h, w, c)).astype(np.uint16)
    mask = (rng.random((n, h, w, c)) < 0.7).astype(np.uint8)
    cols = imf.feature_columns(c, n_angles=4, shape=True, moments=True)
    got = imf.extract_features(img, mask, four_directions=True, shape=True, moments=True)
    want = c_oracle.table(_planar(img), _planar(mask), glcm=True, n_angles=4, shape=True, moments=True)
    compare_tables(got, want, cols, label="odd batch masked")
    got = imf.extract_features(img, four_directions=True, shape=True, moments=True)
    want = c_oracle.table(_planar(img), glcm=True, n_angles=4, shape=True, moments=True)
    compare_tables(got, want, cols, label="odd batch")


def test_synth_device_matches_numpy_mirror(torch_mod):
    from imfeat_b200 import synth
    ex = imf.get_extractor()
    planes, masks, _ = ex.synth(123, 40, 6, 3, 64, 64, with_masks=True)
    planes, masks = planes.cpu().numpy(), masks.cpu().numpy()
    for i in range(6):
        for ch in range(3):
            p, m = synth.synth_plane(123, 40 + i, ch, 64, 64)
            assert (planes[i, ch, :4096].reshape(64, 64) == p).all()
            assert (masks[i, ch, :4096].reshape(64, 64) == m).all()
    planes, masks, sizes = ex.synth(9, 0, 10, 2, 128, 128, with_masks=True, variable=True, hmin=16,
                                    wmin=16, mask_shrink=40)
    planes, masks, sizes = planes.cpu().numpy(), masks.cpu().numpy(), sizes.cpu().numpy()
    for i in range(10):
        h, w = synth.object_size(9, i, 128, 128, True, 16, 16)
        assert tuple(sizes[i]) == (h, w)
        for ch in range(2):
            p, m = synth.synth_plane(9, i, ch, h, w, mask_shrink=40)
            assert (planes[i, ch, :h * w].reshape(h, w) == p).all()
            assert (masks[i, ch, :h * w].reshape(h, w) == m).all()


def test_synthetic_batch_parity(torch_mod):
    """cfg1/cfg2 inputs (64x64x12 + masks), 256 objects, every block, against the C oracle."""
    from imfeat_b200 import synth
    img, mask = synth.synth_batch_hwc(0, 0, 256, 12, 64, 64)
    cols = imf.feature_columns(12, n_angles=4, shape=True, moments=True)
    want = c_oracle.table(_planar(img), _planar(mask), glcm=True, n_angles=4, shape=True, moments=True)
    got = imf.extract_features(img, mask, four_directions=True, shape=True, moments=True)
    compare_tables(got, want, cols, label="synthetic masked")
    want = c_oracle.table(_planar(img))
    got = imf.extract_features(img)
    compare_tables(got, want, imf.feature_columns(12), label="synthetic notebook mode")


def test_host_and_device_paths_agree(torch_mod):
    torch = torch_mod
    from imfeat_b200 import synth
    img, mask = synth.synth_batch_hwc(1, 0, 64, 12, 64, 64)
    ex = imf.get_extractor(four_directions=True)
    host = ex.extract_host_hwc(img, mask)
    planes, pmasks, hs, ws = ex.pack_hwc(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda())
    dev = ex.extract_planar(planes, pmasks, hs=hs, ws=ws).cpu().numpy()
    assert np.array_equal(host, dev, equal_nan=True)
    host_planar = ex.extract_host_planar(_planar(img), _planar(mask))
    assert np.array_equal(host, host_planar, equal_nan=True)
    # device-generated objects == host mirror objects
    dplanes, dmasks, _ = ex.synth(1, 0, 64, 12, 64, 64)
    dev2 = ex.extract_planar(dplanes, dmasks, hs=64, ws=64).cpu().numpy()
    assert np.array_equal(host, dev2, equal_nan=True)


def test_channel_selection_and_ablation_indirection(torch_mod):
    """LOCO == dropping that channel's columns; per-channel object permutation == permuting the
    rows of that channel's column blocks (features are per-channel independent, NB:239/291)."""
    torch = torch_mod
    from imfeat_b200 import schema
    ex = imf.get_extractor(four_directions=False)
    C, N = 6, 48
    planes, masks, _ = ex.synth(5, 0, N, C, 64, 64)
    base = ex.extract_planar(planes, masks, hs=64, ws=64).cpu().numpy()
    for drop in (0, 3, 5):
        keep = [c for c in range(C) if c != drop]
        chan = torch.tensor(keep, dtype=torch.int32, device="cuda")
        t = ex.extract_planar(planes, masks, hs=64, ws=64, chan=chan).cpu().numpy()
        idx = []
        for pos, c in enumerate(keep):
            idx.append((schema.channel_column_index(C - 1, pos), schema.channel_column_index(C, c)))
        for sub, full in idx:
            assert np.array_equal(t[:, sub], base[:, full], equal_nan=True)
    perm = np.random.default_rng(42).permutation(N).astype(np.int32)
    for ch in (1, 4):
        src = np.tile(np.arange(N, dtype=np.int32)[:, None], (1, C))
        src[:, ch] = perm
        t = ex.extract_planar(planes, masks, hs=64, ws=64,
                              src_obj=torch.from_numpy(src).cuda()).cpu().numpy()
        want = base.copy()
        cidx = schema.channel_column_index(C, ch)
        want[:, cidx] = base[perm][:, cidx]
        assert np.array_equal(t, want, equal_nan=True)
    # channels= on the functional API
    img, _ = __import__("imfeat_b200").synth.synth_batch_hwc(5, 0, 4, C, 64, 64)
    sel = imf.extract_features(img, channels=[4, 1])
    full = imf.extract_features(img)
    assert np.array_equal(sel[:, schema.channel_column_index(2, 0)], full[:, schema.channel_column_index(C, 4)])
    assert np.array_equal(sel[:, schema.channel_column_index(2, 1)], full[:, schema.channel_column_index(C, 1)])


def test_argument_errors(torch_mod):
    with pytest.raises(TypeError):
        imf.extract_features(np.zeros((1, 8, 8, 2), np.float32) + 0.5)
    with pytest.raises(imf.ImfeatError):
        imf.extract_features(np.zeros((1, 256, 256, 1), np.uint16))     # > IMFEAT_MAX_PIXELS
    ex = imf.get_extractor()
    import torch
    planes = torch.zeros((2, 3, 64, 64), dtype=torch.uint16, device="cuda")
    with pytest.raises(imf.ImfeatError):
        ex.extract_planar(planes, out=torch.empty((2, 10), dtype=torch.float64, device="cuda"))
    empty = ex.extract_planar(planes[:0])
    assert empty.shape == (0, 69)


def test_full_size_properties(torch_mod):
    """cfg2 at BASELINE size (10,000 objects): size-independent properties + sampled parity."""
    torch = torch_mod
    ex = imf.get_extractor(four_directions=True)
    N, C = 10000, 12
    planes, masks, _ = ex.synth(0, 0, N, C, 64, 64)
    t = ex.extract_planar(planes, None, hs=64, ws=64).cpu().numpy()
    cols = ex.columns(C)
    col = {c: i for i, c in enumerate(cols)}
    for ch in (1, 7, 12):
        g = lambda name: t[:, col["%s_Ch%d" % (name, ch)]]
        assert np.array_equal(g("mean_intensity"), g("total_intensity") / 4096.0)
        assert (g("min_intensity") <= g("percentile10_intensity")).all()
        assert (np.diff(np.stack([g("percentile%d0_intensity" % k) for k in range(1, 10)]), axis=0) >= 0).all()
        assert (g("percentile90_intensity") <= g("max_intensity")).all()
        assert ((g("shannon_entropy") > 0) & (g("shannon_entropy") <= 12.0)).all()
        for tag in ("", "_a45", "_a90", "_a135"):
            a = t[:, col["ASM%s_Ch%d" % (tag, ch)]]
            e = t[:, col["energy%s_Ch%d" % (tag, ch)]]
            np.testing.assert_allclose(e * e, a, rtol=1e-12)
            corr = t[:, col["correlation%s_Ch%d" % (tag, ch)]]
            assert (np.abs(corr) <= 1 + 1e-12).all()
    # idempotence / order independence: reversed object order gives the reversed table
    t2 = ex.extract_planar(planes.flip(0).contiguous(), None, hs=64, ws=64).cpu().numpy()
    assert np.array_equal(t2[::-1], t, equal_nan=True)
    # sampled parity on regenerated objects (host mirror of the device generator)
    from imfeat_b200 import synth
    pick = np.sort(np.random.default_rng(0).choice(N, 1024, replace=False))     # >= 1,024 objects
    img = np.stack([np.stack([synth.synth_plane(0, int(i), ch, 64, 64)[0] for ch in range(C)], axis=2)
                    for i in pick])
    want = c_oracle.table(_planar(img), glcm=True, n_angles=4)
    compare_tables(t[pick], want, cols, label="1024 sampled objects")
    # the same at the headline configuration: masks, four directions, shape, moments (every kernel's masked variant)
    exf = imf.get_extractor(four_directions=True, shape=True, moments=True)
    tm = exf.extract_planar(planes, masks, hs=64, ws=64).cpu().numpy()
    mk = np.stack([np.stack([synth.synth_plane(0, int(i), ch, 64, 64)[1] for ch in range(C)], axis=2) for i in pick])
    want = c_oracle.table(_planar(img), _planar(mk), glcm=True, n_angles=4, shape=True, moments=True)
    compare_tables(tm[pick], want, exf.columns(C), label="1024 sampled objects, masked, all blocks", images=img, masks=mk)


def test_cfg5_shaped_batch(torch_mod):
    """BASELINE.json configs[4]: variable-size objects up to 128x128, 18 channels, sparse masks (1-10% of the tile),
    fixed stride with a size table, every block, device-generated and checked against the C oracle object by object."""
    from imfeat_b200 import synth
    ex = imf.get_extractor(four_directions=True, shape=True, moments=True)
    N, C5 = 96, 18
    planes, masks, sizes = ex.synth(5, 0, N, C5, 128, 128, with_masks=True, variable=True, hmin=16, wmin=16, mask_shrink=32)
    got = ex.extract_planar(planes, masks, sizes, hs=128, ws=128).cpu().numpy()
    cols = ex.columns(C5)
    frac = []
    for i in range(N):
        h, w = synth.object_size(5, i, 128, 128, True, 16, 16)
        pl = [synth.synth_plane(5, i, ch, h, w, mask_shrink=32) for ch in range(C5)]
        img = np.stack([p[0] for p in pl], axis=2)[None]
        mk = np.stack([p[1] for p in pl], axis=2)[None]
        frac.append(mk.mean())
        want = c_oracle.table(_planar(img), _planar(mk), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(got[i:i + 1], want, cols, label="cfg5 object %d (%dx%d)" % (i, h, w), images=img, masks=mk)
    assert 0.01 < np.mean(frac) < 0.10


def test_custom_percentiles_and_distance(torch_mod):
    """Options away from the notebook literals: arbitrary np.percentile arguments (including 0, 50
    and 100) and another GLCM distance, against numpy / the restated greycomatrix directly."""
    rng = np.random.default_rng(21)
    d = parity_distributions(rng, 40, 56)
    names = sorted(d)
    img = np.stack([d[k] for k in names], axis=2)[None]
    qs = (0.0, 1.0, 12.5, 25.0, 50.0, 75.0, 90.0, 99.9, 100.0)
    got = imf.extract_features(img, glcm=True, percentiles=qs, distance=2, four_directions=True)
    cols = imf.feature_columns(len(names), n_angles=4)
    for c, k in enumerate(names):
        plane = d[k]
        for j, q in enumerate(qs):
            assert got[0, cols.index("percentile%d0_intensity_Ch%d" % (j + 1, c + 1))] == np.percentile(plane, q), (k, q)
        qz = orc.quantise(plane)
        P = orc.greycomatrix(qz, [2], list(orc.ANGLES4), levels=256)
        for a, tag in enumerate(orc.ANGLE_TAGS):
            for prop in orc.GLCM_PROPS:
                want = float(orc.greycoprops(P[:, :, :, a:a + 1], prop)[0, 0])
                g = got[0, cols.index("%s%s_Ch%d" % (prop, tag, c + 1))]
                ok = np.isclose(g, want, rtol=1e-9, atol=1e-9) or (prop == "correlation" and g == 1.0 and abs(want) < 1e-6)
                assert ok, (k, tag, prop, g, want)


@pytest.mark.parametrize("C", [1, 3, 18])
def test_channel_counts_and_status_bits(C, torch_mod):
    from imfeat_b200 import synth
    img, mask = synth.synth_batch_hwc(3, 0, 16, C, 64, 64)
    mask[5] = 0                                   # object 5: every mask empty
    img[7, :, :, 0] = 777                         # object 7, channel 0: constant
    want = c_oracle.table(_planar(img), _planar(mask), glcm=True, n_angles=1)
    got, status = imf.extract_features(img, mask, return_status=True)
    compare_tables(got, want, imf.feature_columns(C), label="C=%d" % C)
    assert status[5] & 1 and status[5] & 2        # empty mask, zero GLCM pairs
    assert status[7] & 4                          # constant channel
    assert not (status[0] & 1)


def test_host_path_multiple_slabs(torch_mod):
    """The host entry point pipelines ~64 MiB slabs; 1,500 objects = 3 slabs.  Pageable and pinned
    inputs, with and without masks, must give the table of the device path bit for bit."""
    torch = torch_mod
    ex = imf.get_extractor()
    n = 1500
    planes, masks, _ = ex.synth(11, 0, n, 12, 64, 64)
    dev = ex.extract_planar(planes, None, hs=64, ws=64).cpu().numpy()
    devm = ex.extract_planar(planes, masks, hs=64, ws=64).cpu().numpy()
    hwc = planes[:, :, :4096].reshape(n, 12, 64, 64).permute(0, 2, 3, 1).contiguous().cpu()
    mhwc = masks[:, :, :4096].reshape(n, 12, 64, 64).permute(0, 2, 3, 1).contiguous().cpu()
    assert np.array_equal(ex.extract_host_hwc(hwc.numpy()), dev, equal_nan=True)
    assert np.array_equal(ex.extract_host_hwc(hwc.numpy(), mhwc.numpy()), devm, equal_nan=True)
    pin, mpin = hwc.pin_memory(), mhwc.pin_memory()
    assert np.array_equal(ex.extract_host_hwc(pin.numpy(), mpin.numpy()), devm, equal_nan=True)


def test_ablation_sweep_driver(torch_mod):
    """channel_ablation_sweep (device-side loop) == column-block operations on the base table."""
    from imfeat_b200 import ablation, schema
    ex = imf.get_extractor()
    C, N = 5, 40
    planes, masks, _ = ex.synth(8, 0, N, C, 64, 64)
    base = ex.extract_planar(planes, masks, hs=64, ws=64).cpu().numpy()
    loco = ablation.channel_ablation_sweep(ex, planes, masks, hs=64, ws=64, mode="loco").cpu().numpy()
    assert loco.shape == (C, N, ex.row_width(C - 1))
    for k in range(C):
        keep = [c for c in range(C) if c != k]
        for pos, c in enumerate(keep):
            assert np.array_equal(loco[k][:, schema.channel_column_index(C - 1, pos)],
                                  base[:, schema.channel_column_index(C, c)], equal_nan=True)
    perm = ablation.channel_ablation_sweep(ex, planes, masks, hs=64, ws=64, mode="permute", seed=42).cpu().numpy()
    src = ablation.permutation_sources(N, C, seed=42)
    for k in range(C):
        want = base.copy()
        cidx = schema.channel_column_index(C, k)
        want[:, cidx] = base[src[k][:, k]][:, cidx]
        assert np.array_equal(perm[k], want, equal_nan=True)


def test_ablation_sweep_cuda_graph(torch_mod):
    """The sweep captured into a CUDA graph: replays follow the current contents of the input buffers."""
    torch = torch_mod
    from imfeat_b200 import ablation
    ex = imf.get_extractor(four_directions=True, shape=True, moments=True)
    C, N = 4, 200
    planes, masks, _ = ex.synth(12, 0, N, C, 64, 64)
    for mode in ("loco", "permute"):
        cap = ablation.CapturedSweep(ex, planes, masks, hs=64, ws=64, mode=mode)
        want = ablation.channel_ablation_sweep(ex, planes, masks, hs=64, ws=64, mode=mode)
        assert torch.equal(cap.replay().view(torch.int64), want.view(torch.int64))
        # new objects written into the same buffers, one launch
        p2, m2, _ = ex.synth(13, 0, N, C, 64, 64)
        keep_p, keep_m = planes.clone(), masks.clone()
        planes.copy_(p2); masks.copy_(m2)
        want2 = ablation.channel_ablation_sweep(ex, planes, masks, hs=64, ws=64, mode=mode)
        got2 = cap.replay()
        torch.cuda.synchronize()
        assert torch.equal(got2.view(torch.int64), want2.view(torch.int64))
        assert not torch.equal(want2.view(torch.int64), want.view(torch.int64))
        planes.copy_(keep_p); masks.copy_(keep_m)


def test_repeatability(torch_mod):
    """Same inputs, different launches / streams: identical bits (integer accumulation everywhere
    order could matter)."""
    torch = torch_mod
    ex = imf.get_extractor(four_directions=True, shape=True, moments=True)
    planes, masks, _ = ex.synth(4, 0, 300, 12, 64, 64)
    a = ex.extract_planar(planes, masks, hs=64, ws=64).clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        b = ex.extract_planar(planes, masks, hs=64, ws=64, stream=s).clone()
    s.synchronize()
    c = ex.extract_planar(planes, masks, hs=64, ws=64)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int64), b.view(torch.int64))
    assert torch.equal(a.view(torch.int64), c.view(torch.int64))


def test_bit_packed_host_masks(torch_mod):
    """Host entry points with bit-packed masks (opts.host_mask_bits): same bits as with byte masks, for the
    interleaved (h, w, c) layout with a size table and for the planar layout, pinned and pageable buffers."""
    torch = torch_mod
    rng = np.random.default_rng(9)
    n, h, w, c = 37, 37, 53, 3                      # 37*53*3 and the plane stride are no multiples of 64
    img = rng.integers(0, 4096, (n, h, w, c)).astype(np.uint16)
    mask = (rng.random((n, h, w, c)) < 0.55).astype(np.uint8)
    sizes = np.stack([rng.integers(6, h + 1, n), rng.integers(6, w + 1, n)], axis=1).astype(np.int32)
    ex = imf.get_extractor(four_directions=True, shape=True, moments=True)
    want = ex.extract_host_hwc(img, mask, sizes)
    got = ex.extract_host_hwc(img, imf.pack_mask_bits(mask), sizes, masks_packed=True)
    assert np.array_equal(got, want, equal_nan=True)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    out = torch.empty(want.shape, dtype=torch.float64).pin_memory().numpy()
    ex.extract_host_hwc(pin(img), pin(imf.pack_mask_bits(mask)), sizes, out=out, masks_packed=True)   # direct H2D / D2H
    assert np.array_equal(out, want, equal_nan=True)
    # planar layout: every plane packed on its own (plane stride padded to a multiple of 8 elements)
    stride = imf.plane_stride_for(h, w)
    pl = np.zeros((n, c, stride), np.uint16)
    pm = np.zeros((n, c, stride), np.uint8)
    pl[:, :, :h * w] = _planar(img).reshape(n, c, h * w)
    pm[:, :, :h * w] = _planar(mask).reshape(n, c, h * w)
    want = ex.extract_host_planar(pl, pm, hs=h, ws=w)
    got = ex.extract_host_planar(pl, imf.pack_mask_bits(pm.reshape(n * c, stride)), hs=h, ws=w, masks_packed=True)
    assert np.array_equal(got, want, equal_nan=True)


def test_torch_ops_direct(torch_mod):
    """torch.ops.imfeat.extract / glcm_counts called directly (context kept by the extension, current stream from
    torch) give the same bits as the FeatureExtractor route, also under CUDA-graph capture."""
    torch = torch_mod
    from imfeat_b200 import _lib
    ops = _lib.load_torch_ops()
    ex = imf.get_extractor(four_directions=True, shape=True, moments=True)
    planes, masks, _ = ex.synth(8, 0, 64, 5, 64, 64)
    want = ex.extract_planar(planes, masks, hs=64, ws=64)
    p16 = planes.view(torch.int16)
    got = ops.extract(p16, masks, None, None, None, 64, 64, True, True, 4, 5, True, True, [], None, None)
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    chan = torch.tensor([4, 0, 2], dtype=torch.int32, device=planes.device)
    sub = ops.extract(p16, masks, None, None, chan, 64, 64, True, False, 1, 5, False, False, [], None, None)
    exb = imf.get_extractor(glcm=False)
    assert torch.equal(sub.view(torch.int64), exb.extract_planar(planes, masks, hs=64, ws=64, chan=chan).view(torch.int64))
    c1 = ops.glcm_counts(p16, masks, None, 64, 64, 4, 5)
    assert torch.equal(c1, imf.get_extractor(four_directions=True).glcm_counts(planes, masks, hs=64, ws=64))
    out = torch.empty_like(want)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.extract(p16, masks, None, None, None, 64, 64, True, True, 4, 5, True, True, [], out, None)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), want.view(torch.int64))


def test_repeatability_all_paths(torch_mod):
    """Stand-in for compute-sanitizer's racecheck (closed on this pool, profiles/r2_sanitizer_note.txt): a batch that
    takes every kernel path -- histogram window hit and miss, full-range ring, FP64 moments, 4-bit and 8-bit GLCM
    tables with wrapped counters and their fallbacks, the unmasked ring kernel -- repeated 25 times must give the
    same bits every time, tables and raw GLCM bins alike.  A shared-memory race shows up as a run-to-run difference."""
    torch = torch_mod
    rng = np.random.default_rng(0)
    n, h, w, c = 96, 64, 64, 4
    img = rng.integers(100, 3000, (n, h, w, c)).astype(np.uint16)
    img[:, :, :, 1] = rng.integers(0, 65536, (n, h, w))
    img[:, :, :, 2] = rng.choice([7, 7, 7, 900], (n, h, w))
    mask = (rng.random((n, h, w, c)) < 0.6).astype(np.uint8)
    mask[:, :, :, 2] = 1
    planes = torch.from_numpy(_planar(img)).cuda()
    masks = torch.from_numpy(_planar(mask)).cuda()
    exf = imf.get_extractor(four_directions=True, shape=True, moments=True)
    exn = imf.get_extractor()
    exg = imf.get_extractor(four_directions=True)
    ref = None
    for rep in range(25):
        cur = (exf.extract_planar(planes, masks).view(torch.int64).clone(),
               exn.extract_planar(planes, None).view(torch.int64).clone(),
               exg.glcm_counts(planes[:8], masks[:8]).clone(), exg.glcm_counts(planes[:8]).clone())
        if ref is None:
            ref = cur
        else:
            for k, (a, b) in enumerate(zip(ref, cur)):
                assert torch.equal(a, b), "repetition %d differs in result %d" % (rep, k)
    want = c_oracle.table(_planar(img), _planar(mask), glcm=True, n_angles=4, shape=True, moments=True)
    compare_tables(ref[0].view(torch.float64).cpu().numpy(), want, exf.columns(c), label="all paths", images=img, masks=mask)


def test_pinned_batcher_end_to_end(torch_mod):
    """Variable-size objects through the pinned batcher == per-object oracle rows."""
    rng = np.random.default_rng(17)
    ex = imf.get_extractor()
    objs = []
    for k in range(37):
        h, w = int(rng.integers(8, 97)), int(rng.integers(8, 81))
        img = rng.integers(0, 4096, (h, w, 3)).astype(np.uint16)
        m = rng.random((h, w, 3)) < 0.6
        objs.append((img, m))
    cols = imf.feature_columns(3)
    tables = []
    # synchronous with byte masks; double-buffered in the background with byte masks and with bit-packed masks
    for kw in (dict(asynchronous=False), dict(), dict(packed_masks=True)):
        b = imf.PinnedBatcher(ex, capacity=8, hs=96, ws=80, channels=3, with_masks=True, **kw)
        for k, (img, m) in enumerate(objs):
            b.add(img, m, label=k)
        table, labels = b.finish()
        assert labels == list(range(37)) and table.shape == (37, 69)
        tables.append(table)
    assert np.array_equal(tables[0], tables[1], equal_nan=True) and np.array_equal(tables[0], tables[2], equal_nan=True)
    for k, (img, m) in enumerate(objs):
        want = c_oracle.table(_planar(img[None]), _planar(m[None].astype(np.uint8)))
        compare_tables(tables[0][k:k + 1], want, cols, label="batcher %d" % k, images=[img], masks=[m])


def test_wide_range_tiles_through_host_pipeline(torch_mod):
    """Full 16-bit data (value range >= 4096) takes the worklist path K2c -> K2; 1,400 objects go
    through the two-stream host pipeline, so two worklists are in flight at once."""
    rng = np.random.default_rng(33)
    n = 1400
    img = rng.integers(0, 65536, (n, 64, 64, 12)).astype(np.uint16)
    img[::3] = rng.integers(0, 4096, (len(img[::3]), 64, 64, 12)).astype(np.uint16)   # mixed: some tiles stay compact
    ex = imf.get_extractor(glcm=False)
    got = ex.extract_host_hwc(img)
    pick = rng.choice(n, 40, replace=False)
    want = c_oracle.table(_planar(img[pick]), glcm=False)
    compare_tables(got[pick], want, imf.feature_columns(12, glcm=False), label="wide range")
    import torch
    planes = torch.from_numpy(_planar(img)).cuda()
    dev = ex.extract_planar(planes).cpu().numpy()
    assert np.array_equal(got, dev, equal_nan=True)


def test_minmax_scaler_matches_sklearn(torch_mod):
    """NB:389-394: fit on the training rows, transform training and test rows; against sklearn itself,
    bit for bit (same two roundings), NaN cells and constant / all-NaN columns included."""
    torch = torch_mod
    from sklearn.preprocessing import MinMaxScaler as SkScaler
    import warnings
    rng = np.random.default_rng(21)
    ex = imf.get_extractor(four_directions=True, shape=True, moments=True)
    planes, masks, _ = ex.synth(5, 0, 300, 6, 64, 64, with_masks=True)
    table = ex.extract_planar(planes, masks, hs=64, ws=64)          # stays on the device
    # make the table nasty: a constant column, an all-NaN column, scattered NaN
    table[:, 3] = 7.25
    table[:, 5] = float("nan")
    table[rng.integers(0, 300, 40), rng.integers(0, table.shape[1], 40)] = float("nan")
    X = table.cpu().numpy()
    tr, te = X[:225], X[225:]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sk = SkScaler().fit(tr)
        want_tr, want_te = sk.transform(tr), sk.transform(te)
    sc = imf.MinMaxScaler().fit(table[:225])
    got_tr = sc.transform(table[:225]).cpu().numpy()
    got_te = sc.transform(table[225:]).cpu().numpy()
    for name in ("data_min_", "data_max_", "scale_", "min_"):
        g, w = getattr(sc, name).cpu().numpy(), getattr(sk, name)
        assert np.array_equal(g, w, equal_nan=True), name
    assert np.array_equal(got_tr, want_tr, equal_nan=True)
    assert np.array_equal(got_te, want_te, equal_nan=True)
    # numpy front door
    assert np.array_equal(imf.MinMaxScaler().fit_transform(tr), want_tr, equal_nan=True)


def test_k1_wide_range_falls_back_to_fp64(torch_mod):
    """K1 sums (x - pivot)^k as exact integers while |x - pivot| <= 11,585 and repeats the tile with FP64
    sums otherwise: tiles on both sides of the limit, and just around it, against the C oracle."""
    rng = np.random.default_rng(31)
    h, w = 64, 64
    planes = []
    for spread in (100, 11000, 11580, 11590, 12000, 30000, 65535):
        base = rng.integers(0, 65536 - spread) if spread < 65535 else 0
        x = base + rng.integers(0, spread + 1, (h, w))
        x[0, 0], x[-1, -1] = base, base + spread            # the extremes are really there
        planes.append(x.astype(np.uint16))
    img = np.stack(planes, axis=2)[None]
    mask = (rng.random(img.shape) < 0.8).astype(np.uint8)
    cols = imf.feature_columns(img.shape[3])
    compare_tables(imf.extract_features(img), c_oracle.table(_planar(img)), cols, label="k1 range")
    compare_tables(imf.extract_features(img, mask), c_oracle.table(_planar(img), _planar(mask)), cols,
                   label="k1 range masked")


def test_k1_fourth_power_sum_beyond_64_bits(torch_mod):
    """Every |x - pivot| within the integer limit, but the tile's sum of fourth powers exceeds 2^64: large
    mid-range tiles (128x128 uniform over ~20,000) and a bimodal 64x64 tile (0 / 18,000).  The lane sums
    fit 64 bits, the warp total does not."""
    rng = np.random.default_rng(32)
    big = [rng.integers(0, 20001, (128, 128)), 30000 + rng.integers(0, 22001, (128, 128)),
           np.where(rng.random((128, 128)) < 0.5, 0, 23000)]
    img = np.stack([b.astype(np.uint16) for b in big], axis=2)[None]
    cols = imf.feature_columns(img.shape[3])
    compare_tables(imf.extract_features(img), c_oracle.table(_planar(img)), cols, label="k1 s4 128")
    mask = (rng.random(img.shape) < 0.9).astype(np.uint8)
    compare_tables(imf.extract_features(img, mask), c_oracle.table(_planar(img), _planar(mask)), cols,
                   label="k1 s4 128 masked")
    small = [np.where(rng.random((64, 64)) < 0.5, 0, 18000), np.where(rng.random((64, 64)) < 0.5, 100, 23100)]
    img = np.stack([b.astype(np.uint16) for b in small], axis=2)[None]
    cols = imf.feature_columns(img.shape[3])
    compare_tables(imf.extract_features(img), c_oracle.table(_planar(img)), cols, label="k1 s4 bimodal")


def test_round2_paths_edge_cases(torch_mod):
    """The code paths added late in round 2, on inputs chosen to hit their corners:
    * K4w's general vector pass: row lengths 8..15 and other non-multiples of 8 (chunks that straddle row ends, a
      partial last chunk), next to row lengths below 8 (scalar pass) in the same size-table batch;
    * K3's bounding-box records: masks confined to the last rows / first rows / one row / one column / one pixel,
      a full mask and an empty one;
    * K2's two-level percentile search: full-16-bit data with duplicates, masked, with percentiles from 0 to 100."""
    rng = np.random.default_rng(77)
    shapes = [(9, 8), (11, 9), (16, 15), (5, 13), (64, 10), (7, 5), (3, 3), (33, 71), (128, 12), (12, 128), (31, 100), (64, 64)]
    C = 3
    objs, masks = [], []
    for k, (h, w) in enumerate(shapes):
        o = rng.integers(0, 65536, (h, w, C)).astype(np.uint16)
        o[:, :, 1] = rng.integers(0, 7, (h, w)) * 9000                  # wide range, heavy duplicates
        o[:, :, 2] = rng.integers(100, 3000, (h, w))                    # 12-bit plane
        m = np.zeros((h, w, C), np.uint8)
        kind = k % 6
        if kind == 0: m[h - 2:, :, :] = 1                               # last rows
        elif kind == 1: m[:1, :, :] = 1                                 # one row
        elif kind == 2: m[:, w - 1:, :] = 1                             # one column (no horizontal pair)
        elif kind == 3: m[h // 2, w // 2, :] = 1                        # one pixel
        elif kind == 4: m[:, :, :] = 1                                  # full
        else: m[:, :, 0] = (rng.random((h, w)) < 0.5); m[:, :, 1] = 0; m[:, :, 2] = (rng.random((h, w)) < 0.1)   # channel 1: empty
        objs.append(o)
        masks.append(m)
    qs = (0.0, 0.5, 5.0, 25.0, 50.0, 75.0, 95.0, 99.5, 100.0)
    cols = imf.feature_columns(C, n_angles=4, shape=True, moments=True)
    for use_mask in (True, False):
        got = imf.extract_features(objs, masks if use_mask else None, four_directions=True, shape=True, moments=True,
                                   percentiles=qs)
        for i, (o, m) in enumerate(zip(objs, masks)):
            want = c_oracle.table(_planar(o[None]), _planar(m[None]) if use_mask else None, glcm=True, n_angles=4,
                                  shape=True, moments=True)
            # the oracle's percentile columns are the notebook's literals: take np.percentile for the custom ones
            for c in range(C):
                sel = o[:, :, c][m[:, :, c] > 0] if use_mask else o[:, :, c].ravel()
                for j, q in enumerate(qs):
                    col = cols.index("percentile%d0_intensity_Ch%d" % (j + 1, c + 1))
                    want[0, col] = np.percentile(sel, q) if sel.size else np.nan
            compare_tables(got[i:i + 1], want, cols, label="r2 edge %d %s mask=%s" % (i, o.shape[:2], use_mask),
                           images=[o], masks=[m] if use_mask else None)


def test_k3_tiers_over_several_chunks_and_large_full_range_tiles(torch_mod):
    """Two capacity tiers of K3 (strides above 8,192 pixels with masks) when the batch is worked off in several
    front / bins rounds (IMFEAT_K3_CHUNK, read when the context is created): sparse masks stay in the first tier,
    full and half-plane masks are listed for the second, in every round.  The same batch has full-16-bit planes of
    128x128 pixels, more than one pass of K2's thread groups over a tile."""
    import os
    rng = np.random.default_rng(99)
    C, n = 2, 23
    objs, masks = [], []
    for k in range(n):
        h, w = (128, 128) if k % 5 == 0 else (int(rng.integers(40, 129)), int(rng.integers(40, 129)))
        o = np.empty((h, w, C), np.uint16)
        o[:, :, 0] = rng.integers(0, 65536, (h, w))                      # full range: K12's window fails, K2 takes it
        o[:, :, 1] = rng.integers(200, 2500, (h, w))
        m = np.zeros((h, w, C), np.uint8)
        if k % 3 == 0:
            m[:] = 1                                                    # whole plane: second tier
        elif k % 3 == 1:
            r0, c0 = int(rng.integers(0, h - 12)), int(rng.integers(0, w - 12))
            m[r0:r0 + 12, c0:c0 + 9, :] = 1                             # sparse: first tier
        else:
            m[h // 3:, :, 0] = 1                                        # two thirds of the rows: second tier
            m[:, :, 1] = rng.random((h, w)) < 0.02                      # scattered pixels over all rows: second tier
        objs.append(o)
        masks.append(m)
    old = os.environ.get("IMFEAT_K3_CHUNK")
    os.environ["IMFEAT_K3_CHUNK"] = "5"                                  # 5 objects per round: 5 rounds
    try:
        ex = imf.FeatureExtractor(glcm=True, four_directions=True, shape=True, moments=True)
    finally:
        if old is None:
            os.environ.pop("IMFEAT_K3_CHUNK", None)
        else:
            os.environ["IMFEAT_K3_CHUNK"] = old
    # fixed-stride (h,w,c) slab + size table through this extractor's host entry point
    hs, ws = max(o.shape[0] for o in objs), max(o.shape[1] for o in objs)
    img, msk = np.zeros((n, hs, ws, C), np.uint16), np.zeros((n, hs, ws, C), np.uint8)
    sizes = np.zeros((n, 2), np.int32)
    for i, (o, m) in enumerate(zip(objs, masks)):
        img[i, :o.shape[0], :o.shape[1]] = o
        msk[i, :o.shape[0], :o.shape[1]] = m
        sizes[i] = o.shape[:2]
    got = ex.extract_host_hwc(img, msk, sizes=sizes)
    cols = imf.feature_columns(C, n_angles=4, shape=True, moments=True)
    for i, (o, m) in enumerate(zip(objs, masks)):
        want = c_oracle.table(_planar(o[None]), _planar(m[None]), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(got[i:i + 1], want, cols, label="tiers %d %s" % (i, o.shape[:2]), images=[o], masks=[m])


def test_maximum_plane_size(torch_mod):
    """Planes at the size limit of the C ABI (IMFEAT_MAX_PIXELS = 32,768 pixels: 128x256 and 181x181), with a
    sparse mask (first K3 tier), a full mask (second tier) and without masks; K4 leaves its warp-per-tile path
    above 16,384 pixels."""
    rng = np.random.default_rng(123)
    cols = imf.feature_columns(1, n_angles=4, shape=True, moments=True)
    for (h, w) in ((128, 256), (181, 181)):
        o = np.empty((2, h, w, 1), np.uint16)
        o[0, :, :, 0] = rng.integers(50, 4000, (h, w))
        o[1, :, :, 0] = rng.integers(0, 65536, (h, w))
        m = np.zeros((2, h, w, 1), np.uint8)
        m[0, 40:70, 90:140, 0] = 1
        m[1] = 1
        got = imf.extract_features(o, m, four_directions=True, shape=True, moments=True)
        want = c_oracle.table(_planar(o), _planar(m), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(got, want, cols, label="max size masked %dx%d" % (h, w), images=list(o), masks=list(m))
        got = imf.extract_features(o, four_directions=True, shape=True, moments=True)
        want = c_oracle.table(_planar(o), glcm=True, n_angles=4, shape=True, moments=True)
        compare_tables(got, want, cols, label="max size %dx%d" % (h, w), images=list(o))
