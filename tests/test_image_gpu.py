"""GPU parity for the steps either side of the block path, through the C ABI: ingest == basis_compressor::extract_source_blocks,
decode == basist::unpack_uastc, metrics == image_metrics::calc; plus the raster -> blocks -> UASTC -> texels -> PSNR chain
at 4096^2 checked through size-independent properties."""
import ctypes

import numpy as np
import pytest

import util
from util import _ptr
from basis_universal_b200 import image, uastc
from test_image_cpu import np_histograms, ref_extract, ref_metrics

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    o = image.ImageOps(0)
    yield o
    o.close()


@pytest.fixture(scope="module")
def ref():
    return util.Ref()


@pytest.mark.parametrize("shape", [(8, 8), (5, 7), (1, 1), (13, 4), (64, 37), (255, 257), (512, 768)])
def test_extract_matches_reference(ops, ref, shape):
    h, w = shape
    img = np.ascontiguousarray(util.synth(max(h, w) + 3, 3)[:h, :w])
    assert np.array_equal(ops.extract_source_blocks(img), ref_extract(ref, img))


def test_extract_strided_rows(ops, ref):
    big = util.synth(128, 9)
    view = big[5:5 + 50, 8:8 + 41]          # pitch != width * 4, base not 16-byte aligned
    assert np.array_equal(ops.extract_source_blocks(view), ref_extract(ref, np.ascontiguousarray(view)))


def test_extract_bottom_right_crop_and_flipped_rows(ops, ref):
    """A sub-rectangle that ends at the last byte of the parent allocation: the last row only has width * 4 valid bytes, so the
    host-to-device copy must not read pitch * height (that would run x0 * 4 bytes past the allocation). Rows with a negative
    stride (a vertically flipped view) are copied by the wrapper, not passed through."""
    big = util.synth(64, 3)
    view = big[64 - 23:, 64 - 30:]           # bottom-right corner, 23 x 30
    assert np.array_equal(ops.extract_source_blocks(view), ref_extract(ref, np.ascontiguousarray(view)))
    flipped = big[::-1, 10:50]
    assert np.array_equal(ops.extract_source_blocks(flipped), ref_extract(ref, np.ascontiguousarray(flipped)))
    from basis_universal_b200 import uastc
    enc = uastc.Encoder(0)
    want = enc.encode_uastc(ref_extract(ref, np.ascontiguousarray(view)), 1)
    assert np.array_equal(enc.encode_image(view, 1), want)
    enc.close()


def test_extract_empty(ops):
    assert ops.extract_source_blocks(np.zeros((0, 0, 4), np.uint8)).shape == (0, 64)


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_unpack_matches_reference_transcoder(ops, ref, level):
    blocks = np.concatenate([util.edge_case_blocks(5), util.image_to_blocks(util.synth(128, 40 + level))])
    u = ref.encode_uastc(blocks, level)
    assert np.array_equal(ops.unpack_uastc(u), ref.unpack_uastc(u))


def test_unpack_rejects_invalid_block(ops, ref):
    """Blocks the reference's unpack_uastc refuses (reserved mode code / out-of-range pattern) must fail the call."""
    import ctypes
    rng = np.random.default_rng(4)
    cand = rng.integers(0, 256, (4096, 16), dtype=np.uint8)
    out = np.zeros((1, 64), np.uint8)
    bad = [b for b in cand if not ref.lib.ref_unpack_uastc_blocks(util._ptr(np.ascontiguousarray(b)), ctypes.c_uint32(1), util._ptr(out))]
    assert bad, "no invalid block among 4096 random ones?"
    good = ref.encode_uastc(util.image_to_blocks(util.synth(16, 1)), 2)
    with pytest.raises(Exception):
        ops.unpack_uastc(np.concatenate([good, np.stack(bad[:3])]))
    assert ops.unpack_uastc(good).shape == (16, 64)      # the context stays usable


@pytest.mark.parametrize("shape", [(96, 96), (50, 41), (4, 4), (1, 3)])
def test_metrics_match_image_metrics_calc(ops, ref, shape):
    import torch
    h, w = shape
    a = np.ascontiguousarray(util.synth(128, 11)[:h, :w])
    b = np.clip(a.astype(np.int32) + np.random.default_rng(2).integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    da = torch.from_numpy(ops.extract_source_blocks(a)).cuda()
    db = torch.from_numpy(ops.extract_source_blocks(b)).cuda()
    torch.cuda.synchronize()
    hist, sum_a, sum_b = ops.block_metrics_device(da.data_ptr(), db.data_ptr(), w, h)
    assert np.array_equal(hist, np_histograms(a, b))
    assert [int(v) for v in sum_a] == [int(a[..., c].sum()) for c in range(4)] and [int(v) for v in sum_b] == [int(b[..., c].sum()) for c in range(4)]
    for first, total, use601 in [(0, 3, False), (0, 4, False), (3, 1, False), (0, 0, False), (0, 0, True)]:
        assert image.metrics_from_histograms(hist, w, h, first, total, True, use601) == ref_metrics(ref, a, b, first, total, True, use601)


def test_full_chain_4096(ops, ref):
    """raster -> blocks -> UASTC -> texels -> PSNR, all on the device side of the ABI at BASELINE's 4096^2 size:
    ingest equals the numpy tiling, every texel of the decode equals the reference decode on a sample, and the PSNR figures
    equal image_metrics::calc run by the reference on the full 16.7 Mtexel pair."""
    import torch
    img = util.synth(4096, 1234)
    blocks = ops.extract_source_blocks(img)
    assert np.array_equal(blocks, util.image_to_blocks(img))
    enc = uastc.Encoder(0)
    u = enc.encode_uastc(blocks, 2)
    enc.close()
    dec = ops.unpack_uastc(u)
    idx = np.random.default_rng(1).choice(u.shape[0], 4096, replace=False)
    assert np.array_equal(dec[idx], ref.unpack_uastc(u[idx]))
    da, db = torch.from_numpy(blocks).cuda(), torch.from_numpy(dec).cuda()
    torch.cuda.synchronize()
    hist, _, _ = ops.block_metrics_device(da.data_ptr(), db.data_ptr(), 4096, 4096)
    assert int(hist[0].sum()) == 4096 * 4096
    dec_img = dec.reshape(1024, 1024, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(4096, 4096, 4)
    for first, total in [(0, 3), (0, 4), (0, 0)]:
        assert image.metrics_from_histograms(hist, 4096, 4096, first, total) == ref_metrics(ref, img, np.ascontiguousarray(dec_img), first, total)
    assert 25.0 < image.metrics_from_histograms(hist, 4096, 4096, 0, 3)["psnr"] < 60.0


def test_unpack_random_valid_bit_patterns(ops, ref):
    """Arbitrary 128-bit patterns that the reference accepts decode to the same texels (anchor handling, BISE corner cases)."""
    import ctypes
    cand = np.random.default_rng(9).integers(0, 256, (8192, 16), dtype=np.uint8)
    out = np.zeros((1, 64), np.uint8)
    valid = np.array([bool(ref.lib.ref_unpack_uastc_blocks(util._ptr(np.ascontiguousarray(b)), ctypes.c_uint32(1), util._ptr(out))) for b in cand])
    u = np.ascontiguousarray(cand[valid])
    assert u.shape[0] > 1000
    assert np.array_equal(ops.unpack_uastc(u), ref.unpack_uastc(u))


@pytest.mark.parametrize("shape", [(64, 64), (50, 41), (1, 1), (4, 260)])
def test_encode_image_equals_extract_then_encode(ops, ref, shape):
    """Raster in, UASTC out in one call == the reference's extract_source_blocks followed by encode_uastc per block."""
    h, w = shape
    img = np.ascontiguousarray(util.synth(max(h, w) + 8, 17)[:h, :w])
    enc = uastc.Encoder(0)
    got = enc.encode_image(img, 2)
    assert enc.last_launch_count >= 4
    enc.close()
    assert np.array_equal(got, ref.encode_uastc(ref_extract(ref, img), 2))


def test_unpack_etc1_matches_reference(ops, ref):
    """Arbitrary ETC1 bit patterns (individual / differential, flipped or not, overflowing deltas) and real ETC1S blocks."""
    import ctypes
    from basis_universal_b200 import etc1s
    rnd = np.random.default_rng(12).integers(0, 256, (20000, 8), dtype=np.uint8)
    c = etc1s.Etc1sContext(0)
    c.set_pixel_blocks(util.image_to_blocks(util.synth(256, 8)))
    real = c.encode_etc1s_blocks(True, 16)
    c.close()
    for e in (rnd, real):
        want = np.zeros((e.shape[0], 64), np.uint8)
        ref.lib.ref_unpack_etc1_blocks(util._ptr(np.ascontiguousarray(e)), ctypes.c_uint32(e.shape[0]), util._ptr(want))
        assert np.array_equal(ops.unpack_etc1(e, strict=False), want)
    assert ops.unpack_etc1(real).shape == (real.shape[0], 64)           # ETC1S blocks never overflow
    d = rnd[:, :3].astype(np.int32)
    delta = np.where((d & 7) >= 4, (d & 7) - 8, d & 7)
    overflow = ((rnd[:, 3] & 2) != 0) & ((((d >> 3) + delta) < 0) | (((d >> 3) + delta) > 31)).any(1)
    assert overflow.any()
    with pytest.raises(Exception):
        ops.unpack_etc1(rnd)
    assert ops.unpack_etc1(rnd[~overflow]).shape[0] == int((~overflow).sum())


def _ref_clists(ref, sw, sh, dw, dh, filt, scale, wrap):
    xo = np.zeros(dw + 1, np.uint32); yo = np.zeros(dh + 1, np.uint32)
    assert ref.lib.ref_resampler_clists(sw, sh, dw, dh, filt, ctypes.c_float(scale), wrap, _ptr(xo), None, None, _ptr(yo), None, None)
    xw = np.zeros(int(xo[-1]), np.float32); xp = np.zeros(int(xo[-1]), np.uint32); yw = np.zeros(int(yo[-1]), np.float32); yp = np.zeros(int(yo[-1]), np.uint32)
    assert ref.lib.ref_resampler_clists(sw, sh, dw, dh, filt, ctypes.c_float(scale), wrap, _ptr(xo), _ptr(xw), _ptr(xp), _ptr(yo), _ptr(yw), _ptr(yp))
    return (xo, xw, xp), (yo, yw, yp)


@pytest.mark.parametrize("case", [
    dict(src=(256, 192), dst=(128, 96), filt=b"kaiser", srgb=1, comps=4, wrap=0),     # the compressor's default mip step, sRGB colour + linear alpha
    dict(src=(256, 192), dst=(32, 24), filt=b"kaiser", srgb=0, comps=3, wrap=0),      # level 3 straight from level 0; alpha left alone
    dict(src=(255, 127), dst=(127, 63), filt=b"box", srgb=0, comps=4, wrap=1),        # odd sizes, wrap addressing
    dict(src=(64, 300), dst=(32, 150), filt=b"lanczos4", srgb=1, comps=4, wrap=0),    # tall image: the other axis order
    dict(src=(300, 64), dst=(150, 32), filt=b"tent", srgb=0, comps=4, wrap=0),
    dict(src=(96, 96), dst=(200, 50), filt=b"mitchell", srgb=1, comps=3, wrap=0),     # up along X, down along Y
    dict(src=(17, 9), dst=(1, 1), filt=b"kaiser", srgb=1, comps=4, wrap=0),
])
def test_image_resample_is_byte_identical_to_the_reference(ops, ref, case):
    """b200_image_resample_rgba8 against basisu::image_resample (enc.cpp:1022) with the reference's own contributor lists and sRGB tables."""
    sw, sh = case["src"]; dw, dh = case["dst"]
    rng = np.random.default_rng(sw * 1000 + dh)
    img = util.synth(max(sw, sh), 71)[:sh, :sw].copy()
    img[::7, ::5] = rng.integers(0, 256, img[::7, ::5].shape, dtype=np.uint8)   # speckle: exercises the [0, 1] clamp under negative filter lobes
    init = np.zeros((dh, dw, 4), np.uint8); init[..., 3] = 255                   # what image::resize leaves behind
    want = init.copy()
    assert ref.lib.ref_image_resample(_ptr(np.ascontiguousarray(img)), sw, sh, _ptr(want), dw, dh, case["srgb"], case["filt"], ctypes.c_float(1.0), case["wrap"], 0, case["comps"])
    cx, cy = _ref_clists(ref, sw, sh, dw, dh, case["filt"], 1.0, case["wrap"])
    tables = None
    if case["srgb"]:
        s2l = np.zeros(256, np.float32); l2s = np.zeros(8192, np.uint8)
        ref.lib.ref_srgb_tables(_ptr(s2l), _ptr(l2s))
        tables = (s2l, l2s)
    got = ops.resample(img, init.copy(), cx, cy, 0, case["comps"], tables)
    assert np.array_equal(got, want), int((got != want).sum())
    assert (want[..., :3] != 0).any()
