"""CPU-side checks for the steps either side of the block path: the oracle hooks behave as documented, the host emulation of the
device decode agrees with basist::unpack_uastc, and the host restatement of image_metrics' floating-point tail agrees with
image_metrics::calc. No GPU needed."""
import ctypes

import numpy as np
import pytest

import util
from basis_universal_b200 import image

pytestmark = pytest.mark.skipif(not util.build_ref(), reason="oracle/_ref not buildable here")


@pytest.fixture(scope="module")
def ref():
    return util.Ref()


def ref_extract(ref, img):
    h, w = img.shape[:2]
    out = np.zeros((((h + 3) // 4) * ((w + 3) // 4), 64), np.uint8)
    ref.lib.ref_extract_source_blocks(util._ptr(np.ascontiguousarray(img)), ctypes.c_uint32(w), ctypes.c_uint32(h), util._ptr(out))
    return out


def ref_metrics(ref, a, b, first, total, avg=True, use601=False):
    h, w = a.shape[:2]
    out = np.zeros(5, np.float64)
    ref.lib.ref_image_metrics(util._ptr(np.ascontiguousarray(a)), util._ptr(np.ascontiguousarray(b)), ctypes.c_uint32(w), ctypes.c_uint32(h), ctypes.c_uint32(first), ctypes.c_uint32(total), ctypes.c_uint32(int(avg)), ctypes.c_uint32(int(use601)), util._ptr(out))
    return dict(zip(["max", "mean", "mean_squared", "rms", "psnr"], out.tolist()))


def np_histograms(a, b):
    """numpy restatement of k_block_metrics (the GPU test compares the kernel against this and both against the oracle)."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    hist = np.zeros((6, 256), np.uint64)
    for c in range(4):
        hist[c] = np.bincount(d[..., c].ravel(), minlength=256)

    def luma(x, k):
        x = x.astype(np.uint32)
        return ((k[0] * x[..., 0] + k[1] * x[..., 1] + k[2] * x[..., 2] + 32768) >> 16).astype(np.int32)
    hist[4] = np.bincount(np.abs(luma(a, (13938, 46869, 4729)) - luma(b, (13938, 46869, 4729))).ravel(), minlength=256)
    hist[5] = np.bincount(np.abs(luma(a, (19595, 38470, 7471)) - luma(b, (19595, 38470, 7471))).ravel(), minlength=256)
    return hist


@pytest.mark.parametrize("shape", [(8, 8), (5, 7), (1, 1), (13, 4), (64, 37)])
def test_oracle_extract_clamps_edges(ref, shape):
    h, w = shape
    img = util.synth(64, 3)[:h, :w]
    pad = np.pad(img, ((0, (-h) % 4), (0, (-w) % 4), (0, 0)), mode="edge")
    assert np.array_equal(ref_extract(ref, img), util.image_to_blocks(pad))


@pytest.mark.parametrize("first,total,avg,use601", [(0, 3, True, False), (0, 4, True, False), (1, 1, True, False), (3, 1, True, False), (0, 0, True, False), (0, 0, True, True), (0, 3, False, False)])
def test_metrics_tail_matches_image_metrics_calc(ref, first, total, avg, use601):
    a = util.synth(96, 11)
    b = np.clip(a.astype(np.int32) + np.random.default_rng(2).integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    got = image.metrics_from_histograms(np_histograms(a, b), 96, 96, first, total, avg, use601)
    want = ref_metrics(ref, a, b, first, total, avg, use601)
    assert got == want


def test_identical_images_give_100_db(ref):
    a = util.synth(32, 1)
    assert image.metrics_from_histograms(np_histograms(a, a), 32, 32, 0, 3)["psnr"] == 100.0 == ref_metrics(ref, a, a, 0, 3)["psnr"]


@pytest.mark.skipif(not util.build_emu(), reason="host emulation not buildable")
@pytest.mark.parametrize("level", [0, 2, 3])
def test_emulated_decode_matches_reference_unpack(ref, level):
    blocks = np.concatenate([util.edge_case_blocks(5), util.image_to_blocks(util.synth(64, 40 + level))])
    u = ref.encode_uastc(blocks, level)
    emu = util.Emu()
    out = np.zeros((u.shape[0], 64), np.uint8)
    assert emu.lib.emu_uastc_unpack_blocks(util._ptr(u), ctypes.c_uint32(u.shape[0]), util._ptr(out)) == 1
    assert np.array_equal(out, ref.unpack_uastc(u))


@pytest.mark.skipif(not util.build_emu(), reason="host emulation not buildable")
def test_emulated_decode_random_bit_patterns(ref):
    """Arbitrary 128-bit patterns: the decode accepts exactly the blocks basist::unpack_uastc accepts and yields the same texels."""
    emu = util.Emu()
    cand = np.random.default_rng(4).integers(0, 256, (4096, 16), dtype=np.uint8)
    want, got = np.zeros((1, 64), np.uint8), np.zeros((1, 64), np.uint8)
    invalid = 0
    for b in cand:
        b = np.ascontiguousarray(b)
        r = ref.lib.ref_unpack_uastc_blocks(util._ptr(b), ctypes.c_uint32(1), util._ptr(want))
        e = emu.lib.emu_uastc_unpack_blocks(util._ptr(b), ctypes.c_uint32(1), util._ptr(got))
        assert bool(r) == bool(e)
        if r:
            assert np.array_equal(want, got)
        else:
            invalid += 1
    assert 0 < invalid < 4096
