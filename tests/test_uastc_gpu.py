"""GPU parity tests: everything goes through the C ABI (libbasisu_b200.so) and is compared bit for bit with the committed
golden vectors, and with the compiled reference when its .so travelled to the box."""
import os

import numpy as np
import pytest

import util
from basis_universal_b200 import uastc

pytestmark = pytest.mark.gpu

FLAG_SETS = [0, 1, 2, 3, 4, 2 | 512, 2 | 8, 2 | 16, 2 | 64, 1 | 128, 2 | 256, 3 | 512]


@pytest.fixture(scope="module")
def enc():
    e = uastc.Encoder(0)
    yield e
    e.close()


@pytest.mark.parametrize("flags", FLAG_SETS)
def test_golden_vectors_bit_exact(enc, golden, flags):
    out = enc.encode_uastc(golden["blocks"], flags)
    assert np.array_equal(out, golden[f"uastc_flags_{flags}"])
    assert enc.last_launch_count >= 3


def test_kodim03_level0_slice_bytes(enc):
    """BASELINE.json config 1: bytes identical to the slice data of the reference's .basis (md5 6d98eb72...)."""
    from test_uastc_cpu import basis_slice
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    ofs, size, nbx, nby = basis_slice(g["basis"])
    out = enc.encode_uastc(uastc.extract_blocks(g["image"]), 0)
    assert out.tobytes() == g["basis"].tobytes()[ofs:ofs + size]


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_matches_reference_on_fresh_inputs(enc, ref, level):
    blocks = np.concatenate([util.edge_case_blocks(21 + level), util.image_to_blocks(util.synth(256, 100 + level))])
    assert np.array_equal(enc.encode_uastc(blocks, level), ref.encode_uastc(blocks, level, threads=os.cpu_count()))


@pytest.mark.parametrize("n", [0, 1, 2, 31, 33, 127, 129, 1000])
def test_empty_and_ragged_sizes(enc, golden, n):
    src = np.resize(golden["blocks"], (max(n, 1), 64))[:n]
    want = np.resize(golden["uastc_flags_2"], (max(n, 1), 16))[:n]
    out = enc.encode_uastc(src, 2)
    assert out.shape == (n, 16) and np.array_equal(out, want)


def test_device_pointer_entry_point(enc, golden):
    import torch
    d_in = torch.from_numpy(golden["blocks"]).cuda()
    d_out = torch.empty((d_in.shape[0], 16), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    enc.encode_uastc_device(d_in.data_ptr(), d_in.shape[0], d_out.data_ptr(), 2)
    assert np.array_equal(d_out.cpu().numpy(), golden["uastc_flags_2"])


def test_full_size_properties_4096(enc, ref):
    """BASELINE.json config 2 at full size (1 048 576 blocks): size-independent properties instead of a full oracle run --
    determinism, shard-invariance (block-row halves == whole: what multi-GPU sharding relies on), agreement with the
    reference on a random sample of 4096 blocks, and every output decodable by the reference decoder."""
    blocks = util.image_to_blocks(util.synth(4096, 1234))
    a = enc.encode_uastc(blocks, 2)
    b = enc.encode_uastc(blocks, 2)
    assert np.array_equal(a, b)
    half = blocks.shape[0] // 2
    assert np.array_equal(np.concatenate([enc.encode_uastc(blocks[:half], 2), enc.encode_uastc(blocks[half:], 2)]), a)
    idx = np.random.default_rng(0).choice(blocks.shape[0], 4096, replace=False)
    assert np.array_equal(a[idx], ref.encode_uastc(blocks[idx], 2, threads=os.cpu_count()))
    dec = ref.unpack_uastc(a[idx])
    assert dec.shape == (4096, 64)
    # the whole output against the compiled reference's own run over all 1 048 576 blocks of this image (computed with
    # oracle/_ref on the CPU, 8 threads, 13 s; UASTC without RDO does not depend on the thread count)
    import hashlib
    assert hashlib.md5(a.tobytes()).hexdigest() == "f97bf1b5a5afbd41c5131737ad440e2b"
    assert int(a.astype(np.uint64).sum()) == 2093546678


def test_level4_sample_of_the_bench_image(enc, ref):
    """BASELINE config 5's encoder setting (level 4 + favour-simpler-modes) on 4096 blocks sampled from the 4096^2 bench image,
    against the compiled reference."""
    blocks = util.image_to_blocks(util.synth(4096, 1234))
    idx = np.sort(np.random.default_rng(5).choice(blocks.shape[0], 4096, replace=False))
    sample = np.ascontiguousarray(blocks[idx])
    for flags in (4, 4 | 512):
        assert np.array_equal(enc.encode_uastc(sample, flags), ref.encode_uastc(sample, flags, threads=os.cpu_count()))


def _class_inputs():
    """Inputs that leave one or two of the three slot-class work lists empty, or fill them unevenly."""
    base = util.image_to_blocks(util.synth(128, 77)).reshape(-1, 16, 4).copy()    # 1024 blocks
    opaque = base.copy(); opaque[:, :, 3] = 255
    alpha = base.copy(); alpha[:, :, 3] = (base[:, :, 0] // 2 + 17)
    gray = base.copy(); gray[:, :, 1] = gray[:, :, 0]; gray[:, :, 2] = gray[:, :, 0]; gray[:, :, 3] = 255
    gray_alpha = gray.copy(); gray_alpha[:, :, 3] = base[:, :, 1]
    solid = np.repeat(base[:, :1, :], 16, axis=1)
    rng = np.random.default_rng(3)
    mixed = np.concatenate([opaque[:300], alpha[:300], gray[:100], gray_alpha[:100], solid[:100]])
    mixed = mixed[rng.permutation(mixed.shape[0])]                                 # classes interleaved lane by lane
    return {"opaque": opaque, "alpha": alpha, "gray": gray, "gray_alpha": gray_alpha, "solid": solid, "mixed": mixed}


@pytest.mark.parametrize("kind", ["opaque", "alpha", "gray", "gray_alpha", "solid", "mixed"])
@pytest.mark.parametrize("level", [1, 2, 4])
def test_slot_class_work_lists(enc, ref, kind, level):
    blocks = _class_inputs()[kind].reshape(-1, 64)
    if level == 4:
        blocks = blocks[:160]      # level 4 runs 170 slots per block on the CPU oracle
    assert np.array_equal(enc.encode_uastc(blocks, level), ref.encode_uastc(blocks, level, threads=os.cpu_count()))


def test_multi_chunk_level4(enc):
    """Level 4 processes 2^18 blocks per pass: an input just over one chunk must equal the same blocks encoded separately."""
    tile = util.image_to_blocks(util.synth(64, 5))                                 # 256 distinct blocks
    n = (1 << 18) + 300
    blocks = np.resize(tile, (n, 64))
    out = enc.encode_uastc(blocks, 4)
    small = enc.encode_uastc(tile, 4)
    assert np.array_equal(out, np.resize(small, (n, 16)))


@pytest.mark.parametrize("flags", [0, 1, 2, 3, 2 | 512, 2 | 64 | 256])
def test_golden_real_images_bit_exact(enc, golden_real, flags):
    """2882 blocks sampled from the reference's own test images (kodim photographs, alpha0, wikipedia, xmen, tough, the 1x1
    solids): exercises the LA modes 15-17 and the solid path that the synthetic bench image never reaches."""
    assert np.array_equal(enc.encode_uastc(golden_real["blocks"], flags), golden_real[f"uastc_flags_{flags}"])
