"""GPU parity for b200_uastc_rdo against the reference's uastc_rdo (compiled), through the C ABI."""
import numpy as np
import pytest

import util
from basis_universal_b200 import uastc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc():
    e = uastc.Encoder(0)
    yield e
    e.close()


@pytest.mark.parametrize("level,lam,jobs,extra", [(0, 1.0, 1, 512), (1, 2.0, 4, 512), (2, 1.0, 4, 512), (2, 3.0, 2, 0), (3, 0.5, 4, 512)])
def test_rdo_matches_reference(enc, ref, level, lam, jobs, extra):
    src = util.image_to_blocks(util.rdo_test_image())
    flags = level | extra
    blocks = enc.encode_uastc(src, flags)
    assert np.array_equal(blocks, ref.encode_uastc(src, flags))
    want = util.ref_rdo(ref, blocks, src, lam, flags, jobs)
    got = enc.uastc_rdo(blocks, src, uastc.uastc_rdo_params(lambda_=lam), flags, jobs)
    assert (want != blocks).any()
    assert np.array_equal(got, want)


def test_rdo_larger_image_four_chains(enc, ref):
    src = util.image_to_blocks(util.synth(1024, 4242))
    flags = 2 | 512
    blocks = enc.encode_uastc(src, flags)
    want = util.ref_rdo(ref, blocks, src, 1.0, flags, 4)
    got = enc.uastc_rdo(blocks, src, uastc.uastc_rdo_params(lambda_=1.0), flags, 4)
    print("RDO 1024^2: %.1f ms on the GPU, %d blocks modified" % (enc.last_kernel_ms, int((got != blocks).any(1).sum())))
    assert np.array_equal(got, want)


def test_rdo_level4(enc, ref):
    """BASELINE config 5's flags: level 4 (cPackUASTCLevelVerySlow) + favour-simpler-modes (set by the compressor when RDO is on), lambda 1.0, 4 chains."""
    src = util.image_to_blocks(util.rdo_test_image())
    flags = 4 | 512
    blocks = enc.encode_uastc(src, flags)
    assert np.array_equal(blocks, ref.encode_uastc(src, flags))
    want = util.ref_rdo(ref, blocks, src, 1.0, flags, 4)
    got = enc.uastc_rdo(blocks, src, uastc.uastc_rdo_params(lambda_=1.0), flags, 4)
    assert (want != blocks).any()
    assert np.array_equal(got, want)


def test_rdo_batch_equals_per_slice_reference(enc, ref):
    """Several slices of different sizes in one b200_uastc_rdo_batch call == one reference uastc_rdo per slice (the loop over slices
    at comp.cpp:1996-2089), including a slice too small to be split (<= 8 blocks per job) and an empty one."""
    imgs = [util.rdo_test_image(), util.synth(128, 5), util.rdo_test_image()[:64, :64], util.synth(8, 6)[:, :16]]
    srcs = [util.image_to_blocks(i) for i in imgs]
    flags = 2 | 512
    enc_blocks = [enc.encode_uastc(s, flags) for s in srcs]
    want = np.concatenate([util.ref_rdo(ref, b, s, 1.5, flags, 4) for b, s in zip(enc_blocks, srcs)])
    counts = [s.shape[0] for s in srcs[:2]] + [0] + [s.shape[0] for s in srcs[2:]]
    got = enc.uastc_rdo_batch(np.concatenate(enc_blocks), np.concatenate(srcs), counts, uastc.uastc_rdo_params(lambda_=1.5), flags, 4)
    assert np.array_equal(got, want)


def test_rdo_4096_four_chains_device_resident(enc):
    """Full-size slice (1 048 576 blocks, 4 chains of 262 144): encode and RDO back to back in HBM through the _device entry
    points. Pinned against the compiled reference's own run over the whole image (oracle/_ref on the CPU: encode_uastc with
    flags 2|512 -> md5 f2d51fcf..., then uastc_rdo(lambda 1.0, total_jobs 4) -> md5 df47a1bc..., 12 blocks modified; 14 s + 45 s
    on 8 cores), so the GPU test does not have to repeat the 3-minute CPU pass."""
    import hashlib
    import torch
    src = util.image_to_blocks(util.synth(4096, 1234))
    flags = 2 | 512
    n = src.shape[0]
    d_src = torch.from_numpy(src).cuda()
    d_out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    enc.encode_uastc_device(d_src.data_ptr(), n, d_out.data_ptr(), flags)
    blocks = d_out.cpu().numpy()
    assert hashlib.md5(blocks.tobytes()).hexdigest() == "f2d51fcf380bb2243d0143ae3c077fed"
    enc.uastc_rdo_batch_device(d_out.data_ptr(), d_src.data_ptr(), [n], uastc.uastc_rdo_params(lambda_=1.0), flags, 4)
    ms = enc.last_kernel_ms
    got = d_out.cpu().numpy()
    print("RDO 4096^2, 4 chains: %.1f ms on the GPU, %d blocks modified" % (ms, int((got != blocks).any(1).sum())))
    assert int((got != blocks).any(1).sum()) == 12
    assert hashlib.md5(got.tobytes()).hexdigest() == "df47a1bcfab09916169b2b2d8f8ee479"


def test_rdo_2048_natural_content_four_chains(enc, ref):
    """A slice where RDO actually rewrites blocks (kodim03 tiled to 2048 x 1024: 131 072 blocks, 4 chains), against the reference."""
    import os
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    img = np.ascontiguousarray(np.tile(g["image"], (2, 3, 1))[:1024, :2048])
    src = util.image_to_blocks(img)
    flags = 2 | 512
    blocks = enc.encode_uastc(src, flags)
    want = util.ref_rdo(ref, blocks, src, 1.0, flags, 4)
    got = enc.uastc_rdo(blocks, src, uastc.uastc_rdo_params(lambda_=1.0), flags, 4)
    print("RDO 2048x1024 natural: %.1f ms on the GPU, %d of %d blocks modified" % (enc.last_kernel_ms, int((got != blocks).any(1).sum()), len(blocks)))
    assert (want != blocks).any(1).sum() > len(blocks) // 20
    assert np.array_equal(got, want)
