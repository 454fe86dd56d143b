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
