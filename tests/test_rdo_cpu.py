"""UASTC RDO post-pass: host-emulation of the device code (sequential rendition of the chain) vs the reference's uastc_rdo,
bit for bit, including the dependence of the output on the chain split (`total_jobs`)."""
import numpy as np
import pytest

import util


@pytest.mark.parametrize("level,lam,jobs,extra", [(0, 1.0, 1, 512), (1, 2.0, 4, 512), (2, 1.0, 4, 512), (2, 3.0, 2, 0), (3, 0.5, 4, 512)])
def test_hostemu_rdo_matches_reference(ref, emu, level, lam, jobs, extra):
    src = util.image_to_blocks(util.rdo_test_image())
    flags = level | extra
    enc = ref.encode_uastc(src, flags)
    want = util.ref_rdo(ref, enc, src, lam, flags, jobs)
    got = util.emu_rdo(emu, enc, src, lam, flags, jobs)
    assert (want != enc).any(), "fixture too easy: RDO changed nothing"
    assert np.array_equal(got, want)


def test_rdo_output_depends_on_chain_split(ref):
    """SURVEY.md section 7 hard part 3: 1, 2 and 4 chains give different bytes, so `total_jobs` is part of the contract."""
    src = util.image_to_blocks(util.rdo_test_image())
    enc = ref.encode_uastc(src, 1 | 512)
    outs = [util.ref_rdo(ref, enc, src, 2.0, 1 | 512, j).tobytes() for j in (1, 2, 4)]
    assert len(set(outs)) == 3
