"""CPU-only: the host-emulation build of the ETC1S device code against the reference's own OpenCL C kernels compiled for the
host (oracle/_ref/libocl_ref.so), plus the oracle's internal consistency with the reference's CPU decoder."""
import ctypes
import os

import numpy as np
import pytest

import util
from util import _ptr


@pytest.fixture(scope="module")
def ocl():
    util.build_ref()
    if not os.path.exists(util.OCL_SO):
        pytest.skip("host-compiled OpenCL kernels not available")
    return util.OclRef()


@pytest.fixture(scope="module")
def blocks():
    return np.concatenate([util.edge_case_blocks(5), util.image_to_blocks(util.synth(96, 31))])


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("perms", [4, 16, 64, 165])
def test_encode_blocks_matches_ocl_kernel(ocl, emu, blocks, perceptual, perms):
    want = ocl.encode_etc1s_blocks(blocks, perceptual, perms)
    got = np.zeros_like(want)
    emu.lib.emu_etc1s_encode_blocks(_ptr(blocks), blocks.shape[0], _ptr(got), perceptual, perms)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("perceptual", [0, 1])
def test_determine_selectors_and_refine_match_ocl_kernels(ocl, emu, blocks, perceptual):
    inp = util.etc1s_stage_inputs(blocks, 17)
    want = ocl.determine_selectors(blocks, inp["color5_inten"], perceptual)
    got = np.zeros_like(want)
    emu.lib.emu_etc1s_determine_selectors(_ptr(blocks), blocks.shape[0], _ptr(inp["color5_inten"]), _ptr(got), perceptual)
    assert np.array_equal(got, want)
    want = ocl.refine(blocks, inp["block_info"], inp["cluster_info"], inp["sorted_idx"], perceptual)
    got = np.zeros(blocks.shape[0], np.uint32)
    emu.lib.emu_etc1s_refine(_ptr(blocks), blocks.shape[0], _ptr(inp["block_info"]), _ptr(inp["cluster_info"]), _ptr(got), perceptual)
    assert np.array_equal(got, want)
    assert (want != inp["block_info"]["cur_cluster_index"]).any()  # the stage actually moves blocks in this fixture


def test_ocl_kernel_blocks_decode_like_reference_cpu(ocl, ref, blocks):
    """Ties the host-compiled kernels to the rest of the reference: blocks they emit decode (reference unpack_etc1) to texels
    whose PSNR against the source matches what the reference's CPU etc1_optimizer achieves, within the reference's own
    CPU-vs-OpenCL tolerance band (basisu_tool.cpp:6884-6922 allows 0.2 dB)."""
    def psnr(enc):
        dec = np.zeros((blocks.shape[0], 64), np.uint8)
        ref.lib.ref_unpack_etc1_blocks(_ptr(np.ascontiguousarray(enc)), ctypes.c_uint32(blocks.shape[0]), _ptr(dec))
        d = dec.reshape(-1, 16, 4)[:, :, :3].astype(np.float64) - blocks.reshape(-1, 16, 4)[:, :, :3].astype(np.float64)
        return 10 * np.log10(255 ** 2 / np.mean(d ** 2))
    cpu = np.zeros((blocks.shape[0], 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(cpu), 1, 1)
    assert abs(psnr(ocl.encode_etc1s_blocks(blocks, 1, 16)) - psnr(cpu)) < 0.2
