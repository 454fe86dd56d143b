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


def test_plain_c_port_is_pinned_to_both_reference_builds(ocl, ref, blocks):
    """oracle/etc1s_port.c (kind "port") agrees with the compiled reference CPU code and with the reference's OpenCL kernels."""
    port_so = os.path.join(util.ROOT, "oracle", "_ref", "liboracle_port.so")
    if not os.path.exists(port_so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(util.ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    port = ctypes.CDLL(port_so)
    port.port_color_distance.restype = ctypes.c_uint32
    rng = np.random.default_rng(1)
    for _ in range(2000):
        a = rng.integers(0, 256, 4, dtype=np.uint8); b = rng.integers(0, 256, 4, dtype=np.uint8)
        for p in (0, 1):
            assert port.port_color_distance(p, _ptr(a), _ptr(b)) == ref.lib.ref_color_distance(p, _ptr(a), _ptr(b), 0)
    inp = util.etc1s_stage_inputs(blocks, 23)
    n = 400
    for p in (0, 1):
        want_cl = ocl.determine_selectors(blocks, inp["color5_inten"], p)
        want_cpu = np.zeros((blocks.shape[0], 8), np.uint8)
        ref.lib.ref_etc1s_determine_selectors(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(inp["color5_inten"]), _ptr(want_cpu), p)
        assert np.array_equal(want_cl, want_cpu)  # the reference's two implementations agree with each other
        for i in range(n):
            out = np.zeros(8, np.uint8)
            port.port_determine_selectors(_ptr(blocks[i]), _ptr(inp["color5_inten"][i]), p, _ptr(out))
            assert np.array_equal(out, want_cpu[i])


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (2, 64), (6, 165)])
def test_cpu_flavour_matches_reference_cpu_optimizer(ref, emu, blocks, perceptual, comp_level, perms):
    """ETC1S_FLAVOUR_CPU (the library default) is bit-exact with the reference CPU etc1_optimizer (incl. its Bloom filter)."""
    want = np.zeros((blocks.shape[0], 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(want), perceptual, comp_level)
    got = np.zeros_like(want)
    emu.lib.emu_etc1s_encode_blocks_flavour(_ptr(blocks), blocks.shape[0], _ptr(got), perceptual, perms, 1)
    assert np.array_equal(got, want)


def test_training_vectors_from_colliding_keys_are_merged():
    """Distinct 18-bit keys can give the same vec6F: with intensity table 7 (+-183) every channel value 9..22 (5 bits, i.e. 74..181 at
    8 bits) clamps to low 0 / high 255, so 14^3 keys collapse to (0,0,0,1,1,1). The clusterer's std::map merges them
    (enc.h:2228-2260); merge_training_vectors is that merge (lexicographic order, summed weights)."""
    import numpy as np
    from basis_universal_b200 import etc1s
    hist = np.zeros(1 << 18, np.uint32)
    n = 0
    for r5 in (9, 15, 22):
        for g5 in (9, 22):
            for b5 in (10, 21):
                hist[(r5 << 13) | (g5 << 8) | (b5 << 3) | 7] = 2 * (n + 1)
                n += 1
    hist[(3 << 13) | (3 << 8) | (3 << 3) | 0] = 4     # one ordinary key
    keys, vecs, weights = etc1s.training_vectors_from_histogram(hist)
    assert len(keys) == n + 1
    u, w = etc1s.merge_training_vectors(vecs, weights)
    assert u.shape[0] == 2
    assert np.array_equal(u[0], np.array([0, 0, 0, 1, 1, 1], np.float32)) or np.array_equal(u[1], np.array([0, 0, 0, 1, 1, 1], np.float32))
    assert int(w.sum()) == int(hist.sum()) and sorted(int(x) for x in w) == sorted([4, int(hist.sum()) - 4])
    assert np.all(np.lexsort(u.T[::-1]) == np.arange(u.shape[0]))   # lexicographic order, the std::map's
