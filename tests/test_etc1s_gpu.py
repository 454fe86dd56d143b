"""GPU parity for the five ETC1S seam entry points, through the C ABI, against the reference's OpenCL C kernels compiled for
the host (bit-exact: these stages are integer arithmetic plus one rounding of the block average)."""
import os

import numpy as np
import pytest

import util
from basis_universal_b200 import etc1s

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ocl():
    if not os.path.exists(util.OCL_SO):
        pytest.skip("host-compiled OpenCL kernels (oracle/_ref/libocl_ref.so) did not travel")
    return util.OclRef()


@pytest.fixture(scope="module")
def blocks():
    return np.concatenate([util.edge_case_blocks(5), util.image_to_blocks(util.synth(512, 31))])


@pytest.fixture(scope="module")
def ctx(blocks):
    c = etc1s.Etc1sContext(0)
    c.set_flavour(etc1s.FLAVOUR_OPENCL_KERNELS)  # the oracle in this module is the reference's OpenCL kernel source
    c.set_pixel_blocks(blocks)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_cpu(blocks):
    c = etc1s.Etc1sContext(0)                    # default flavour: the reference's CPU etc1_optimizer
    c.set_pixel_blocks(blocks)
    yield c
    c.close()


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (2, 64), (6, 165)])
def test_encode_blocks_cpu_flavour_matches_reference_cpu_optimizer(ctx_cpu, ref, blocks, perceptual, comp_level, perms):
    """Default flavour: bit-exact with basisu_frontend::init_etc1_images' CPU path (etc1_optimizer, frontend.cpp:775-815)."""
    import ctypes
    want = np.zeros((blocks.shape[0], 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(util._ptr(blocks), ctypes.c_uint32(blocks.shape[0]), util._ptr(want), perceptual, comp_level)
    assert np.array_equal(ctx_cpu.encode_etc1s_blocks(perceptual, perms), want)


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (2, 64)])
def test_pixel_clusters_cpu_flavour_matches_reference_cpu_optimizer(ctx_cpu, ref, perceptual, comp_level, perms):
    """Weighted unique-colour clusters vs the CPU path of generate_endpoint_codebook (frontend.cpp:1523-1549) fed the same
    texels with repetition (what the CPU path sees)."""
    import ctypes
    clusters, px, weights = util.pixel_cluster_inputs(13, 40)
    weights = np.minimum(weights, 6).astype(np.uint32)
    got = ctx_cpu.encode_etc1s_pixel_clusters(clusters, px, weights, perceptual, perms)
    for k, (n, first) in enumerate(clusters):
        n, first = int(n), int(first)
        rep = np.ascontiguousarray(np.repeat(px[first:first + n], weights[first:first + n], axis=0))
        out4 = np.zeros(4, np.uint8)
        ref.lib.ref_etc1s_encode_cluster(util._ptr(rep), ctypes.c_uint32(rep.shape[0]), perceptual, comp_level, util._ptr(out4))
        want = np.array([out4[0] << 3, out4[1] << 3, out4[2] << 3, (int(out4[3]) << 5) | (int(out4[3]) << 2) | 3], np.uint8)
        assert np.array_equal(got[k, :4], want), k


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("perms", [4, 16, 64, 165])
def test_encode_etc1s_blocks(ctx, ocl, blocks, perceptual, perms):
    assert np.array_equal(ctx.encode_etc1s_blocks(perceptual, perms), ocl.encode_etc1s_blocks(blocks, perceptual, perms))


@pytest.mark.parametrize("perceptual", [0, 1])
def test_refine_fosc_determine(ctx, ocl, blocks, perceptual):
    inp = util.etc1s_stage_inputs(blocks, 17, parents=16, clusters_per_parent=(3, 200), selectors_per_parent=(1, 900))
    assert np.array_equal(ctx.determine_selectors(inp["color5_inten"], perceptual), ocl.determine_selectors(blocks, inp["color5_inten"], perceptual))
    got = ctx.refine_endpoint_clusterization(inp["block_info"], inp["cluster_info"], inp["sorted_idx"], perceptual)
    assert np.array_equal(got, ocl.refine(blocks, inp["block_info"], inp["cluster_info"], inp["sorted_idx"], perceptual))
    got = ctx.find_optimal_selector_clusters_for_each_block(inp["fosc_blocks"], inp["selectors"], inp["sel_cluster_idx"], perceptual)
    assert np.array_equal(got, ocl.fosc(blocks, inp["fosc_blocks"], inp["selectors"], inp["sel_cluster_idx"], perceptual))


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("perms", [16, 64])
def test_encode_pixel_clusters(ctx, ocl, perceptual, perms):
    clusters, px, weights = util.pixel_cluster_inputs(9, 60)
    got = ctx.encode_etc1s_pixel_clusters(clusters, px, weights, perceptual, perms)
    want = ocl.encode_pixel_clusters(clusters, px, weights, perceptual, perms)
    assert np.array_equal(got[:, :4], want[:, :4])  # the kernel leaves the selector bytes undefined


def test_calls_fail_loudly_without_blocks():
    c = etc1s.Etc1sContext(0)
    with pytest.raises(Exception):
        c.encode_etc1s_blocks(True, 16)
    c.close()


def test_endpoint_histogram_kernel(ctx_cpu, blocks):
    enc = ctx_cpu.encode_etc1s_blocks(True, 16)
    hist = ctx_cpu.endpoint_histogram(enc)
    want = np.bincount(util.endpoint_keys(enc), minlength=1 << 18).astype(np.uint32) * 2
    assert np.array_equal(hist, want) and int(hist.sum()) == 2 * blocks.shape[0]
    keys, vecs, weights = etc1s.training_vectors_from_histogram(hist)
    assert len(keys) > 100 and float(vecs.max()) <= 1.0 and int(weights.sum()) == 2 * blocks.shape[0]


@pytest.mark.parametrize("perceptual", [0, 1])
def test_selector_training_kernel(ctx_cpu, ref, blocks, perceptual):
    """Per-block selector training vectors == the loop body of basisu_frontend::generate_selector_clusters."""
    import ctypes
    enc = ctx_cpu.encode_etc1s_blocks(bool(perceptual), 16)
    rnd = np.random.default_rng(6).integers(0, 256, (2000, 8), dtype=np.uint8)
    rnd[:, 3] |= 2                                   # differential mode (the only one ETC1S uses); deltas and selectors arbitrary
    for b in (enc, rnd):
        b = np.ascontiguousarray(b)
        keys, weights = ctx_cpu.selector_training(b, perceptual)
        wk, ww = np.zeros(b.shape[0], np.uint32), np.zeros(b.shape[0], np.uint32)
        ref.lib.ref_selector_training(util._ptr(b), ctypes.c_uint32(b.shape[0]), ctypes.c_uint32(perceptual), util._ptr(wk), util._ptr(ww))
        assert np.array_equal(keys, wk) and np.array_equal(weights, ww)
    u, w = etc1s.merge_selector_training(*ctx_cpu.selector_training(enc, perceptual))
    assert int(w.sum()) == int(ctx_cpu.selector_training(enc, perceptual)[1].astype(np.uint64).sum()) and np.all(np.diff(u.astype(np.int64)) > 0)
