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
    c.set_pixel_blocks(blocks)
    yield c
    c.close()


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
