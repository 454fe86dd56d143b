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


def _cluster_cases(seed):
    """Ragged clusters of the synthetic image's blocks (similar blocks together, like real endpoint clusters) + selectors per block."""
    img = util.synth(128, seed)
    blocks = util.image_to_blocks(img)
    rng = np.random.default_rng(seed)
    lum = blocks.reshape(blocks.shape[0], 16, 4)[:, :, :3].astype(np.int64).sum(axis=(1, 2))
    order = np.argsort(lum, kind="stable").astype(np.uint32)
    cuts = np.sort(rng.choice(np.arange(1, blocks.shape[0]), 40, replace=False))
    return blocks, [c for c in np.split(order, cuts)], rng


def emu_optimize_cluster(emu, blk, sels, cur4, perceptual, perms, flavour=1):
    out4 = np.zeros(4, np.uint8)
    cur_err = ctypes.c_uint64(0)
    emu.lib.emu_etc1s_optimize_cluster.restype = ctypes.c_uint64
    err = emu.lib.emu_etc1s_optimize_cluster(_ptr(blk), ctypes.c_uint32(blk.shape[0]), _ptr(sels) if sels is not None else None,
                                             _ptr(cur4) if cur4 is not None else None, perceptual, ctypes.c_uint32(perms), flavour, _ptr(out4), ctypes.byref(cur_err))
    return tuple(int(v) for v in out4), int(err), int(cur_err.value)


def ref_reoptimize_cluster(ref, blk, sels, cur4, perceptual, comp_level):
    out4 = np.zeros(4, np.uint8)
    cur_err = ctypes.c_uint64(0)
    ref.lib.ref_etc1s_reoptimize_cluster.restype = ctypes.c_uint64
    err = ref.lib.ref_etc1s_reoptimize_cluster(_ptr(blk), ctypes.c_uint32(blk.shape[0]), _ptr(sels), _ptr(cur4), perceptual, comp_level, _ptr(out4), ctypes.byref(cur_err))
    return tuple(int(v) for v in out4), int(err), int(cur_err.value)


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (2, 64), (6, 165)])
def test_cluster_optimizer_matches_reference_cpu_optimizer(ref, emu, perceptual, comp_level, perms):
    """The team-templated cluster optimiser (what the endpoint-cluster kernels run) against etc1_optimizer over the gathered texels."""
    blocks, clusters, _ = _cluster_cases(31)
    ref.lib.ref_etc1s_encode_cluster.restype = ctypes.c_uint64
    for c in clusters:
        blk = np.ascontiguousarray(blocks[c])
        want4 = np.zeros(4, np.uint8)
        want_err = ref.lib.ref_etc1s_encode_cluster(_ptr(blk), ctypes.c_uint32(blk.shape[0] * 16), perceptual, comp_level, _ptr(want4))
        got4, got_err, _ = emu_optimize_cluster(emu, blk, None, None, perceptual, perms)
        assert got4 == tuple(int(v) for v in want4) and got_err == want_err, len(c)


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(2, 64), (6, 165)])
def test_reoptimize_cluster_with_forced_selectors_matches_reference(ref, emu, perceptual, comp_level, perms):
    """reoptimize_remapped_endpoints' per-cluster body (frontend.cpp:3018-3090): imposed selectors, new and current error."""
    blocks, clusters, rng = _cluster_cases(37)
    n = blocks.shape[0]
    etc = np.zeros((n, 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(n), _ptr(etc), perceptual, 1)
    # selectors as the frontend has them: selector INDEX (0..3 = darkest..brightest) per texel, bits 2 * (x + 4 * y); taken from
    # each block's own encoding so that they are meaningful, then a few blocks get random ones
    sel_words = np.zeros(n, np.uint32)
    raw_to_sel = np.array([2, 3, 1, 0], np.uint32)
    for b in range(n):
        msb = (int(etc[b, 4]) << 8) | int(etc[b, 5]); lsb = (int(etc[b, 6]) << 8) | int(etc[b, 7])
        w = 0
        for y in range(4):
            for x in range(4):
                bit = x * 4 + y
                raw = ((lsb >> bit) & 1) | (((msb >> bit) & 1) << 1)
                w |= int(raw_to_sel[raw]) << (2 * (x + 4 * y))
        sel_words[b] = w
    rnd = rng.choice(n, n // 10, replace=False)
    sel_words[rnd] = rng.integers(0, 1 << 32, rnd.shape[0], dtype=np.uint64).astype(np.uint32)
    improved = 0
    for c in clusters:
        blk = np.ascontiguousarray(blocks[c]); sels = np.ascontiguousarray(sel_words[c])
        cur4 = np.array([etc[c[0], 0] >> 3, etc[c[0], 1] >> 3, etc[c[0], 2] >> 3, etc[c[0], 3] >> 5], np.uint8)  # some member's endpoint
        want = ref_reoptimize_cluster(ref, blk, sels, cur4, perceptual, comp_level)
        got = emu_optimize_cluster(emu, blk, sels, cur4, perceptual, perms)
        assert got == want, (len(c), got, want)
        improved += want[1] < want[2]
    assert improved > 0


def test_cluster_average_past_2_pow_24_follows_the_reference_float_sum(ref, emu):
    """etc1_optimizer::init sums texels into a float (etc.cpp:1022-1040); past 2^24 (> 4112 bright blocks) that sum rounds, and the
    optimiser's starting colour follows it. Here the red channel's exact mean rounds to 31 and the float-summed one to 30, and at comp_level 1 the
    optimiser ends on red = 30 where an exact mean would end on 31 with a lower error."""
    n = 4500
    rng = np.random.default_rng(76)
    red = rng.integers(248, 255, n * 16)
    red = (red - (rng.random(n * 16) < 0.1135)).astype(np.uint8)
    to5 = lambda a: int(np.float32(np.float32(np.float32(a * np.float32(31)) / np.float32(255)) + np.float32(.5)))
    exact = np.float32(int(red.astype(np.int64).sum())) / np.float32(n * 16)
    serial = np.cumsum(red.astype(np.float32), dtype=np.float32)[-1] / np.float32(n * 16)
    assert (to5(exact), to5(serial)) == (31, 30)
    blocks = np.empty((n * 16, 4), np.uint8)
    blocks[:, 0] = red; blocks[:, 1] = rng.integers(100, 140, n * 16); blocks[:, 2] = rng.integers(50, 60, n * 16); blocks[:, 3] = 255
    blocks = np.ascontiguousarray(blocks.reshape(n, 64))
    ref.lib.ref_etc1s_encode_cluster.restype = ctypes.c_uint64
    for perceptual, comp_level, perms in ((0, 1, 16), (1, 1, 16), (1, 2, 64)):
        want4 = np.zeros(4, np.uint8)
        want_err = ref.lib.ref_etc1s_encode_cluster(_ptr(blocks), ctypes.c_uint32(n * 16), perceptual, comp_level, _ptr(want4))
        got4, got_err, _ = emu_optimize_cluster(emu, blocks, None, None, perceptual, perms)
        assert got4 == tuple(int(v) for v in want4) and got_err == want_err
    assert tuple(int(v) for v in want4)[0] in (30, 31)


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (4, 64)])
def test_refit_cluster_with_free_selectors_matches_reference(ref, emu, perceptual, comp_level, perms):
    """generate_endpoint_codebook at step >= 1: optimiser result + error of the previous endpoint (best of four colours per texel)."""
    blocks, clusters, rng = _cluster_cases(41)
    ref.lib.ref_etc1s_refit_cluster.restype = ctypes.c_uint64
    for c in clusters:
        blk = np.ascontiguousarray(blocks[c])
        prev4 = rng.integers(0, [32, 32, 32, 8], 4).astype(np.uint8)
        want4 = np.zeros(4, np.uint8); want_prev = ctypes.c_uint64(0)
        want_new = ref.lib.ref_etc1s_refit_cluster(_ptr(blk), ctypes.c_uint32(len(c)), _ptr(prev4), perceptual, comp_level, _ptr(want4), ctypes.byref(want_prev))
        got = emu_optimize_cluster(emu, blk, None, prev4, perceptual, perms)
        assert got == (tuple(int(v) for v in want4), want_new, want_prev.value), len(c)


@pytest.mark.parametrize("perceptual,thresh", [(0, 1.5), (1, 1.5), (0, 0.375), (1, 3.0), (0, 0.0)])
def test_backend_endpoint_prediction_decisions_match_reference_scan(ref, emu, perceptual, thresh):
    """The per-block decision functions of the endpoint-prediction wavefront kernel (bu_etc1s.h), run in raster order on the host,
    against the restatement of basisu_backend::create_encoder_blocks' scan built from the reference's own primitives (backend.cpp:437-600)."""
    img = util.synth(192, 67)
    blocks = util.image_to_blocks(img)
    w, h = 48, 37
    n = w * h
    blocks = np.ascontiguousarray(blocks[:n])
    etc = np.zeros((n, 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(n), _ptr(etc), perceptual, 1)
    key = ((etc[:, 0] >> 5).astype(np.uint32) << 9) | ((etc[:, 1] >> 5).astype(np.uint32) << 6) | ((etc[:, 2] >> 5).astype(np.uint32) << 3) | (etc[:, 3] >> 5)
    uniq, idx0 = np.unique(key, return_inverse=True)
    cb = np.stack([((uniq >> 9) & 7) * 4 + 2, ((uniq >> 6) & 7) * 4 + 2, ((uniq >> 3) & 7) * 4 + 2, uniq & 7], -1).astype(np.uint8)
    e = cb[idx0]
    etc[:, 0] = e[:, 0] << 3; etc[:, 1] = e[:, 1] << 3; etc[:, 2] = e[:, 2] << 3; etc[:, 3] = (e[:, 3] << 5) | (e[:, 3] << 2) | 3
    want_idx = idx0.astype(np.uint32).copy(); want_pred = np.zeros(n, np.uint8)
    ref.lib.ref_backend_endpoint_prediction(_ptr(blocks), _ptr(etc), w, h, _ptr(cb), ctypes.c_float(thresh), perceptual, _ptr(want_idx), _ptr(want_pred))
    got_idx = idx0.astype(np.uint32).copy(); got_pred = np.zeros(n, np.uint8)
    emu.lib.emu_backend_endpoint_prediction(_ptr(blocks), _ptr(etc), w, h, _ptr(cb), ctypes.c_float(thresh), perceptual, _ptr(got_idx), _ptr(got_pred))
    assert np.array_equal(got_pred, want_pred) and np.array_equal(got_idx, want_idx)
    if thresh > 0:
        assert (want_idx != idx0).sum() > 20


@pytest.mark.parametrize("kind,n_syms,m", [("walk", 2400, 16000), ("walk", 5000, 120000), ("uniform", 500, 20000), ("few", 3000, 50000),
                                            ("constant", 40, 1000), ("uniform", 2, 50), ("walk", 1, 10), ("uniform", 700, 2)])
def test_sparse_palette_ordering_matches_reference(ref, emu, kind, n_syms, m):
    """The sparse formulation of palette_index_reorderer::init that the device kernel walks (b200_backend.cu), run serially on the host."""
    rng = np.random.default_rng(n_syms + m)
    if kind == "walk":
        steps = np.where(rng.random(m) < 0.75, rng.integers(0, 48, m), rng.integers(0, n_syms, m))
        idx = (np.cumsum(steps) % n_syms).astype(np.uint32)
    elif kind == "uniform":
        idx = rng.integers(0, n_syms, m).astype(np.uint32)
    elif kind == "few":
        idx = np.repeat(rng.integers(0, max(n_syms // 50, 2), m // 7 + 1), 7)[:m].astype(np.uint32)
    else:
        idx = np.full(m, 3 % n_syms, np.uint32)
    want = np.zeros(n_syms, np.uint32); got = np.zeros(n_syms, np.uint32)
    ref.lib.ref_palette_reorder(ctypes.c_uint32(m), _ptr(idx), ctypes.c_uint32(n_syms), _ptr(want))
    emu.lib.emu_palette_reorder(ctypes.c_uint32(m), _ptr(idx), ctypes.c_uint32(n_syms), _ptr(got))
    assert np.array_equal(got, want), int((got != want).sum())
