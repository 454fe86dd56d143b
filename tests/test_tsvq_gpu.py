"""GPU parity for the ETC1S codebook stages behind the wider seam (include/basisu_b200.h):
  b200_tsvq_generate                     vs generate_hierarchical_codebook_threaded (enc.h:2219) in the compiled reference
  b200_etc1s_encode_endpoint_clusters    vs the CPU path of generate_endpoint_codebook (frontend.cpp:1493-1549)
  b200_etc1s_optimize_selector_codebook  vs create_optimized_selector_codebook (frontend.cpp:2259-2345)
All bit-exact: the last two are integer stages, and the clusterer accumulates every float sum in the reference's member order
(b200_tsvq.cu, "serial sums"), so clusters, their order and the order of their members equal the CPU's."""
import ctypes

import numpy as np
import pytest

import util
from util import _ptr
from basis_universal_b200 import etc1s

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = etc1s.Etc1sContext(0)
    yield c
    c.close()


def ref_tsvq(ref, vecs, weights, max_codebook, max_parent, max_threads=1, even_odd=False):
    vecs = np.ascontiguousarray(vecs, np.float32); weights = np.ascontiguousarray(weights, np.uint64)
    n, dim = vecs.shape
    cl_off = np.zeros(n + 1, np.uint32); cl_idx = np.zeros(n, np.uint32); pa_off = np.zeros(n + 1, np.uint32); pa_idx = np.zeros(n, np.uint32)
    nc = ctypes.c_uint32(0); npar = ctypes.c_uint32(0)
    ok = ref.lib.ref_tsvq(dim, n, _ptr(vecs), _ptr(weights), int(max_codebook), int(max_parent), int(max_threads), int(even_odd), _ptr(cl_off), _ptr(cl_idx), ctypes.byref(nc),
                          _ptr(pa_off), _ptr(pa_idx), ctypes.byref(npar))
    assert ok
    return [cl_idx[cl_off[i]:cl_off[i + 1]] for i in range(nc.value)], [pa_idx[pa_off[i]:pa_off[i + 1]] for i in range(npar.value)]


def distortion(vecs, weights, clusters):
    """weighted sum of squared distances to the cluster centroids (what the tree minimises)"""
    v = vecs.astype(np.float64); w = weights.astype(np.float64)
    total = 0.0
    for c in clusters:
        ww = w[c][:, None]
        mean = (v[c] * ww).sum(0) / ww.sum()
        total += float((((v[c] - mean) ** 2) * ww).sum())
    return total


def labels(n, clusters):
    lab = np.full(n, -1, np.int64)
    for i, c in enumerate(clusters):
        lab[c] = i
    return lab


def same_clusters(got, want):
    w = {c.tobytes() for c in want}
    return sum(1 for c in got if c.tobytes() in w)


def assert_same_codebook(got, want):
    """same clusters, in the same order, each listing its training vectors in the same order"""
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), f"cluster {i} differs"


def check_partition(n, clusters):
    allidx = np.concatenate(clusters)
    assert allidx.shape[0] == n and np.array_equal(np.sort(allidx), np.arange(n)), "clusters must partition the training set"


def endpoint_training(ref, img, comp_level=1):
    """init_endpoint_training_vectors (frontend.cpp:825-866): per block, low/high block colours / 255, inserted twice with weight 1."""
    blocks = util.image_to_blocks(img)
    etc = np.zeros((blocks.shape[0], 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(etc), 0, comp_level)
    hist = np.bincount(util.endpoint_keys(etc), minlength=1 << 18).astype(np.uint32)
    keys = util.endpoint_keys(etc)
    _, kvecs, _ = etc1s.training_vectors_from_histogram(hist)
    lut = np.zeros((1 << 18, 6), np.float32)
    lut[np.nonzero(hist)[0]] = kvecs
    vecs = np.repeat(lut[keys], 2, axis=0)
    return blocks, etc, vecs, np.ones(vecs.shape[0], np.uint64)


@pytest.mark.parametrize("max_codebook,max_parent", [(300, 16), (37, 0), (2416, 16)])
def test_tsvq_endpoints_matches_reference(ctx, ref, max_codebook, max_parent):
    g = np.load(util.GOLDEN + "/kodim03_uastc_l0.npz")
    img = np.ascontiguousarray(np.concatenate([g["image"][:256, :512], util.synth(512, 5)[:128]], 0))
    _, _, vecs, weights = endpoint_training(ref, img)
    want, want_par = ref_tsvq(ref, vecs, weights, max_codebook, max_parent, 1, True)
    got, got_par, info = ctx.tsvq_generate(vecs, weights, max_codebook, max_parent, 1, True)
    n = vecs.shape[0]
    check_partition(n, got)
    d_ref, d_gpu = distortion(vecs, weights, want), distortion(vecs, weights, got)
    print(f"endpoints {max_codebook}/{max_parent}: {info['num_unique']} unique of {n}, {len(got)} clusters, {info['rounds']} rounds, {info['nodes_split']} splits, "
          f"{info['device_ms']:.2f} ms; distortion ref {d_ref:.6f} gpu {d_gpu:.6f}; identical clusters {same_clusters(got, want)} of {len(want)}")
    assert_same_codebook(got, want)
    assert_same_codebook(got_par, want_par)
    # every child cluster lies inside one parent cluster (frontend.cpp:918-940 verifies this and aborts otherwise)
    if got_par:
        plab = labels(n, got_par)
        for c in got:
            assert np.all(plab[c] == plab[c[0]])


def selector_training(ref, etc):
    keys = np.zeros(etc.shape[0], np.uint32); wts = np.zeros(etc.shape[0], np.uint32)
    ref.lib.ref_selector_training(_ptr(etc), ctypes.c_uint32(etc.shape[0]), 0, _ptr(keys), _ptr(wts))
    vecs = ((keys[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).astype(np.float32)
    return vecs, wts.astype(np.uint64)


@pytest.mark.parametrize("max_codebook,max_parent", [(400, 16), (2731, 32)])
def test_tsvq_selectors_matches_reference(ctx, ref, max_codebook, max_parent):
    g = np.load(util.GOLDEN + "/kodim03_uastc_l0.npz")
    img = np.ascontiguousarray(g["image"][:384, :768])
    blocks, etc, _, _ = endpoint_training(ref, img)
    c5i = np.stack([etc[:, 0] >> 3, etc[:, 1] >> 3, etc[:, 2] >> 3, etc[:, 3] >> 5], -1).astype(np.uint8)
    enc = np.zeros_like(etc)
    ref.lib.ref_etc1s_determine_selectors(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(np.ascontiguousarray(c5i)), _ptr(enc), 0)
    vecs, weights = selector_training(ref, enc)
    want, want_par = ref_tsvq(ref, vecs, weights, max_codebook, max_parent, 1, False)
    got, got_par, info = ctx.tsvq_generate(vecs, weights, max_codebook, max_parent, 1, False)
    n = vecs.shape[0]
    check_partition(n, got)
    d_ref, d_gpu = distortion(vecs, weights, want), distortion(vecs, weights, got)
    print(f"selectors {max_codebook}/{max_parent}: {info['num_unique']} unique of {n}, {len(got)} clusters, {info['rounds']} rounds, {info['nodes_split']} splits, "
          f"{info['device_ms']:.2f} ms; distortion ref {d_ref:.4f} gpu {d_gpu:.4f}; identical clusters {same_clusters(got, want)} of {len(want)}")
    assert_same_codebook(got, want)
    assert_same_codebook(got_par, want_par)


def test_tsvq_small_and_degenerate_sets(ctx, ref):
    rng = np.random.default_rng(3)
    # fewer unique vectors than the codebook; a single vector; two vectors; many duplicates
    for n, dim, cb in [(1, 6, 8), (2, 16, 8), (50, 6, 100), (500, 16, 7)]:
        vecs = rng.integers(0, 4, (n, dim)).astype(np.float32)
        weights = rng.integers(1, 9, n).astype(np.uint64)
        want, _ = ref_tsvq(ref, vecs, weights, cb, 0)
        got, _, _ = ctx.tsvq_generate(vecs, weights, cb, 0)
        check_partition(n, got)
        assert_same_codebook(got, want)


def test_tsvq_two_level_path(ctx, ref):
    """>= 2^18 unique vectors and max_threads > 1: max_threads top clusters, one sub-tree each (enc.h:2112-2214)."""
    rng = np.random.default_rng(11)
    n = 280000
    vecs = rng.integers(0, 4, (n, 16)).astype(np.float32)
    weights = rng.integers(1, 30, n).astype(np.uint64)
    got, got_par, info = ctx.tsvq_generate(vecs, weights, 1024, 32, 8, False)
    assert info["num_unique"] >= 1 << 18
    check_partition(n, got)
    check_partition(n, got_par)
    want, want_par = ref_tsvq(ref, vecs, weights, 1024, 32, 8, False)
    print(f"two-level: {len(got)} clusters / {len(got_par)} parents (reference {len(want)} / {len(want_par)}), {info['rounds']} rounds, {info['device_ms']:.1f} ms")
    print(f"identical clusters {same_clusters(got, want)} of {len(want)}")
    assert_same_codebook(got, want)
    assert_same_codebook(got_par, want_par)


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (2, 64)])
def test_endpoint_clusters_match_cpu_optimizer(ctx, ref, perceptual, comp_level, perms):
    img = util.synth(256, 9)
    blocks = util.image_to_blocks(img)
    ctx.set_pixel_blocks(blocks)
    rng = np.random.default_rng(17)
    n = blocks.shape[0]
    order = rng.permutation(n).astype(np.uint32)
    cuts = np.sort(rng.choice(np.arange(1, n), 60, replace=False))
    clusters = [c for c in np.split(order, cuts)] + [np.zeros(0, np.uint32)]   # ragged sizes plus an empty cluster
    got = ctx.encode_endpoint_clusters(clusters, perceptual, perms)
    for i, c in enumerate(clusters):
        if not len(c):
            continue
        px = np.ascontiguousarray(blocks[c].reshape(-1, 4))
        out4 = np.zeros(4, np.uint8)
        ref.lib.ref_etc1s_encode_cluster.restype = ctypes.c_uint64
        ref.lib.ref_etc1s_encode_cluster(_ptr(px), ctypes.c_uint32(px.shape[0]), perceptual, comp_level, _ptr(out4))
        assert (got[i, 0] >> 3, got[i, 1] >> 3, got[i, 2] >> 3, got[i, 3] >> 5) == tuple(int(v) for v in out4), (i, len(c))


@pytest.mark.parametrize("perceptual", [0, 1])
def test_selector_codebook_matches_reference(ctx, ref, perceptual):
    img = util.synth(256, 21)
    blocks = util.image_to_blocks(img)
    n = blocks.shape[0]
    ctx.set_pixel_blocks(blocks)
    etc = np.zeros((n, 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(n), _ptr(etc), perceptual, 1)
    rng = np.random.default_rng(23)
    order = rng.permutation(n).astype(np.uint32)
    cuts = np.sort(rng.choice(np.arange(1, n), 99, replace=False))
    clusters = [c for c in np.split(order, cuts)] + [np.zeros(0, np.uint32)]
    got = ctx.optimize_selector_codebook(etc, clusters, perceptual)
    off = np.zeros(len(clusters) + 1, np.uint32); np.cumsum([len(c) for c in clusters], out=off[1:])
    idx = np.ascontiguousarray(np.concatenate(clusters), np.uint32)
    want = np.zeros(len(clusters), np.uint32)
    ref.lib.ref_optimize_selector_codebook(_ptr(blocks), _ptr(etc), len(clusters), _ptr(off), _ptr(idx), perceptual, _ptr(want))
    assert np.array_equal(got, want)


def _selector_words(etc):
    """Selector INDEX words (texel (x, y) at bits 2 * (x + 4 * y)) of encoded etc_blocks, as etc_block::get_selector gives them."""
    raw_to_sel = np.array([2, 3, 1, 0], np.uint32)
    msb = (etc[:, 4].astype(np.uint32) << 8) | etc[:, 5]
    lsb = (etc[:, 6].astype(np.uint32) << 8) | etc[:, 7]
    w = np.zeros(etc.shape[0], np.uint32)
    for y in range(4):
        for x in range(4):
            bit = x * 4 + y
            raw = ((lsb >> bit) & 1) | (((msb >> bit) & 1) << 1)
            w |= raw_to_sel[raw] << np.uint32(2 * (x + 4 * y))
    return w


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(2, 64), (6, 165)])
def test_reoptimize_endpoint_clusters_match_reference(ctx, ref, perceptual, comp_level, perms):
    """b200_etc1s_reoptimize_endpoint_clusters against reoptimize_remapped_endpoints' per-cluster body (frontend.cpp:3018-3090):
    ragged clusters incl. an empty one and one large enough for the CTA-per-cluster kernel."""
    img = util.synth(256, 41)
    blocks = util.image_to_blocks(img)
    n = blocks.shape[0]
    ctx.set_pixel_blocks(blocks)
    etc = np.zeros((n, 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(n), _ptr(etc), perceptual, 1)
    sels = _selector_words(etc)
    rng = np.random.default_rng(43)
    rnd = rng.choice(n, n // 10, replace=False)
    sels[rnd] = rng.integers(0, 1 << 32, rnd.shape[0], dtype=np.uint64).astype(np.uint32)
    lum = blocks.reshape(n, 16, 4)[:, :, :3].astype(np.int64).sum(axis=(1, 2))
    order = np.argsort(lum, kind="stable").astype(np.uint32)
    cuts = np.sort(rng.choice(np.arange(1, n - 400), 70, replace=False))
    clusters = [c for c in np.split(order[:n - 400], cuts)] + [np.zeros(0, np.uint32), order[n - 400:]]
    cur = np.zeros((len(clusters), 4), np.uint8)
    for i, c in enumerate(clusters):
        if len(c):
            e = etc[c[len(c) // 2]]
            cur[i] = (e[0] >> 3, e[1] >> 3, e[2] >> 3, e[3] >> 5)
    got4, got_new, got_cur = ctx.reoptimize_endpoint_clusters(clusters, sels, cur, perceptual, perms)
    ref.lib.ref_etc1s_reoptimize_cluster.restype = ctypes.c_uint64
    improved = 0
    for i, c in enumerate(clusters):
        if not len(c):
            assert (int(got_new[i]), int(got_cur[i])) == (0, 0)
            continue
        blk = np.ascontiguousarray(blocks[c]); s = np.ascontiguousarray(sels[c])
        want4 = np.zeros(4, np.uint8); want_cur = ctypes.c_uint64(0)
        want_new = ref.lib.ref_etc1s_reoptimize_cluster(_ptr(blk), ctypes.c_uint32(len(c)), _ptr(s), _ptr(cur[i]), perceptual, comp_level, _ptr(want4), ctypes.byref(want_cur))
        assert (tuple(int(v) for v in got4[i]), int(got_new[i]), int(got_cur[i])) == (tuple(int(v) for v in want4), want_new, want_cur.value), (i, len(c))
        improved += want_new < want_cur.value
    assert improved > 0


def test_endpoint_cluster_past_2_pow_24_follows_the_reference_float_sum(ctx, ref):
    """A 4500-block bright cluster: etc1_optimizer::init's float sum rounds past 2^24 and the CTA-per-cluster kernel follows it
    (the same construction as tests/test_etc1s_cpu.py, where the exact mean is shown to end on another colour)."""
    n = 4500
    rng = np.random.default_rng(76)
    red = rng.integers(248, 255, n * 16)
    red = (red - (rng.random(n * 16) < 0.1135)).astype(np.uint8)
    px = np.empty((n * 16, 4), np.uint8)
    px[:, 0] = red; px[:, 1] = rng.integers(100, 140, n * 16); px[:, 2] = rng.integers(50, 60, n * 16); px[:, 3] = 255
    blocks = np.ascontiguousarray(px.reshape(n, 64))
    ctx.set_pixel_blocks(blocks)
    clusters = [np.arange(n, dtype=np.uint32), np.arange(0, 40, dtype=np.uint32)]
    ref.lib.ref_etc1s_encode_cluster.restype = ctypes.c_uint64
    for perceptual, comp_level, perms in ((1, 1, 16), (0, 2, 64)):
        got = ctx.encode_endpoint_clusters(clusters, perceptual, perms)
        for i, c in enumerate(clusters):
            blk = np.ascontiguousarray(blocks[c])
            out4 = np.zeros(4, np.uint8)
            ref.lib.ref_etc1s_encode_cluster(_ptr(blk), ctypes.c_uint32(len(c) * 16), perceptual, comp_level, _ptr(out4))
            assert (got[i, 0] >> 3, got[i, 1] >> 3, got[i, 2] >> 3, got[i, 3] >> 5) == tuple(int(v) for v in out4), (i, perceptual)


@pytest.mark.parametrize("perceptual", [0, 1])
@pytest.mark.parametrize("comp_level,perms", [(1, 16), (4, 64), (6, 165)])
def test_refit_endpoint_clusters_free_selectors_match_reference(ctx, ref, perceptual, comp_level, perms):
    """generate_endpoint_codebook at step >= 1 (frontend.cpp:1493-1606): new endpoint, its error and the previous endpoint's error."""
    img = util.synth(256, 47)
    blocks = util.image_to_blocks(img)
    n = blocks.shape[0]
    ctx.set_pixel_blocks(blocks)
    etc = np.zeros((n, 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(n), _ptr(etc), perceptual, 1)
    rng = np.random.default_rng(53)
    lum = blocks.reshape(n, 16, 4)[:, :, :3].astype(np.int64).sum(axis=(1, 2))
    order = np.argsort(lum, kind="stable").astype(np.uint32)
    cuts = np.sort(rng.choice(np.arange(1, n - 300), 50, replace=False))
    clusters = [c for c in np.split(order[:n - 300], cuts)] + [order[n - 300:]]
    prev = np.zeros((len(clusters), 4), np.uint8)
    for i, c in enumerate(clusters):
        e = etc[c[0]]
        prev[i] = (e[0] >> 3, e[1] >> 3, e[2] >> 3, e[3] >> 5)
    got4, got_new, got_prev = ctx.reoptimize_endpoint_clusters(clusters, None, prev, perceptual, perms)
    ref.lib.ref_etc1s_refit_cluster.restype = ctypes.c_uint64
    kept = 0
    for i, c in enumerate(clusters):
        blk = np.ascontiguousarray(blocks[c])
        want4 = np.zeros(4, np.uint8); want_prev = ctypes.c_uint64(0)
        want_new = ref.lib.ref_etc1s_refit_cluster(_ptr(blk), ctypes.c_uint32(len(c)), _ptr(prev[i]), perceptual, comp_level, _ptr(want4), ctypes.byref(want_prev))
        assert (tuple(int(v) for v in got4[i]), int(got_new[i]), int(got_prev[i])) == (tuple(int(v) for v in want4), want_new, want_prev.value), (i, len(c))
        kept += want_prev.value <= want_new
    assert kept < len(clusters)


@pytest.mark.parametrize("perceptual", [0, 1])
def test_subblock_errors_match_reference(ctx, ref, perceptual):
    """compute_endpoint_subblock_error_vec's per-subblock error (frontend.cpp:1022-1066), including its unscaled-base behaviour."""
    img = util.synth(192, 59)
    blocks = util.image_to_blocks(img)
    n = blocks.shape[0]
    ctx.set_pixel_blocks(blocks)
    rng = np.random.default_rng(61)
    c5i = np.stack([rng.integers(0, 32, n), rng.integers(0, 32, n), rng.integers(0, 32, n), rng.integers(0, 8, n)], -1).astype(np.uint8)
    got = ctx.subblock_errors(c5i, perceptual)
    want = np.zeros((n, 2), np.uint64)
    ref.lib.ref_etc1s_subblock_errors(_ptr(blocks), ctypes.c_uint32(n), _ptr(c5i), perceptual, _ptr(want))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("perceptual,thresh", [(0, 1.5), (1, 1.5), (0, 0.375), (1, 3.0), (0, 0.0)])
def test_backend_endpoint_prediction_matches_reference_scan(ctx, ref, perceptual, thresh):
    """create_encoder_blocks' endpoint prediction / endpoint RDO (backend.cpp:437-600): the anti-diagonal wavefront on the device against
    the reference's raster scan, on a quantised codebook where neighbours often share or nearly share endpoints. Two slices, odd sizes."""
    img = util.synth(256, 67)
    blocks_all = util.image_to_blocks(img)                       # 64 x 64 blocks
    shapes = [(37, 23), (64, 40)]
    n = sum(w * h for w, h in shapes)
    blocks = np.ascontiguousarray(blocks_all[:n])
    ctx.set_pixel_blocks(blocks)
    etc = np.zeros((n, 8), np.uint8)
    ref.lib.ref_etc1s_encode_blocks(_ptr(blocks), ctypes.c_uint32(n), _ptr(etc), perceptual, 1)
    # a coarse endpoint codebook: colours quantised to 3 bits, so that many blocks share an endpoint and neighbours are near misses
    key = ((etc[:, 0] >> 5).astype(np.uint32) << 9) | ((etc[:, 1] >> 5).astype(np.uint32) << 6) | ((etc[:, 2] >> 5).astype(np.uint32) << 3) | (etc[:, 3] >> 5)
    uniq, idx0 = np.unique(key, return_inverse=True)
    cb = np.stack([((uniq >> 9) & 7) * 4 + 2, ((uniq >> 6) & 7) * 4 + 2, ((uniq >> 3) & 7) * 4 + 2, uniq & 7], -1).astype(np.uint8)
    # the blocks carry their cluster's endpoint, as the frontend's output blocks do
    e = cb[idx0]
    etc[:, 0] = e[:, 0] << 3; etc[:, 1] = e[:, 1] << 3; etc[:, 2] = e[:, 2] << 3; etc[:, 3] = (e[:, 3] << 5) | (e[:, 3] << 2) | 3
    slices, first = [], 0
    for w, h in shapes:
        slices.append((first, w, h)); first += w * h
    got_idx, got_pred = ctx.backend_endpoint_prediction(slices, etc, cb, idx0.astype(np.uint32), thresh, perceptual)
    want_idx = idx0.astype(np.uint32).copy(); want_pred = np.zeros(n, np.uint8)
    for first, w, h in slices:
        wi = np.ascontiguousarray(want_idx[first:first + w * h]); wp = np.zeros(w * h, np.uint8)
        ref.lib.ref_backend_endpoint_prediction(_ptr(np.ascontiguousarray(blocks[first:first + w * h])), _ptr(np.ascontiguousarray(etc[first:first + w * h])), w, h, _ptr(cb),
                                                ctypes.c_float(thresh), perceptual, _ptr(wi), _ptr(wp))
        want_idx[first:first + w * h] = wi; want_pred[first:first + w * h] = wp
    assert np.array_equal(got_pred, want_pred)
    assert np.array_equal(got_idx, want_idx)
    if thresh > 0:
        assert (want_idx != idx0).sum() > 50, "the fixture does not exercise endpoint RDO"
    assert ((want_pred & 3) != 3).sum() > 100


def _index_stream(kind, n_syms, m, seed):
    rng = np.random.default_rng(seed)
    if kind == "walk":          # neighbours mostly close, occasional jumps: what a block stream of endpoint indices looks like
        steps = np.where(rng.random(m) < 0.75, rng.integers(0, 48, m), rng.integers(0, n_syms, m))
        return (np.cumsum(steps) % n_syms).astype(np.uint32)
    if kind == "uniform":
        return rng.integers(0, n_syms, m).astype(np.uint32)
    if kind == "few":           # many symbols never used, long runs of equal indices, heavy ties
        return np.repeat(rng.integers(0, max(n_syms // 50, 2), m // 7 + 1), 7)[:m].astype(np.uint32)
    if kind == "constant":
        return np.full(m, 3 % n_syms, np.uint32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind,n_syms,m", [("walk", 2400, 16000), ("walk", 8080, 300000), ("uniform", 500, 20000), ("few", 3000, 50000),
                                            ("constant", 40, 1000), ("uniform", 2, 50), ("walk", 1, 10), ("uniform", 700, 2)])
def test_palette_reorder_matches_reference(ctx, ref, kind, n_syms, m):
    """b200_palette_reorder against palette_index_reorderer::init (enc.cpp:1785-1915): identical remap table, including the degenerate
    inputs (no two different indices adjacent; one symbol; two indices)."""
    idx = _index_stream(kind, n_syms, m, n_syms + m)
    want = np.zeros(n_syms, np.uint32)
    ref.lib.ref_palette_reorder(ctypes.c_uint32(m), _ptr(idx), ctypes.c_uint32(n_syms), _ptr(want))
    got = ctx.palette_reorder(idx, n_syms)
    assert np.array_equal(got, want), int((got != want).sum())
