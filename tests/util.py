"""Shared test helpers: synthetic inputs, 4x4 block marshalling, ctypes loaders for the checkers.

The reference .so (oracle/_ref/libbasisu_ref.so) and the host-emulation .so (tests/hostemu/_build) are TEST
infrastructure; the product package never imports this module.
"""
import ctypes
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbasisu_ref.so")
EMU_SO = os.path.join(ROOT, "tests", "hostemu", "_build", "libbu_hostemu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def synth(n, seed, h=None):
    """SURVEY.md 9.6 generator: n x n (or h x n) RGBA8 image."""
    h = n if h is None else h
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:n].astype(np.float32)
    r = 127.5 + 100 * np.sin(x / 97.0) * np.cos(y / 131.0)
    g = 127.5 + 100 * np.sin((x + y) / 61.0)
    b = 127.5 + 100 * np.cos(x / 23.0) * np.sin(y / 17.0)
    rgb = np.stack([r, g, b], -1) + rng.normal(0, 12, size=(h, n, 3)).astype(np.float32)
    a = 255 * ((np.sin(x / 257.0) * np.sin(y / 311.0)) > -0.3).astype(np.float32)
    a = np.where(((x // 64 + y // 64) % 7) == 0, 128 + 100 * np.sin(x / 11.0), a)
    return np.concatenate([rgb, a[..., None]], -1).clip(0, 255).astype(np.uint8)


def image_to_blocks(img):
    """(H, W, 4) uint8, H and W multiples of 4 -> (nblocks, 64) uint8 in raster block order, [y][x] RGBA per block
    (pixel_block layout, encoder/basisu_enc.h:4156)."""
    h, w, _ = img.shape
    assert h % 4 == 0 and w % 4 == 0
    return np.ascontiguousarray(img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 64))


def edge_case_blocks(seed=7):
    """Blocks that hit the special paths: solid, solid-but-alpha, LA, LA opaque, two-colour, gradients, random, 1-texel outliers."""
    rng = np.random.default_rng(seed)
    out = []

    def blk(f):
        b = np.zeros((4, 4, 4), np.uint8)
        for y in range(4):
            for x in range(4):
                b[y, x] = f(x, y)
        out.append(b.reshape(64))

    for c in [(0, 0, 0, 255), (255, 255, 255, 255), (0, 0, 0, 0), (13, 200, 77, 255), (13, 200, 77, 9), (1, 1, 1, 255), (254, 255, 0, 128)]:
        blk(lambda x, y: c)
    blk(lambda x, y: (x * 60, x * 60, x * 60, 255))                      # opaque luminance ramp
    blk(lambda x, y: (x * 60, x * 60, x * 60, 255 - y * 70))             # LA block
    blk(lambda x, y: (40, 40, 40, 255 if (x + y) & 1 else 0))            # LA, binary alpha
    blk(lambda x, y: (255, 0, 0, 255) if x < 2 else (0, 0, 255, 255))    # two colours, partition friendly
    blk(lambda x, y: (255, 0, 0, 255) if y < 2 else (0, 255, 0, 255))
    blk(lambda x, y: (x * 80, y * 80, 128, 255))
    blk(lambda x, y: (x * 80, y * 80, 128, 64 + x * 40))
    blk(lambda x, y: (200, 200, 200, 255) if (x, y) != (3, 3) else (0, 0, 0, 255))
    blk(lambda x, y: (200, 10, 30, 255) if (x, y) != (0, 0) else (200, 10, 30, 254))
    blk(lambda x, y: (x * 85, x * 85, x * 85, x * 85))
    blk(lambda x, y: (17 * (x + 4 * y) % 256, 255 - 17 * (x + 4 * y) % 256, (x * y * 20) % 256, 255))
    for _ in range(24):
        out.append(rng.integers(0, 256, 64, dtype=np.uint8))
    for _ in range(8):   # random opaque
        b = rng.integers(0, 256, (16, 4), dtype=np.uint8); b[:, 3] = 255; out.append(b.reshape(64))
    for _ in range(8):   # low-variance noise around a colour
        base = rng.integers(20, 235, 4); b = (base + rng.integers(-6, 7, (16, 4))).clip(0, 255).astype(np.uint8); b[:, 3] = 255; out.append(b.reshape(64))
    for _ in range(8):   # random LA
        l = rng.integers(0, 256, 16, dtype=np.uint8); a = rng.integers(0, 256, 16, dtype=np.uint8)
        out.append(np.stack([l, l, l, a], -1).reshape(64))
    return np.ascontiguousarray(np.stack(out))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def build_ref():
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/encoder"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(REF_SO)


def build_emu():
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    deps = [src] + [os.path.join(ROOT, "basis_universal_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "basis_universal_b200", "csrc")) if f.startswith("bu_") or f.endswith(".inc")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        subprocess.check_call([os.path.join(ROOT, "tests", "hostemu", "build.sh")])
    return os.path.exists(EMU_SO)


class Ref:
    """ctypes view of oracle/_ref/libbasisu_ref.so (the compiled, unmodified reference)."""

    def __init__(self):
        self.lib = ctypes.CDLL(REF_SO)
        self.lib.ref_init()
        self.lib.ref_color_cell_compression.restype = ctypes.c_uint64
        self.lib.ref_compress_image.restype = ctypes.c_void_p
        self.lib.ref_compress_image.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
        self.lib.ref_free.argtypes = [ctypes.c_void_p]
        self.lib.ref_uastc_rdo.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]

    def encode_uastc(self, blocks, flags, threads=8):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((blocks.shape[0], 16), np.uint8)
        self.lib.ref_encode_uastc_blocks(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(out), ctypes.c_uint32(flags), ctypes.c_uint32(threads))
        return out

    def unpack_uastc(self, ublocks):
        ublocks = np.ascontiguousarray(ublocks, np.uint8)
        out = np.zeros((ublocks.shape[0], 64), np.uint8)
        ok = self.lib.ref_unpack_uastc_blocks(_ptr(ublocks), ctypes.c_uint32(ublocks.shape[0]), _ptr(out))
        assert ok
        return out

    def ccc(self, px, wtab, rng, has_alpha, uber, ls):
        px = np.ascontiguousarray(px, np.uint8)
        lo = np.zeros(4, np.uint8); hi = np.zeros(4, np.uint8); sel = np.zeros(16, np.uint8)
        err = self.lib.ref_color_cell_compression(_ptr(px), ctypes.c_uint32(px.shape[0]), ctypes.c_uint32(wtab), ctypes.c_uint32(rng), ctypes.c_uint32(has_alpha), ctypes.c_uint32(uber), ctypes.c_uint32(ls), _ptr(lo), _ptr(hi), _ptr(sel))
        return err, lo, hi, sel[:px.shape[0]]

    def compress_image(self, fmt, img, flags_and_quality, rdo_quality=0.0):
        img = np.ascontiguousarray(img, np.uint8)
        size = ctypes.c_size_t(0)
        p = self.lib.ref_compress_image(fmt, _ptr(img), img.shape[1], img.shape[0], flags_and_quality, ctypes.c_float(rdo_quality), ctypes.byref(size))
        assert p
        data = ctypes.string_at(p, size.value)
        self.lib.ref_free(p)
        return data


class Emu:
    """ctypes view of the host-emulation build of the device code."""

    def __init__(self):
        self.lib = ctypes.CDLL(EMU_SO)
        self.lib.emu_color_cell_compression.restype = ctypes.c_uint64

    def encode_uastc(self, blocks, flags, threads=8):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((blocks.shape[0], 16), np.uint8)
        self.lib.emu_encode_uastc_blocks(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(out), ctypes.c_uint32(flags), ctypes.c_uint32(threads))
        return out

    def ccc(self, px, wtab, rng, has_alpha, uber, ls):
        px = np.ascontiguousarray(px, np.uint8)
        lo = np.zeros(4, np.uint8); hi = np.zeros(4, np.uint8); sel = np.zeros(16, np.uint8)
        err = self.lib.emu_color_cell_compression(_ptr(px), ctypes.c_uint32(px.shape[0]), ctypes.c_uint32(wtab), ctypes.c_uint32(rng), ctypes.c_uint32(has_alpha), ctypes.c_uint32(uber), ctypes.c_uint32(ls), _ptr(lo), _ptr(hi), _ptr(sel))
        return err, lo, hi, sel[:px.shape[0]]


def uastc_fields(ref, block16):
    """Human-readable field dump of one packed UASTC block via the reference unpacker (diagnostics)."""
    out = np.zeros(61, np.uint8)
    b = np.ascontiguousarray(block16, np.uint8)
    ok = ref.lib.ref_uastc_fields(_ptr(b), _ptr(out))
    if not ok:
        return None
    names = ["mode", "pattern", "bc1h0", "bc1h1", "flip", "diff", "inten0", "inten1", "bias", "etc2", "ccs"]
    d = {n: int(out[i]) for i, n in enumerate(names)}
    d["ep"] = out[11:29].tolist()
    d["w"] = out[29:61].tolist()
    return d


OCL_SO = os.path.join(ROOT, "oracle", "_ref", "libocl_ref.so")


class OclRef:
    """The reference's own OpenCL C kernels (bin/ocl_kernels.cl) compiled for the host by oracle/Makefile (see oracle/ocl_host.cpp)."""

    def __init__(self):
        self.lib = ctypes.CDLL(OCL_SO)
        self.threads = os.cpu_count() or 1

    def encode_etc1s_blocks(self, blocks, perceptual, total_perms):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((blocks.shape[0], 8), np.uint8)
        self.lib.oclref_encode_etc1s_blocks(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(out), int(perceptual), int(total_perms), ctypes.c_uint32(self.threads))
        return out

    def encode_pixel_clusters(self, clusters, pixels, weights, perceptual, total_perms):
        clusters = np.ascontiguousarray(clusters, np.uint64)  # (n, 2): total_pixels, first_pixel_index
        pixels = np.ascontiguousarray(pixels, np.uint8); weights = np.ascontiguousarray(weights, np.uint32)
        out = np.zeros((clusters.shape[0], 8), np.uint8)
        self.lib.oclref_encode_etc1s_pixel_clusters(_ptr(clusters), ctypes.c_uint32(clusters.shape[0]), _ptr(pixels), _ptr(weights), _ptr(out), int(perceptual), int(total_perms), ctypes.c_uint32(self.threads))
        return out

    def refine(self, blocks, block_info, cluster_info, sorted_idx, perceptual):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros(blocks.shape[0], np.uint32)
        self.lib.oclref_refine_endpoint_clusterization(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(block_info), _ptr(cluster_info), _ptr(sorted_idx), _ptr(out), int(perceptual), ctypes.c_uint32(self.threads))
        return out

    def fosc(self, blocks, block_info, selectors, cluster_indices, perceptual):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros(blocks.shape[0], np.uint32)
        self.lib.oclref_find_optimal_selector_clusters_for_each_block(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(block_info), _ptr(selectors), _ptr(cluster_indices), _ptr(out), int(perceptual), ctypes.c_uint32(self.threads))
        return out

    def determine_selectors(self, blocks, color5_inten, perceptual):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((blocks.shape[0], 8), np.uint8)
        self.lib.oclref_determine_selectors(_ptr(blocks), ctypes.c_uint32(blocks.shape[0]), _ptr(color5_inten), _ptr(out), int(perceptual), ctypes.c_uint32(self.threads))
        return out


BLOCK_INFO_DT = np.dtype([("first_cluster_ofs", "<u2"), ("num_clusters", "<u2"), ("cur_cluster_index", "<u2"), ("cur_cluster_etc_inten", "u1")])          # cl_block_info_struct, 7 B
ENDPOINT_CLUSTER_DT = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1"), ("etc_inten", "u1"), ("cluster_index", "<u2")])                       # cl_endpoint_cluster_struct, 7 B
FOSC_BLOCK_DT = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("inten", "u1"), ("first_selector", "<u4"), ("num_selectors", "<u4")])                  # fosc_block_struct, 12 B


def etc1s_stage_inputs(blocks, seed, parents=8, clusters_per_parent=(3, 40), selectors_per_parent=(5, 300)):
    """Synthetic but structurally faithful inputs for the refine / fosc / determine_selectors stages."""
    rng = np.random.default_rng(seed)
    n = blocks.shape[0]
    sizes = rng.integers(clusters_per_parent[0], clusters_per_parent[1], parents)
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    total = int(sizes.sum())
    ci = np.zeros(total, ENDPOINT_CLUSTER_DT)
    mean = blocks.reshape(n, 16, 4)[:, :, :3].mean(1)
    pick = rng.integers(0, n, total)
    base = np.clip(np.round(mean[pick] * 31 / 255 + rng.integers(-2, 3, (total, 3))), 0, 31).astype(np.uint8)
    ci["r"], ci["g"], ci["b"], ci["a"] = base[:, 0], base[:, 1], base[:, 2], 255
    ci["etc_inten"] = rng.integers(0, 8, total)
    ci["cluster_index"] = rng.permutation(total).astype(np.uint16)
    bi = np.zeros(n, BLOCK_INFO_DT)
    par = rng.integers(0, parents, n)
    bi["first_cluster_ofs"] = first[par]; bi["num_clusters"] = sizes[par]
    own = first[par] + rng.integers(0, 1 << 30, n) % sizes[par]
    bi["cur_cluster_index"] = ci["cluster_index"][own]; bi["cur_cluster_etc_inten"] = ci["etc_inten"][own]
    sorted_idx = np.argsort(bi["cur_cluster_index"], kind="stable").astype(np.uint32)
    ssz = rng.integers(selectors_per_parent[0], selectors_per_parent[1], parents)
    sfirst = np.concatenate([[0], np.cumsum(ssz)[:-1]])
    stotal = int(ssz.sum())
    sels = rng.integers(0, 1 << 32, stotal, dtype=np.uint64).astype(np.uint32)
    sels[rng.integers(0, stotal, stotal // 4)] = np.uint32(0xAAAAAAAA)
    sel_cluster_idx = rng.permutation(stotal).astype(np.uint32)
    fb = np.zeros(n, FOSC_BLOCK_DT)
    c5 = np.clip(np.round(mean * 31 / 255), 0, 31).astype(np.uint8)
    fb["r"], fb["g"], fb["b"] = c5[:, 0], c5[:, 1], c5[:, 2]
    fb["inten"] = rng.integers(0, 8, n)
    spar = rng.integers(0, parents, n)
    fb["first_selector"] = sfirst[spar]; fb["num_selectors"] = ssz[spar]
    color5_inten = np.stack([c5[:, 0], c5[:, 1], c5[:, 2], fb["inten"]], -1).astype(np.uint8)
    return dict(block_info=bi, cluster_info=ci, sorted_idx=sorted_idx, fosc_blocks=fb, selectors=sels, sel_cluster_idx=sel_cluster_idx, color5_inten=np.ascontiguousarray(color5_inten))


def pixel_cluster_inputs(seed, n_clusters=40):
    rng = np.random.default_rng(seed)
    sizes = np.concatenate([rng.integers(1, 40, n_clusters // 2), rng.integers(40, 3000, n_clusters - n_clusters // 2)])
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    total = int(sizes.sum())
    px = np.zeros((total, 4), np.uint8)
    for s, f in zip(sizes, first):
        base = rng.integers(0, 256, 3)
        spread = rng.integers(1, 80)
        px[f:f + s, :3] = np.clip(base + rng.integers(-spread, spread + 1, (s, 3)), 0, 255)
    px[:, 3] = 255
    weights = rng.integers(1, 50, total).astype(np.uint32)
    clusters = np.stack([sizes, first], -1).astype(np.uint64)
    return clusters, px, weights


class RdoParams(ctypes.Structure):
    """b200_uastc_rdo_params / basisu::uastc_rdo_params with the reference defaults."""
    _fields_ = [("lz_dict_size", ctypes.c_uint32), ("lam", ctypes.c_float), ("ratio", ctypes.c_float), ("skip", ctypes.c_float), ("refine", ctypes.c_uint32),
                ("sd", ctypes.c_float), ("scale", ctypes.c_float), ("lit", ctypes.c_uint32)]


def ref_rdo(ref, uastc, src, lam, flags, jobs):
    out = np.ascontiguousarray(uastc).copy()
    ok = ref.lib.ref_uastc_rdo(len(out), _ptr(out), _ptr(np.ascontiguousarray(src)), lam, 4096, 10.0, 8.0, 18.0, 10.0, flags, jobs, max(jobs, 4))
    assert ok
    return out


def emu_rdo(emu, uastc, src, lam, flags, jobs):
    out = np.ascontiguousarray(uastc).copy()
    p = RdoParams(4096, lam, 10.0, 8.0, 1, 18.0, 10.0, 100)
    ok = emu.lib.emu_uastc_rdo(len(out), _ptr(out), _ptr(np.ascontiguousarray(src)), ctypes.byref(p), flags, jobs)
    assert ok
    return out


def rdo_test_image():
    """A crop of a natural image (RDO modifies ~half of its blocks) tiled with part of the synthetic image (alpha modes)."""
    g = np.load(os.path.join(GOLDEN, "kodim03_uastc_l0.npz"))
    return np.ascontiguousarray(np.concatenate([g["image"][:128, :256], synth(256, 77)[:64]], 0))


def endpoint_keys(etc_blocks):
    """numpy restatement of k_etc1s_endpoint_histogram's key extraction (r5<<13 | g5<<8 | b5<<3 | inten from bytes 0..3 of an
    ETC1S etc_block): the checker for the histogram kernel and the stand-in for it in the CPU-only gloo test."""
    b = np.asarray(etc_blocks, np.uint8).astype(np.uint32)
    return ((b[:, 0] >> 3) << 13) | ((b[:, 1] >> 3) << 8) | ((b[:, 2] >> 3) << 3) | (b[:, 3] >> 5)
