"""Host-side mirror of the reference's ETC1S GPU seam (encoder/basisu_opencl.h:46-141): one method per `opencl_*` entry
point, same argument meaning, same failure behaviour (a False/exception tells the caller to use its own CPU path).

Packed argument structs as numpy dtypes (opencl.h:49-135):"""
import ctypes
import numpy as np
from ._lib import lib, B200Error

pixel_cluster_dtype = np.dtype([("total_pixels", "<u8"), ("first_pixel_index", "<u8")])                                                           # cl_pixel_cluster
block_info_dtype = np.dtype([("first_cluster_ofs", "<u2"), ("num_clusters", "<u2"), ("cur_cluster_index", "<u2"), ("cur_cluster_etc_inten", "u1")])  # cl_block_info_struct
endpoint_cluster_dtype = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1"), ("etc_inten", "u1"), ("cluster_index", "<u2")])               # cl_endpoint_cluster_struct
fosc_block_dtype = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("inten", "u1"), ("first_selector", "<u4"), ("num_selectors", "<u4")])          # fosc_block_struct

OPENCL_ENCODE_ETC1S_MAX_PERMS = 165  # opencl.h:44
FLAVOUR_OPENCL_KERNELS, FLAVOUR_CPU_OPTIMIZER = 0, 1


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


class Etc1sContext:
    """opencl_context equivalent: one device, one stream, the slice's source blocks resident in HBM."""

    def __init__(self, device=0):
        self._lib = lib()
        self.device = int(device)
        self._ctx = self._lib.b200_create_context(int(device))
        if not self._ctx:
            raise B200Error("b200_create_context failed: " + self._lib.b200_last_error(None).decode())
        self.total_blocks = 0

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.b200_destroy_context(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, ok, what):
        if not ok:
            raise B200Error(f"{what} failed: {self._lib.b200_last_error(self._ctx).decode()}")

    @property
    def last_kernel_ms(self):
        return float(self._lib.b200_last_kernel_ms(self._ctx))

    def set_flavour(self, flavour):
        """FLAVOUR_CPU_OPTIMIZER (default) or FLAVOUR_OPENCL_KERNELS: which of the reference's two optimiser implementations to reproduce."""
        self._check(self._lib.b200_etc1s_set_flavour(self._ctx, int(flavour)), "b200_etc1s_set_flavour")

    def set_pixel_blocks(self, blocks):
        """opencl_set_pixel_blocks(ctx, total_blocks, pPixel_blocks)"""
        blocks = np.ascontiguousarray(blocks, np.uint8)
        assert blocks.ndim == 2 and blocks.shape[1] == 64
        self._check(self._lib.b200_etc1s_set_pixel_blocks(self._ctx, blocks.shape[0], _p(blocks)), "b200_etc1s_set_pixel_blocks")
        self.total_blocks = blocks.shape[0]

    def encode_etc1s_blocks(self, perceptual, total_perms):
        """opencl_encode_etc1s_blocks(ctx, pOutput_blocks, perceptual, total_perms) -> (N, 8) uint8 etc_blocks"""
        out = np.empty((self.total_blocks, 8), np.uint8)
        self._check(self._lib.b200_etc1s_encode_blocks(self._ctx, _p(out), int(bool(perceptual)), int(total_perms)), "b200_etc1s_encode_blocks")
        return out

    def encode_etc1s_pixel_clusters(self, clusters, pixels, weights, perceptual, total_perms):
        """opencl_encode_etc1s_pixel_clusters(...) -> (total_clusters, 8) uint8; only bytes 0..3 (colour, intensity, flags) are defined."""
        clusters = np.ascontiguousarray(clusters); pixels = np.ascontiguousarray(pixels, np.uint8); weights = np.ascontiguousarray(weights, np.uint32)
        n = clusters.shape[0]
        out = np.zeros((n, 8), np.uint8)
        self._check(self._lib.b200_etc1s_encode_pixel_clusters(self._ctx, _p(out), n, _p(clusters), ctypes.c_uint64(pixels.shape[0]), _p(pixels), _p(weights),
                                                                int(bool(perceptual)), int(total_perms)), "b200_etc1s_encode_pixel_clusters")
        return out

    def refine_endpoint_clusterization(self, block_info, cluster_info, sorted_block_indices, perceptual):
        """opencl_refine_endpoint_clusterization(...) -> (N,) uint32 best cluster index per block"""
        block_info = np.ascontiguousarray(block_info); cluster_info = np.ascontiguousarray(cluster_info)
        sorted_block_indices = np.ascontiguousarray(sorted_block_indices, np.uint32)
        out = np.empty(self.total_blocks, np.uint32)
        self._check(self._lib.b200_etc1s_refine_endpoint_clusterization(self._ctx, _p(block_info), cluster_info.shape[0], _p(cluster_info), _p(sorted_block_indices), _p(out),
                                                                         int(bool(perceptual))), "b200_etc1s_refine_endpoint_clusterization")
        return out

    def find_optimal_selector_clusters_for_each_block(self, block_info, selectors, selector_cluster_indices, perceptual):
        """opencl_find_optimal_selector_clusters_for_each_block(...) -> (N,) uint32 selector cluster index per block"""
        block_info = np.ascontiguousarray(block_info); selectors = np.ascontiguousarray(selectors, np.uint32)
        selector_cluster_indices = np.ascontiguousarray(selector_cluster_indices, np.uint32)
        out = np.empty(self.total_blocks, np.uint32)
        self._check(self._lib.b200_etc1s_find_optimal_selector_clusters_for_each_block(self._ctx, _p(block_info), selectors.shape[0], _p(selectors), _p(selector_cluster_indices),
                                                                                        _p(out), int(bool(perceptual))), "b200_etc1s_find_optimal_selector_clusters_for_each_block")
        return out

    def determine_selectors(self, color5_and_inten, perceptual):
        """opencl_determine_selectors(ctx, pInput_etc_color5_and_inten, pOutput_blocks, perceptual) -> (N, 8) uint8 etc_blocks"""
        c = np.ascontiguousarray(color5_and_inten, np.uint8)
        out = np.empty((self.total_blocks, 8), np.uint8)
        self._check(self._lib.b200_etc1s_determine_selectors(self._ctx, _p(c), _p(out), int(bool(perceptual))), "b200_etc1s_determine_selectors")
        return out


    def endpoint_histogram(self, etc_blocks):
        """(N, 8) uint8 ETC1S blocks (host) -> (2**18,) uint32 counts per endpoint training key (two per block)."""
        b = np.ascontiguousarray(etc_blocks, np.uint8)
        hist = np.empty(1 << 18, np.uint32)
        self._check(self._lib.b200_etc1s_endpoint_histogram(self._ctx, _p(b), b.shape[0], _p(hist)), "b200_etc1s_endpoint_histogram")
        return hist

    def endpoint_histogram_device(self, d_etc_blocks_ptr, num_blocks, d_hist_ptr):
        """Device form: accumulates into a caller-zeroed device buffer of 2**18 uint32 (ready for an NCCL all-reduce)."""
        ok = self._lib.b200_etc1s_endpoint_histogram_device(self._ctx, ctypes.c_void_p(d_etc_blocks_ptr), int(num_blocks), ctypes.c_void_p(d_hist_ptr))
        self._check(ok, "b200_etc1s_endpoint_histogram_device")


    def selector_training(self, etc_blocks, perceptual):
        """(N, 8) uint8 ETC1S blocks (host) -> (keys (N,) uint32, weights (N,) uint32): generate_selector_clusters' per-block part."""
        b = np.ascontiguousarray(etc_blocks, np.uint8)
        keys = np.empty(b.shape[0], np.uint32)
        weights = np.empty(b.shape[0], np.uint32)
        self._check(self._lib.b200_etc1s_selector_training(self._ctx, _p(b), b.shape[0], int(bool(perceptual)), _p(keys), _p(weights)), "b200_etc1s_selector_training")
        return keys, weights


class TsvqResult(ctypes.Structure):
    """b200_tsvq_result (include/basisu_b200.h)"""
    _fields_ = [("num_unique", ctypes.c_uint32), ("num_clusters", ctypes.c_uint32), ("cluster_offsets", ctypes.POINTER(ctypes.c_uint32)),
                ("cluster_indices", ctypes.POINTER(ctypes.c_uint32)), ("num_parent_clusters", ctypes.c_uint32), ("parent_offsets", ctypes.POINTER(ctypes.c_uint32)),
                ("parent_indices", ctypes.POINTER(ctypes.c_uint32)), ("rounds", ctypes.c_uint32), ("nodes_split", ctypes.c_uint32)]


def _csr(clusters):
    """list of index arrays -> (offsets (n+1,) uint32, indices uint32)"""
    sizes = np.array([len(c) for c in clusters], np.uint32)
    off = np.zeros(len(clusters) + 1, np.uint32)
    np.cumsum(sizes, out=off[1:])
    idx = np.concatenate([np.asarray(c, np.uint32) for c in clusters]) if len(clusters) and off[-1] else np.zeros(0, np.uint32)
    return off, np.ascontiguousarray(idx, np.uint32)


def _tsvq_generate(self, vecs, weights, max_codebook_size, max_parent_codebook_size=0, max_threads=1, even_odd_input_pairs_equal=False):
    """generate_hierarchical_codebook_threaded (enc.h:2219) on the device. vecs: (N, 6) or (N, 16) float32, weights: (N,) uint64.
    Returns (clusters, parent_clusters, info): lists of uint32 arrays of training-vector indices in the reference's order."""
    vecs = np.ascontiguousarray(vecs, np.float32)
    n, dim = vecs.shape
    rec = np.zeros(n, np.dtype([("v", "<f4", (dim,)), ("w", "<u8")], align=True))  # the layout of std::pair<vecNF, uint64_t>
    rec["v"] = vecs
    rec["w"] = np.asarray(weights, np.uint64)
    res = TsvqResult()
    ok = self._lib.b200_tsvq_generate(self._ctx, dim, n, _p(rec), ctypes.c_size_t(rec.dtype.itemsize), ctypes.c_size_t(rec.dtype.fields["w"][1]),
                                      int(max_codebook_size), int(max_parent_codebook_size), int(max_threads), int(bool(even_odd_input_pairs_equal)), ctypes.byref(res))
    self._check(ok, "b200_tsvq_generate")

    def unpack(count, off_p, idx_p):
        if not count:
            return []
        off = np.ctypeslib.as_array(off_p, (count + 1,)).copy()
        idx = np.ctypeslib.as_array(idx_p, (int(off[-1]),)).copy()
        return [idx[off[i]:off[i + 1]] for i in range(count)]
    info = {"num_unique": res.num_unique, "rounds": res.rounds, "nodes_split": res.nodes_split, "device_ms": self.last_kernel_ms}
    return unpack(res.num_clusters, res.cluster_offsets, res.cluster_indices), unpack(res.num_parent_clusters, res.parent_offsets, res.parent_indices), info


def _encode_endpoint_clusters(self, clusters, perceptual, total_perms):
    """generate_endpoint_codebook step 0 (frontend.cpp:1214): clusters = list of block-index arrays -> (n, 8) uint8 etc_blocks (colour + intensity table)."""
    off, idx = _csr(clusters)
    out = np.zeros((len(clusters), 8), np.uint8)
    self._check(self._lib.b200_etc1s_encode_endpoint_clusters(self._ctx, _p(out), len(clusters), _p(off), _p(idx), int(bool(perceptual)), int(total_perms)),
                "b200_etc1s_encode_endpoint_clusters")
    return out


def _optimize_selector_codebook(self, etc_blocks, clusters, perceptual):
    """create_optimized_selector_codebook (frontend.cpp:2259): -> (n,) uint32, texel (x, y) of cluster c at bits 2 * (x + 4 * y)."""
    b = np.ascontiguousarray(etc_blocks, np.uint8)
    assert b.shape == (self.total_blocks, 8)
    off, idx = _csr(clusters)
    out = np.zeros(len(clusters), np.uint32)
    self._check(self._lib.b200_etc1s_optimize_selector_codebook(self._ctx, _p(b), len(clusters), _p(off), _p(idx), _p(out), int(bool(perceptual))),
                "b200_etc1s_optimize_selector_codebook")
    return out


def _reoptimize_endpoint_clusters(self, clusters, block_selectors, cluster_color5_inten, perceptual, total_perms):
    """reoptimize_remapped_endpoints' per-cluster loop (frontend.cpp:3008-3090). clusters = list of block-index arrays;
    block_selectors = one packed selector word per block OF THE SLICE (texel (x, y) at bits 2 * (x + 4 * y)); cluster_color5_inten
    = (n, 4) uint8 current endpoints. -> (new (n, 4) uint8, new_err (n,) uint64, cur_err (n,) uint64)."""
    off, idx = _csr(clusters)
    if block_selectors is None:   # free selectors: generate_endpoint_codebook at step >= 1 (frontend.cpp:1493-1606)
        sels = None
    else:
        sels = np.ascontiguousarray(np.asarray(block_selectors, np.uint32)[idx]) if idx.shape[0] else np.zeros(1, np.uint32)
    cur = np.ascontiguousarray(cluster_color5_inten, np.uint8)
    assert cur.shape == (len(clusters), 4)
    out = np.zeros((len(clusters), 4), np.uint8)
    new_err = np.zeros(len(clusters), np.uint64)
    cur_err = np.zeros(len(clusters), np.uint64)
    self._check(self._lib.b200_etc1s_reoptimize_endpoint_clusters(self._ctx, len(clusters), _p(off), _p(idx), _p(sels) if sels is not None else None, _p(cur), _p(out), _p(new_err), _p(cur_err),
                                                                    int(bool(perceptual)), int(total_perms)), "b200_etc1s_reoptimize_endpoint_clusters")
    return out, new_err, cur_err


def _subblock_errors(self, block_color5_inten, perceptual):
    """compute_endpoint_subblock_error_vec (frontend.cpp:1006): (n, 4) uint8 endpoint per block -> (n, 2) uint64 subblock errors."""
    e = np.ascontiguousarray(block_color5_inten, np.uint8)
    assert e.shape == (self.total_blocks, 4)
    out = np.zeros((self.total_blocks, 2), np.uint64)
    self._check(self._lib.b200_etc1s_subblock_errors(self._ctx, _p(e), _p(out), int(bool(perceptual))), "b200_etc1s_subblock_errors")
    return out


def _backend_endpoint_prediction(self, slices, etc_blocks, endpoint_color5_inten, block_endpoint_indices, thresh, perceptual):
    """basisu_backend::create_encoder_blocks' endpoint prediction / endpoint RDO (backend.cpp:405-617). slices: list of
    (first_block, blocks_x, blocks_y); -> (endpoint index per block (n,) uint32, predictor per block (n,) uint8; 3 = none)."""
    sl = np.ascontiguousarray(np.asarray(slices, np.uint32).reshape(-1, 3))
    etc = np.ascontiguousarray(etc_blocks, np.uint8)
    assert etc.shape == (self.total_blocks, 8)
    cb = np.ascontiguousarray(endpoint_color5_inten, np.uint8)
    assert cb.ndim == 2 and cb.shape[1] == 4
    idx = np.array(block_endpoint_indices, np.uint32, copy=True)
    assert idx.shape == (self.total_blocks,)
    pred = np.zeros(self.total_blocks, np.uint8)
    self._check(self._lib.b200_etc1s_backend_endpoint_prediction(self._ctx, sl.shape[0], _p(sl), _p(etc), cb.shape[0], _p(cb), ctypes.c_float(thresh), int(bool(perceptual)),
                                                                   _p(idx), _p(pred)), "b200_etc1s_backend_endpoint_prediction")
    return idx, pred


def _palette_reorder(self, indices, num_syms):
    """palette_index_reorderer::init + get_remap_table (enc.cpp:1785), no distance function: index stream -> (num_syms,) uint32 old -> new."""
    idx = np.ascontiguousarray(indices, np.uint32)
    out = np.zeros(int(num_syms), np.uint32)
    self._check(self._lib.b200_palette_reorder(self._ctx, idx.shape[0], _p(idx), int(num_syms), _p(out)), "b200_palette_reorder")
    return out


def comm_unique_id():
    """128-byte NCCL id (rank 0 calls this, then broadcasts the bytes to the other ranks)."""
    buf = np.zeros(128, np.uint8)
    if not lib().b200_comm_unique_id(_p(buf)):
        raise B200Error("b200_comm_unique_id failed: " + lib().b200_comm_last_error().decode())
    return buf


def _comm_init(self, rank, world, unique_id):
    """Attaches the NCCL communicator: afterwards every stage call computes this rank's share and merges with the other ranks."""
    uid = np.ascontiguousarray(unique_id, np.uint8)
    assert uid.shape == (128,)
    self._check(self._lib.b200_comm_init(self._ctx, int(rank), int(world), _p(uid)), "b200_comm_init")


def _comm_stats(self):
    ms, by, ca = ctypes.c_float(0), ctypes.c_uint64(0), ctypes.c_uint32(0)
    self._lib.b200_comm_stats(self._ctx, ctypes.byref(ms), ctypes.byref(by), ctypes.byref(ca))
    return {"ms": float(ms.value), "bytes": int(by.value), "calls": int(ca.value)}


Etc1sContext.comm_init = _comm_init
Etc1sContext.comm_stats = _comm_stats
Etc1sContext.tsvq_generate = _tsvq_generate
Etc1sContext.encode_endpoint_clusters = _encode_endpoint_clusters
Etc1sContext.optimize_selector_codebook = _optimize_selector_codebook
Etc1sContext.reoptimize_endpoint_clusters = _reoptimize_endpoint_clusters
Etc1sContext.subblock_errors = _subblock_errors
Etc1sContext.backend_endpoint_prediction = _backend_endpoint_prediction
Etc1sContext.palette_reorder = _palette_reorder


def merge_selector_training(keys, weights):
    """Unique keys in ascending order with summed weights: what the clusterer's duplicate merge leaves (enc.h:2228-2260)."""
    u, inv = np.unique(np.asarray(keys, np.uint32), return_inverse=True)
    w = np.zeros(u.shape[0], np.uint64)
    np.add.at(w, inv, np.asarray(weights, np.uint64))
    return u, w


_INTEN_LOW_HIGH = np.array([[-8, 8], [-17, 17], [-29, 29], [-42, 42], [-60, 60], [-80, 80], [-106, 106], [-183, 183]], np.int32)  # g_etc1_inten_tables[t][0], [t][3]


def training_vectors_from_histogram(hist):
    """Endpoint training vectors from the (all-reduced) key histogram (frontend.cpp:843-857): per non-zero key, vec6F = (low rgb,
    high rgb) * (1/255) and weight = count, in ascending key order. Distinct keys can give the same vector: merge_training_vectors
    (or the device clusterer, which merges duplicates itself) produces the reference's unique set."""
    keys = np.nonzero(hist)[0].astype(np.uint32)
    r5, g5, b5, inten = (keys >> 13) & 31, (keys >> 8) & 31, (keys >> 3) & 31, keys & 7
    base = np.stack([(r5 << 3) | (r5 >> 2), (g5 << 3) | (g5 >> 2), (b5 << 3) | (b5 >> 2)], -1).astype(np.int32)
    low = np.clip(base + _INTEN_LOW_HIGH[inten, 0:1], 0, 255)
    high = np.clip(base + _INTEN_LOW_HIGH[inten, 1:2], 0, 255)
    vecs = np.concatenate([low, high], -1).astype(np.float32) * np.float32(1.0 / 255.0)
    return keys, vecs, hist[keys].astype(np.uint64)


def merge_training_vectors(vecs, weights):
    """Distinct keys can clamp to the same vector (e.g. intensity table 7 around mid grey gives (0,0,0,1,1,1) for thousands of keys);
    the clusterer's std::map merges those (enc.h:2228-2260). Returns the unique vectors in lexicographic order with summed weights."""
    u, inv = np.unique(np.asarray(vecs, np.float32), axis=0, return_inverse=True)
    w = np.zeros(u.shape[0], np.uint64)
    np.add.at(w, inv.reshape(-1), np.asarray(weights, np.uint64))
    return u, w
