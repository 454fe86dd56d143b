"""Host-side mirror of the reference's UASTC LDR 4x4 encoder interface (encoder/basisu_uastc_enc.h), batch form.

Flag names and values are the reference's (uastc_enc.h:24-63).  `Encoder.encode_uastc` is the array form of
`basisu::encode_uastc(const uint8_t* pRGBAPixels, basist::uastc_block&, uint32_t flags)` (uastc_enc.h:68).
"""
import ctypes
import numpy as np
from ._lib import lib, B200Error

cPackUASTCLevelFastest = 0
cPackUASTCLevelFaster = 1
cPackUASTCLevelDefault = 2
cPackUASTCLevelSlower = 3
cPackUASTCLevelVerySlow = 4
cPackUASTCLevelMask = 0xF
cPackUASTCFavorUASTCError = 8
cPackUASTCFavorBC7Error = 16
cPackUASTCETC1FasterHints = 64
cPackUASTCETC1FastestHints = 128
cPackUASTCETC1DisableFlipAndIndividual = 256
cPackUASTCFavorSimplerModes = 512


class uastc_rdo_params(ctypes.Structure):
    """basisu::uastc_rdo_params (encoder/basisu_uastc_enc.h:82-128) with the reference's defaults (clear(), :89-100)."""
    _fields_ = [("lz_dict_size", ctypes.c_uint32), ("lambda_", ctypes.c_float), ("max_allowed_rms_increase_ratio", ctypes.c_float),
                ("skip_block_rms_thresh", ctypes.c_float), ("endpoint_refinement", ctypes.c_uint32), ("max_smooth_block_std_dev", ctypes.c_float),
                ("smooth_block_max_error_scale", ctypes.c_float), ("lz_literal_cost", ctypes.c_uint32)]

    def __init__(self, lambda_=0.5, lz_dict_size=4096, max_allowed_rms_increase_ratio=10.0, skip_block_rms_thresh=8.0, endpoint_refinement=True,
                 max_smooth_block_std_dev=18.0, smooth_block_max_error_scale=10.0, lz_literal_cost=100):
        super().__init__(lz_dict_size, lambda_, max_allowed_rms_increase_ratio, skip_block_rms_thresh, int(bool(endpoint_refinement)),
                         max_smooth_block_std_dev, smooth_block_max_error_scale, lz_literal_cost)


def extract_blocks(image):
    """(H, W, 4) uint8 -> (num_blocks, 64) uint8 in raster block order; edge texels are clamped like
    image::extract_block_clamped (encoder/basisu_enc.h:3168). Plain numpy tiling for callers that marshal pixel_blocks on the
    host, as the reference's compressor does; it is not used by any encode call (the device form is
    image.ImageOps.extract_source_blocks, or Encoder.encode_image which ingests and encodes in one call)."""
    image = np.asarray(image, np.uint8)
    h, w, c = image.shape
    if c != 4:
        raise ValueError("expected RGBA")
    ph, pw = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    if (ph, pw) != (h, w):
        image = np.pad(image, ((0, ph - h), (0, pw - w), (0, 0)), mode="edge")
    return np.ascontiguousarray(image.reshape(ph // 4, 4, pw // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 64))


class Encoder:
    """One device context (one CUDA device, one stream, its scratch buffers)."""

    def __init__(self, device=0):
        self._lib = lib()
        self._ctx = self._lib.b200_create_context(int(device))
        if not self._ctx:
            raise B200Error("b200_create_context failed: " + self._lib.b200_last_error(None).decode())
        self.device = device

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

    @property
    def last_launch_count(self):
        return int(self._lib.b200_last_launch_count(self._ctx))

    def stage_ms(self, stage):
        return float(self._lib.b200_last_stage_ms(self._ctx, int(stage)))

    def timer_start(self):
        self._check(self._lib.b200_timer_start(self._ctx), "b200_timer_start")

    def timer_stop_ms(self):
        ms = float(self._lib.b200_timer_stop_ms(self._ctx))
        self._check(ms >= 0, "b200_timer_stop_ms")
        return ms

    def encode_uastc_host_ptr(self, host_in_ptr, num_blocks, host_out_ptr, flags=cPackUASTCLevelDefault):
        """Host-pointer form for callers that own (ideally pinned) buffers: H2D, kernels and D2H all happen inside the call."""
        ok = self._lib.b200_uastc_encode_blocks(self._ctx, ctypes.c_void_p(host_in_ptr), int(num_blocks), ctypes.c_void_p(host_out_ptr), int(flags))
        self._check(ok, "b200_uastc_encode_blocks")

    def encode_uastc(self, blocks, flags=cPackUASTCLevelDefault, out=None):
        """blocks: (N, 64) uint8 host array of pixel_blocks -> (N, 16) uint8 UASTC blocks (host). Copies are inside the call."""
        blocks = np.ascontiguousarray(blocks, np.uint8)
        if blocks.ndim != 2 or blocks.shape[1] != 64:
            raise ValueError("blocks must be (N, 64) uint8")
        n = blocks.shape[0]
        if out is None:
            out = np.empty((n, 16), np.uint8)
        ok = self._lib.b200_uastc_encode_blocks(self._ctx, blocks.ctypes.data, n, out.ctypes.data, int(flags))
        self._check(ok, "b200_uastc_encode_blocks")
        return out

    def uastc_rdo(self, uastc_blocks, source_blocks, params=None, flags=cPackUASTCLevelDefault, total_jobs=4):
        """basisu::uastc_rdo (uastc_enc.h:139): in-place RDO post-pass. `total_jobs` is the reference's chain split
        (comp.cpp:2078 passes min(4, pool threads)); the output depends on it. Returns the modified (N, 16) array."""
        blocks = np.ascontiguousarray(uastc_blocks, np.uint8).copy()
        src = np.ascontiguousarray(source_blocks, np.uint8)
        if blocks.shape[0] != src.shape[0] or blocks.shape[1] != 16 or src.shape[1] != 64:
            raise ValueError("uastc_blocks must be (N,16) and source_blocks (N,64)")
        p = params if params is not None else uastc_rdo_params()
        ok = self._lib.b200_uastc_rdo(self._ctx, blocks.shape[0], blocks.ctypes.data, src.ctypes.data, ctypes.byref(p), int(flags), int(total_jobs))
        self._check(ok, "b200_uastc_rdo")
        return blocks

    def uastc_rdo_batch(self, uastc_blocks, source_blocks, slice_num_blocks, params=None, flags=cPackUASTCLevelDefault, total_jobs=4):
        """uastc_rdo over several slices laid end to end (the per-slice loop of comp.cpp:1996-2089): every chain of every slice
        runs concurrently. Same result as one uastc_rdo call per slice. Returns the modified (N, 16) array."""
        blocks = np.ascontiguousarray(uastc_blocks, np.uint8).copy()
        src = np.ascontiguousarray(source_blocks, np.uint8)
        counts = np.ascontiguousarray(slice_num_blocks, np.uint32)
        if blocks.shape[0] != src.shape[0] or blocks.shape[1] != 16 or src.shape[1] != 64 or int(counts.sum()) != blocks.shape[0]:
            raise ValueError("uastc_blocks must be (N,16), source_blocks (N,64) and the slice sizes must add up to N")
        p = params if params is not None else uastc_rdo_params()
        ok = self._lib.b200_uastc_rdo_batch(self._ctx, counts.shape[0], counts.ctypes.data, blocks.ctypes.data, src.ctypes.data, ctypes.byref(p), int(flags), int(total_jobs))
        self._check(ok, "b200_uastc_rdo_batch")
        return blocks

    def uastc_rdo_batch_device(self, d_uastc_ptr, d_source_ptr, slice_num_blocks, params=None, flags=cPackUASTCLevelDefault, total_jobs=4):
        """Device-resident form of uastc_rdo_batch: the encoder's output is post-processed in place in HBM."""
        counts = np.ascontiguousarray(slice_num_blocks, np.uint32)
        p = params if params is not None else uastc_rdo_params()
        ok = self._lib.b200_uastc_rdo_batch_device(self._ctx, counts.shape[0], counts.ctypes.data, ctypes.c_void_p(d_uastc_ptr), ctypes.c_void_p(d_source_ptr),
                                                   ctypes.byref(p), int(flags), int(total_jobs))
        self._check(ok, "b200_uastc_rdo_batch_device")

    def encode_rdo_host_ptr(self, host_in_ptr, slice_num_blocks, host_out_ptr, params=None, flags=cPackUASTCLevelDefault, total_jobs=4):
        """b200_uastc_encode_rdo_blocks: source blocks in, encoded (and, with params, RDO post-processed) UASTC blocks out; host pointers."""
        counts = np.ascontiguousarray(slice_num_blocks, np.uint32)
        ok = self._lib.b200_uastc_encode_rdo_blocks(self._ctx, counts.shape[0], counts.ctypes.data, ctypes.c_void_p(host_in_ptr), ctypes.c_void_p(host_out_ptr), int(flags),
                                                    ctypes.byref(params) if params is not None else None, int(total_jobs))
        self._check(ok, "b200_uastc_encode_rdo_blocks")

    def encode_uastc_device(self, d_blocks_ptr, num_blocks, d_out_ptr, flags=cPackUASTCLevelDefault):
        """Device-resident form: raw device pointers (e.g. torch tensor .data_ptr()) to (N,64) and (N,16) uint8 buffers."""
        ok = self._lib.b200_uastc_encode_blocks_device(self._ctx, ctypes.c_void_p(d_blocks_ptr), int(num_blocks), ctypes.c_void_p(d_out_ptr), int(flags))
        self._check(ok, "b200_uastc_encode_blocks_device")

    def encode_image(self, image, flags=cPackUASTCLevelDefault):
        """(H, W, 4) uint8 raster (rows may be strided) -> (ceil(H/4) * ceil(W/4), 16) uint8 UASTC blocks: ingest and encode both on
        the device (b200_uastc_encode_image)."""
        img = image
        if img.ndim != 3 or img.shape[2] != 4 or img.dtype != np.uint8 or img.strides[2] != 1 or img.strides[1] != 4 or img.strides[0] < img.shape[1] * 4:  # negative or overlapping row strides are copied
            img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape[:2]
        out = np.empty((((h + 3) // 4) * ((w + 3) // 4), 16), np.uint8)
        ok = self._lib.b200_uastc_encode_image(self._ctx, img.ctypes.data_as(ctypes.c_void_p), w, h, ctypes.c_size_t(img.strides[0] if h else 0), out.ctypes.data_as(ctypes.c_void_p), int(flags))
        self._check(ok, "b200_uastc_encode_image")
        return out
