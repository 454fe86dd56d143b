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


def extract_blocks(image):
    """(H, W, 4) uint8 -> (num_blocks, 64) uint8 in raster block order; edge texels are clamped like
    image::extract_block_clamped (encoder/basisu_enc.h:3168)."""
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

    def encode_uastc_device(self, d_blocks_ptr, num_blocks, d_out_ptr, flags=cPackUASTCLevelDefault):
        """Device-resident form: raw device pointers (e.g. torch tensor .data_ptr()) to (N,64) and (N,16) uint8 buffers."""
        ok = self._lib.b200_uastc_encode_blocks_device(self._ctx, ctypes.c_void_p(d_blocks_ptr), int(num_blocks), ctypes.c_void_p(d_out_ptr), int(flags))
        self._check(ok, "b200_uastc_encode_blocks_device")
