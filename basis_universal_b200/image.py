"""Host-side mirror of the steps either side of the per-block path (include/basisu_b200.h, "either side" section):
raster -> pixel_blocks (basis_compressor::extract_source_blocks), UASTC blocks -> texels (basist::unpack_uastc), and the
reference's image_metrics::calc from device-built histograms. No CPU fallback: every call goes through libbasisu_b200.so."""
import ctypes

import numpy as np

from ._lib import B200Error, lib


class _BlockMetrics(ctypes.Structure):
    _fields_ = [("hist", (ctypes.c_uint64 * 256) * 6), ("sum_a", ctypes.c_uint64 * 4), ("sum_b", ctypes.c_uint64 * 4)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class ImageOps:
    def __init__(self, device=0):
        self._lib = lib()
        self._ctx = self._lib.b200_create_context(int(device))
        if not self._ctx:
            raise B200Error(f"b200_create_context({device}) failed: {self._lib.b200_last_error(None).decode()}")

    def close(self):
        if self._ctx:
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

    def extract_source_blocks(self, img):
        """(H, W, 4) uint8 (rows may be strided) -> (ceil(H/4) * ceil(W/4), 64) uint8 pixel_blocks, edges clamped."""
        if img.ndim != 3 or img.shape[2] != 4 or img.dtype != np.uint8 or img.strides[2] != 1 or img.strides[1] != 4 or img.strides[0] < img.shape[1] * 4:  # negative or overlapping row strides are copied
            img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape[:2]
        out = np.empty((((h + 3) // 4) * ((w + 3) // 4), 64), np.uint8)
        self._check(self._lib.b200_extract_source_blocks(self._ctx, _p(img), w, h, ctypes.c_size_t(img.strides[0] if h else 0), _p(out)), "b200_extract_source_blocks")
        return out

    def unpack_uastc(self, ublocks):
        """(N, 16) uint8 UASTC blocks -> (N, 64) uint8 RGBA texels, [y][x] within the block."""
        u = np.ascontiguousarray(ublocks, np.uint8).reshape(-1, 16)
        out = np.empty((u.shape[0], 64), np.uint8)
        self._check(self._lib.b200_uastc_unpack_blocks(self._ctx, _p(u), u.shape[0], _p(out)), "b200_uastc_unpack_blocks")
        return out

    def unpack_etc1(self, eblocks, strict=True):
        """(N, 8) uint8 ETC1 blocks -> (N, 64) uint8 RGBA texels. strict=False returns the texels even when some block's
        differential colour overflows (basisu::unpack_etc1 returns false for such a block and writes nothing: zeros here)."""
        e = np.ascontiguousarray(eblocks, np.uint8).reshape(-1, 8)
        out = np.empty((e.shape[0], 64), np.uint8)
        ok = self._lib.b200_etc1_unpack_blocks(self._ctx, _p(e), e.shape[0], _p(out))
        if strict:
            self._check(ok, "b200_etc1_unpack_blocks")
        return out

    def resample(self, src, dst, clist_x, clist_y, first_comp=0, num_comps=4, srgb_tables=None):
        """basisu::image_resample (enc.cpp:1022) with the reference's contributor lists: src (H, W, 4) uint8, dst (h, w, 4) uint8 holding
        the destination's initial contents (modified in place and returned); clist_? = (offsets uint32 (n + 1,), weights float32, pixels
        uint32) as Resampler::get_clist_x / _y give them; srgb_tables = (srgb_to_linear float32 (256,), linear_to_srgb uint8 (8192,)) or None."""
        src = np.ascontiguousarray(src, np.uint8)
        assert dst.dtype == np.uint8 and dst.flags.c_contiguous and dst.ndim == 3 and dst.shape[2] == 4 and src.ndim == 3 and src.shape[2] == 4

        def pack(cl, n):
            off, wts, pix = cl
            off = np.ascontiguousarray(off, np.uint32)
            assert off.shape == (n + 1,)
            c = np.zeros(int(off[-1]), np.dtype([("weight", np.float32), ("pixel", np.uint32)]))
            c["weight"] = wts; c["pixel"] = pix
            return off, c

        xo, xc = pack(clist_x, dst.shape[1])
        yo, yc = pack(clist_y, dst.shape[0])
        s2l = l2s = None
        if srgb_tables is not None:
            s2l = np.ascontiguousarray(srgb_tables[0], np.float32); l2s = np.ascontiguousarray(srgb_tables[1], np.uint8)
            assert s2l.shape == (256,) and l2s.shape == (8192,)
        self._check(self._lib.b200_image_resample_rgba8(self._ctx, _p(src), src.shape[1], src.shape[0], ctypes.c_size_t(src.shape[1] * 4), _p(dst), dst.shape[1], dst.shape[0],
                                                         ctypes.c_size_t(dst.shape[1] * 4), _p(xo), _p(xc), _p(yo), _p(yc), int(first_comp), int(num_comps),
                                                         _p(s2l) if s2l is not None else None, _p(l2s) if l2s is not None else None), "b200_image_resample_rgba8")
        return dst

    def block_metrics_device(self, d_blocks_a, d_blocks_b, width, height):
        """Device pointers to two block arrays of a width x height image -> (hist (6, 256) uint64, sum_a (4,), sum_b (4,))."""
        m = _BlockMetrics()
        ok = self._lib.b200_block_metrics_device(self._ctx, ctypes.c_void_p(d_blocks_a), ctypes.c_void_p(d_blocks_b), int(width), int(height), ctypes.byref(m))
        self._check(ok, "b200_block_metrics_device")
        return np.ctypeslib.as_array(m.hist).astype(np.uint64).reshape(6, 256), np.ctypeslib.as_array(m.sum_a).copy(), np.ctypeslib.as_array(m.sum_b).copy()


def metrics_from_histograms(hist, width, height, first_chan=0, total_chans=0, avg_comp_error=True, use_601_luma=False):
    """The floating-point tail of image_metrics::calc (encoder/basisu_enc.cpp:2205-2224), operation for operation:
    returns dict(max, mean, mean_squared, rms, psnr). total_chans == 0 selects the luma histogram."""
    if total_chans:
        h = hist[first_chan:first_chan + total_chans].sum(0)
    else:
        h = hist[5 if use_601_luma else 4]
    m_max, s, s2 = 0.0, 0.0, 0.0
    for i in range(256):
        if h[i]:
            m_max = max(m_max, float(i))
            v = float(i) * float(h[i])
            s += v
            s2 += float(i) * v
    total_values = float(width) * float(height)
    if avg_comp_error:
        total_values *= float(min(max(total_chans, 1), 4))
    mean = float(np.float32(min(max(s / total_values, 0.0), 255.0)))
    mean_squared = float(np.float32(min(max(s2 / total_values, 0.0), 255.0 * 255.0)))
    rms = float(np.float32(np.sqrt(mean_squared)))
    psnr = float(np.float32(min(max(np.log10(255.0 / rms) * 20.0, 0.0), 100.0))) if rms else 100.0
    return {"max": m_max, "mean": mean, "mean_squared": mean_squared, "rms": rms, "psnr": psnr}
