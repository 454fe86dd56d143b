"""End-to-end drop-in check on the GPU: the UNMODIFIED reference encoder library with one translation unit swapped
(encoder/basisu_opencl.cpp -> integration/basisu_opencl_b200.cpp) compresses ETC1S through the reference's own
basis_compress() API with cFlagUseOpenCL, i.e. its frontend runs the five per-block stages on the B200 through the
opencl_* seam.  Gate (BASELINE.json north_star): PSNR within +-0.02 dB of the reference CPU encoder at identical -q."""
import ctypes
import os

import numpy as np
import pytest

import util
from util import _ptr

pytestmark = pytest.mark.gpu

DROPIN = os.path.join(util.ROOT, "integration", "_build", "libbasisu_dropin.so")
cFlagUseOpenCL, cFlagThreaded = 1 << 8, 1 << 9


@pytest.fixture(scope="module")
def dropin():
    if not os.path.exists(DROPIN):
        pytest.skip("integration/_build/libbasisu_dropin.so did not travel (built by integration/Makefile where /root/reference exists)")
    lib = ctypes.CDLL(DROPIN)
    lib.ref_compress_image.restype = ctypes.c_void_p
    lib.ref_compress_image.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
    lib.ref_free.argtypes = [ctypes.c_void_p]
    assert lib.ref_init_gpu_seam() == 1, "the reference's opencl_is_available() is false: the seam did not come up"
    return lib


def compress(lib, img, flags):
    size = ctypes.c_size_t(0)
    p = lib.ref_compress_image(0, _ptr(img), img.shape[1], img.shape[0], flags, ctypes.c_float(0), ctypes.byref(size))
    assert p
    data = ctypes.string_at(p, size.value)
    lib.ref_free(p)
    return data


def psnr(lib, data, img):
    out = np.zeros(img.shape, np.uint8)
    buf = np.frombuffer(data, np.uint8)
    assert lib.ref_transcode_basis_to_rgba(_ptr(buf), ctypes.c_uint32(len(data)), _ptr(out), ctypes.c_uint32(img.shape[0] * img.shape[1]))
    a, b = out[..., :3].astype(np.float64), img[..., :3].astype(np.float64)
    rgb = 10 * np.log10(255 ** 2 / np.mean((a - b) ** 2))
    la = a @ np.array([0.2126, 0.7152, 0.0722]); lb = b @ np.array([0.2126, 0.7152, 0.0722])
    return rgb, 10 * np.log10(255 ** 2 / np.mean((la - lb) ** 2))


@pytest.mark.parametrize("quality", [128, 255])
def test_etc1s_through_the_reference_frontend(dropin, quality):
    from basis_universal_b200 import lib as b200lib
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    img = np.ascontiguousarray(g["image"])
    launches0 = b200lib().b200_global_launch_count()
    gpu = compress(dropin, img, quality | cFlagThreaded | cFlagUseOpenCL)
    launches = b200lib().b200_global_launch_count() - launches0
    cpu = compress(dropin, img, quality | cFlagThreaded)
    (gpu_rgb, gpu_y), (cpu_rgb, cpu_y) = psnr(dropin, gpu, img), psnr(dropin, cpu, img)
    print(f"q{quality}: CPU {cpu_rgb:.4f} dB RGB / {cpu_y:.4f} dB Y, {len(cpu)} B; B200 seam {gpu_rgb:.4f} / {gpu_y:.4f}, {len(gpu)} B; {launches} kernel launches")
    assert launches >= 4, "the frontend did not reach the B200 kernels"
    assert abs(gpu_y - cpu_y) <= 0.02 and abs(gpu_rgb - cpu_rgb) <= 0.02
    assert abs(len(gpu) - len(cpu)) <= 0.045 * len(cpu)  # the reference's own KAT size tolerance (basisu_tool.cpp:6793)
