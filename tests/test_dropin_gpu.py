"""End-to-end drop-in checks on the GPU: the reference encoder library with encoder/basisu_opencl.cpp swapped for
integration/basisu_opencl_b200.cpp and the call-site patches of integration/patches/ applied (integration/Makefile), driven
through the reference's own public API, basis_compress(), with cFlagUseOpenCL.
  ETC1S: the frontend's per-block stages, both VQ clusterers, the endpoint-cluster optimiser and the selector codebook run on
         the B200. Gate (BASELINE.json north_star): PSNR within +-0.02 dB of the reference CPU encoder at identical -q.
  UASTC: encode + RDO run on the B200. Gate: the .basis file is byte-identical (SURVEY 8(c) md5 for kodim03 level 0)."""
import hashlib
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


def compress(lib, img, flags, fmt=0, rdo_quality=0.0):
    size = ctypes.c_size_t(0)
    p = lib.ref_compress_image(fmt, _ptr(img), img.shape[1], img.shape[0], flags, ctypes.c_float(rdo_quality), ctypes.byref(size))
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


def kodim_images():
    """kodim01..24 from the copy build() leaves next to the compiled reference (oracle/_ref/test_files, git-ignored, travels to the box)."""
    from PIL import Image
    d = os.path.join(util.ROOT, "oracle", "_ref", "test_files")
    out = []
    for i in range(1, 25):
        f = os.path.join(d, f"kodim{i:02d}.png")
        if not os.path.exists(f):
            pytest.skip("oracle/_ref/test_files/kodim*.png did not travel (copied by __graft_entry__.build() where /root/reference exists)")
        out.append((i, np.ascontiguousarray(np.array(Image.open(f).convert("RGBA")))))
    return out


def test_etc1s_kodim_batch_q128_within_gate(dropin):
    """BASELINE config 3: kodim01-24, ETC1S q128, every image within +-0.02 dB (RGB and luma) and 4.5 % size of the CPU encoder."""
    from basis_universal_b200 import lib as b200lib
    worst = (0.0, 0)
    launches0 = b200lib().b200_global_launch_count()
    for i, img in kodim_images():
        gpu = compress(dropin, img, 128 | cFlagThreaded | cFlagUseOpenCL)
        cpu = compress(dropin, img, 128 | cFlagThreaded)
        (g_rgb, g_y), (c_rgb, c_y) = psnr(dropin, gpu, img), psnr(dropin, cpu, img)
        d = max(abs(g_y - c_y), abs(g_rgb - c_rgb))
        print(f"kodim{i:02d}: CPU {c_rgb:.4f}/{c_y:.4f} dB {len(cpu)} B; B200 {g_rgb:.4f}/{g_y:.4f} dB {len(gpu)} B; delta {g_rgb - c_rgb:+.4f}/{g_y - c_y:+.4f}")
        if d > worst[0]:
            worst = (d, i)
        assert abs(g_y - c_y) <= 0.02 and abs(g_rgb - c_rgb) <= 0.02, f"kodim{i:02d}"
        assert abs(len(gpu) - len(cpu)) <= 0.045 * len(cpu), f"kodim{i:02d}"
    assert b200lib().b200_global_launch_count() - launches0 >= 24 * 8
    print(f"worst |delta PSNR| {worst[0]:.4f} dB on kodim{worst[1]:02d}")


def test_uastc_kodim03_level0_md5_through_basis_compress(dropin):
    """BASELINE config 1 through the patched compressor: `basisu -uastc -uastc_level 0 kodim03.png` -> 393347 B,
    md5 6d98eb72a9a3112ff55344132a28b042, with the blocks encoded by the B200."""
    from basis_universal_b200 import lib as b200lib
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    img = np.ascontiguousarray(g["image"])
    launches0 = b200lib().b200_global_launch_count()
    data = compress(dropin, img, 0 | cFlagThreaded | cFlagUseOpenCL, fmt=1)
    assert b200lib().b200_global_launch_count() - launches0 >= 4, "encode_slices_to_uastc_4x4_ldr did not reach the B200 kernels"
    assert len(data) == 393347
    assert hashlib.md5(data).hexdigest() == "6d98eb72a9a3112ff55344132a28b042"


@pytest.mark.parametrize("level", [1, 2])
def test_uastc_rdo_file_identical_through_basis_compress(dropin, level):
    """`-uastc -uastc_rdo_l 1.0`: encode + RDO on the B200 give the CPU encoder's file byte for byte (threaded: 4 RDO chains)."""
    from basis_universal_b200 import lib as b200lib
    img = util.rdo_test_image()
    launches0 = b200lib().b200_global_launch_count()
    gpu = compress(dropin, img, level | cFlagThreaded | cFlagUseOpenCL, fmt=1, rdo_quality=1.0)
    assert b200lib().b200_global_launch_count() - launches0 >= 6
    cpu = compress(dropin, img, level | cFlagThreaded, fmt=1, rdo_quality=1.0)
    assert gpu == cpu


def test_bu_c_api_on_top_of_the_dropin(dropin):
    """The reference's pure-C whole-texture API (encoder/basisu_wasm_api.h:1-58) sits unchanged on the patched library:
    bu_compress_texture with the OpenCL flag runs the B200 path and returns the same ETC1S file as basis_compress."""
    from basis_universal_b200 import lib as b200lib
    L = dropin
    u64, u32 = ctypes.c_uint64, ctypes.c_uint32
    L.bu_init.restype = None
    L.bu_new_comp_params.restype = u64
    L.bu_alloc.restype = u64; L.bu_alloc.argtypes = [u64]
    L.bu_free.argtypes = [u64]
    L.bu_comp_params_set_image_rgba32.restype = u32
    L.bu_comp_params_set_image_rgba32.argtypes = [u64, u32, u64, u32, u32, u32]
    L.bu_compress_texture.restype = u32
    L.bu_compress_texture.argtypes = [u64, u32, ctypes.c_int, ctypes.c_int, u64, ctypes.c_float]
    L.bu_comp_params_get_comp_data_ofs.restype = u64; L.bu_comp_params_get_comp_data_ofs.argtypes = [u64]
    L.bu_comp_params_get_comp_data_size.restype = u64; L.bu_comp_params_get_comp_data_size.argtypes = [u64]
    L.bu_delete_comp_params.argtypes = [u64]
    L.bu_init()
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    img = np.ascontiguousarray(g["image"][:256, :256])
    h, w = img.shape[:2]
    params = L.bu_new_comp_params()
    buf = L.bu_alloc(img.nbytes)
    ctypes.memmove(buf, img.ctypes.data, img.nbytes)
    assert L.bu_comp_params_set_image_rgba32(params, 0, buf, w, h, w * 4)
    launches0 = b200lib().b200_global_launch_count()
    # basis_tex_format 0 = ETC1S; unified quality/effort -1 = "use the low-level flags": quality 128 | cFlagUseOpenCL | cFlagThreaded
    assert L.bu_compress_texture(params, 0, -1, -1, 128 | cFlagUseOpenCL | cFlagThreaded, ctypes.c_float(0.0))
    assert b200lib().b200_global_launch_count() - launches0 >= 8
    size = L.bu_comp_params_get_comp_data_size(params)
    data = ctypes.string_at(L.bu_comp_params_get_comp_data_ofs(params), size)
    L.bu_free(buf)
    L.bu_delete_comp_params(params)
    assert data == compress(dropin, img, 128 | cFlagThreaded | cFlagUseOpenCL), "bu_compress_texture and basis_compress disagree on the same input"
    (g_rgb, g_y) = psnr(dropin, data, img)
    cpu = compress(dropin, img, 128 | cFlagThreaded)
    (c_rgb, c_y) = psnr(dropin, cpu, img)
    assert abs(g_y - c_y) <= 0.02 and abs(g_rgb - c_rgb) <= 0.02


def _run_tool(args, cwd):
    import subprocess
    tool = os.path.join(util.ROOT, "integration", "_build", "basisu")
    if not os.path.exists(tool):
        pytest.skip("integration/_build/basisu did not travel (built by integration/Makefile where /root/reference exists)")
    r = subprocess.run([tool] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_basisu_tool_cli_on_the_dropin(tmp_path):
    """BASELINE config 1 with the reference's own CLI, unmodified (basisu_tool.cpp linked against the drop-in library):
    `basisu -uastc -uastc_level 0 -opencl kodim03.png` -> 393347 B, md5 6d98eb72..., blocks encoded by the B200; and the ETC1S default
    (`basisu -opencl kodim03.png`, -q 128, comp_level 1) gives the same file as the CPU run of the same binary."""
    png = os.path.join(util.ROOT, "oracle", "_ref", "test_files", "kodim03.png")
    if not os.path.exists(png):
        pytest.skip("oracle/_ref/test_files/kodim03.png did not travel")
    out = _run_tool(["-uastc", "-uastc_level", "0", "-basis", "-opencl", "-output_file", str(tmp_path / "u.basis"), png], tmp_path)
    assert "OpenCL: 1" in out, "the tool did not see the GPU seam"
    data = open(tmp_path / "u.basis", "rb").read()
    assert len(data) == 393347 and hashlib.md5(data).hexdigest() == "6d98eb72a9a3112ff55344132a28b042"
    out = _run_tool(["-basis", "-opencl", "-output_file", str(tmp_path / "e_gpu.basis"), png], tmp_path)
    assert "OpenCL: 1" in out
    _run_tool(["-basis", "-output_file", str(tmp_path / "e_cpu.basis"), png], tmp_path)
    gpu, cpu = open(tmp_path / "e_gpu.basis", "rb").read(), open(tmp_path / "e_cpu.basis", "rb").read()
    print(f"basisu CLI ETC1S q128 kodim03: CPU {len(cpu)} B, B200 {len(gpu)} B")
    assert gpu == cpu


@pytest.mark.parametrize("comp_level", [0, 3, 4, 5, 6])
def test_basisu_tool_cli_etc1s_comp_levels_identical_to_cpu(tmp_path, comp_level):
    """Every ETC1S effort level through the reference's CLI on the drop-in library: -comp_level 0 (no refinement), 3 / 5 / 6 (flat
    codebooks), 4-6 (endpoint / selector refinement iterations: introduce_new_endpoint_clusters with the device's subblock errors,
    generate_endpoint_codebook at steps >= 1 on the device, refine_block_endpoints_given_selectors on the host). The GPU file must equal
    the CPU file of the same binary, byte for byte (except level 0, where the reference's own seam asks the GPU stage for more
    permutations than its CPU path runs, frontend.cpp:745-783: PSNR gate there)."""
    from PIL import Image
    png = os.path.join(util.ROOT, "oracle", "_ref", "test_files", "kodim03.png")
    if not os.path.exists(png):
        pytest.skip("oracle/_ref/test_files/kodim03.png did not travel")
    crop = tmp_path / "crop.png"
    Image.open(png).convert("RGB").crop((256, 128, 256 + 320, 128 + 256)).save(crop)
    out = _run_tool(["-basis", "-comp_level", str(comp_level), "-q", "200", "-opencl", "-debug", "-output_file", str(tmp_path / "gpu.basis"), str(crop)], tmp_path)
    assert "OpenCL: 1" in out and "failed! Using CPU" not in out
    if comp_level >= 4:
        assert "opencl_b200_compute_subblock_errors" in out and out.count("opencl_b200_reoptimize_endpoint_clusters") >= 3
    _run_tool(["-basis", "-comp_level", str(comp_level), "-q", "200", "-output_file", str(tmp_path / "cpu.basis"), str(crop)], tmp_path)
    gpu, cpu = open(tmp_path / "gpu.basis", "rb").read(), open(tmp_path / "cpu.basis", "rb").read()
    print(f"basisu CLI ETC1S comp_level {comp_level} q200 320x256: CPU {len(cpu)} B, B200 {len(gpu)} B")
    if comp_level == 0:
        assert abs(len(gpu) - len(cpu)) <= 0.045 * len(cpu)
    else:
        assert gpu == cpu


@pytest.mark.parametrize("mode", ["etc1s", "etc1s_linear_clamp_slow", "uastc"])
def test_basisu_tool_cli_mipmaps_identical_to_cpu(tmp_path, mode):
    """`basisu -mipmap`: every mip level filtered on the device (basis_compressor::generate_mipmaps -> image_resample, comp.cpp:2146-2230),
    then the multi-slice texture through the ETC1S frontend + backend (one wavefront CTA per slice) or the UASTC encoder. The file must
    equal the CPU run's, byte for byte."""
    from PIL import Image
    png = os.path.join(util.ROOT, "oracle", "_ref", "test_files", "kodim03.png")
    if not os.path.exists(png):
        pytest.skip("oracle/_ref/test_files/kodim03.png did not travel")
    crop = tmp_path / "crop.png"
    Image.open(png).convert("RGB").crop((200, 100, 200 + 300, 100 + 212)).save(crop)   # 300 x 212: odd mip sizes down to 1 x 1
    args = ["-basis", "-mipmap"]
    if mode == "uastc":
        args += ["-uastc", "-uastc_level", "1"]
    if mode == "etc1s_linear_clamp_slow":   # defaults are sRGB filtering, wrap addressing, each level from the previous one
        args += ["-mip_linear", "-mip_clamp", "-mip_slow", "-mip_filter", "lanczos4"]
    out = _run_tool(args + ["-opencl", "-debug", "-output_file", str(tmp_path / "gpu.basis"), str(crop)], tmp_path)
    assert "OpenCL: 1" in out and "failed! Using CPU" not in out
    assert out.count("opencl_b200_image_resample") >= 8, "the mip levels were not filtered on the device"
    _run_tool(args + ["-output_file", str(tmp_path / "cpu.basis"), str(crop)], tmp_path)
    gpu, cpu = open(tmp_path / "gpu.basis", "rb").read(), open(tmp_path / "cpu.basis", "rb").read()
    print(f"basisu CLI -mipmap {mode} 300x212: CPU {len(cpu)} B, B200 {len(gpu)} B")
    assert gpu == cpu
