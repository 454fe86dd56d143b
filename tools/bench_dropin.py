"""Whole-encoder timing through the reference's own basis_compress() (development/measurement tool): the unmodified reference
encoder linked with integration/basisu_opencl_b200.cpp (integration/_build/libbasisu_dropin.so), ETC1S at -q 128, with the
seam on (cFlagUseOpenCL: the five per-block frontend stages run on the B200) and off (stock CPU path, all host threads).
Prints one JSON line per image size: seconds, file size and PSNR of both."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from test_dropin_gpu import DROPIN, cFlagThreaded, cFlagUseOpenCL, compress, psnr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs="+", default=[1024, 2048])
    ap.add_argument("--quality", type=int, default=128)
    ap.add_argument("--debug", action="store_true", help="turn the encoder's debug_printf on: per-stage timers of both runs go to stdout")
    args = ap.parse_args()
    lib = ctypes.CDLL(DROPIN)
    lib.ref_compress_image.restype = ctypes.c_void_p
    lib.ref_compress_image.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
    lib.ref_free.argtypes = [ctypes.c_void_p]
    assert lib.ref_init_gpu_seam() == 1
    for dim in args.dims:
        img = util.synth(dim, 1234)
        img[..., 3] = 255
        img = np.ascontiguousarray(img)
        compress(lib, img[:64, :64].copy(), args.quality | cFlagThreaded | cFlagUseOpenCL)  # warm-up: context, module load
        if args.debug:
            lib.ref_enable_debug_printf(1); print(f"==== B200 seam run, {dim}x{dim}", flush=True)
        t0 = time.perf_counter(); gpu = compress(lib, img, args.quality | cFlagThreaded | cFlagUseOpenCL); t_gpu = time.perf_counter() - t0
        if args.debug:
            print(f"==== CPU run, {dim}x{dim}", flush=True)
        t0 = time.perf_counter(); cpu = compress(lib, img, args.quality | cFlagThreaded); t_cpu = time.perf_counter() - t0
        if args.debug:
            lib.ref_enable_debug_printf(0)
        (g_rgb, g_y), (c_rgb, c_y) = psnr(lib, gpu, img), psnr(lib, cpu, img)
        print(json.dumps({"config": f"ETC1S q{args.quality} {dim}x{dim} through basis_compress()", "cpu_s": round(t_cpu, 3), "b200_seam_s": round(t_gpu, 3),
                          "cpu_bytes": len(cpu), "b200_bytes": len(gpu), "cpu_psnr_rgb_y": [round(c_rgb, 4), round(c_y, 4)], "b200_psnr_rgb_y": [round(g_rgb, 4), round(g_y, 4)],
                          "host_threads": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}), flush=True)


if __name__ == "__main__":
    main()
