"""A/B timing of libbasisu_b200.so build variants on one GPU: python tools/ab_bench.py a.so b.so ...
Each library encodes the same 4096^2 synthetic image (level 2 unless --level) a few times; prints per-stage CUDA-event times
and a hash of the output so that variants can be compared for speed and for identical bytes. Development tool, not a bench."""
import argparse
import ctypes
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth, to_blocks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--dim", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    blocks = torch.from_numpy(to_blocks(synth(args.dim, 1234))).cuda()
    n = blocks.shape[0]
    out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    for path in args.libs:
        L = ctypes.CDLL(os.path.abspath(path))
        L.b200_create_context.restype = vp
        L.b200_create_context.argtypes = [ctypes.c_int]
        L.b200_uastc_encode_blocks_device.restype = ctypes.c_int
        L.b200_uastc_encode_blocks_device.argtypes = [vp, vp, u32, vp, u32]
        L.b200_last_stage_ms.restype = ctypes.c_float
        L.b200_last_stage_ms.argtypes = [vp, u32]
        L.b200_last_kernel_ms.restype = ctypes.c_float
        L.b200_last_kernel_ms.argtypes = [vp]
        L.b200_destroy_context.argtypes = [vp]
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_last_error.argtypes = [vp]
        ctx = L.b200_create_context(0)
        if not ctx:
            print(f"{os.path.basename(path):40s} create_context failed: {L.b200_last_error(None)}", flush=True)
            continue
        best = None
        for r in range(args.reps):
            out.zero_()
            torch.cuda.synchronize()
            if not L.b200_uastc_encode_blocks_device(ctx, blocks.data_ptr(), n, out.data_ptr(), args.level):
                print(f"{os.path.basename(path):40s} encode failed: {L.b200_last_error(ctx)}", flush=True)
                break
            t = [L.b200_last_stage_ms(ctx, i) for i in range(3)] + [L.b200_last_kernel_ms(ctx)]
            if best is None or t[3] < best[3]:
                best = t
        if best is None:
            continue
        h = hashlib.md5(out.cpu().numpy().tobytes()).hexdigest()[:12]
        print(f"{os.path.basename(path):40s} classify {best[0]:7.2f}  candidates {best[1]:7.2f}  finish {best[2]:7.2f}  total {best[3]:7.2f} ms  md5 {h}", flush=True)
        L.b200_destroy_context(ctx)


if __name__ == "__main__":
    main()
