"""Stage timings of the ETC1S seam kernels on one GPU against the CPU oracles (development/measurement tool; the contract
bench is bench.py). For a DIM x DIM synthetic image: per b200_etc1s_* entry point the kernel time (CUDA events inside the
library), the host-visible time of the call (host buffers, copies included), and the time of the matching CPU oracle --
the reference's etc1_optimizer (one thread, on a block sample) and the reference's OpenCL kernels compiled for the host
(all cores) -- plus a bit-exactness check of each output. Prints one JSON line per stage."""
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
from basis_universal_b200 import etc1s  # noqa: E402


def timed(f, reps=3):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--perms", type=int, default=16)
    args = ap.parse_args()
    img = util.synth(args.dim, 99)
    img[..., 3] = 255
    blocks = util.image_to_blocks(img)
    n = blocks.shape[0]
    ctx = etc1s.Etc1sContext(0)
    ocl = util.OclRef() if os.path.exists(util.OCL_SO) else None
    ref = util.Ref() if os.path.exists(util.REF_SO) else None
    st = util.etc1s_stage_inputs(blocks, 5, parents=64, clusters_per_parent=(8, 64), selectors_per_parent=(32, 512))
    clusters, cpx, cw = util.pixel_cluster_inputs(3, n_clusters=4096)

    def line(stage, units, unit_name, gpu_wall, kernel_ms, cpu_s, cpu_kind, cpu_units, exact):
        print(json.dumps({"stage": stage, "units": units, "unit": unit_name, "gpu_kernel_ms": round(kernel_ms, 3), "gpu_call_ms": round(gpu_wall * 1e3, 3),
                          "gpu_units_per_s": round(units / gpu_wall), "cpu_kind": cpu_kind, "cpu_units_per_s": None if cpu_s is None else round(cpu_units / cpu_s),
                          "speedup_call": None if cpu_s is None else round((cpu_s / cpu_units) / (gpu_wall / units), 1), "bit_exact": exact}), flush=True)

    ctx.set_pixel_blocks(blocks)
    # encode_etc1s_blocks, both flavours
    for flavour, name in ((etc1s.FLAVOUR_CPU_OPTIMIZER, "cpu_optimizer"), (etc1s.FLAVOUR_OPENCL_KERNELS, "opencl_kernels")):
        ctx.set_flavour(flavour)
        ctx.encode_etc1s_blocks(True, args.perms)
        wall, out = timed(lambda: ctx.encode_etc1s_blocks(True, args.perms))
        kms = ctx.last_kernel_ms
        if name == "cpu_optimizer" and ref is not None:
            m = min(n, 16384)
            comp_level = {16: 1, 64: 2, 165: 6}.get(args.perms, 1)  # basis_compressor: comp_level -> total_perms (frontend.cpp:739-751)
            want = np.zeros((m, 8), np.uint8)
            t0 = time.perf_counter()
            ref.lib.ref_etc1s_encode_blocks(util._ptr(np.ascontiguousarray(blocks[:m])), ctypes.c_uint32(m), util._ptr(want), 1, comp_level)
            cpu_s = time.perf_counter() - t0
            line(f"encode_etc1s_blocks[{name}]", n, "blocks", wall, kms, cpu_s, "reference etc1_optimizer, 1 thread", m, bool(np.array_equal(out[:m], want)))
        elif name == "opencl_kernels" and ocl is not None:
            cpu_s, want = timed(lambda: ocl.encode_etc1s_blocks(blocks, True, args.perms), 1)
            line(f"encode_etc1s_blocks[{name}]", n, "blocks", wall, kms, cpu_s, f"reference ocl_kernels.cl on host, {ocl.threads} threads", n, bool(np.array_equal(out, want)))
    ctx.set_flavour(etc1s.FLAVOUR_OPENCL_KERNELS)
    if ocl is not None:
        ctx.encode_etc1s_pixel_clusters(clusters, cpx, cw, True, args.perms)
        wall, out = timed(lambda: ctx.encode_etc1s_pixel_clusters(clusters, cpx, cw, True, args.perms))
        kms = ctx.last_kernel_ms
        cpu_s, want = timed(lambda: ocl.encode_pixel_clusters(clusters, cpx, cw, True, args.perms), 1)
        line("encode_etc1s_pixel_clusters", int(cpx.shape[0]), "texels", wall, kms, cpu_s, f"ocl_kernels.cl on host, {ocl.threads} threads", int(cpx.shape[0]), bool(np.array_equal(out[:, :4], want[:, :4])))  # selector bytes are undefined

        ctx.refine_endpoint_clusterization(st["block_info"], st["cluster_info"], st["sorted_idx"], True)
        wall, out = timed(lambda: ctx.refine_endpoint_clusterization(st["block_info"], st["cluster_info"], st["sorted_idx"], True))
        kms = ctx.last_kernel_ms
        cpu_s, want = timed(lambda: ocl.refine(blocks, st["block_info"], st["cluster_info"], st["sorted_idx"], True), 1)
        line("refine_endpoint_clusterization", n, "blocks", wall, kms, cpu_s, f"ocl_kernels.cl on host, {ocl.threads} threads", n, bool(np.array_equal(out, want)))

        ctx.find_optimal_selector_clusters_for_each_block(st["fosc_blocks"], st["selectors"], st["sel_cluster_idx"], True)
        wall, out = timed(lambda: ctx.find_optimal_selector_clusters_for_each_block(st["fosc_blocks"], st["selectors"], st["sel_cluster_idx"], True))
        kms = ctx.last_kernel_ms
        cpu_s, want = timed(lambda: ocl.fosc(blocks, st["fosc_blocks"], st["selectors"], st["sel_cluster_idx"], True), 1)
        line("find_optimal_selector_clusters_for_each_block", n, "blocks", wall, kms, cpu_s, f"ocl_kernels.cl on host, {ocl.threads} threads", n, bool(np.array_equal(out, want)))

        ctx.determine_selectors(st["color5_inten"], True)
        wall, out = timed(lambda: ctx.determine_selectors(st["color5_inten"], True))
        kms = ctx.last_kernel_ms
        cpu_s, want = timed(lambda: ocl.determine_selectors(blocks, st["color5_inten"], True), 1)
        line("determine_selectors", n, "blocks", wall, kms, cpu_s, f"ocl_kernels.cl on host, {ocl.threads} threads", n, bool(np.array_equal(out, want)))

    enc = ctx.encode_etc1s_blocks(True, args.perms)
    ctx.endpoint_histogram(enc)
    wall, hist = timed(lambda: ctx.endpoint_histogram(enc))
    want = np.bincount(util.endpoint_keys(enc), minlength=1 << 18).astype(np.uint32) * 2
    line("endpoint_histogram", n, "blocks", wall, ctx.last_kernel_ms, None, None, n, bool(np.array_equal(hist, want)))
    ctx.close()


if __name__ == "__main__":
    main()
