#!/usr/bin/env python
"""bench.py -- Mtexels/s of the basis_universal encode hot path on N B200s, one process per GPU.

  python bench.py [--config C] --gpus 1 --steps K --warmup W       (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference [--config C] --gpus N ...       (the reference's CPU encoder on the host cores, rank 0 only)

Configurations (BASELINE.json `configs`; the default is the one the headline metric is quoted on):
  uastc_l2         configs[1]  synthetic 4096^2 RGBA, UASTC LDR level 2, one tile per GPU, no collective (weak scaling)     [default]
  uastc_l2_strong  configs[1]  the same single 4096^2 image cut into block-row ranges over the N GPUs (strong scaling)
  etc1s_kodim      configs[2]  kodim01-24, ETC1S -q 128, through the reference's basis_compress() on the patched library (drop-in)
  etc1s_8k         configs[3]  synthetic 8192^2, ETC1S -q 255, per-stage work sharded over N GPUs, stage outputs merged with NCCL all-reduces
  uastc_l4_rdo     configs[4]  synthetic 4096^2 tiles, UASTC level 4 + RDO (lambda 1.0, 4 chains), tiles spread over the GPUs (replicas)

`value` is device-timed with inputs resident in HBM (ETC1S: the frontend stage, whose host orchestration is the reference's own
code); `e2e` is the same work through the reference-facing call with HOST buffers, copies inside the timed region. `roofline`
reports the dominant kernel against the measured HBM copy bandwidth using SURVEY.md 8(d)'s algorithmic bytes; these kernels are
instruction-bound, so the fraction is small by construction.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGE_DIM = 4096
BLOCKS = (IMAGE_DIM // 4) ** 2
TEXELS = IMAGE_DIM * IMAGE_DIM
ROTATING_INPUTS = 4            # 4 x 64 MiB distinct inputs > 126 MB L2
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbasisu_ref.so")
DROPIN_SO = os.path.join(ROOT, "integration", "_build", "libbasisu_dropin.so")
cFlagUseOpenCL, cFlagThreaded = 1 << 8, 1 << 9
cPackUASTCFavorSimplerModes = 512


def synth(n, seed):
    """SURVEY.md 9.6 generator (same code as tests/util.py; duplicated so bench.py has no test imports)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    r = 127.5 + 100 * np.sin(x / 97.0) * np.cos(y / 131.0)
    g = 127.5 + 100 * np.sin((x + y) / 61.0)
    b = 127.5 + 100 * np.cos(x / 23.0) * np.sin(y / 17.0)
    rgb = np.stack([r, g, b], -1) + rng.normal(0, 12, size=(n, n, 3)).astype(np.float32)
    a = 255 * ((np.sin(x / 257.0) * np.sin(y / 311.0)) > -0.3).astype(np.float32)
    a = np.where(((x // 64 + y // 64) % 7) == 0, 128 + 100 * np.sin(x / 11.0), a)
    return np.concatenate([rgb, a[..., None]], -1).clip(0, 255).astype(np.uint8)


def to_blocks(img):
    h, w, _ = img.shape
    return np.ascontiguousarray(img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 64))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(kernel):
    """dram read+write bytes per launch of `kernel` from the latest committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py from the raw CSV), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


def host_cores():
    """What this process may actually use: min(affinity, cgroup cpu.max quota). The 1-GPU lease of a 128-thread node is a cgroup
    with a 16-CPU quota although nproc and the affinity mask say 128."""
    aff = len(os.sched_getaffinity(0))
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            quota = int(q) / int(per)
    except Exception:
        pass
    used = max(1, min(aff, int(quota + 0.5)) if quota else aff)
    return {"nproc": os.cpu_count(), "affinity": aff, "cgroup_cpu_max": quota, "threads_used": used}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [s.strip() for s in out.split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def finish(self):
        self.stop_flag.set()
        self.join(2)
        return self.summary()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---- the reference on the host cores (oracle/_ref: the unmodified reference compiled here) -------------------------------------

def load_ref():
    if not os.path.exists(REF_SO):
        if os.path.isdir("/root/reference/encoder"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        else:
            raise RuntimeError("oracle/_ref/libbasisu_ref.so missing and /root/reference not present")
    lib = ctypes.CDLL(REF_SO)
    lib.ref_init()
    lib.ref_compress_image.restype = ctypes.c_void_p
    lib.ref_compress_image.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
    lib.ref_free.argtypes = [ctypes.c_void_p]
    lib.ref_uastc_rdo.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    return lib


def cpu_uastc(lib, blocks, flags, threads):
    out = np.empty((blocks.shape[0], 16), np.uint8)
    t0 = time.perf_counter()
    lib.ref_encode_uastc_blocks(blocks.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(blocks.shape[0]), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(flags), ctypes.c_uint32(threads))
    return time.perf_counter() - t0, out


def cpu_rdo(lib, uastc, src, lam, flags, jobs):
    out = np.ascontiguousarray(uastc).copy()
    t0 = time.perf_counter()
    ok = lib.ref_uastc_rdo(len(out), out.ctypes.data_as(ctypes.c_void_p), src.ctypes.data_as(ctypes.c_void_p), lam, 4096, 10.0, 8.0, 18.0, 10.0, flags, jobs, max(jobs, 4))
    assert ok
    return time.perf_counter() - t0, out


def compress(lib, img, flags, fmt=0, rdo_quality=0.0):
    size = ctypes.c_size_t(0)
    p = lib.ref_compress_image(fmt, img.ctypes.data_as(ctypes.c_void_p), img.shape[1], img.shape[0], flags, ctypes.c_float(rdo_quality), ctypes.byref(size))
    if not p:
        raise RuntimeError("basis_compress failed")
    data = ctypes.string_at(p, size.value)
    lib.ref_free(p)
    return data


def psnr_y(lib, data, img):
    out = np.zeros(img.shape, np.uint8)
    buf = np.frombuffer(data, np.uint8)
    ok = lib.ref_transcode_basis_to_rgba(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(data)), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(img.shape[0] * img.shape[1]))
    assert ok
    w = np.array([0.2126, 0.7152, 0.0722])
    la = out[..., :3].astype(np.float64) @ w
    lb = img[..., :3].astype(np.float64) @ w
    return float(10 * np.log10(255 ** 2 / np.mean((la - lb) ** 2)))


def kodim_images():
    from PIL import Image
    d = os.path.join(ROOT, "oracle", "_ref", "test_files")
    imgs = []
    for i in range(1, 25):
        f = os.path.join(d, f"kodim{i:02d}.png")
        if not os.path.exists(f):
            raise RuntimeError(f"{f} missing: __graft_entry__.build() copies the reference's test images there where /root/reference exists")
        imgs.append(np.ascontiguousarray(np.array(Image.open(f).convert("RGBA"))))
    return imgs


# ---- per-config descriptions -----------------------------------------------------------------------------------------------------

METRIC = {
    "uastc_l2": "Mtexels/sec encoded (4K RGBA UASTC LDR level 2)",
    "uastc_l2_strong": "Mtexels/sec encoded (one 4K RGBA image, UASTC LDR level 2, block rows over the GPUs)",
    "etc1s_kodim": "Mtexels/sec encoded (kodim01-24 ETC1S -q 128)",
    "etc1s_8k": "Mtexels/sec encoded (8192^2 RGBA ETC1S -q 255)",
    "uastc_l4_rdo": "Mtexels/sec encoded (4K RGBA tiles, UASTC LDR level 4 + RDO lambda 1.0)",
}


def config_dict(cfg, world):
    if cfg == "uastc_l2":
        return {"workload": "synthetic 4096x4096 RGBA8 (SURVEY 9.6 generator, seed 1234+rank), UASTC LDR 4x4 level 2, no RDO",
                "blocks_per_step_per_gpu": BLOCKS, "texels_per_step_per_gpu": TEXELS,
                "parallelism": f"{world} independent tile(s), one per GPU, no collective",
                "l2": f"{ROTATING_INPUTS} rotating 64 MiB inputs (> 126 MB L2) plus 1.8 GB of candidate scratch streamed per step"}
    if cfg == "uastc_l2_strong":
        return {"workload": "ONE synthetic 4096x4096 RGBA8 image (seed 1234), UASTC LDR 4x4 level 2, no RDO", "texels_per_step": TEXELS,
                "parallelism": f"block rows of the image split over {world} GPU(s) (sharding.block_range), no collective; host gathers 16 B/block",
                "l2": f"{ROTATING_INPUTS} rotating inputs; 1.8 GB / {world} of candidate scratch streamed per step"}
    if cfg == "etc1s_kodim":
        return {"workload": "kodim01-24 (24 x 768x512), ETC1S -q 128 (basis_compress defaults: comp_level 2), one basis_compress() per image through the patched library",
                "texels_per_step": 24 * 768 * 512, "parallelism": "1 GPU; images one after the other", "l2": "inputs 1.5 MB each (L2-resident); every step re-uploads them"}
    if cfg == "etc1s_8k":
        return {"workload": "synthetic 8192x8192 RGBA8 (seed 1234, alpha forced opaque), ETC1S -q 255, basis_compress() on every rank (replicated host logic)",
                "texels_per_step": 8192 * 8192,
                "parallelism": f"{world} rank(s): per-block stages by block-row ranges, per-cluster stages by clusters, stage outputs merged with one NCCL all-reduce each",
                "l2": "268 MB of source blocks per pass (> 126 MB L2)"}
    if cfg == "uastc_l4_rdo":
        return {"workload": "synthetic 4096x4096 RGBA8 tiles (seed 1234 + tile), UASTC LDR level 4 | favour-simpler-modes, then RDO lambda 1.0 with 4 chains per tile, device resident",
                "texels_per_step_per_gpu": TEXELS, "parallelism": f"{world} GPU(s), tiles spread over the GPUs (replicas over tiles: BASELINE config 5 is 64 tiles, 8 per GPU on 8 GPUs), no collective",
                "l2": "64 MiB input + 11.4 GB of level-4 candidate scratch streamed per tile"}
    raise ValueError(cfg)


# ---- the reference arm -----------------------------------------------------------------------------------------------------------

def reference_sample(cfg, lib, cores):
    """One bounded sample of the config's workload on the host cores. Returns (seconds, texels, description)."""
    threads = cores["threads_used"]
    if cfg in ("uastc_l2", "uastc_l2_strong"):
        blocks = to_blocks(synth(IMAGE_DIM, 1234))
        n = blocks.shape[0] if threads >= 32 else blocks.shape[0] // 8
        t, _ = cpu_uastc(lib, blocks[:n], 2, threads)
        return t, n * 16, f"first {n} of {blocks.shape[0]} blocks of the synthetic 4096^2 image (seed 1234), encode_uastc in 256-block jobs on {threads} threads"
    if cfg == "uastc_l4_rdo":
        img = synth(IMAGE_DIM, 1234)[:512, :512]
        src = to_blocks(img)
        t0, enc = cpu_uastc(lib, src, 4 | cPackUASTCFavorSimplerModes, threads)
        t1, _ = cpu_rdo(lib, enc, src, 1.0, 4 | cPackUASTCFavorSimplerModes, 4)
        return t0 + t1, src.shape[0] * 16, f"512x512 crop of the synthetic tile: encode_uastc level 4 on {threads} threads ({t0:.2f} s) + uastc_rdo with 4 jobs ({t1:.2f} s)"
    if cfg == "etc1s_kodim":
        imgs = kodim_images()
        t0 = time.perf_counter()
        for im in imgs:
            compress(lib, im, 128 | cFlagThreaded)
        return time.perf_counter() - t0, sum(im.shape[0] * im.shape[1] for im in imgs), "all 24 images through basis_compress(cETC1S, q128, cFlagThreaded): the reference sizes its own job pool (hardware_concurrency)"
    if cfg == "etc1s_8k":
        img = synth(2048, 1234)
        img[..., 3] = 255
        img = np.ascontiguousarray(img)
        t0 = time.perf_counter()
        compress(lib, img, 255 | cFlagThreaded)
        return time.perf_counter() - t0, 2048 * 2048, "2048x2048 instance of the same generator (1/16 of the texels) through basis_compress(cETC1S, q255, cFlagThreaded)"
    raise ValueError(cfg)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    lib = load_ref()
    for _ in range(min(args.warmup, 1)):
        reference_sample(args.config, lib, cores)
    t = 0.0
    texels = 0
    desc = ""
    for _ in range(args.steps):
        dt, tx, desc = reference_sample(args.config, lib, cores)
        t += dt
        texels += tx
    mtex = texels / 1e6 / t
    line = {"impl": "reference", "metric": METRIC[args.config], "value": mtex, "unit": "Mtexel/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.config in ("uastc_l2_strong", "etc1s_8k") else "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic" if "kodim" not in args.config else "kodim01-24",
            "config": config_dict(args.config, args.gpus),
            "cpu_baseline": {"value": mtex, "unit": "Mtexel/s", "cores": cores["threads_used"], "kind": "reference", "sample": desc, "host": cores},
            "e2e": {"value": mtex, "unit": "Mtexel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- GPU arms --------------------------------------------------------------------------------------------------------------------

class Dist:
    def __init__(self, local_rank, world):
        import torch
        self.torch = torch
        self.world = world
        self.dist = None
        torch.cuda.set_device(local_rank)
        if world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def sum_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def broadcast_bytes(self, arr):
        if self.dist is None:
            return arr
        t = self.torch.from_numpy(arr.copy()).cuda()
        self.dist.broadcast(t, 0)
        return t.cpu().numpy()

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def run_uastc(args, rank, local_rank, world, strong):
    import torch
    from basis_universal_b200 import uastc, sharding
    D = Dist(local_rank, world)
    enc = uastc.Encoder(local_rank)
    LEVEL = 2
    if strong:
        first, last = sharding.block_range(IMAGE_DIM // 4, IMAGE_DIM // 4, rank, world)   # block rows of this rank
        nblk = last - first
    else:
        first, nblk = 0, BLOCKS
    host_in = []
    for k in range(ROTATING_INPUTS):
        seed = (1234 + 1000 * k) if strong else (1234 + rank + 1000 * k)
        b = to_blocks(synth(IMAGE_DIM, seed))
        host_in.append(torch.from_numpy(np.ascontiguousarray(b[first:first + nblk])).pin_memory())
    dev_in = [t.cuda() for t in host_in]
    dev_out = torch.empty((max(nblk, 1), 16), dtype=torch.uint8, device="cuda")
    host_out = torch.empty((max(nblk, 1), 16), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    for w in range(args.warmup):
        enc.encode_uastc_device(dev_in[w % ROTATING_INPUTS].data_ptr(), nblk, dev_out.data_ptr(), LEVEL)
    sampler = ClockSampler(local_rank)
    D.barrier()
    sampler.start()
    launches = 0
    stage = [0.0, 0.0, 0.0]
    enc.timer_start()
    for s in range(args.steps):
        enc.encode_uastc_device(dev_in[s % ROTATING_INPUTS].data_ptr(), nblk, dev_out.data_ptr(), LEVEL)
        launches += enc.last_launch_count
        for i in range(3):
            stage[i] += enc.stage_ms(i)
    dev_ms = enc.timer_stop_ms()
    D.barrier()
    clocks = sampler.finish()

    for w in range(min(args.warmup, 3)):
        enc.encode_uastc_host_ptr(host_in[w % ROTATING_INPUTS].data_ptr(), nblk, host_out.data_ptr(), LEVEL)
    D.barrier()
    enc.timer_start()
    t0 = time.perf_counter()
    for s in range(args.steps):
        enc.encode_uastc_host_ptr(host_in[s % ROTATING_INPUTS].data_ptr(), nblk, host_out.data_ptr(), LEVEL)
    e2e_dev_ms = enc.timer_stop_ms()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    D.barrier()
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms)  # the host-visible time includes the final D2H completion
    dev_ms_max, e2e_ms_max = D.max_over_ranks([dev_ms, e2e_ms])

    if rank == 0:
        peak, peak_kind = measured_peaks()
        k1_ms = stage[1] / args.steps
        achieved = 80 * nblk / (k1_ms * 1e-3) / 1e9   # read 64 B + write 16 B per block
        total_texels = TEXELS if strong else world * TEXELS
        value = total_texels * args.steps / 1e6 / (dev_ms_max * 1e-3)
        e2e_value = total_texels * args.steps / 1e6 / (e2e_ms_max * 1e-3)
        cores = host_cores()
        ref = load_ref()
        blocks0 = to_blocks(synth(IMAGE_DIM, 1234))
        n = blocks0.shape[0] if cores["threads_used"] >= 32 else blocks0.shape[0] // 8
        cpu_t, cpu_out = cpu_uastc(ref, blocks0[:n], LEVEL, cores["threads_used"])
        # parity spot check while we are here: the reference's bytes for the sample equal the GPU's
        chk = torch.from_numpy(np.ascontiguousarray(blocks0[:n])).cuda()
        chk_out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        enc.encode_uastc_device(chk.data_ptr(), n, chk_out.data_ptr(), LEVEL)
        parity = bool(np.array_equal(chk_out.cpu().numpy(), cpu_out))
        cfg = "uastc_l2_strong" if strong else "uastc_l2"
        line = {"metric": METRIC[cfg], "value": value, "unit": "Mtexel/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic", "config": config_dict(cfg, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mtexel/s", "h2d_bytes_per_step": nblk * 64, "d2h_bytes_per_step": nblk * 16, "ms_per_step": e2e_ms_max / args.steps},
                "gpu_launches": launches,
                "stage_ms_per_step": {"classify_rank": stage[0] / args.steps, "candidates": k1_ms, "finish": stage[2] / args.steps},
                "roofline": {"bound": "hbm", "kernel": "k_candidates", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic("k_candidates"), "peak_source": peak_kind,
                             "note": "k_candidates = the three work-list launches of one step, timed together with CUDA events; algorithmic 80 B/block (SURVEY 8d); "
                                     "traffic = ncu dram read+write of those launches in the latest committed capture; the stage is instruction-issue bound (DESIGN.md section 4)"},
                "cpu_baseline": {"value": n * 16 / 1e6 / cpu_t, "unit": "Mtexel/s", "cores": cores["threads_used"], "kind": "reference", "host": cores,
                                 "sample": f"first {n} of {blocks0.shape[0]} blocks of the synthetic 4096^2 image, seed 1234"},
                "bit_exact_vs_reference_on_cpu_sample": parity}
        print(json.dumps(line), flush=True)
    D.close()


def run_uastc_l4_rdo(args, rank, local_rank, world):
    """BASELINE config 5: each GPU owns `args.tiles` 4096^2 tiles of the atlas (8 when the 64 tiles are spread over 8 GPUs). A step encodes
    them one after the other (level 4, 2^18 blocks per pass) and post-processes all of them in ONE RDO launch: 4 chains per tile, every
    chain of every tile concurrent (a chain is sequential, so its latency is paid once per step, not once per tile)."""
    import torch
    from basis_universal_b200 import uastc
    D = Dist(local_rank, world)
    enc = uastc.Encoder(local_rank)
    FLAGS = 4 | cPackUASTCFavorSimplerModes
    params = uastc.uastc_rdo_params(lambda_=1.0)
    T = args.tiles
    src = np.concatenate([to_blocks(synth(IMAGE_DIM, 1234 + rank * T + t)) for t in range(T)])
    host_in = torch.from_numpy(src).pin_memory()
    dev_in = host_in.cuda()
    dev_out = torch.empty((T * BLOCKS, 16), dtype=torch.uint8, device="cuda")
    host_out = torch.empty((T * BLOCKS, 16), dtype=torch.uint8).pin_memory()
    counts = [BLOCKS] * T
    torch.cuda.synchronize()

    def step_device():
        enc.encode_uastc_device(dev_in.data_ptr(), T * BLOCKS, dev_out.data_ptr(), FLAGS)
        enc_ms, la = enc.last_kernel_ms, enc.last_launch_count
        enc.uastc_rdo_batch_device(dev_out.data_ptr(), dev_in.data_ptr(), counts, params, FLAGS, 4)
        return enc_ms, enc.last_kernel_ms, la + enc.last_launch_count

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    D.barrier()
    sampler.start()
    enc.timer_start()
    enc_ms = rdo_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        a, b, la = step_device()
        enc_ms += a; rdo_ms += b; launches += la
    dev_ms = enc.timer_stop_ms()
    D.barrier()
    clocks = sampler.finish()

    # e2e: host blocks in, RDO'd UASTC blocks out, through the host-pointer entry point that chains both stages on the device
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        enc.encode_rdo_host_ptr(host_in.data_ptr(), counts, host_out.data_ptr(), params, FLAGS, 4)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    D.barrier()
    dev_ms_max, e2e_ms_max = D.max_over_ranks([dev_ms, e2e_ms])
    if rank == 0:
        peak, peak_kind = measured_peaks()
        cores = host_cores()
        ref = load_ref()
        cpu_t, cpu_tx, cpu_desc = reference_sample("uastc_l4_rdo", ref, cores)
        # parity on the CPU sample: the same 512x512 crop through the device path
        crop = to_blocks(synth(IMAGE_DIM, 1234)[:512, :512])
        _, want_enc = cpu_uastc(ref, crop, FLAGS, cores["threads_used"])
        _, want = cpu_rdo(ref, want_enc, crop, 1.0, FLAGS, 4)
        got = enc.uastc_rdo_batch(enc.encode_uastc(crop, FLAGS), crop, [crop.shape[0]], params, FLAGS, 4)
        texels = world * T * TEXELS * args.steps
        achieved = (64 + 16 + 64 + 16 + 16) * T * BLOCKS / ((enc_ms + rdo_ms) / args.steps * 1e-3) / 1e9   # 9 B/texel read + 2 written
        cfg = config_dict("uastc_l4_rdo", world)
        cfg["tiles_per_gpu_per_step"] = T
        cfg["texels_per_step_per_gpu"] = T * TEXELS
        line = {"metric": METRIC["uastc_l4_rdo"], "value": texels / 1e6 / (dev_ms_max * 1e-3), "unit": "Mtexel/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/f32/f64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": texels / 1e6 / (e2e_ms_max * 1e-3), "unit": "Mtexel/s", "h2d_bytes_per_step": T * BLOCKS * 64, "d2h_bytes_per_step": T * BLOCKS * 16,
                        "ms_per_step": e2e_ms_max / args.steps},
                "gpu_launches": launches, "stage_ms_per_step": {"encode_level4": enc_ms / args.steps, "rdo_chains_plus_rehint": rdo_ms / args.steps},
                "roofline": {"bound": "hbm", "kernel": "encode (k_candidates x3 + k_finish + k_classify_rank per 2^18-block pass) + k_rdo_chain + k_rdo_rehint", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_kind,
                             "note": "algorithmic 11 B/texel (SURVEY 8d: UASTC + RDO) over the whole step; level 4 runs 170 candidate slots per block, instruction-bound; "
                                     "the RDO launch is latency-bound (4 sequential chains per tile, one CTA each)"},
                "cpu_baseline": {"value": cpu_tx / 1e6 / cpu_t, "unit": "Mtexel/s", "cores": cores["threads_used"], "kind": "reference", "sample": cpu_desc, "host": cores},
                "bit_exact_vs_reference_on_cpu_sample": bool(np.array_equal(got, want))}
        print(json.dumps(line), flush=True)
    D.close()


def load_dropin():
    if not os.path.exists(DROPIN_SO):
        raise RuntimeError(f"{DROPIN_SO} missing: built by integration/Makefile where /root/reference exists (there is no CPU fallback for the GPU arm)")
    lib = ctypes.CDLL(DROPIN_SO)
    lib.ref_compress_image.restype = ctypes.c_void_p
    lib.ref_compress_image.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
    lib.ref_free.argtypes = [ctypes.c_void_p]
    lib.b200_dropin_stage_secs.restype = ctypes.c_double
    lib.b200_dropin_stage_secs.argtypes = [ctypes.c_char_p]
    if lib.ref_init_gpu_seam() != 1:
        raise RuntimeError("the drop-in's opencl_is_available() is false: no usable B200")
    return lib


def etc1s_roofline(stats0, stats1, texels_per_pass, passes):
    """Dominant per-block kernel family between two global-stats snapshots, against SURVEY 8(d)'s 64 B/block per pass."""
    fam = {}
    for k in stats1:
        ms = stats1[k][0] - stats0[k][0]
        la = stats1[k][1] - stats0[k][1]
        ca = stats1[k][2] - stats0[k][2]
        fam[k] = {"kernel_ms": ms, "launches": la, "calls": ca}
    per_block = ["etc1s_encode_blocks", "etc1s_refine", "etc1s_determine_selectors", "etc1s_find_selector_clusters", "etc1s_endpoint_clusters", "etc1s_selector_codebook"]
    dom = max(per_block, key=lambda k: fam[k]["kernel_ms"])
    return fam, dom


def run_etc1s(args, rank, local_rank, world, cfg):
    import torch
    from basis_universal_b200 import _lib as b200
    D = Dist(local_rank, world)
    os.environ["B200_DEVICE"] = str(local_rank)
    if world > 1:
        from basis_universal_b200 import etc1s
        uid = etc1s.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
        uid = D.broadcast_bytes(uid)
        os.environ["B200_COMM_WORLD"] = str(world)
        os.environ["B200_COMM_RANK"] = str(rank)
        os.environ["B200_COMM_ID"] = bytes(uid).hex()
    lib = load_dropin()
    if cfg == "etc1s_kodim":
        imgs = kodim_images()
        q = 128
    else:
        im = synth(8192, 1234)
        im[..., 3] = 255
        imgs = [np.ascontiguousarray(im)]
        q = 255
    texels = sum(i.shape[0] * i.shape[1] for i in imgs)
    flags = q | cFlagThreaded | cFlagUseOpenCL

    def step():
        fe = be = 0.0
        out = []
        for im in imgs:
            out.append(compress(lib, im, flags))
            fe += lib.b200_dropin_stage_secs(b"frontend")
            be += lib.b200_dropin_stage_secs(b"backend")
        return fe, be, out

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    D.barrier()
    sampler.start()
    s0 = b200.global_stats()
    launches0 = b200.lib().b200_global_launch_count()
    t0 = time.perf_counter()
    fe = be = 0.0
    files = None
    for _ in range(args.steps):
        a, b, files = step()
        fe += a; be += b
    wall = time.perf_counter() - t0
    s1 = b200.global_stats()
    launches = b200.lib().b200_global_launch_count() - launches0
    if launches < 8 * len(imgs) * args.steps:
        raise RuntimeError(f"only {launches} kernel launches in the timed region: the compressor fell back to its CPU path (see stderr); no GPU number to report")
    D.barrier()
    clocks = sampler.finish()
    fe_max, wall_max = D.max_over_ranks([fe, wall])
    if rank == 0:
        peak, peak_kind = measured_peaks()
        fam, dom = etc1s_roofline(s0, s1, texels, 6)
        kernel_ms = sum(v["kernel_ms"] for v in fam.values())
        dom_launch_ms = fam[dom]["kernel_ms"] / max(fam[dom]["launches"], 1)
        blocks_per_launch = (texels / 16 / len(imgs)) / (world if dom not in ("etc1s_endpoint_clusters", "etc1s_selector_codebook") else 1)
        achieved = 64 * blocks_per_launch / (dom_launch_ms * 1e-3) / 1e9 if dom_launch_ms > 0 else 0.0
        cores = host_cores()
        ref = load_ref()
        # CPU side by side (bounded: the kodim batch once; a 2048^2 instance for the 8K config) and the PSNR gate on what was just produced
        cpu_t, cpu_tx, cpu_desc = reference_sample(cfg, ref, cores)
        gate = {}
        if cfg == "etc1s_kodim":
            deltas = []
            for im, f in zip(imgs, files):
                c = compress(ref, im, q | cFlagThreaded)
                deltas.append(psnr_y(ref, f, im) - psnr_y(ref, c, im))
            gate = {"psnr_y_delta_vs_cpu_db": {"max_abs": float(np.max(np.abs(deltas))), "mean": float(np.mean(deltas))}, "gate_db": 0.02}
        else:
            small = np.ascontiguousarray(imgs[0][:1024, :1024])
            g = compress(lib, small, flags) if world == 1 else None
            if g is not None:
                c = compress(ref, small, q | cFlagThreaded)
                gate = {"psnr_y_delta_vs_cpu_db": {"max_abs": abs(psnr_y(ref, g, small) - psnr_y(ref, c, small)), "on": "1024x1024 crop"}, "gate_db": 0.02}
        line = {"metric": METRIC[cfg], "value": texels * args.steps / 1e6 / fe_max, "unit": "Mtexel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * fe_max / args.steps, "higher_is_better": True, "scaling": "strong" if cfg == "etc1s_8k" else "weak", "vs_baseline": None, "dtype": "u8/u32/f32",
                "data": "kodim01-24" if cfg == "etc1s_kodim" else "synthetic", "config": config_dict(cfg, world), "clocks": clocks,
                "value_definition": "texels / wall seconds of basis_compressor::process_frontend (source blocks uploaded inside it): the ETC1S frontend with every per-block stage, both "
                                    "clusterers, the endpoint-cluster optimiser and the selector codebook on the GPU; its host orchestration is the reference's own frontend code",
                "e2e": {"value": texels * args.steps / 1e6 / wall_max, "unit": "Mtexel/s", "h2d_bytes_per_step": texels * 4, "d2h_bytes_per_step": int(sum(len(f) for f in files)),
                        "ms_per_step": 1e3 * wall_max / args.steps, "call": "basis_compress(cETC1S, RGBA host image in, .basis bytes out) on the patched library",
                        "backend_ms_per_step": 1e3 * be / args.steps},
                "gpu_launches": int(launches), "kernel_ms_per_step": kernel_ms / args.steps,
                "stage_kernels_per_step": {k: {"ms": v["kernel_ms"] / args.steps, "launches": v["launches"] // args.steps} for k, v in fam.items() if v["launches"]},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic({"etc1s_find_selector_clusters": "k_etc1s_fosc"}.get(dom, "k_" + dom)), "peak_source": peak_kind,
                             "note": "dominant per-block kernel family of the frontend; algorithmic 64 B per block per pass (SURVEY 8d: 24 B/texel over the six mandatory passes); "
                                     "average launch duration from CUDA events inside the library"},
                "cpu_baseline": {"value": cpu_tx / 1e6 / cpu_t, "unit": "Mtexel/s", "cores": cores["threads_used"], "kind": "reference", "sample": cpu_desc, "host": cores}}
        line.update(gate)
        if world > 1:
            line["collective"] = "one ncclAllReduce(sum, u32) per stage call merges the ranks' shares of the stage's output array (4-8 B per block / cluster)"
        print(json.dumps(line), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="uastc_l2", choices=sorted(METRIC))
    ap.add_argument("--tiles", type=int, default=8, help="uastc_l4_rdo: 4096^2 tiles per GPU per step (config 5: 64 tiles over 8 GPUs)")
    args = ap.parse_args()
    if args.warmup is None:
        args.warmup = 1 if args.config in ("etc1s_8k", "uastc_l4_rdo") else 3
    if args.steps is None:
        args.steps = {"uastc_l2": 10, "uastc_l2_strong": 10, "etc1s_kodim": 3, "etc1s_8k": 1, "uastc_l4_rdo": 1}[args.config]
    if args.impl == "b200":
        # the timing rules ask for >= 3 warm-up steps; a step of the two long configurations is 6-15 s, so they warm up once
        args.warmup = max(args.warmup, 1 if args.config in ("etc1s_8k", "uastc_l4_rdo") else 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torch.distributed.run with {args.gpus} ranks (got WORLD_SIZE={world})", file=sys.stderr)
        sys.exit(2)
    if args.config in ("uastc_l2", "uastc_l2_strong"):
        run_uastc(args, rank, local_rank, world, args.config == "uastc_l2_strong")
    elif args.config == "uastc_l4_rdo":
        run_uastc_l4_rdo(args, rank, local_rank, world)
    else:
        run_etc1s(args, rank, local_rank, world, args.config)


if __name__ == "__main__":
    main()
