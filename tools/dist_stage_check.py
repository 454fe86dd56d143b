"""N ranks, one GPU each, NCCL through the library's own communicator (b200_comm_*): every sharded ETC1S stage call returns
on EVERY rank exactly what a single GPU returns for the whole slice, and a whole basis_compress() through the drop-in
produces the same .basis bytes on every rank as the single-GPU run and as the CPU encoder's PSNR gate allows.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_stage_check.py"""
import ctypes
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from basis_universal_b200 import etc1s  # noqa: E402
from bench import synth, to_blocks, load_dropin, compress, cFlagThreaded, cFlagUseOpenCL  # noqa: E402
import util  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = etc1s.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
    t = torch.from_numpy(uid.copy()).cuda()
    dist.broadcast(t, 0)
    uid = t.cpu().numpy()

    blocks = to_blocks(synth(1024, 4321))
    n = blocks.shape[0]
    single = etc1s.Etc1sContext(local)          # no communicator: the whole slice on this GPU
    sharded = etc1s.Etc1sContext(local)
    sharded.comm_init(rank, world, uid)
    single.set_pixel_blocks(blocks)
    sharded.set_pixel_blocks(blocks)
    checks = {}
    a, b = single.encode_etc1s_blocks(False, 16), sharded.encode_etc1s_blocks(False, 16)
    checks["encode_etc1s_blocks"] = np.array_equal(a, b)
    inp = util.etc1s_stage_inputs(blocks, 3)
    checks["refine"] = np.array_equal(single.refine_endpoint_clusterization(inp["block_info"], inp["cluster_info"], inp["sorted_idx"], False),
                                      sharded.refine_endpoint_clusterization(inp["block_info"], inp["cluster_info"], inp["sorted_idx"], False))
    checks["find_selector_clusters"] = np.array_equal(single.find_optimal_selector_clusters_for_each_block(inp["fosc_blocks"], inp["selectors"], inp["sel_cluster_idx"], False),
                                                      sharded.find_optimal_selector_clusters_for_each_block(inp["fosc_blocks"], inp["selectors"], inp["sel_cluster_idx"], False))
    checks["determine_selectors"] = np.array_equal(single.determine_selectors(inp["color5_inten"], False), sharded.determine_selectors(inp["color5_inten"], False))
    rng = np.random.default_rng(5)
    order = rng.permutation(n).astype(np.uint32)
    clusters = [c for c in np.split(order, np.sort(rng.choice(np.arange(1, n), 300, replace=False)))]
    checks["endpoint_clusters"] = np.array_equal(single.encode_endpoint_clusters(clusters, False, 16), sharded.encode_endpoint_clusters(clusters, False, 16))
    checks["selector_codebook"] = np.array_equal(single.optimize_selector_codebook(a, clusters, False), sharded.optimize_selector_codebook(a, clusters, False))
    sels = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    cur = np.stack([a[[c[0] for c in clusters], 0] >> 3, a[[c[0] for c in clusters], 1] >> 3, a[[c[0] for c in clusters], 2] >> 3, a[[c[0] for c in clusters], 3] >> 5], -1).astype(np.uint8)
    for name, sw in (("reoptimize_forced", sels), ("reoptimize_free", None)):
        ra, rb = single.reoptimize_endpoint_clusters(clusters, sw, cur, False, 64), sharded.reoptimize_endpoint_clusters(clusters, sw, cur, False, 64)
        checks[name] = all(np.array_equal(x, y) for x, y in zip(ra, rb))
    per_block = np.stack([a[:, 0] >> 3, a[:, 1] >> 3, a[:, 2] >> 3, a[:, 3] >> 5], -1).astype(np.uint8)
    checks["subblock_errors"] = np.array_equal(single.subblock_errors(per_block, False), sharded.subblock_errors(per_block, False))
    key = ((a[:, 0] >> 5).astype(np.uint32) << 9) | ((a[:, 1] >> 5).astype(np.uint32) << 6) | ((a[:, 2] >> 5).astype(np.uint32) << 3) | (a[:, 3] >> 5)
    uniq, idx0 = np.unique(key, return_inverse=True)
    cb = np.stack([((uniq >> 9) & 7) * 4 + 2, ((uniq >> 6) & 7) * 4 + 2, ((uniq >> 3) & 7) * 4 + 2, uniq & 7], -1).astype(np.uint8)
    pa = single.backend_endpoint_prediction([(0, 256, 256)], a, cb, idx0.astype(np.uint32), 1.5, False)
    pb = sharded.backend_endpoint_prediction([(0, 256, 256)], a, cb, idx0.astype(np.uint32), 1.5, False)
    checks["backend_endpoint_prediction"] = np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
    cs = sharded.comm_stats()
    single.close()
    sharded.close()

    # whole encoder: single-GPU file first (no communicator in the environment), then the sharded run
    os.environ["B200_DEVICE"] = str(local)
    lib = load_dropin()
    img = synth(1024, 99)
    img[..., 3] = 255
    img = np.ascontiguousarray(img)
    f_single = compress(lib, img, 200 | cFlagThreaded | cFlagUseOpenCL)
    os.environ["B200_COMM_WORLD"], os.environ["B200_COMM_RANK"] = str(world), str(rank)
    uid2 = etc1s.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
    t = torch.from_numpy(uid2.copy()).cuda()
    dist.broadcast(t, 0)
    os.environ["B200_COMM_ID"] = bytes(t.cpu().numpy()).hex()
    f_sharded = compress(lib, img, 200 | cFlagThreaded | cFlagUseOpenCL)
    checks["basis_compress_bytes"] = f_single == f_sharded
    digest = torch.tensor(list(hashlib.md5(f_sharded).digest()), dtype=torch.int32, device="cuda")
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    checks["same_file_on_every_rank"] = bool((lo == hi).all().item())

    ok = all(checks.values())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_stage_check: world {world}: {checks}; merges so far on rank 0: {cs}; all ranks ok: {bool(flag.item())}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
