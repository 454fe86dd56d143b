"""Two (or N) ranks, one GPU each, NCCL: the sharded ETC1S endpoint training set equals the single-GPU one.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_etc1s_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from basis_universal_b200 import distributed, etc1s  # noqa: E402
from bench import synth, to_blocks  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dim = 1024
    img = synth(dim, 4321)
    blocks = to_blocks(img)
    nbx = nby = dim // 4
    ctx = etc1s.Etc1sContext(local)
    local_blocks, (keys, vecs, weights) = distributed.global_endpoint_training_set(ctx, nbx, nby, blocks, True, 16, rank, world)
    # single-GPU answer for the whole image, computed redundantly on every rank
    ctx.set_pixel_blocks(blocks)
    whole = ctx.encode_etc1s_blocks(True, 16)
    hist = ctx.endpoint_histogram(whole)
    k2, v2, w2 = etc1s.training_vectors_from_histogram(hist)
    ok = np.array_equal(keys, k2) and np.array_equal(vecs, v2) and np.array_equal(weights, w2) and int(weights.sum()) == 2 * blocks.shape[0]
    # second exchange: selector training set (all-gatherv of unique keys + weights)
    lk, lw = etc1s.merge_selector_training(*ctx.selector_training(local_blocks, True))
    gk, gw = distributed.allgather_selector_training(torch.from_numpy(lk.astype(np.int64)).cuda(), torch.from_numpy(lw.astype(np.int64)).cuda())
    wk, ww = etc1s.merge_selector_training(*ctx.selector_training(whole, True))
    ok_sel = np.array_equal(gk.cpu().numpy(), wk.astype(np.int64)) and np.array_equal(gw.cpu().numpy(), ww.astype(np.int64))
    ok = ok and ok_sel
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_etc1s_check: world {world}, {len(keys)} unique training vectors, total weight {int(weights.sum())}, {len(wk)} unique selector vectors; identical on all ranks: {bool(flag.item())}", flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
