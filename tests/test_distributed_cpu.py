"""world_size-2 gloo test (CPU) of the N>1 host logic: each rank owns a contiguous block-row range, encodes nothing here
(no GPU) but proves the partition is a disjoint cover and that per-rank unit counts reduce to the whole job, which is what
bench.py reports as `value` = all units / max-over-ranks time."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from basis_universal_b200 import sharding

import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nbx, nby, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = sharding.block_range(nbx, nby, rank, world)
    owned = torch.zeros(nbx * nby, dtype=torch.int32)
    owned[first:last] = 1
    dist.all_reduce(owned)                       # every block owned exactly once
    units = torch.tensor([float((last - first) * 16)])
    t = torch.tensor([0.5 + rank])               # fake per-rank step time
    dist.all_reduce(units)                       # whole-job texels
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks
    if rank == 0:
        q.put((bool((owned == 1).all()), float(units.item()), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_block_row_sharding_gloo():
    world, nbx, nby = 2, 37, 51
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nbx, nby, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, units, tmax = q.get()
    assert ok and units == nbx * nby * 16 and tmax == 1.5


def _hist_worker(rank, world, port, q):
    """gloo all-reduce of per-rank endpoint-key histograms == histogram of the whole image (the ETC1S exchange step)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from basis_universal_b200 import distributed, etc1s
    rng = np.random.default_rng(5)           # same stream on every rank: the "image" of ETC1S blocks
    nbx, nby = 24, 19
    blocks = rng.integers(0, 256, (nbx * nby, 8), dtype=np.uint8)
    first, last = sharding.block_range(nbx, nby, rank, world)
    local = torch.from_numpy(np.bincount(util.endpoint_keys(blocks[first:last]), minlength=1 << 18).astype(np.int32) * 2)
    distributed.allreduce_endpoint_histogram(local)
    whole = np.bincount(util.endpoint_keys(blocks), minlength=1 << 18) * 2
    keys, vecs, weights = etc1s.training_vectors_from_histogram(local.numpy().astype(np.uint32))
    ok = bool(np.array_equal(local.numpy(), whole)) and int(weights.sum()) == 2 * nbx * nby and vecs.shape == (len(keys), 6)
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


def test_two_rank_endpoint_histogram_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_hist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get()


def _sel_worker(rank, world, port, q):
    """gloo all-gatherv of per-rank unique selector keys + weights == merge over the whole image (second ETC1S exchange)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from basis_universal_b200 import distributed, etc1s
    rng = np.random.default_rng(8)
    nbx, nby = 40, 31
    keys = rng.integers(0, 50, nbx * nby).astype(np.uint32) * np.uint32(0x01010101)   # few distinct keys, many duplicates
    weights = rng.integers(1, 4097, nbx * nby).astype(np.uint32)
    first, last = sharding.block_range(nbx, nby, rank, world)
    lk, lw = etc1s.merge_selector_training(keys[first:last], weights[first:last])
    gk, gw = distributed.allgather_selector_training(torch.from_numpy(lk.astype(np.int64)), torch.from_numpy(lw.astype(np.int64)))
    wk, ww = etc1s.merge_selector_training(keys, weights)
    ok = np.array_equal(gk.numpy(), wk.astype(np.int64)) and np.array_equal(gw.numpy(), ww.astype(np.int64))
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


def test_two_rank_selector_training_allgather_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sel_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get()


def _merge_worker(rank, world, port, n_blocks, n_clusters, q):
    """The multi-GPU ETC1S pattern of b200_dist.cu on the CPU: every rank fills ITS share of a stage's output array (per-block
    stages: the contiguous range of b200_shard_range; per-cluster stages: every world-th entry) into a zeroed buffer, and one SUM
    all-reduce over the array as u32 words is a merge, because each element is written by exactly one rank (no carries between the
    two halves of a 64-bit etc_block either)."""
    import ctypes
    from basis_universal_b200 import lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)                              # same "full result" on every rank
    full_blocks = rng.integers(0, 1 << 63, n_blocks, dtype=np.uint64)   # 8-byte per-block outputs (etc_blocks)
    full_clusters = rng.integers(0, 1 << 32, n_clusters, dtype=np.uint64).astype(np.uint32)
    first, last = ctypes.c_uint32(0), ctypes.c_uint32(0)
    lib().b200_shard_range(n_blocks, rank, world, ctypes.byref(first), ctypes.byref(last))
    mine = np.zeros(n_blocks, np.uint64)
    mine[first.value:last.value] = full_blocks[first.value:last.value]
    words = torch.from_numpy(mine.view(np.uint32).astype(np.int64))     # gloo has no u32 sum; int64 holds every partial sum exactly
    dist.all_reduce(words)
    merged_blocks = words.numpy().astype(np.uint32).view(np.uint64)
    minec = np.zeros(n_clusters, np.uint32)
    minec[rank::world] = full_clusters[rank::world]
    wc = torch.from_numpy(minec.astype(np.int64))
    dist.all_reduce(wc)
    sizes = torch.tensor([last.value - first.value], dtype=torch.int64)
    dist.all_reduce(sizes)
    if rank == 0:
        q.put((bool(np.array_equal(merged_blocks, full_blocks)), bool(np.array_equal(wc.numpy().astype(np.uint32), full_clusters)), int(sizes.item())))
    dist.destroy_process_group()


def test_stage_output_merge_by_allreduce_gloo():
    world, n_blocks, n_clusters = 2, 1001, 77
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, n_blocks, n_clusters, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_blocks, ok_clusters, covered = q.get()
    assert ok_blocks and ok_clusters and covered == n_blocks


def test_shard_range_is_a_disjoint_cover():
    import ctypes
    from basis_universal_b200 import lib
    for n in (0, 1, 7, 64, 1001, 1 << 20):
        for world in (1, 2, 3, 8):
            covered = 0
            prev_last = 0
            for rank in range(world):
                first, last = ctypes.c_uint32(0), ctypes.c_uint32(0)
                lib().b200_shard_range(n, rank, world, ctypes.byref(first), ctypes.byref(last))
                assert first.value == prev_last or first.value == n
                assert first.value <= last.value <= n
                covered += last.value - first.value
                prev_last = last.value
            assert covered == n
