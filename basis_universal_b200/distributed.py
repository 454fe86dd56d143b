"""torch.distributed helpers for hosts that shard the ETC1S training sets themselves: a SUM all-reduce of the
2^18-bin endpoint-key histogram and an all-gatherv of the selector training pairs give every rank the global training set, after
which the (deterministic) clusterer runs replicated. The drop-in path does not need them -- there the library's own communicator
(b200_comm_*, csrc/b200_dist.cu) merges each stage's output -- and UASTC needs no collective at all.
Contract of the *_device entry points these helpers call: they run on the context's own (non-blocking) stream, so the caller
synchronises its stream before the call (torch.cuda.synchronize below) and the call returns only after its own stream is idle."""
import numpy as np

from . import etc1s, sharding


def allreduce_endpoint_histogram(local_hist, group=None):
    """local_hist: torch tensor (2**18,) of int32/int64 counts on this rank's device (cuda for NCCL, cpu for gloo).
    In-place SUM all-reduce; a no-op when torch.distributed is not initialised (single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_hist, op=dist.ReduceOp.SUM, group=group)
    return local_hist


def global_endpoint_training_set(ctx, blocks_x, blocks_y, all_blocks, perceptual, total_perms, rank, world_size, group=None):
    """Rank-local ETC1S block optimisation of this rank's block rows on the GPU, histogram of its endpoint keys on the GPU,
    NCCL all-reduce, then the global (keys, vec6F, weights) -- identical on every rank."""
    import torch
    first, last = sharding.block_range(blocks_x, blocks_y, rank, world_size)
    local = np.ascontiguousarray(all_blocks[first:last])
    dev = torch.device("cuda", ctx.device)                  # the context's device, not torch's current one
    d_hist = torch.zeros(1 << 18, dtype=torch.int32, device=dev)
    if local.shape[0]:
        ctx.set_pixel_blocks(local)
        etc_blocks = ctx.encode_etc1s_blocks(perceptual, total_perms)
        d_blocks = torch.from_numpy(etc_blocks).to(dev)
        torch.cuda.synchronize(dev)
        ctx.endpoint_histogram_device(d_blocks.data_ptr(), etc_blocks.shape[0], d_hist.data_ptr())
    else:
        etc_blocks = np.zeros((0, 8), np.uint8)              # more ranks than block rows: this rank contributes a zero histogram
    allreduce_endpoint_histogram(d_hist, group)
    hist = d_hist.cpu().numpy().astype(np.uint32)
    return etc_blocks, etc1s.training_vectors_from_histogram(hist)


def allgather_selector_training(local_keys, local_weights, group=None):
    """local_keys / local_weights: torch int64 tensors (any length, on this rank's device) holding this shard's UNIQUE selector
    keys and their summed weights. Returns the global unique keys (ascending) and summed weights, identical on every rank.
    Shards differ in length, so lengths are exchanged first and the payload is padded to the longest (an all-gatherv)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        keys, inv = torch.unique(local_keys, return_inverse=True)
        return keys, torch.zeros_like(keys).index_add_(0, inv, local_weights)
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_keys.numel()], dtype=torch.int64, device=local_keys.device)
    lengths = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(lengths, n_local, group=group)
    longest = int(max(int(v.item()) for v in lengths))
    payload = torch.zeros((2, longest), dtype=torch.int64, device=local_keys.device)
    payload[0, :local_keys.numel()] = local_keys
    payload[1, :local_keys.numel()] = local_weights
    gathered = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    keys = torch.cat([g[0, :int(n.item())] for g, n in zip(gathered, lengths)])
    weights = torch.cat([g[1, :int(n.item())] for g, n in zip(gathered, lengths)])
    ukeys, inv = torch.unique(keys, return_inverse=True)
    return ukeys, torch.zeros_like(ukeys).index_add_(0, inv, weights)
