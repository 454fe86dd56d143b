"""Multi-GPU partitioning of the block array: contiguous block-row ranges per rank, no data-path
collective for UASTC (blocks are independent).  Slices/images are distributed round-robin for the RDO configuration,
whose chains must not be split (the reference chunks each slice into `total_jobs` sequential chains)."""


def block_row_range(num_block_rows, rank, world_size):
    """Contiguous [first, last) block-row range of `rank`; remainders go to the lowest ranks so sizes differ by <= 1."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(num_block_rows, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def block_range(num_blocks_x, num_blocks_y, rank, world_size):
    """[first_block, last_block) in raster block order for `rank`'s block rows."""
    r0, r1 = block_row_range(num_blocks_y, rank, world_size)
    return r0 * num_blocks_x, r1 * num_blocks_x


def slices_for_rank(num_slices, rank, world_size):
    """Whole slices (images / tiles) owned by `rank`, round-robin."""
    return list(range(rank, num_slices, world_size))
