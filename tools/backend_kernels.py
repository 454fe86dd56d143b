"""Drives the two ETC1S backend kernels on their own (for ncu): endpoint prediction over 1024 x 1024 blocks with a coarse synthetic
codebook, and the palette ordering of the resulting index stream. python tools/backend_kernels.py [blocks_per_side]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth, to_blocks  # noqa: E402
from basis_universal_b200 import etc1s  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
blocks = to_blocks(synth(side * 4, 1234))
n = blocks.shape[0]
ctx = etc1s.Etc1sContext(0)
ctx.set_pixel_blocks(blocks)
etc = ctx.encode_etc1s_blocks(False, 16)
key = ((etc[:, 0] >> 4).astype(np.uint32) << 12) | ((etc[:, 1] >> 4).astype(np.uint32) << 8) | ((etc[:, 2] >> 4).astype(np.uint32) << 4) | (etc[:, 3] >> 5)
uniq, idx0 = np.unique(key, return_inverse=True)
cb = np.stack([((uniq >> 12) & 15) * 2 + 1, ((uniq >> 8) & 15) * 2 + 1, ((uniq >> 4) & 15) * 2 + 1, uniq & 7], -1).astype(np.uint8)
e = cb[idx0]
etc[:, 0] = e[:, 0] << 3; etc[:, 1] = e[:, 1] << 3; etc[:, 2] = e[:, 2] << 3; etc[:, 3] = (e[:, 3] << 5) | (e[:, 3] << 2) | 3
for _ in range(2):
    t0 = time.perf_counter()
    idx, pred = ctx.backend_endpoint_prediction([(0, side, side)], etc, cb, idx0.astype(np.uint32), 1.5, False)
    t1 = time.perf_counter()
    stream = idx[(pred & 3) == 3]
    remap = ctx.palette_reorder(stream, cb.shape[0])
    t2 = time.perf_counter()
print(f"{n} blocks, {cb.shape[0]} endpoints: prediction call {1e3 * (t1 - t0):.1f} ms ({(pred & 3 != 3).mean():.1%} predicted, {(idx != idx0).mean():.1%} remapped), "
      f"palette ordering of {stream.shape[0]} indices {1e3 * (t2 - t1):.1f} ms, kernels {ctx.last_kernel_ms:.1f} ms")
