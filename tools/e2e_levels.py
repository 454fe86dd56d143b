"""Host-pointer b200_uastc_encode_blocks (pinned buffers, H2D + kernels + D2H inside the call) against the device-resident call, per
effort level, on the 4096^2 bench image: what the copy/kernel pipelining of the host-pointer entry point buys at levels 0-1."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth, to_blocks  # noqa: E402
from basis_universal_b200 import uastc  # noqa: E402

enc = uastc.Encoder(0)
blocks = torch.from_numpy(to_blocks(synth(4096, 1234))).pin_memory()
n = blocks.shape[0]
out = torch.empty((n, 16), dtype=torch.uint8).pin_memory()
d_in = blocks.cuda()
d_out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for level in (0, 1, 2, 3):
    for _ in range(2):
        enc.encode_uastc_host_ptr(blocks.data_ptr(), n, out.data_ptr(), level)
        enc.encode_uastc_device(d_in.data_ptr(), n, d_out.data_ptr(), level)
    t0 = time.perf_counter()
    for _ in range(5):
        enc.encode_uastc_host_ptr(blocks.data_ptr(), n, out.data_ptr(), level)
    host_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    for _ in range(5):
        enc.encode_uastc_device(d_in.data_ptr(), n, d_out.data_ptr(), level)
    dev_ms = (time.perf_counter() - t0) / 5 * 1e3
    same = bool((out == d_out.cpu()).all())
    print(f"level {level}: host-pointer call {host_ms:7.2f} ms ({n * 16 / 1e3 / host_ms:7.1f} Mtexel/s), device-resident call {dev_ms:7.2f} ms, copies 80 MiB, identical output {same}", flush=True)
