"""Generates tests/golden/*.npz from the compiled, unmodified reference (oracle/_ref/libbasisu_ref.so).
Run here (where /root/reference exists); the fixtures are committed so the GPU box needs neither the reference sources
nor the reference .so.  Usage: python tests/golden/make_golden.py"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from util import GOLDEN, Ref, build_ref, edge_case_blocks, image_to_blocks, synth  # noqa: E402

UASTC_FLAG_SETS = [0, 1, 2, 3, 4, 2 | 512, 2 | 8, 2 | 16, 2 | 64, 1 | 128, 2 | 256, 3 | 512]


def main():
    assert build_ref(), "reference .so could not be built"
    ref = Ref()
    blocks = np.concatenate([edge_case_blocks(), image_to_blocks(synth(64, 4321))])
    out = {"blocks": blocks}
    for f in UASTC_FLAG_SETS:
        out[f"uastc_flags_{f}"] = ref.encode_uastc(blocks, f)
    np.savez_compressed(os.path.join(GOLDEN, "uastc_blocks.npz"), **out)

    # Config 1 of BASELINE.json: kodim03 UASTC level 0 .basis via the reference's own basis_compress API.
    kodim = "/root/reference/test_files/kodim03.png"
    if os.path.exists(kodim):
        from PIL import Image
        img = np.array(Image.open(kodim).convert("RGBA"))
        data = ref.compress_image(1, img, 0 | (1 << 9))  # cUASTC_LDR_4x4, level 0, cFlagThreaded
        # keep the source blocks and the file (393 KB compresses well: mostly mode-0 blocks)
        np.savez_compressed(os.path.join(GOLDEN, "kodim03_uastc_l0.npz"), image=img, basis=np.frombuffer(data, np.uint8))
        print("kodim03 L0 .basis", len(data), "bytes md5", hashlib.md5(data).hexdigest())
    print("wrote", os.listdir(GOLDEN))


if __name__ == "__main__":
    main()
