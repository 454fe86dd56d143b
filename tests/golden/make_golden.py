"""Generates tests/golden/*.npz from the compiled, unmodified reference (oracle/_ref/libbasisu_ref.so).
Run here (where /root/reference exists); the fixtures are committed so the GPU box needs neither the reference sources
nor the reference .so.  Usage: python tests/golden/make_golden.py"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from util import GOLDEN, Ref, build_ref, edge_case_blocks, image_to_blocks, synth, uastc_fields  # noqa: E402

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
    # Real content from the reference's own test set: 320 blocks sampled from each of several images (photographs, an
    # alpha-channel image, line art, the 1x1 solid images), encoded at levels 0-3 and level 2 with the alpha-in-RGB flag bits.
    tf = "/root/reference/test_files"
    names = ["kodim01.png", "kodim08.png", "kodim13.png", "kodim20.png", "kodim23.png", "alpha0.png", "wikipedia.png", "xmen.png", "tough.png", "black_1x1.png", "white_1x1.png"]
    if all(os.path.exists(os.path.join(tf, n)) for n in names):
        from PIL import Image
        rng = np.random.default_rng(2024)
        picked = []
        for n in names:
            img = np.array(Image.open(os.path.join(tf, n)).convert("RGBA"))
            h, w = img.shape[:2]
            pad = np.pad(img, ((0, (-h) % 4), (0, (-w) % 4), (0, 0)), mode="edge")
            b = image_to_blocks(pad)
            picked.append(b[rng.choice(b.shape[0], min(320, b.shape[0]), replace=False)])
        blocks = np.concatenate(picked)
        out = {"blocks": blocks}
        for f in [0, 1, 2, 3, 2 | 512, 2 | 64 | 256]:
            out[f"uastc_flags_{f}"] = ref.encode_uastc(blocks, f)
        np.savez_compressed(os.path.join(GOLDEN, "uastc_real_images.npz"), **out)
        modes = np.bincount([(uastc_fields(ref, x) or {"mode": 8})["mode"] for x in out["uastc_flags_2"]], minlength=19)
        print("real-image fixture:", blocks.shape[0], "blocks; level-2 mode histogram", modes.tolist())
    print("wrote", os.listdir(GOLDEN))


if __name__ == "__main__":
    main()
