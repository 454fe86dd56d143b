"""CPU-only checks (no GPU): the oracle is pinned to the reference's own golden numbers, the host-emulation build of the
device code matches the oracle bit for bit, and the C-ABI library loads and exports what include/basisu_b200.h declares."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

import util
from basis_universal_b200 import sharding, uastc

FLAG_SETS = [0, 1, 2, 3, 4, 2 | 512, 2 | 8, 2 | 16, 2 | 64, 1 | 128, 2 | 256, 3 | 512]


def basis_slice(data):
    """(offset, size, blocks_x, blocks_y) of slice 0 of a .basis file (transcoder/basisu_file_headers.h:16-33, 208-250)."""
    b = bytes(data)
    assert b[0:2] == b"sB"  # cBASISSigValue little-endian
    desc = int.from_bytes(b[65:69], "little")
    nbx = int.from_bytes(b[desc + 9:desc + 11], "little")
    nby = int.from_bytes(b[desc + 11:desc + 13], "little")
    ofs = int.from_bytes(b[desc + 13:desc + 17], "little")
    size = int.from_bytes(b[desc + 17:desc + 21], "little")
    return ofs, size, nbx, nby


def test_oracle_reproduces_golden_vectors(ref, golden):
    """Pins oracle/_ref to the committed vectors (and so the vectors to the reference build that made them)."""
    for f in FLAG_SETS:
        assert np.array_equal(ref.encode_uastc(golden["blocks"], f), golden[f"uastc_flags_{f}"]), f"flags {f}"


def test_oracle_kodim03_level0_basis_md5(ref):
    """`basisu -uastc -uastc_level 0 kodim03.png` -> 393347 bytes, md5 6d98eb72a9a3112ff55344132a28b042."""
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    data = ref.compress_image(1, g["image"], 0 | (1 << 9))
    assert len(data) == 393347
    assert hashlib.md5(data).hexdigest() == "6d98eb72a9a3112ff55344132a28b042"
    assert data == g["basis"].tobytes()


@pytest.mark.parametrize("flags", FLAG_SETS)
def test_hostemu_matches_golden(emu, golden, flags):
    out = emu.encode_uastc(golden["blocks"], flags)
    assert np.array_equal(out, golden[f"uastc_flags_{flags}"])


def test_hostemu_kodim03_level0_slice_bytes(emu):
    """BASELINE.json config 1 (bit-exact gate): the slice data inside the reference's .basis equals our block bytes."""
    g = np.load(os.path.join(util.GOLDEN, "kodim03_uastc_l0.npz"))
    ofs, size, nbx, nby = basis_slice(g["basis"])
    blocks = uastc.extract_blocks(g["image"])
    assert blocks.shape[0] == nbx * nby and size == nbx * nby * 16
    out = emu.encode_uastc(blocks, 0)
    assert out.tobytes() == g["basis"].tobytes()[ofs:ofs + size]


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_hostemu_matches_reference_on_fresh_inputs(ref, emu, level):
    blocks = util.image_to_blocks(util.synth(128, 777 + level))
    assert np.array_equal(emu.encode_uastc(blocks, level), ref.encode_uastc(blocks, level))


def test_hostemu_level4_matches_reference(ref, emu):
    blocks = np.concatenate([util.edge_case_blocks(11), util.image_to_blocks(util.synth(32, 5))])
    assert np.array_equal(emu.encode_uastc(blocks, 4), ref.encode_uastc(blocks, 4))


def test_colour_cell_compression_differential(ref, emu):
    """Per-function fuzz of the innermost routine (bc7enc.cpp:1364) across every (weights, range, alpha) combination UASTC uses."""
    rng = np.random.default_rng(3)
    combos = [(4, 19, 0), (5, 11, 0), (2, 20, 0), (3, 8, 0), (2, 7, 0), (2, 12, 0), (3, 20, 0), (2, 18, 0), (2, 8, 1), (4, 13, 1), (2, 13, 0),
              (3, 19, 1), (1, 20, 0), (2, 20, 1), (4, 20, 1), (2, 20, 0)]
    for wtab, rngidx, alpha in combos:
        for trial in range(60):
            n = int(rng.integers(1, 17))
            kind = trial % 4
            if kind == 0:
                px = rng.integers(0, 256, (n, 4), dtype=np.uint8)
            elif kind == 1:
                base = rng.integers(0, 256, 4)
                px = (base + rng.integers(-10, 11, (n, 4))).clip(0, 255).astype(np.uint8)
            elif kind == 2:
                px = np.repeat(rng.integers(0, 256, (1, 4), dtype=np.uint8), n, 0)
            else:
                a, b = rng.integers(0, 256, 4), rng.integers(0, 256, 4)
                t = rng.random((n, 1))
                px = (a + (b - a) * t).astype(np.uint8)
            if not alpha:
                px[:, 3] = 255
            for uber, ls in [(0, 1), (1, 1), (3, 2), (6, 2)]:
                r = ref.ccc(px, wtab, rngidx, alpha, uber, ls)
                e = emu.ccc(px, wtab, rngidx, alpha, uber, ls)
                assert r[0] == e[0] and np.array_equal(r[1], e[1]) and np.array_equal(r[2], e[2]) and np.array_equal(r[3], e[3]), (wtab, rngidx, alpha, uber, ls, px.tolist())


def test_decoded_quality_is_sane(ref, golden):
    """Round trip through the reference decoder: level-2 blocks reconstruct the source closely (guards against a green-but-garbage encoder)."""
    blocks = util.image_to_blocks(util.synth(64, 4321))
    dec = ref.unpack_uastc(golden["uastc_flags_2"][-blocks.shape[0]:])
    mse = np.mean((dec.astype(np.float64) - blocks.astype(np.float64)) ** 2)
    assert 10 * np.log10(255 ** 2 / mse) > 30.0


def test_extract_blocks_clamps_edges():
    img = np.arange(5 * 6 * 4, dtype=np.uint8).reshape(5, 6, 4)
    b = uastc.extract_blocks(img)
    assert b.shape == (4, 64)
    last = b[3].reshape(4, 4, 4)
    assert np.array_equal(last[0, 0], img[4, 4]) and np.array_equal(last[3, 3], img[4, 5]) and np.array_equal(last[0, 3], img[4, 5])


def test_sharding_partitions_rows_exactly():
    for rows in [0, 1, 7, 8, 1024, 2049]:
        for world in [1, 2, 3, 4, 8]:
            spans = [sharding.block_row_range(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.slices_for_rank(64, 3, 8) == list(range(3, 64, 8))


def test_c_abi_library_exports_every_declared_symbol():
    """The product library builds for sm_100a without a GPU and exports exactly the header's entry points. No compute calls here."""
    from basis_universal_b200 import build, _lib
    path = build.build()
    header = open(os.path.join(util.ROOT, "include", "basisu_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no entry points parsed"
    L = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/basisu_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    # Without a GPU the library must fail loudly rather than fall back to a CPU path.
    L.b200_device_count.restype = ctypes.c_int
    if L.b200_device_count() <= 0:
        L.b200_create_context.restype = ctypes.c_void_p
        assert not L.b200_create_context(0)
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_last_error.argtypes = [ctypes.c_void_p]
        assert L.b200_last_error(None)


REAL_FLAG_SETS = [0, 1, 2, 3, 2 | 512, 2 | 64 | 256]


@pytest.mark.parametrize("flags", REAL_FLAG_SETS)
def test_hostemu_matches_golden_real_images(emu, golden_real, flags):
    """Same device headers, real content from the reference's test set (LA and solid modes included)."""
    blocks = golden_real["blocks"]
    sel = slice(None) if flags in (0, 2) else slice(0, None, 4)     # keep the CPU suite short: full set at levels 0 and 2
    assert np.array_equal(emu.encode_uastc(blocks[sel], flags), golden_real[f"uastc_flags_{flags}"][sel])


def test_oracle_reproduces_golden_real_images(ref, golden_real):
    assert np.array_equal(ref.encode_uastc(golden_real["blocks"][::8], 2), golden_real["uastc_flags_2"][::8])


def test_python_wrappers_raise_without_a_gpu():
    """No silent CPU path: on a machine without a CUDA device every wrapper constructor fails with the library's message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from basis_universal_b200 import B200Error, etc1s, image, uastc
    for ctor in (uastc.Encoder, etc1s.Etc1sContext, image.ImageOps):
        with pytest.raises(B200Error):
            ctor(0)


def test_reference_cli_on_the_dropin_library_cpu_path(tmp_path):
    """integration/_build/basisu = the reference's basisu_tool.cpp, unmodified, linked against the drop-in library (patched
    compressor + GPU seam adapter). Without -opencl (and without a GPU) it must still be the reference encoder: config 1's KAT."""
    import subprocess
    tool = os.path.join(util.ROOT, "integration", "_build", "basisu")
    png = os.path.join(util.ROOT, "oracle", "_ref", "test_files", "kodim03.png")
    if not (os.path.exists(tool) and os.path.exists(png)):
        pytest.skip("drop-in CLI or test image not built here")
    r = subprocess.run([tool, "-uastc", "-uastc_level", "0", "-basis", "-output_file", str(tmp_path / "k.basis"), png], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    data = open(tmp_path / "k.basis", "rb").read()
    assert len(data) == 393347 and hashlib.md5(data).hexdigest() == "6d98eb72a9a3112ff55344132a28b042"
