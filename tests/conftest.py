import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref). Built on demand where /root/reference exists; otherwise the
    prebuilt .so that travelled with the repo; otherwise tests that need it skip and the committed golden vectors carry parity."""
    import util
    if not util.build_ref():
        pytest.skip("reference .so not available (no /root/reference and no prebuilt oracle/_ref)")
    try:
        return util.Ref()
    except OSError as e:
        pytest.skip(f"reference .so not loadable: {e}")


@pytest.fixture(scope="session")
def emu():
    import util
    assert util.build_emu()
    return util.Emu()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    import util
    return np.load(os.path.join(util.GOLDEN, "uastc_blocks.npz"))


@pytest.fixture(scope="session")
def golden_real():
    """Blocks sampled from the reference's own test images (photographs, alpha, line art, 1x1 solids) with the reference's
    output at several flag sets (tests/golden/make_golden.py)."""
    import numpy as np
    import util
    return np.load(os.path.join(util.GOLDEN, "uastc_real_images.npz"))
