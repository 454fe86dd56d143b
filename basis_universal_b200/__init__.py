"""basis_universal_b200: B200 (sm_100a) implementation of basis_universal's per-block texture-encode hot path
(UASTC LDR 4x4, ETC1S frontend stages) behind the reference's own interfaces. See DESIGN.md / INTEGRATION.md."""
from ._lib import B200Error, lib, LIB_PATH  # noqa: F401
from . import uastc, etc1s, image  # noqa: F401
from .uastc import Encoder  # noqa: F401
