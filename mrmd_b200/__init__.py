"""mrmd_b200 -- B200-native implementation of MRMD's per-step force + neighbour hot path.

Layers: hand-written sm_100a kernels + C ABI (csrc/, include/mrmd_b200.h) -> host mirrors of the reference
interface (include/mrmd/ for C++20, mrmd_b200.api for Python).  No CPU fallback.
"""
from . import api  # noqa: F401
from ._lib import LIB_PATH, MrmdB200Error, load  # noqa: F401
