"""CPU-only: the C-ABI library loads, exports every symbol include/mrmd_b200.h declares, the ctypes table
covers the header one to one, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "mrmd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrmd_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mrmd_b200 import _lib, build

    build.build()
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) > 60
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mrmd_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names


def test_no_cpu_fallback():
    from mrmd_b200 import api
    from mrmd_b200._lib import MrmdB200Error

    if api.L().mrmd_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(MrmdB200Error, match="no CPU fallback"):
        api.Atoms(16)
    s = api.Subdomain([0, 0, 0], [2, 4, 6], 0.5)  # plain host struct: works without a device
    assert list(s.minInnerCorner) == [0.5, 0.5, 0.5] and s.getVolume() == 48.0


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under mrmd_b200/ or include/ may reference it."""
    for base in ("mrmd_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "pyoracle" not in text and "mrmd_oracle.h" not in text and "libmrmd_oracle" not in text, f
