#!/usr/bin/env python
"""Builds side-by-side variants of libmrmd_b200.so with other tuning constants (-D overrides of the MRMD_* knobs in
csrc/tiled.cu) into mrmd_b200/variants/, for A/B measurements on the GPU box:
    python profiles/variants.py name=-DMRMD_LJT_PREFETCH=24 other=-DMRMD_TL_THREADS_FORCE=256,-DMRMD_LJT_PREFETCH=24
    MRMD_B200_LIB_VARIANT=name python bench.py ...
Measurement tooling only: the product is mrmd_b200/libmrmd_b200.so built by mrmd_b200/build.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrmd_b200 import build as b  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "mrmd_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        name, flags = spec.split("=", 1)
        flags = [f for f in flags.split(",") if f]
        obj_dir = os.path.join(ROOT, "mrmd_b200", "build", "variant_" + name)
        os.makedirs(obj_dir, exist_ok=True)
        procs, objs = [], []
        for src in b.sources():
            obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            procs.append(subprocess.Popen([b.NVCC] + b.NVCC_FLAGS + flags + ["-c", src, "-o", obj]))
        for p in procs:
            if p.wait() != 0:
                raise SystemExit("nvcc failed for variant " + name)
        lib = os.path.join(out_dir, f"libmrmd_b200_{name}.so")
        subprocess.check_call([b.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", b.HOST_CXX, "-o",
                               lib] + objs)
        print(lib)


if __name__ == "__main__":
    main()
