"""Builds A/B variants of libapg_b200.so into build/variants/ (git-ignored, travels with gpurun): one source file
recompiled with extra -D flags, linked with the objects of the regular build.  Select one with APG_B200_LIB=<path>.

    python tools/build_variants.py name=apg_image_tma.cu:-DAPG_TMA_MINB=3,-DAPG_TMA_REPLICAS=8 ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "active-perception-gym_b200"))
from ap_gym_b200 import _build  # noqa: E402


def main(specs):
    _build.build()
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    obj_dir = os.path.join(_build.LIB_DIR, "obj")
    for spec in specs:
        name, rest = spec.split("=", 1)
        src, flags = rest.split(":", 1)
        obj = os.path.join(out_dir, f"{name}.o")
        cmd = [_build._nvcc(), *_build.NVCC_FLAGS, *flags.split(","), "-c", os.path.join(_build.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        objs = [os.path.join(obj_dir, o) for o in sorted(os.listdir(obj_dir)) if o.endswith(".o") and o != src[:-3] + ".o"]
        lib = os.path.join(out_dir, f"libapg_{name}.so")
        r = subprocess.run([_build._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, obj, *objs,
                            "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        print(lib)


if __name__ == "__main__":
    main(sys.argv[1:])
