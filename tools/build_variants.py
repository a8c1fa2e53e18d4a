"""Build tuning variants of libmpmvs_b200.so into mp-mvs_b200/variants/ (selected at run time with MPMVS_LIB_VARIANT)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mp-mvs_b200", "csrc")
OUT = os.path.join(ROOT, "mp-mvs_b200", "variants")
BASE = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--use_fast_math", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

# "default" = the in-tree defaults (PM_BH=4, PM_MIN_BLOCKS=3, PM_CA_SMEM=1, PM_UNIFORM_VIEWS=1, PM_EARLY_OUT=1)
VARIANTS = {
    "default": [],
    "v1": ["-DPM_BH=8", "-DPM_MIN_BLOCKS=3", "-DPM_CA_SMEM=0", "-DPM_UNIFORM_VIEWS=0", "-DPM_EARLY_OUT=0"],   # first measured kernel
    "noeo": ["-DPM_EARLY_OUT=0"],
    "nouv": ["-DPM_UNIFORM_VIEWS=0"],
    "local": ["-DPM_CA_SMEM=0", "-DPM_BH=8", "-DPM_MIN_BLOCKS=2"],
    "mb2": ["-DPM_MIN_BLOCKS=2"],
    "mb4": ["-DPM_MIN_BLOCKS=4"],
    "mb5": ["-DPM_MIN_BLOCKS=5"],
    "bh2mb6": ["-DPM_BH=2", "-DPM_MIN_BLOCKS=6"],
    "bh2mb8": ["-DPM_BH=2", "-DPM_MIN_BLOCKS=8"],
    "bh8mb2": ["-DPM_BH=8", "-DPM_MIN_BLOCKS=2"],
    "bh8mb1": ["-DPM_BH=8", "-DPM_MIN_BLOCKS=1"],
    # fidelity experiments (pm_core.cuh): the reference's own operation order for the tap coordinates / for the whole NCC
    # and the geometric cost. Use with float32 view storage (--tex f32): 8-bit storage filters differently.
    "litwarp": ["-DPM_LITERAL_WARP=1"],
    "literal": ["-DPM_LITERAL_NCC=1"],
    "literal_mb2": ["-DPM_LITERAL_NCC=1", "-DPM_MIN_BLOCKS=2"],
    # level 2: the same arithmetic with unrolled taps, every rounding pinned to what the reference's SASS does
    "literal2": ["-DPM_LITERAL_NCC=2"],
    "literal2_mb2": ["-DPM_LITERAL_NCC=2", "-DPM_MIN_BLOCKS=2"],
}


def build(name):
    out = os.path.join(OUT, f"libmpmvs_b200_{name}.so")
    cmd = ["/usr/local/cuda/bin/nvcc"] + BASE + VARIANTS[name] + ["-o", out, os.path.join(CSRC, "pm_kernels.cu"), os.path.join(CSRC, "pm_prior.cu"), os.path.join(CSRC, "pm_fusion.cu"), os.path.join(CSRC, "pm_sky.cu"), os.path.join(CSRC, "pm_capi.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-400:]


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    failed = 0
    with ThreadPoolExecutor(8) as ex:
        for name, rc, err in ex.map(build, names):
            print(name, "ok" if rc == 0 else "FAILED " + err)
            failed += rc != 0
    sys.exit(1 if failed else 0)
