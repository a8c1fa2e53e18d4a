"""Build tuning variants of libmpmvs_b200.so into mp-mvs_b200/variants/ (selected at run time with MPMVS_LIB_VARIANT).

Every variant is the whole library -- both arithmetics -- compiled with extra -D flags through the one Makefile
(`make OUT=... OBJ=... EXTRA=...`), so a variant differs from the shipped library only in the knob under test."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mp-mvs_b200", "csrc")
OUT = os.path.join(ROOT, "mp-mvs_b200", "variants")

# "default" = the in-tree defaults (PM_BH=4, PM_MIN_BLOCKS=3, PM_CA_SMEM=1, PM_UNIFORM_VIEWS=1, PM_EARLY_OUT=1, PM_WTAB=0)
VARIANTS = {
    "default": [],
    "nopacked": ["-DPM_PACKED=0"],                    # scalar FP32 instructions in the NCC tap loop (before FFMA2 / FADD2 / FMUL2)
    "noeo": ["-DPM_EARLY_OUT=0"],
    "nouv": ["-DPM_UNIFORM_VIEWS=0"],
    "mb2": ["-DPM_MIN_BLOCKS=2"],
    "mb4": ["-DPM_MIN_BLOCKS=4"],
    "bh2mb6": ["-DPM_BH=2", "-DPM_MIN_BLOCKS=6"],
    "bh8mb1": ["-DPM_BH=8", "-DPM_MIN_BLOCKS=1"],
    "vm": ["-DPM_VIEW_MAJOR=1"],                      # candidate costs view by view (two inlined copies of the NCC)
    "vmni": ["-DPM_VIEW_MAJOR=1", "-DPM_NCC_NOINLINE=1"],   # ... with the NCC as a real function (one copy)
    "ni": ["-DPM_NCC_NOINLINE=1"],
    "wtab": ["-DPM_WTAB=1"],                         # bilateral weights from a per-thread shared-memory table (measured: slower)
    "wtab_mb2": ["-DPM_WTAB=1", "-DPM_MIN_BLOCKS=2"],
}


def build(name):
    out = os.path.join(OUT, f"libmpmvs_b200_{name}.so")
    obj = os.path.join(ROOT, "mp-mvs_b200", "build", "variants", name)
    cmd = ["make", "-s", "-j4", "-C", CSRC, f"OUT={out}", f"OBJ={obj}", "EXTRA=" + " ".join(VARIANTS[name]), out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, (r.stdout + r.stderr)[-400:]


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    failed = 0
    with ThreadPoolExecutor(4) as ex:
        for name, rc, err in ex.map(build, names):
            print(name, "ok" if rc == 0 else "FAILED " + err)
            failed += rc != 0
    sys.exit(1 if failed else 0)
