"""Generate tests/golden/sky.npz: outputs of the reference's OWN sky-mask kernel (oracle/_ref/libmpmvs_ref_sky.so =
/root/reference/SkySegment/src/SkyRegionDetect.cu compiled where it lies) on seeded synthetic inputs.
Run on a GPU box:  python tests/golden/make_golden_sky.py   (the oracle .so travels with the snapshot)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
import sky_oracle  # noqa: E402

out = {}
for name, (w, h, div, seed) in {"a": (160, 120, 4, 7), "b": (75, 53, 3, 8), "c": (96, 64, 1, 9)}.items():
    bgr, lo, truth = sky_oracle.make_sky_case(w, h, div, seed)
    full = sky_oracle.resize_linear(lo, w, h)
    res = sky_oracle.ref_sky_filter(bgr, full)
    out[f"{name}_bgr"], out[f"{name}_mask_lo"], out[f"{name}_mask_full"] = bgr, lo, full
    out[f"{name}_result"] = (res > 0).astype(np.uint8)
    print(name, bgr.shape, "sky fraction truth", float(truth.mean()), "reference", float((res > 0).mean()))
dst = os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "tests/golden", "sky.npz")
np.savez_compressed(dst, **out)
print("wrote", dst)
