"""Sky-mask joint-bilateral upsampling at the ETH3D-shaped frame size: device time of pm_sky_filter_kernel against its
shared-memory roof, next to the reference kernel when oracle/_ref is present.   python tests/tests/tools/sky_bench.py [W H reps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pkgload  # noqa: E402

pkgload.load_package()
import sky_oracle  # noqa: E402
from mpmvs_b200 import capi  # noqa: E402

W, H, reps = (int(a) for a in (sys.argv[1:4] + ["3200", "2130", "5"][len(sys.argv) - 1:]))
bgr, lo, truth = sky_oracle.make_sky_case(W, H, 8, seed=5)
ms = []
for _ in range(reps):
    res, _, t = capi.sky_mask_refine(bgr, lo)
    ms.append(t)
taps = W * H * 37 * 37
sm_mhz = 1965.0
try:
    sm_mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", sm_mhz))
except Exception:
    pass
lds_peak = 128 * 148 * sm_mhz * 1e6            # bytes/s of shared-memory bandwidth
xu_peak = 16 * 148 * sm_mhz * 1e6              # MUFU results/s (16 lanes/clk/SM)
best = min(ms[1:]) if len(ms) > 1 else ms[0]
out = {"size": f"{W}x{H}", "ms": [round(m, 3) for m in ms], "taps": taps, "gtaps_per_s": round(taps / best / 1e6, 1),
       "mufu_per_tap": 2, "xu_roof_frac": round(taps * 2 / (best * 1e-3) / xu_peak, 3),
       "smem_bytes_per_tap": 10, "smem_roof_frac": round(taps * 10 / (best * 1e-3) / lds_peak, 3),
       "agreement_with_truth": round(float(((res > 0) == truth).mean()), 5)}
if sky_oracle.ref_available():
    full = sky_oracle.resize_linear(lo, W, H)
    ref, ref_ms = sky_oracle.ref_sky_filter(bgr, full, reps=3)
    out.update({"reference_kernel_ms": round(ref_ms, 3), "pixels_differing_from_reference": int(((res > 0) != (ref > 0)).sum())})
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sky_bench.json"), "w"), indent=1)
