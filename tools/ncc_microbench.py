"""BASELINE.json config 5: NCC cost kernel microbenchmark sweep -- window 5..11 px (3..6 taps per side at step 2, plus the
reference's 21 and 41 px dilations), 4..20 source views, 1600x1200 -- against the L1/TEX roof (4 bilinear fetches/clk/SM,
see tools/tex_microbench.cu) and the FP32 roof. Prints one line per point and a JSON summary.

    python tools/ncc_microbench.py [--tex u8|f32] [--out gpurun_out/ncc_microbench.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import PKG, gt_planes_cam  # noqa: E402
from mpmvs_b200 import capi, io_formats, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tex", default="u8")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ncc_microbench.json"))
args = ap.parse_args()

N_SRC = 20
probe = synth.make_dtu_scene(views=[])
ref = 24
ids = [ref] + [i for i, _ in synth._nearest_pairs(np.array([c.C for c in probe.cams]), np.array([0.1, 0.05, 0.15]), N_SRC)[ref]]
sc = synth.make_dtu_scene(views=ids, workers=min(16, os.cpu_count() or 1), n_src=N_SRC)
imgs = [sc.images[i] for i in ids]
cams = io_formats.pack_cameras([sc.cams[i] for i in ids])
fmt = capi.TEX_U8 if args.tex == "u8" else capi.TEX_F32
pm = capi.PatchMatch(0).set_tex_format(fmt).set_problem(imgs if args.tex == "u8" else [i.astype(np.float32) for i in imgs], cams)
planes = gt_planes_cam(sc, ref)            # fixed planes near ground truth (coherent fetches: the kernel's best case)
rng = np.random.default_rng(1)
planes_rnd = planes.copy()
planes_rnd[..., 3] *= rng.uniform(0.8, 1.2, planes.shape[:2]).astype(np.float32)   # per-pixel random depths: incoherent fetches
W, H = sc.width, sc.height
SM, CLK = 148, 1965e6
tex_roof = 4 * SM * CLK            # taps/s
fp32_roof = SM * 128 * 2 * CLK     # flop/s
rows = []
for label, pl in (("gt planes", planes), ("random depth", planes_rnd)):
    for scale, taps in ((0, 3), (0, 4), (0, 5), (0, 6), (1, 6), (2, 6)):
        for nv in (4, 8, 10, 16, 20):
            reps = max(1, 40 // nv)
            ms, cnt = pm.ncc_bench(pl, scale, taps, nv, reps)
            t = cnt * taps * taps
            gtaps = t / ms / 1e6
            step = 2 << scale
            rows.append(dict(planes=label, scale=scale, taps_per_side=taps, window_px=(taps - 1) * step + 1, n_views=nv, reps=reps,
                             ms=round(ms, 3), gtaps_per_s=round(gtaps, 1), frac_tex_roof=round(gtaps * 1e9 / tex_roof, 3),
                             frac_fp32_roof_at_12flop=round(gtaps * 1e9 * 12 / fp32_roof, 3)))
            print(f"{label:12s} window {rows[-1]['window_px']:2d}px ({taps}x{taps} taps, step {step}) views {nv:2d}: {ms:8.3f} ms  {gtaps:7.1f} Gtaps/s  "
                  f"{rows[-1]['frac_tex_roof']:.3f} of the TEX roof, {rows[-1]['frac_fp32_roof_at_12flop']:.3f} of the FP32 roof", flush=True)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump({"image": [W, H], "view_storage": args.tex, "tex_roof_gtaps": tex_roof / 1e9, "rows": rows}, open(args.out, "w"), indent=1)
pm.destroy()
