"""Whole-scene parity and speed on ONE GPU: the full stage schedule (main.cpp:20-41) through mp-mvs_b200/pipeline.py and
through the reference's own kernels (oracle/_ref) + restated host prior stage, identical seeds, Jacobi order on both sides.
Reports accuracy at 2/5/10 cm against the synthetic ground truth for both, their agreement, and wall times.

    python tests/tests/tools/scene_parity.py --scene dtu            # BASELINE config 2: 49 views 1600x1200, photometric + 2 geom
    python tests/tests/tools/scene_parity.py --scene eth3d --planar 1 --geom-planar 1   # config 3, shipped default schedule
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import PKG, problem_arrays  # noqa: E402
import oracle_py  # noqa: E402
import prior_oracle  # noqa: E402
from mpmvs_b200 import capi, io_formats, pipeline, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="dtu")
ap.add_argument("--planar", type=int, default=0)
ap.add_argument("--geom-planar", type=int, default=0)
ap.add_argument("--views", type=int, default=0)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--tex", default="u8")
ap.add_argument("--skip-reference", action="store_true")
ap.add_argument("--in-flight", type=int, default=8)
args = ap.parse_args()
workers = min(32, os.cpu_count() or 1)
t = time.time()
if args.scene == "dtu":
    sc = synth.make_dtu_scene(width=int(1600 * args.scale), height=int(1200 * args.scale), workers=workers)
    label = "dtu-shaped 49 views"
else:
    sc = synth.make_eth3d_scene(width=int(3200 * args.scale), height=int(2130 * args.scale), n_views=args.views or 11, workers=workers)
    label = "eth3d-shaped"
print(f"rendered {sc.num_views} views {sc.width}x{sc.height} in {time.time() - t:.1f} s", flush=True)
n, W, H = sc.num_views, sc.width, sc.height
entries = [io_formats.SceneEntry(ref_id=i, src_ids=[i] + [j for j, _ in sc.pairs[i]], estimate=True) for i in range(n)]
cams = {i: c for i, c in enumerate(sc.cams)}
images = {i: im for i, im in enumerate(sc.images)}
cfg = pipeline.PipelineConfig(geom_iterations=2, max_src=10, seed=33, planar_prior=bool(args.planar), geom_planar_prior=bool(args.geom_planar),
                              tex_format=capi.TEX_U8 if args.tex == "u8" else capi.TEX_F32, in_flight=args.in_flight)
p = pipeline.DensePipeline(entries, cams, images, cfg)
p.setup()
t = time.time()
stats = p.run()
t_ours = time.time() - t
ours = p.results()
prior_stats = [p.engines[r].prior_stats for r in p.my_refs if getattr(p.engines[r], "prior_stats", None)]
if prior_stats:
    print("prior stage per image (last pass with a prior): vertices", [s["n_vertices"] for s in prior_stats], "pick ms", [round(s["pick_ms"], 1) for s in prior_stats],
          "delaunay ms", [round(s["delaunay_ms"], 1) for s in prior_stats], "raster ms", [round(s["raster_ms"], 1) for s in prior_stats], flush=True)
p.destroy()
print("ours:", [(s.name, round(s.device_ms, 1), round(s.exchange_ms, 1)) for s in stats], f"total {t_ours:.2f} s", flush=True)


def acc_table(res):
    a = np.array([synth.accuracy_at(res[i][0][..., 3], sc.gt_depth[i]) for i in range(n)])
    return [round(float(v), 3) for v in a.mean(0)]


out = {"scene": f"{label} {n} views {W}x{H}, 10 src", "planar": args.planar, "geom_planar": args.geom_planar, "view_storage": args.tex,
       "ours_s": round(t_ours, 3), "ours_mpix_per_s": round(n * W * H / 1e6 / t_ours, 3), "ours_accuracy_2_5_10cm": acc_table(ours),
       "ours_passes": [(s.name, round(s.device_ms, 1)) for s in stats]}
if not args.skip_reference and oracle_py.available("ref"):
    def ref_process(ref_id, stage, geom, with_prior, state, depth_maps):
        ids, imgs, packed = problem_arrays(sc, ref_id, 10)
        R = oracle_py.Oracle("ref").set_problem(imgs, packed)
        R.set_geom_consistency_params(geom, with_prior)
        if geom:
            R.set_src_depths([depth_maps[i] for i in ids[1:]])
            R.set_state(*state)
        seed = pipeline.stage_seed(cfg.seed, ref_id, stage)
        R.run(seed)
        res = R.result(geom=True)
        if with_prior:
            dmin, dmax = R.depth_range
            prior, mask, _, _, _ = prior_oracle.build_prior_fast(res[0], res[1], sc.cams[ref_id].K, dmin, dmax, res[2] if geom else None)
            R.set_planar_prior_params()
            R.set_geom_consistency_params(False, True)
            R.set_prior(prior, mask)
            R.run(seed ^ 0x5DEECE66D)
            res = R.result(geom=True)
        R.destroy()
        return res[0], res[1]

    t = time.time()
    state = {i: ref_process(i, 0, False, bool(args.planar) and not args.geom_planar, None, None) for i in range(n)}
    for g in range(2):
        depth_maps = {i: state[i][0][..., 3].copy() for i in range(n)}
        state = {i: ref_process(i, 1 + g, True, bool(args.geom_planar) and g != 1, state[i], depth_maps) for i in range(n)}
    t_ref = time.time() - t
    agree = []
    for i in range(n):
        valid = (sc.gt_depth[i] > 0) & (state[i][1] < 0.5)
        agree.append(synth.depth_normal_agreement(ours[i][0][..., 3], ours[i][0][..., :3], state[i][0][..., 3], state[i][0][..., :3], valid))
    out.update({"reference_s": round(t_ref, 3), "reference_mpix_per_s": round(n * W * H / 1e6 / t_ref, 3), "reference_accuracy_2_5_10cm": acc_table(state),
                "speedup_wall": round(t_ref / t_ours, 2), "agreement_median": round(float(np.median(agree)), 4), "agreement_min": round(float(min(agree)), 4),
                "accuracy_delta_points": [round(a - b, 3) for a, b in zip(acc_table(ours), acc_table(state))]})
print(json.dumps(out), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"scene_parity_{args.scene}_p{args.planar}g{args.geom_planar}.json"), "w"), indent=1)
