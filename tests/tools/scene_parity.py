"""Whole-scene parity and speed on ONE GPU: the full stage schedule (main.cpp:20-41) through mp-mvs_b200/pipeline.py and
through the reference's own kernels (oracle/_ref) with its host stage restated (oracle/prior_oracle.py: OpenCV's Subdiv2D +
plain C), identical seeds, the SAME order on both sides (--order jacobi | gauss_seidel; the reference's own is gauss_seidel).

With the exact arithmetic the kernels are bit-identical, so a whole scene can only differ where the HOST stage differs (our
exact-integer Delaunay + closed-form plane against cv::Subdiv2D + cv::SVD::solveZ): --share-prior feeds the reference run
with the priors OUR planar-prior stage built, which separates the two -- with it, every plane of every image must come out
bit-identical. Reports per image the fraction of bit-identical planes, the north-star agreement (1 % depth, 5 deg normal),
accuracy and completeness at 2/5/10 cm against the synthetic ground truth for both sides, and wall times.

    python tests/tools/scene_parity.py --scene dtu                          # BASELINE config 2: 49 views 1600x1200, photometric + 2 geom
    python tests/tools/scene_parity.py --scene eth3d --planar 1 --geom-planar 1 [--share-prior]   # config 3, shipped default schedule
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
ap.add_argument("--arithmetic", default="exact", choices=["exact", "fast"])
ap.add_argument("--tex", default=None)
ap.add_argument("--order", default="jacobi", choices=["jacobi", "gauss_seidel"])
ap.add_argument("--share-prior", action="store_true", help="the reference run uses the priors our planar-prior stage built")
ap.add_argument("--plane-fit", default="svd", choices=["svd", "closed"], help="the reference side's plane fit: cv2's SVD per triangle (the reference's own, default) or the closed form")
ap.add_argument("--skip-reference", action="store_true")
ap.add_argument("--in-flight", type=int, default=8)
ap.add_argument("--out", default=None)
args = ap.parse_args()
tex = args.tex or ("f32" if args.arithmetic == "exact" else "u8")
workers = min(32, os.cpu_count() or 1)
t = time.time()
if args.scene == "dtu":
    sc = synth.make_dtu_scene(width=int(1600 * args.scale), height=int(1200 * args.scale), workers=workers, **({"grid": 3, "n_src": 4} if args.views == 9 else {}))
    label = "dtu-shaped"
else:
    sc = synth.make_eth3d_scene(width=int(3200 * args.scale), height=int(2130 * args.scale), n_views=args.views or 11, workers=workers)
    label = "eth3d-shaped"
print(f"rendered {sc.num_views} views {sc.width}x{sc.height} in {time.time() - t:.1f} s", flush=True)
n, W, H = sc.num_views, sc.width, sc.height
n_src = min(10, len(sc.pairs[0]))
entries = [io_formats.SceneEntry(ref_id=i, src_ids=[i] + [j for j, _ in sc.pairs[i]], estimate=True) for i in range(n)]
cams = {i: c for i, c in enumerate(sc.cams)}
images = {i: im for i, im in enumerate(sc.images)}
any_prior = bool(args.planar) or bool(args.geom_planar)
cfg = pipeline.PipelineConfig(geom_iterations=2, max_src=10, seed=33, planar_prior=bool(args.planar), geom_planar_prior=bool(args.geom_planar),
                              tex_format={"u8": capi.TEX_U8, "f32": capi.TEX_F32}[tex], in_flight=args.in_flight, arithmetic=args.arithmetic,
                              order=args.order, keep_priors=args.share_prior and any_prior)
p = pipeline.DensePipeline(entries, cams, images, cfg)
p.setup()
t = time.time()
stats = p.run()
t_ours = time.time() - t
ours = p.results()
priors = {r: list(p.engines[r].saved_priors) for r in p.my_refs} if cfg.keep_priors else {}
p.destroy()
print("ours:", [(s.name, round(s.device_ms, 1), round(s.exchange_ms, 1)) for s in stats], f"total {t_ours:.2f} s", flush=True)


def metrics(res):
    a = np.array([synth.accuracy_completeness_at(res[i][0][..., 3], res[i][1], sc.gt_depth[i]) for i in range(n)])      # [n, 2, 3]
    return [round(float(v), 3) for v in a[:, 0].mean(0)], [round(float(v), 3) for v in a[:, 1].mean(0)]


acc_o, comp_o = metrics(ours)
out = {"scene": f"{label} {n} views {W}x{H}, {n_src} src", "planar": args.planar, "geom_planar": args.geom_planar, "arithmetic": args.arithmetic,
       "view_storage": tex, "order": args.order, "share_prior": bool(args.share_prior),
       "ours_s": round(t_ours, 3), "ours_mpix_per_s": round(n * W * H / 1e6 / t_ours, 3), "ours_accuracy_2_5_10cm": acc_o, "ours_completeness_2_5_10cm": comp_o,
       "ours_passes": [(s.name, round(s.device_ms, 1)) for s in stats]}
if not args.skip_reference and oracle_py.available("ref"):
    prior_turn = {i: 0 for i in range(n)}

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
            if args.share_prior:
                prior, mask = priors[ref_id][prior_turn[ref_id]]
                prior_turn[ref_id] += 1
            else:
                dmin, dmax = R.depth_range
                prior, mask, _, _, _ = prior_oracle.build_prior_fast(res[0], res[1], sc.cams[ref_id].K, dmin, dmax, res[2] if geom else None, plane_fit=args.plane_fit)
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
        new = {}
        for i in range(n):
            new[i] = ref_process(i, 1 + g, True, bool(args.geom_planar) and g != 1, state[i], depth_maps)
            if args.order == "gauss_seidel":
                depth_maps[i] = new[i][0][..., 3].copy()         # depths.dmb overwritten in place (PatchMatch.cpp:620-633)
        state = new
    t_ref = time.time() - t
    agree, identical = [], []
    for i in range(n):
        valid = (sc.gt_depth[i] > 0) & (state[i][1] < 0.5)
        agree.append(synth.depth_normal_agreement(ours[i][0][..., 3], ours[i][0][..., :3], state[i][0][..., 3], state[i][0][..., :3], valid))
        identical.append(float((np.all(ours[i][0] == state[i][0], -1) & (ours[i][1] == state[i][1])).mean()))
    acc_r, comp_r = metrics(state)
    out.update({"reference_s": round(t_ref, 3), "reference_mpix_per_s": round(n * W * H / 1e6 / t_ref, 3),
                "reference_accuracy_2_5_10cm": acc_r, "reference_completeness_2_5_10cm": comp_r, "speedup_wall": round(t_ref / t_ours, 2),
                "agreement_median": round(float(np.median(agree)), 5), "agreement_min": round(float(min(agree)), 5),
                "bit_identical_pixels_per_image_min": round(min(identical), 6), "bit_identical_pixels_per_image_mean": round(float(np.mean(identical)), 6),
                "images_fully_bit_identical": int(sum(v == 1.0 for v in identical)), "images": n,
                "accuracy_delta_points": [round(a - b, 3) for a, b in zip(acc_o, acc_r)],
                "completeness_delta_points": [round(a - b, 3) for a, b in zip(comp_o, comp_r)]})
print(json.dumps(out), flush=True)
dst = args.out or os.path.join(ROOT, "gpurun_out", f"scene_parity_{args.scene}_p{args.planar}g{args.geom_planar}_{args.arithmetic}_{args.order}{'_shared' if args.share_prior else ''}.json")
os.makedirs(os.path.dirname(dst), exist_ok=True)
json.dump(out, open(dst, "w"), indent=1)
