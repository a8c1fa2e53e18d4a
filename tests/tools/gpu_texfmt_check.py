"""Storage-format check on a GPU: NCC maps and whole runs with f32 / f16 / u8 view storage vs the reference (prints)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import PKG, gt_planes_cam, problem_arrays  # noqa: E402
from cases import CASES, make_case, random_planes  # noqa: E402
import oracle_py  # noqa: E402
from mpmvs_b200 import capi, synth  # noqa: E402

for name in CASES:
    c = make_case(name)
    ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
    rnd = random_planes(c)
    want = ref.ncc_map(rnd, 1)
    for fmt, tag in ((0, "f32"), (1, "f16"), (2, "u8")):
        pm = capi.PatchMatch(0).set_tex_format(fmt).set_problem(c["images"], c["cams"])
        d = np.abs(pm.ncc_map(rnd, 1) - want)
        print(f"{name} ncc s1 {tag}: max {d.max():.3e} mean {d.mean():.3e} frac>2e-3 {(d > 2e-3).mean():.5f}")
        pm.destroy()
    ref.destroy()

for scene, refv, tag in ((synth.make_plane_scene(), 1, "plane640"),
                         (synth.make_dtu_scene(width=800, height=600, grid=3, n_src=6, jpeg=True), 4, "dtu800")):
    ids, imgs, cams = problem_arrays(scene, refv)
    gt, gtn = scene.gt_depth[refv], scene.gt_normal[refv]
    res = {}
    R = oracle_py.Oracle("ref").set_problem(imgs, cams)
    R.set_geom_consistency_params(False, False)
    for seed in (1, 2, 3):
        R.run(seed)
        res["ref", seed] = R.result()
    R.destroy()
    for fmt, ft in ((0, "f32"), (1, "f16"), (2, "u8")):
        pm = capi.PatchMatch(0).set_tex_format(fmt).set_problem(imgs, cams)
        pm.set_geom_consistency_params(False, False)
        for seed in (1, 2):
            pm.run(seed)
            res[ft, seed] = pm.result()
        pm.destroy()

    def agree(a, b, m):
        return synth.depth_normal_agreement(a[0][..., 3], a[0][..., :3], b[0][..., 3], b[0][..., :3], m)

    r1 = res["ref", 1]
    gt_ok = np.abs(r1[0][..., 3] - gt) <= 0.01 * gt
    for vname, valid in (("gt&cost<0.5", (gt > 0) & (r1[1] < 0.5)), ("gt&cost<0.1", (gt > 0) & (r1[1] < 0.1)),
                         ("gt&cost<0.05", (gt > 0) & (r1[1] < 0.05)), ("gt&ref within 1% of gt", (gt > 0) & gt_ok),
                         ("gt&cost<0.1&ref within 1% gt", (gt > 0) & gt_ok & (r1[1] < 0.1))):
        line = f"[{tag}] valid={vname} ({valid.mean():.3f}): ref1-ref2 {agree(r1, res['ref', 2], valid):.4f} ref1-ref3 {agree(r1, res['ref', 3], valid):.4f}"
        for ft in ("f32", "f16", "u8"):
            line += f" | {ft}: same-seed {agree(res[ft, 1], r1, valid):.4f} other-seed {agree(res[ft, 2], r1, valid):.4f}"
        print(line)
    for k, v in res.items():
        print(f"[{tag}] {k}: acc {['%.2f' % a for a in synth.accuracy_at(v[0][..., 3], gt)]} mean cost {v[1].mean():.4f}")
