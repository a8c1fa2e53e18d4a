"""Small end-to-end case for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
Photometric Run, planar-prior stage + Run, geometric Run on a 96x72 5-view problem, both view storage formats."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import PKG, problem_arrays  # noqa: E402
from mpmvs_b200 import capi  # noqa: E402

sc = PKG.synth.make_dtu_scene(width=96, height=72, grid=3, n_src=4, seed=2, jpeg=False)
ids, imgs, cams = problem_arrays(sc, 4)
for fmt in (capi.TEX_F32, capi.TEX_U8):
    pm = capi.PatchMatch(0).set_tex_format(fmt).set_problem(imgs, cams)
    pm.set_geom_consistency_params(False, True)
    pm.run(1)
    pm.set_planar_prior_params()
    pm.set_geom_consistency_params(False, True)
    st = pm.build_prior()
    pm.run(2)
    planes, costs = pm.result()
    pm.reset_params()
    pm.set_geom_consistency_params(True, True)
    pm.set_src_depths([sc.gt_depth[i] for i in ids[1:]])
    pm.run(3)
    p2, c2, g2 = pm.result(geom=True)
    print("fmt", fmt, "prior", st["n_vertices"], st["n_triangles"], "mean cost", float(costs.mean()), float(c2.mean()), "finite", bool(np.isfinite(p2).all()))
    pm.destroy()
print("sanitize case done")
