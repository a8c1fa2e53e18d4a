"""Full-size (BASELINE.json configs[2] shape, 3200x2130, 10 sources) timing of one arithmetic of the library and, with
--check-ref, bit identity against the live reference of the WHOLE bench step: a same-seed photometric Run(), then -- with
the same prior planes, those our planar-prior stage built -- a same-seed planar-prior Run(). No torch; prints one JSON line.

    [MPMVS_ARITHMETIC=fast] [MPMVS_LIB_VARIANT=<tuning variant>] python tests/tools/variant_fullsize.py [--check-ref] [--tex u8|f16|f32]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

import bench  # noqa: E402  (workload + /dev/shm cache; imports no torch at module level)
from mpmvs_b200 import capi  # noqa: E402


def main():
    t0 = time.time()
    prob = bench.load_problem("eth3d", 0, 1, lambda: None)
    tex = sys.argv[sys.argv.index("--tex") + 1] if "--tex" in sys.argv else "f32"
    assert tex in ("u8", "f16", "f32")
    imgs = [i.astype(np.uint8) for i in prob["images"]] if tex == "u8" else prob["images"]
    out = {"variant": os.environ.get("MPMVS_LIB_VARIANT", "shipped"), "arithmetic": capi.default_arithmetic(), "tex": tex, "load_s": round(time.time() - t0, 1)}
    pm = capi.PatchMatch(0).set_tex_format({"u8": capi.TEX_U8, "f16": capi.TEX_F16, "f32": capi.TEX_F32}[tex]).set_problem(imgs, prob["cams"])
    pm.set_geom_consistency_params(False, False)
    pm.run(1)                                           # warm-up
    out["photometric_run_ms"] = round(float(pm.run(2)), 2)
    planes, costs = pm.result()
    if "--check-ref" in sys.argv:
        import oracle_py

        ref = oracle_py.Oracle("ref").set_problem(prob["images"], prob["cams"])
        ref.set_geom_consistency_params(False, False)
        out["reference_run_ms"] = round(float(ref.run(2)), 2)
        rp, rc = ref.result()
        out["run_planes_identical"] = float(np.all(planes == rp, -1).mean())
        out["run_costs_identical"] = float((costs == rc).mean())
        out["max_abs_depth_diff"] = float(np.abs(planes[..., 3] - rp[..., 3]).max())
    # the rest of the bench step: planar-prior stage + prior Run() (ProcessProblem(geom=0, planar=1))
    pm.set_planar_prior_params()
    pm.set_geom_consistency_params(False, True)
    st = pm.build_prior()
    out["prior_run_ms"] = round(float(pm.run(3)), 2)
    out["delaunay_ms"] = round(float(st["delaunay_ms"]), 1)
    if "--check-ref" in sys.argv:
        # the reference's second Run() of ProcessProblem with the SAME prior (CudaPlanarPriorInitialization, cpp:978-996),
        # on top of its own photometric result (bit-identical to ours if the check above passed)
        prior, mask = pm.get_prior()
        ref.set_planar_prior_params()
        ref.set_geom_consistency_params(False, True)
        ref.set_prior(prior, mask)
        out["reference_prior_run_ms"] = round(float(ref.run(3)), 2)
        (pp, pc), (rp, rc) = pm.result(), ref.result()
        out["prior_run_planes_identical"] = float(np.all(pp == rp, -1).mean())
        out["prior_run_costs_identical"] = float((pc == rc).mean())
        out["prior_pixels_fraction"] = float((mask > 0).mean())
        ref.destroy()
    out["step_device_ms"] = round(out["photometric_run_ms"] + out["prior_run_ms"], 2)
    out["total_s"] = round(time.time() - t0, 1)
    pm.destroy()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
