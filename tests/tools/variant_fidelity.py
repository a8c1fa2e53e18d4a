"""Bit-fidelity of ONE library build against the live reference (oracle/_ref) on the GPU; prints one JSON line.

The library is chosen at import time (MPMVS_LIB_VARIANT=<name> -> mp-mvs_b200/variants/libmpmvs_b200_<name>.so, unset = the
shipped in-tree library), so every build gets its own process:

    MPMVS_ARITHMETIC=exact python tests/tools/variant_fidelity.py
    python tests/tools/variant_fidelity.py                    # the shipped kernels, for comparison
    ... --modes                                                # also whole planar-prior and geometric-consistency runs

For the three parity cases, float32 view storage: every half-sweep of a photometric Run() started from the reference's own
state (fraction of bit-identical planes / costs / view masks among the updated pixels, minimum and mean over the 18
half-sweeps), then a whole same-seed Run() (fraction of bit-identical result planes and costs), and the device time of that
Run() for both. tests/test_zz_fidelity_build_gpu.py reads the JSON.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

from cases import CASES, SEED, make_case  # noqa: E402  (conftest loads the package)
import oracle_py  # noqa: E402
from mpmvs_b200 import capi  # noqa: E402
from parity_checks import colour_mask  # noqa: E402


def main():
    out = {"variant": os.environ.get("MPMVS_LIB_VARIANT", "shipped"), "library": capi.LIB_PATH, "arithmetic": capi.default_arithmetic(), "cases": {}}
    for name in CASES:
        c = make_case(name)
        pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
        ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        for o in (pm, ref):
            o.set_geom_consistency_params(False, False)
            o.init_only(SEED)
        sr = ref.get_state()
        h, w = sr["costs"].shape
        planes, costs, views = [], [], []
        for scale in (2, 1, 0):
            for it in range(3):
                for red in (0, 1):
                    pm.set_dev_state(sr)
                    pm.half_sweep(red, it, scale)
                    ref.half_sweep(red, it, scale)
                    sr = ref.get_state()
                    so = pm.get_state()
                    m = colour_mask(h, w, red)
                    planes.append(float(np.all(so["planes"] == sr["planes"], -1)[m].mean()))
                    costs.append(float((so["costs"] == sr["costs"])[m].mean()))
                    views.append(float((so["views"] == sr["views"])[m].mean()))
        pm.destroy(); ref.destroy()
        pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
        ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        for o in (pm, ref):
            o.set_geom_consistency_params(False, False)
        ms = pm.run(SEED)
        ms_ref = ref.run(SEED)
        (pa, ca), (pb, cb) = pm.result(), ref.result()
        out["cases"][name] = {
            "half_sweep_planes_min": min(planes), "half_sweep_planes_mean": float(np.mean(planes)),
            "half_sweep_costs_min": min(costs), "half_sweep_views_min": min(views),
            "run_planes_identical": float(np.all(pa == pb, -1).mean()), "run_costs_identical": float((ca == cb).mean()),
            "run_ms": float(ms), "run_ms_reference": float(ms_ref)}
        pm.destroy(); ref.destroy()
    if "--modes" in sys.argv:
        out["modes"] = other_modes()
    print(json.dumps(out))


def other_modes():
    """Whole same-seed planar-prior Run() (on top of a photometric one, state resident) and geometric-consistency Run() on
    the three cases: fractions of bit-identical planes / costs / geometric costs against the reference."""
    from cases import prior_planes, src_depths, world_state_from_gt

    res = {}
    for name in CASES:
        c = make_case(name)
        pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
        ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        for o in (pm, ref):
            o.set_geom_consistency_params(False, False)
            o.run(SEED)
            o.set_planar_prior_params()
            o.set_geom_consistency_params(False, True)
            o.set_prior(*prior_planes(c))
            o.run(SEED + 1)
        (pa, ca), (pb, cb) = pm.result(), ref.result()
        r = {"prior_run_planes_identical": float(np.all(pa == pb, -1).mean()), "prior_run_costs_identical": float((ca == cb).mean())}
        pm.destroy(); ref.destroy()
        pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
        ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        for o in (pm, ref):
            o.set_geom_consistency_params(True, False)
            o.set_src_depths(src_depths(c, 0.002))
            o.set_state(*world_state_from_gt(c))
            o.run(SEED + 2)
        ga, gb = pm.result(geom=True), ref.result(geom=True)
        r.update({"geom_run_planes_identical": float(np.all(ga[0] == gb[0], -1).mean()), "geom_run_costs_identical": float((ga[1] == gb[1]).mean()),
                  "geom_run_geom_costs_identical": float((ga[2] == gb[2]).mean())})
        pm.destroy(); ref.destroy()
        res[name] = r
    return res


if __name__ == "__main__":
    main()
