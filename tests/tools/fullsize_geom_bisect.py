"""Geometric-consistency mode at scale (DTU-shaped 1600x1200, 10 sources: BASELINE.json configs[1] shape) against the live
reference (oracle/_ref), stage by stage: InitializeScore from (world normal, depth) planes, the geometric-cost map, every
half-sweep restarted from the reference's own state, the finalize kernels, and a whole Run(). Prints one JSON line.

    python tests/tools/fullsize_geom_bisect.py [--out file.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

from conftest import problem_arrays  # noqa: E402
from cases import src_depths, world_state_from_gt  # noqa: E402
import oracle_py  # noqa: E402
from mpmvs_b200 import capi, synth  # noqa: E402
from parity_checks import colour_mask  # noqa: E402


def compare(so, sr, m):
    same_p = np.all(so["planes"] == sr["planes"], -1)
    same_c = so["costs"] == sr["costs"]
    same_v = so["views"] == sr["views"]
    same_g = so["geom"] == sr["geom"]
    same_r = np.all(so["rng"] == sr["rng"], -1)
    out = {"planes": float(same_p[m].mean()), "costs": float(same_c[m].mean()), "views": float(same_v[m].mean()),
           "geom": float(same_g[m].mean()), "rng": float(same_r[m].mean())}
    bad = m & ~(same_p & same_c & same_v & same_g)
    out["n_diff"] = int(bad.sum())
    ys, xs = np.nonzero(bad)
    ex = []
    for y, x in list(zip(ys, xs))[:5]:
        ex.append({"xy": [int(x), int(y)],
                   "ours": {"plane": [float(v) for v in so["planes"][y, x]], "cost": float(so["costs"][y, x]), "geom": float(so["geom"][y, x])},
                   "ref": {"plane": [float(v) for v in sr["planes"][y, x]], "cost": float(sr["costs"][y, x]), "geom": float(sr["geom"][y, x])}})
    if ex:
        out["examples"] = ex
    return out


def main():
    w, h = 1600, 1200
    if "--size" in sys.argv:
        w, h = (int(t) for t in sys.argv[sys.argv.index("--size") + 1].split("x"))
    probe = synth.make_dtu_scene(views=[])
    ids = [24] + [i for i, _ in probe.pairs[24]][:10]
    sc = synth.make_dtu_scene(width=w, height=h, views=ids, workers=min(16, os.cpu_count() or 1), jpeg=False)
    pids, imgs, cams = problem_arrays(sc, 24, 10)
    c = dict(scene=sc, ref=24, ids=pids, images=imgs, cams=cams)
    pm = capi.PatchMatch(0).set_problem(imgs, cams)
    ref = oracle_py.Oracle("ref").set_problem(imgs, cams)
    for o in (pm, ref):
        o.set_geom_consistency_params(True, False)
        o.set_src_depths(src_depths(c, 0.002))
        o.set_state(*world_state_from_gt(c))
        o.init_only(7)
    sr = ref.get_state()
    hh, ww = sr["costs"].shape
    full = np.ones((hh, ww), bool)
    out = {"arithmetic": capi.default_arithmetic(), "size": [ww, hh], "init": compare(pm.get_state(), sr, full)}
    ga, gb = pm.geom_map(sr["planes"]), ref.geom_map(sr["planes"])
    out["geom_map"] = {"identical": float((ga == gb).mean()), "max_abs_diff": float(np.nanmax(np.abs(ga - gb)))}
    for it in range(2):
        for red in (0, 1):
            pm.set_dev_state(sr)
            pm.half_sweep(red, it, 0)
            ref.half_sweep(red, it, 0)
            sr = ref.get_state()
            out[f"i{it}r{red}"] = compare(pm.get_state(), sr, colour_mask(hh, ww, red))
    pm.set_dev_state(sr)
    for o in (pm, ref):
        o.finalize()
    sa, sb = pm.get_state(), ref.get_state()
    out["finalize"] = {"planes": float(np.all(sa["planes"] == sb["planes"], -1).mean()), "costs": float((sa["costs"] == sb["costs"]).mean())}
    for o in (pm, ref):
        o.set_state(*world_state_from_gt(c))
    ms, ms_ref = pm.run(8), ref.run(8)
    ra, rb = pm.result(geom=True), ref.result(geom=True)
    out["whole_run"] = {"planes": float(np.all(ra[0] == rb[0], -1).mean()), "costs": float((ra[1] == rb[1]).mean()), "geom": float((ra[2] == rb[2]).mean()),
                        "run_ms": float(ms), "reference_run_ms": float(ms_ref)}
    line = json.dumps(out)
    if "--out" in sys.argv:
        with open(sys.argv[sys.argv.index("--out") + 1], "w") as f:
            f.write(line + "\n")
    print(line)


if __name__ == "__main__":
    main()
