"""Stage-by-stage bit comparison of ONE library build against the live reference (oracle/_ref) in the planar-prior and the
geometric-consistency modes; prints one JSON line (and writes it to the path given with --out).

    MPMVS_ARITHMETIC=exact python tests/tools/modes_bisect.py [--out gpurun_out/x.json] [--cases room6,dtu5]

Per case and mode: InitializeScore from the same state, the geometric-cost map on the reference's planes (geom mode),
every half-sweep restarted from the reference's own state (so a difference is local to one launch), then the finalize
kernels. For every stage: fraction of bit-identical planes / costs / view masks / RNG states / geometric costs among the
pixels the stage may write, and a few of the differing pixels with both sides' values (how far apart, which field first).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

from cases import CASES, SEED, make_case, prior_planes, src_depths, world_state_from_gt  # noqa: E402
import oracle_py  # noqa: E402
from mpmvs_b200 import capi  # noqa: E402
from parity_checks import colour_mask  # noqa: E402


def compare(so, sr, m, with_geom):
    out = {
        "planes": float(np.all(so["planes"] == sr["planes"], -1)[m].mean()),
        "costs": float((so["costs"] == sr["costs"])[m].mean()),
        "views": float((so["views"] == sr["views"])[m].mean()),
        "rng": float(np.all(so["rng"] == sr["rng"], -1)[m].mean()),
    }
    if with_geom:
        out["geom"] = float((so["geom"] == sr["geom"])[m].mean())
    bad = m & ~(np.all(so["planes"] == sr["planes"], -1) & (so["costs"] == sr["costs"]) & (so["views"] == sr["views"]))
    if with_geom:
        bad |= m & (so["geom"] != sr["geom"])
    ex = []
    ys, xs = np.nonzero(bad)
    for y, x in list(zip(ys, xs))[:4]:
        ex.append({"xy": [int(x), int(y)],
                   "ours": {"plane": [float(v) for v in so["planes"][y, x]], "cost": float(so["costs"][y, x]), "views": int(so["views"][y, x]),
                            "geom": float(so["geom"][y, x])},
                   "ref": {"plane": [float(v) for v in sr["planes"][y, x]], "cost": float(sr["costs"][y, x]), "views": int(sr["views"][y, x]),
                           "geom": float(sr["geom"][y, x])},
                   "rng_same": bool((so["rng"][y, x] == sr["rng"][y, x]).all())})
    out["n_diff"] = int(bad.sum())
    if ex:
        out["examples"] = ex
    return out


def sweeps(pm, ref, iters, with_geom):
    res = {}
    sr = ref.get_state()
    h, w = sr["costs"].shape
    for it in range(iters):
        for red in (0, 1):
            pm.set_dev_state(sr)
            pm.half_sweep(red, it, 0)
            ref.half_sweep(red, it, 0)
            sr = ref.get_state()
            res[f"i{it}r{red}"] = compare(pm.get_state(), sr, colour_mask(h, w, red), with_geom)
    return res


def prior_mode(c):
    pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
    ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
    for o in (pm, ref):
        o.set_geom_consistency_params(False, False)
        o.run(SEED)
    (pa, ca), (pb, cb) = pm.result(), ref.result()
    res = {"photometric_run": {"planes": float(np.all(pa == pb, -1).mean()), "costs": float((ca == cb).mean())}}
    for o in (pm, ref):
        o.set_planar_prior_params()
        o.set_geom_consistency_params(False, True)
        o.set_prior(*prior_planes(c))
    pm.set_state(pb, cb)        # identical start even if the photometric runs differed
    ref.set_state(pb, cb)
    for o in (pm, ref):
        o.init_only(SEED + 1)
    sr = ref.get_state()
    full = np.ones(sr["costs"].shape, bool)
    res["init"] = compare(pm.get_state(), sr, full, False)
    res.update(sweeps(pm, ref, 3, False))
    pm.set_dev_state(ref.get_state())
    for o in (pm, ref):
        o.finalize()
    sa, sb = pm.get_state(), ref.get_state()
    res["finalize"] = {"planes": float(np.all(sa["planes"] == sb["planes"], -1).mean()), "costs": float((sa["costs"] == sb["costs"]).mean())}
    # and the whole run, as the caller does it
    for o in (pm, ref):
        o.set_state(sb["planes"], sb["costs"])
        o.run(SEED + 3)
    (pa, ca), (pb, cb) = pm.result(), ref.result()
    res["whole_run"] = {"planes": float(np.all(pa == pb, -1).mean()), "costs": float((ca == cb).mean())}
    pm.destroy(); ref.destroy()
    return res


def geom_mode(c):
    pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
    ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
    for o in (pm, ref):
        o.set_geom_consistency_params(True, False)
        o.set_src_depths(src_depths(c, 0.002))
        o.set_state(*world_state_from_gt(c))
        o.init_only(SEED + 2)
    sr = ref.get_state()
    full = np.ones(sr["costs"].shape, bool)
    res = {"init": compare(pm.get_state(), sr, full, False)}
    ga, gb = pm.geom_map(sr["planes"]), ref.geom_map(sr["planes"])
    res["geom_map"] = {"identical": float((ga == gb).mean()), "max_abs_diff": float(np.nanmax(np.abs(ga - gb)))}
    res.update(sweeps(pm, ref, 2, True))
    pm.set_dev_state(ref.get_state())
    for o in (pm, ref):
        o.finalize()
    sa, sb = pm.get_state(), ref.get_state()
    res["finalize"] = {"planes": float(np.all(sa["planes"] == sb["planes"], -1).mean()), "costs": float((sa["costs"] == sb["costs"]).mean())}
    for o in (pm, ref):
        o.set_state(*world_state_from_gt(c))
        o.run(SEED + 2)
    ra, rb = pm.result(geom=True), ref.result(geom=True)
    res["whole_run"] = {"planes": float(np.all(ra[0] == rb[0], -1).mean()), "costs": float((ra[1] == rb[1]).mean()),
                        "geom": float((ra[2] == rb[2]).mean())}
    pm.destroy(); ref.destroy()
    return res


def main():
    names = CASES
    if "--cases" in sys.argv:
        names = sys.argv[sys.argv.index("--cases") + 1].split(",")
    out = {"library": capi.LIB_PATH, "arithmetic": capi.default_arithmetic(), "prior": {}, "geom": {}}
    def save():
        if "--out" in sys.argv:
            with open(sys.argv[sys.argv.index("--out") + 1], "w") as f:
                f.write(json.dumps(out) + "\n")

    for name in names:
        c = make_case(name)
        out["prior"][name] = prior_mode(c)
        save()
        out["geom"][name] = geom_mode(c)
        save()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
