"""Where does a planar-prior Run() at full size (bench step, 3200x2130, 10 sources) leave the reference's bits? Stage by stage
against the live reference (oracle/_ref), every half-sweep restarted from the reference's own state; prints one JSON line.

    python tests/tools/fullsize_prior_bisect.py [--size WxH] [--out file.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

import bench  # noqa: E402
import oracle_py  # noqa: E402
from mpmvs_b200 import capi  # noqa: E402
from parity_checks import colour_mask  # noqa: E402


def compare(so, sr, m, prior, mask, start=None):
    same_p = np.all(so["planes"] == sr["planes"], -1)
    same_c = so["costs"] == sr["costs"]
    same_v = so["views"] == sr["views"]
    same_r = np.all(so["rng"] == sr["rng"], -1)
    out = {"planes": float(same_p[m].mean()), "costs": float(same_c[m].mean()), "views": float(same_v[m].mean()), "rng": float(same_r[m].mean())}
    bad = m & ~(same_p & same_c & same_v)
    out["n_diff"] = int(bad.sum())
    out["n_diff_with_prior"] = int((bad & (mask > 0)).sum())
    out["n_diff_rng_differs"] = int((bad & ~same_r).sum())
    out["n_diff_plane_only_cost_same"] = int((bad & same_c & ~same_p).sum())
    ys, xs = np.nonzero(bad)
    ex = []
    for y, x in list(zip(ys, xs))[:6]:
        e = {"xy": [int(x), int(y)], "mask": int(mask[y, x]), "prior": [float(v) for v in prior[y, x]],
             "ours": {"plane": [float(v) for v in so["planes"][y, x]], "cost": float(so["costs"][y, x]), "views": int(so["views"][y, x])},
             "ref": {"plane": [float(v) for v in sr["planes"][y, x]], "cost": float(sr["costs"][y, x]), "views": int(sr["views"][y, x])},
             "rng_same": bool(same_r[y, x])}
        if start is not None:
            e["start"] = {"plane": [float(v) for v in start["planes"][y, x]], "cost": float(start["costs"][y, x])}
        ex.append(e)
    if ex:
        out["examples"] = ex
    return out


def main():
    prob = bench.load_problem("eth3d", 0, 1, lambda: None)
    imgs, cams = prob["images"], prob["cams"]
    if "--size" in sys.argv:          # centre crop: a smaller problem with the same content
        w, h = (int(t) for t in sys.argv[sys.argv.index("--size") + 1].split("x"))
        W, H = prob["width"], prob["height"]
        x0, y0 = (W - w) // 2, (H - h) // 2
        imgs = [np.ascontiguousarray(i[y0:y0 + h, x0:x0 + w]) for i in imgs]
        cams = cams.copy()
        for c in cams:
            c["K"][2] -= x0
            c["K"][5] -= y0
            c["width"], c["height"] = w, h
    pm = capi.PatchMatch(0).set_problem(imgs, cams)
    ref = oracle_py.Oracle("ref").set_problem(imgs, cams)
    for o in (pm, ref):
        o.set_geom_consistency_params(False, False)
        o.run(2)
    (pa, ca), (pb, cb) = pm.result(), ref.result()
    out = {"arithmetic": capi.default_arithmetic(), "photometric_run": {"planes": float(np.all(pa == pb, -1).mean()), "costs": float((ca == cb).mean())}}
    pm.set_planar_prior_params()
    pm.set_geom_consistency_params(False, True)
    pm.build_prior()
    prior, mask = pm.get_prior()
    out["prior_pixels"] = float((mask > 0).mean())
    out["prior_nonfinite"] = int((~np.isfinite(prior)).any(-1).sum())
    ref.set_planar_prior_params()
    ref.set_geom_consistency_params(False, True)
    ref.set_prior(prior, mask)
    ref.set_state(pb, cb)
    pm.set_state(pb, cb)
    for o in (pm, ref):
        o.init_only(3)
    sr = ref.get_state()
    h, w = sr["costs"].shape
    full = np.ones((h, w), bool)
    out["init"] = compare(pm.get_state(), sr, full, prior, mask)
    for it in range(3):
        for red in (0, 1):
            start = sr
            pm.set_dev_state(sr)
            pm.half_sweep(red, it, 0)
            ref.half_sweep(red, it, 0)
            sr = ref.get_state()
            out[f"i{it}r{red}"] = compare(pm.get_state(), sr, colour_mask(h, w, red), prior, mask, start)
    pm.set_dev_state(sr)
    for o in (pm, ref):
        o.finalize()
    sa, sb = pm.get_state(), ref.get_state()
    out["finalize"] = {"planes": float(np.all(sa["planes"] == sb["planes"], -1).mean()), "costs": float((sa["costs"] == sb["costs"]).mean())}
    line = json.dumps(out)
    if "--out" in sys.argv:
        with open(sys.argv[sys.argv.index("--out") + 1], "w") as f:
            f.write(line + "\n")
    print(line)


if __name__ == "__main__":
    main()
