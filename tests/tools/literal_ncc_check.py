"""How far is the product's per-pixel arithmetic from the literal restatement, and what closes the gap? (CPU only.)

Instantiates mp-mvs_b200/csrc/pm_core.cuh on the host (tests/emul) three ways -- shipped form, -DPM_LITERAL_WARP=1 (the
reference's operation order for the homography and the tap coordinates), -DPM_LITERAL_NCC=1 (additionally its bilateral
weights, row-wise sums, products and the geometric cost) -- and compares each with oracle/pm_oracle.c on the three parity
cases: NCC maps at all scales, then whole runs in the three modes. Results of the round it was written in:
profiles/r01_literal_variant.md.

    python tests/tools/literal_ncc_check.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle_py  # noqa: E402
from cases import CASES, SEED, make_case, prior_planes, random_planes, src_depths, world_state_from_gt  # noqa: E402
from conftest import build_emul, gt_planes_cam  # noqa: E402

VARIANTS = {"shipped": ("", []), "literal-warp": ("_litwarp", ["-DPM_LITERAL_WARP=1"]), "literal": ("_literal", ["-DPM_LITERAL_NCC=1"])}


def main():
    oracle_py.build("cpu")
    for tag, (suffix, flags) in VARIANTS.items():
        oracle_py.LIBS["emul:" + tag] = (build_emul(suffix, flags), "emu_")
    print("== NCC maps: |product - oracle| (mean, max) and fraction of bit-equal costs")
    for name in CASES:
        c = make_case(name)
        o = oracle_py.Oracle("cpu").set_problem(c["images"], c["cams"])
        es = {t: oracle_py.Oracle("emul:" + t).set_problem(c["images"], c["cams"]) for t in VARIANTS}
        for what, pl in (("gt planes", gt_planes_cam(c["scene"], c["ref"])), ("random planes", random_planes(c))):
            for s in (0, 1, 2):
                a = o.ncc_map(pl, s)
                row = [f"{t}: {np.abs(a - e.ncc_map(pl, s)).mean():.1e} {np.abs(a - e.ncc_map(pl, s)).max():.1e} {np.mean(a == e.ncc_map(pl, s)):.3f}"
                       for t, e in es.items()]
                print(f"{name:7s} {what:13s} scale {s} | " + " | ".join(row))
    print("== whole runs, same seed: fraction of bit-identical planes (photometric / prior / geom run) and geom costs")
    for name in CASES:
        c = make_case(name)
        for t in VARIANTS:
            o = oracle_py.Oracle("cpu").set_problem(c["images"], c["cams"])
            e = oracle_py.Oracle("emul:" + t).set_problem(c["images"], c["cams"])
            for x in (o, e):
                x.set_geom_consistency_params(False, False)
                x.run(SEED)
            photo = np.all(o.result()[0] == e.result()[0], -1).mean()
            for x in (o, e):
                x.set_planar_prior_params()
                x.set_geom_consistency_params(False, True)
                x.set_prior(*prior_planes(c))
                x.run(SEED + 1)
            prior = np.all(o.result()[0] == e.result()[0], -1).mean()
            o.destroy(); e.destroy()
            o = oracle_py.Oracle("cpu").set_problem(c["images"], c["cams"])
            e = oracle_py.Oracle("emul:" + t).set_problem(c["images"], c["cams"])
            for x in (o, e):
                x.set_geom_consistency_params(True, False)
                x.set_src_depths(src_depths(c, 0.002))
                x.set_state(*world_state_from_gt(c))
                x.run(SEED + 2)
            ro, re_ = o.result(geom=True), e.result(geom=True)
            print(f"{name:7s} {t:13s} planes {photo:.4f} / {prior:.4f} / {np.all(ro[0] == re_[0], -1).mean():.4f}   geom costs {np.mean(ro[2] == re_[2]):.4f}")
            o.destroy(); e.destroy()


if __name__ == "__main__":
    main()
