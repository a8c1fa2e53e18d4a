"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libmpmvs_ref.so = /root/reference/src/PatchMatch.cu
compiled in place for sm_100 with the fixed-seed redirection of oracle/ref_harness.cu). Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py && cp tests/golden/*.npz gpurun_out/golden/'

The reference ships no tests or golden vectors; these files are what pins the CPU oracle
(tests/test_oracle_golden.py, runs without a GPU) and, through it and directly, the CUDA product.

Per case (tests/cases.py) the file holds the uint8 input images, packed cameras and the reference's outputs:
  uniform      first 64 curand_uniform draws of three pixels under SEED
  ncc_gt_s{0,1,2}, ncc_rnd_s{0,1,2}   ComputeBilateralNCC maps for GT planes / perturbed planes
  geom_rnd     ComputeGeomConsistencyCost map for the perturbed planes
  init_*       state after InitializeScore (photometric mode)
  sweep_*      state after one Black half-sweep at scale 2 from the init state
  gsweep_*     geom mode: init from a (world normal, depth) state + one Black half-sweep at scale 0
  psweep_*     planar-prior mode: init + one Black half-sweep at scale 0
  run_depth / run_cost / run_normal   full photometric Run()
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import PKG  # noqa: E402,F401
from cases import (CASES, SEED, make_case, rng_hash, prior_planes, random_planes, src_depths, world_state_from_gt)  # noqa: E402
from conftest import gt_planes_cam  # noqa: E402
import oracle_py  # noqa: E402

STATE_KEYS = ("planes", "costs", "views", "rng", "geom")
PIX = ((0, 0), (17, 5), (63, 47))


def put_state(out, prefix, st, geom=False):
    for k in STATE_KEYS:
        if k == "geom" and not geom:
            continue
        out[f"{prefix}_{k}"] = rng_hash(st[k]) if k == "rng" else st[k]


def main():
    os.makedirs(HERE, exist_ok=True)
    for name in CASES:
        c = make_case(name)
        out = {"images": np.stack([i.astype(np.uint8) for i in c["images"]]), "cams": c["cams"], "ref": c["ref"]}
        R = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        out["uniform"] = np.stack([R.uniform_stream(SEED, x, y, 64) for x, y in PIX])
        gt, rnd = gt_planes_cam(c["scene"], c["ref"]), random_planes(c)
        for s in (0, 1, 2):
            out[f"ncc_gt_s{s}"] = R.ncc_map(gt, s)
            out[f"ncc_rnd_s{s}"] = R.ncc_map(rnd, s)
        R.set_geom_consistency_params(False, False)
        R.init_only(SEED)
        put_state(out, "init", R.get_state())
        R.half_sweep(0, 0, 2)
        put_state(out, "sweep", R.get_state())
        R.run(SEED)
        planes, costs = R.result()
        out["run_planes"], out["run_costs"] = planes, costs
        # planar prior on top of the photometric result (device state stays resident, as in ProcessProblem)
        R.set_planar_prior_params()
        R.set_geom_consistency_params(False, True)
        R.set_prior(*prior_planes(c))
        R.init_only(SEED + 1)
        put_state(out, "pinit", R.get_state())
        R.half_sweep(0, 0, 0)
        put_state(out, "psweep", R.get_state())
        R.destroy()
        # geometric consistency pass from a stored (world normal, depth) state
        R = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        R.set_geom_consistency_params(True, False)
        R.set_src_depths(src_depths(c, 0.002))
        R.set_state(*world_state_from_gt(c))
        out["geom_rnd"] = R.geom_map(rnd)
        R.init_only(SEED + 2)
        put_state(out, "ginit", R.get_state(), geom=True)
        R.half_sweep(0, 0, 0)
        put_state(out, "gsweep", R.get_state(), geom=True)
        R.destroy()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
