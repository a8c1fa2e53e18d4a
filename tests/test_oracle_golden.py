"""Pins the CPU oracle (oracle/pm_oracle.c) against outputs of the REFERENCE ITSELF: tests/golden/*.npz were produced
on a B200 by oracle/_ref/libmpmvs_ref.so (= /root/reference/src/PatchMatch.cu compiled in place, fixed-seed redirection
only) with tests/golden/make_golden.py. Runs without a GPU. The reference ships no tests or vectors of its own.
"""
import os

import numpy as np
import pytest

from cases import CASES, SEED, make_case, prior_planes, random_planes, rng_hash, src_depths, world_state_from_gt
from conftest import ROOT, gt_planes_cam
from parity_checks import T_GEOM, check_cost_map, check_state, colour_mask, frac_within

GOLD = os.path.join(ROOT, "tests", "golden")
PIX = ((0, 0), (17, 5), (63, 47))


def load(name):
    p = os.path.join(GOLD, f"{name}.npz")
    if not os.path.exists(p):
        pytest.fail(f"golden fixture {p} missing: run tests/golden/make_golden.py on a GPU box")
    return np.load(p)


def state_of(g, prefix):
    return {k: g[f"{prefix}_{k}"] for k in ("planes", "costs", "views", "rng")}


@pytest.mark.parametrize("name", CASES)
def test_inputs_reproducible(name):
    """The renderer must still produce the images the golden outputs were computed from."""
    g, c = load(name), make_case(name)
    np.testing.assert_array_equal(g["images"], np.stack([i.astype(np.uint8) for i in c["images"]]))
    now = c["cams_rendered"]
    for k in ("K", "R", "t", "height", "width", "depth_min", "depth_max"):
        assert g["cams"][k].tobytes() == now[k].tobytes(), k
    # C = -R^T t: the stored records hold a BLAS product's rounding, Camera.C the reference's own order of operations (cases.py)
    np.testing.assert_allclose(g["cams"]["C"], now["C"], rtol=3e-7, atol=1e-7)


@pytest.mark.parametrize("name", CASES)
def test_xorwow_stream_bit_exact(oracle_cpu, name):
    g = load(name)
    o = oracle_cpu.Oracle("cpu")
    for (x, y), want in zip(PIX, g["uniform"]):
        np.testing.assert_array_equal(o.uniform_stream(SEED, x, y, 64), want)   # cuRAND XORWOW, bit for bit


@pytest.mark.parametrize("name", CASES)
def test_ncc_and_geom_maps(oracle_cpu, name):
    g, c = load(name), make_case(name)
    o = oracle_cpu.Oracle("cpu").set_problem(c["images"], c["cams"])
    gt, rnd = gt_planes_cam(c["scene"], c["ref"]), random_planes(c)
    for s in (0, 1, 2):
        check_cost_map(name, o.ncc_map(gt, s), g[f"ncc_gt_s{s}"])
        check_cost_map(name, o.ncc_map(rnd, s), g[f"ncc_rnd_s{s}"])
    o.set_geom_consistency_params(True, False)
    o.set_src_depths(src_depths(c, 0.002))
    assert frac_within(o.geom_map(rnd), g["geom_rnd"], T_GEOM) >= 0.999


@pytest.mark.parametrize("name", CASES)
def test_photometric_init_and_sweep(oracle_cpu, name):
    g, c = load(name), make_case(name)
    o = oracle_cpu.Oracle("cpu").set_problem(c["images"], c["cams"])
    o.set_geom_consistency_params(False, False)
    o.init_only(SEED)
    check_state(name, o.get_state(), state_of(g, "init"), "cpu", rng_digest=rng_hash, planes_exact=True)
    # continue from the reference's own init state so the sweep comparison starts bit-identical (rng comes from ours,
    # which the previous check proved identical to the reference's)
    st = o.get_state()
    st.update(planes=g["init_planes"], costs=g["init_costs"], views=g["init_views"])
    o.set_dev_state(st)
    o.half_sweep(0, 0, 2)
    h, w = g["init_costs"].shape
    check_state(name, o.get_state(), state_of(g, "sweep"), "cpu", upd=colour_mask(h, w, 0), rng_digest=rng_hash)


@pytest.mark.parametrize("name", CASES)
def test_geom_and_prior_sweeps(oracle_cpu, name):
    g, c = load(name), make_case(name)
    h, w = g["init_costs"].shape
    o = oracle_cpu.Oracle("cpu").set_problem(c["images"], c["cams"])
    o.set_geom_consistency_params(True, False)
    o.set_src_depths(src_depths(c, 0.002))
    o.set_state(*world_state_from_gt(c))
    o.init_only(SEED + 2)
    check_state(name, o.get_state(), state_of(g, "ginit"), "cpu", rng_digest=rng_hash, planes_exact=True)
    st = o.get_state()
    st.update(planes=g["ginit_planes"], costs=g["ginit_costs"], views=g["ginit_views"])
    o.set_dev_state(st)
    o.half_sweep(0, 0, 0)
    got = o.get_state()
    check_state(name, got, state_of(g, "gsweep"), "cpu", upd=colour_mask(h, w, 0), rng_digest=rng_hash)
    close = np.all(np.abs(got["planes"] - g["gsweep_planes"]) <= 1e-4 * (1 + np.abs(g["gsweep_planes"])), -1)
    assert np.abs(got["geom"] - g["gsweep_geom"])[close].mean() < 2e-3
    # planar prior, starting from the reference's photometric result
    o = oracle_cpu.Oracle("cpu").set_problem(c["images"], c["cams"])
    o.set_geom_consistency_params(False, False)
    o.set_state(g["run_planes"], g["run_costs"])
    o.set_planar_prior_params()
    o.set_geom_consistency_params(False, True)
    o.set_prior(*prior_planes(c))
    o.init_only(SEED + 1)
    check_state(name, o.get_state(), state_of(g, "pinit"), "cpu", rng_digest=rng_hash, planes_exact=True)
    st = o.get_state()
    st.update(planes=g["pinit_planes"], costs=g["pinit_costs"], views=g["pinit_views"])
    o.set_dev_state(st)
    o.half_sweep(0, 0, 0)
    check_state(name, o.get_state(), state_of(g, "psweep"), "cpu", upd=colour_mask(h, w, 0), rng_digest=rng_hash)


@pytest.mark.parametrize("name", CASES)
def test_full_run_statistics(oracle_cpu, pkg, name):
    """Whole photometric Run(): trajectories diverge (chaotic), so compare statistics against the reference's result."""
    g, c = load(name), make_case(name)
    o = oracle_cpu.Oracle("cpu").set_problem(c["images"], c["cams"])
    o.set_geom_consistency_params(False, False)
    o.run(SEED)
    planes, costs = o.result()
    gt = c["scene"].gt_depth[c["ref"]]
    acc_o = pkg.synth.accuracy_at(planes[..., 3], gt)
    acc_r = pkg.synth.accuracy_at(g["run_planes"][..., 3], gt)
    assert all(abs(a - b) < 3.0 for a, b in zip(acc_o, acc_r)), (acc_o, acc_r)      # tiny images: 1 point ~ 30-60 pixels
    assert abs(float(costs.mean()) - float(g["run_costs"].mean())) < 0.03
