"""Host instantiation of the product's per-pixel templates (mp-mvs_b200/csrc/pm_core.cuh via tests/emul) against the
CPU oracle (oracle/pm_oracle.c), stage by stage. This is how the kernel logic is checked where there is no GPU; the
GPU parity tests repeat the same comparisons with the real kernels and the real reference.

Tolerances: the product evaluates the homography in the hoisted form A + b m^T and folds constants, so NCC costs differ
from the literal restatement by float rounding amplified by the texture unit's 8-bit interpolation weights: |dcost| is
required < 2e-3 for >= 97% of samples (weak-texture scenes are the worst case); RNG streams must stay bit-identical
(same number of draws per pixel), planes produced by pure RNG paths must be bit-identical.
"""
import numpy as np
import pytest

from cases import CASES, SEED, make_case, prior_planes, random_planes, src_depths, world_state_from_gt
from conftest import build_emul, gt_planes_cam


@pytest.fixture(scope="module")
def libs(oracle_cpu):
    build_emul()
    return oracle_cpu


def frac_close(a, b, tol):
    return float((np.abs(a - b) <= tol).mean())


@pytest.mark.parametrize("name", CASES)
def test_ncc_and_geom_maps(libs, name):
    c = make_case(name)
    o = libs.Oracle("cpu").set_problem(c["images"], c["cams"])
    e = libs.Oracle("emul").set_problem(c["images"], c["cams"])
    for pl in (gt_planes_cam(c["scene"], c["ref"]), random_planes(c)):
        for s in (0, 1, 2):
            a, b = o.ncc_map(pl, s), e.ncc_map(pl, s)
            assert a.min() >= 0 and a.max() <= 2
            assert frac_close(a, b, 2e-3) > 0.95, (name, s)
            assert np.abs(a - b).mean() < 1e-3
    for x in (o, e):
        x.set_geom_consistency_params(True, False)
        x.set_src_depths(src_depths(c, 0.002))
    a, b = o.geom_map(random_planes(c)), e.geom_map(random_planes(c))
    assert a.max() <= 3.0 and frac_close(a, b, 1e-3) > 0.999


@pytest.mark.parametrize("name", CASES)
def test_photometric_stages(libs, name):
    c = make_case(name)
    o = libs.Oracle("cpu").set_problem(c["images"], c["cams"])
    e = libs.Oracle("emul").set_problem(c["images"], c["cams"])
    for x in (o, e):
        x.set_geom_consistency_params(False, False)
        x.init_only(SEED)
    so, se = o.get_state(), e.get_state()
    np.testing.assert_array_equal(so["planes"], se["planes"])       # pure RNG path
    np.testing.assert_array_equal(so["rng"], se["rng"])
    assert frac_close(so["costs"], se["costs"], 2e-3) > 0.93
    assert (so["views"] == se["views"]).mean() > 0.99
    for scale in (2, 1, 0):
        for red in (0, 1):
            e.set_dev_state(so)
            o.half_sweep(red, 0, scale)
            e.half_sweep(red, 0, scale)
            so, se = o.get_state(), e.get_state()
            np.testing.assert_array_equal(so["rng"], se["rng"])     # identical control flow -> identical draw counts
            h, w = so["costs"].shape
            yy, xx = np.mgrid[0:h, 0:w]
            upd = ((xx + yy) & 1) == red
            np.testing.assert_array_equal(so["planes"][~upd], se["planes"][~upd])   # the other colour is untouched
            same = np.all(so["planes"] == se["planes"], -1)
            assert same[upd].mean() > 0.6, (name, scale, red, same[upd].mean())
            assert frac_close(so["costs"][upd & same], se["costs"][upd & same], 5e-3) > 0.9
    o.finalize(); e.finalize()
    po, pe = o.result()[0], e.result()[0]
    same = np.all(so["planes"] == se["planes"], -1)
    assert np.isfinite(pe).all()


@pytest.mark.parametrize("name", CASES)
def test_geom_and_prior_stages(libs, name):
    c = make_case(name)
    o = libs.Oracle("cpu").set_problem(c["images"], c["cams"])
    e = libs.Oracle("emul").set_problem(c["images"], c["cams"])
    for x in (o, e):
        x.set_geom_consistency_params(True, False)
        x.set_src_depths(src_depths(c, 0.002))
        x.set_state(*world_state_from_gt(c))
        x.init_only(SEED + 2)
    so, se = o.get_state(), e.get_state()
    assert np.abs(so["planes"] - se["planes"]).max() < 1e-5         # (world n, depth) -> (cam n, distance): same arithmetic
    e.set_dev_state(so)
    for red in (0, 1):
        o.half_sweep(red, 0, 0); e.half_sweep(red, 0, 0)
        so, se = o.get_state(), e.get_state()
        np.testing.assert_array_equal(so["rng"], se["rng"])
        same = np.all(so["planes"] == se["planes"], -1)
        assert same.mean() > 0.6
        assert np.abs(so["geom"] - se["geom"])[same].mean() < 5e-3
        e.set_dev_state(so)
    # planar prior on top of a photometric state
    o = libs.Oracle("cpu").set_problem(c["images"], c["cams"])
    e = libs.Oracle("emul").set_problem(c["images"], c["cams"])
    for x in (o, e):
        x.set_geom_consistency_params(False, False)
        x.init_only(SEED)
    o.half_sweep(0, 0, 0); o.half_sweep(1, 0, 0); o.finalize()
    e.set_dev_state(o.get_state())
    for x in (o, e):
        x.set_planar_prior_params()
        x.set_geom_consistency_params(False, True)
        x.set_prior(*prior_planes(c))
        x.init_only(SEED + 1)
    so, se = o.get_state(), e.get_state()
    np.testing.assert_array_equal(so["rng"], se["rng"])
    assert np.abs(so["planes"] - se["planes"]).max() < 1e-4
    e.set_dev_state(so)
    for red in (0, 1):
        o.half_sweep(red, 0, 0); e.half_sweep(red, 0, 0)
        so, se = o.get_state(), e.get_state()
        np.testing.assert_array_equal(so["rng"], se["rng"])
        same = np.all(so["planes"] == se["planes"], -1)
        assert same.mean() > 0.75
        e.set_dev_state(so)


def test_uniform_stream_is_xorwow(libs):
    a = libs.Oracle("cpu").uniform_stream(SEED, 3, 4, 256)
    b = libs.Oracle("emul").uniform_stream(SEED, 3, 4, 256)
    np.testing.assert_array_equal(a, b)
    assert a.min() > 0 and a.max() <= 1.0
    c = libs.Oracle("cpu").uniform_stream(SEED, 4, 3, 256)
    assert not np.array_equal(a, c)       # (x, y) are hashed into the seed, not used as stream offsets
    assert abs(float(a.mean()) - 0.5) < 0.08


def test_restructurings_are_bit_identical(libs):
    """The kernel's tuning switches only reorder or skip work whose result cannot matter: early-out of refinement
    proposals (PM_EARLY_OUT) and warp-uniform view order (PM_UNIFORM_VIEWS).
    Whole runs must be bit-identical with every combination, in all three modes."""
    import ctypes as C

    from conftest import build_emul as be

    variants = {"_plain": ["-DPM_EARLY_OUT=0", "-DPM_UNIFORM_VIEWS=0"], "_eo": ["-DPM_EARLY_OUT=1", "-DPM_UNIFORM_VIEWS=1"],
                "_nouv": ["-DPM_EARLY_OUT=1", "-DPM_UNIFORM_VIEWS=0"]}
    c = make_case("room6")
    results = {}
    for tag, flags in variants.items():
        path = be(tag, flags)
        libs.LIBS["emul" + tag] = (path, "emu_")
        e = libs.Oracle("emul" + tag).set_problem(c["images"], c["cams"])
        e.set_geom_consistency_params(False, False)
        e.run(SEED)
        out = [e.result()]
        e.set_planar_prior_params()
        e.set_geom_consistency_params(False, True)
        e.set_prior(*prior_planes(c))
        e.run(SEED + 1)
        out.append(e.result())
        e.destroy()
        g = libs.Oracle("emul" + tag).set_problem(c["images"], c["cams"])
        g.set_geom_consistency_params(True, False)
        g.set_src_depths(src_depths(c, 0.002))
        g.set_state(*world_state_from_gt(c))
        g.run(SEED + 2)
        out.append(g.result(geom=True))
        g.destroy()
        results[tag] = out
    for tag in ("_eo", "_nouv"):
        for a, b in zip(results["_plain"], results[tag]):
            for x, y in zip(a, b):
                np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("name", CASES)
def test_exact_arithmetic_is_bit_identical_to_the_oracle(libs, name):
    """-DPM_EXACT=1 (pm_core.cuh, namespace pm_exact: one of the two arithmetics compiled into the library): homography, tap
    coordinates, bilateral weights, NCC sums and the geometric cost in the reference's own operation order, every rounding
    written down. Instantiated on the host (unfused arithmetic, as the plain-C oracle), the product's per-pixel templates must
    reproduce the literal restatement (oracle/pm_oracle.c) BIT FOR BIT over whole runs in all three modes -- every plane, cost,
    geometric cost and RNG draw. This pins the control flow and every piece of arithmetic outside the NCC of the fast
    arithmetic too (the two differ only inside `#if PM_EXACT`). What the GPU adds -- which products nvcc fuses -- is checked
    on the SASS (test_sass_equivalence.py) and on the B200 (test_zz_fidelity_build_gpu.py)."""
    from conftest import build_emul as be

    libs.LIBS["emul_literal"] = (be("_exact", ["-DPM_EXACT=1"]), "emu_")
    c = make_case(name)
    o = libs.Oracle("cpu").set_problem(c["images"], c["cams"])
    e = libs.Oracle("emul_literal").set_problem(c["images"], c["cams"])
    pl = random_planes(c)
    for s in (0, 1, 2):
        np.testing.assert_array_equal(o.ncc_map(pl, s), e.ncc_map(pl, s))
    for x in (o, e):
        x.set_geom_consistency_params(False, False)
        x.run(SEED)
    for a, b in zip(o.result(), e.result()):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(o.get_state()["rng"], e.get_state()["rng"])
    for x in (o, e):
        x.set_planar_prior_params()
        x.set_geom_consistency_params(False, True)
        x.set_prior(*prior_planes(c))
        x.run(SEED + 1)
    for a, b in zip(o.result(), e.result()):
        np.testing.assert_array_equal(a, b)
    o.destroy(); e.destroy()
    o = libs.Oracle("cpu").set_problem(c["images"], c["cams"])
    e = libs.Oracle("emul_literal").set_problem(c["images"], c["cams"])
    for x in (o, e):
        x.set_geom_consistency_params(True, False)
        x.set_src_depths(src_depths(c, 0.002))
        x.set_state(*world_state_from_gt(c))
    np.testing.assert_array_equal(o.geom_map(pl), e.geom_map(pl))
    for x in (o, e):
        x.run(SEED + 2)
    for a, b in zip(o.result(geom=True), e.result(geom=True)):
        np.testing.assert_array_equal(a, b)
    o.destroy(); e.destroy()
