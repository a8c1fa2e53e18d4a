"""Bit-identity of the exact arithmetic to the reference's kernels across problem SHAPES the three parity cases do not
cover: odd image sizes (partial blocks on both axes), 1, 2 and 12 source views, views of different sizes (the soft-clamp
kernel variant: the layered texture is as large as the largest view, clamp-to-edge is done on the coordinates), in all three
modes -- whole same-seed Run()s against oracle/_ref."""
import numpy as np
import pytest

from conftest import PKG, problem_arrays

pytestmark = pytest.mark.gpu


def build_case(kind):
    synth = PKG.synth
    if kind == "odd_size":
        sc = synth.make_dtu_scene(width=101, height=77, grid=3, n_src=4, seed=2, jpeg=False)
        ref = 4
    elif kind == "one_source":
        sc = synth.make_plane_scene(width=70, height=50, n_views=2, seed=1, jpeg=False, n_src=1)
        ref = 0
    elif kind == "two_sources":
        sc = synth.make_plane_scene(width=64, height=48, n_views=3, seed=1, jpeg=False)
        ref = 1
    elif kind == "twelve_sources":
        sc = synth.make_dtu_scene(width=96, height=72, grid=4, n_src=12, seed=2, jpeg=False)
        ref = 5
    elif kind == "mixed_sizes":
        sc = synth.make_dtu_scene(width=96, height=72, grid=3, n_src=4, seed=2, jpeg=False)
        ref = 4
    else:
        raise KeyError(kind)
    ids, imgs, cams = problem_arrays(sc, ref)
    gt = {i: sc.gt_depth[i] for i in ids}
    if kind == "mixed_sizes":          # crop two sources at the right / bottom: K stays valid, the sizes differ
        cams = cams.copy()
        for k, (dw, dh) in ((1, (7, 0)), (3, (0, 5))):
            h, w = imgs[k].shape
            imgs[k] = np.ascontiguousarray(imgs[k][:h - dh, :w - dw])
            gt[ids[k]] = np.ascontiguousarray(gt[ids[k]][:h - dh, :w - dw])
            cams[k]["width"], cams[k]["height"] = w - dw, h - dh
    return dict(scene=sc, ref=ref, ids=ids, images=imgs, cams=cams, gt=gt)


@pytest.mark.parametrize("kind", ["odd_size", "one_source", "two_sources", "twelve_sources", "mixed_sizes"])
def test_whole_runs_are_bit_identical_in_all_modes(kind):
    import oracle_py
    from cases import SEED, prior_planes, world_state_from_gt

    from mpmvs_b200 import capi

    if not oracle_py.available("ref"):
        pytest.skip("oracle/_ref/libmpmvs_ref.so not built on this box")
    c = build_case(kind)
    rng = np.random.default_rng(5)
    depths = [(c["gt"][i] * rng.uniform(0.998, 1.002, c["gt"][i].shape)).astype(np.float32) for i in c["ids"][1:]]
    pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
    ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
    assert pm.arithmetic == "exact"
    # photometric, then the planar-prior run on top of it (state resident)
    for o in (pm, ref):
        o.set_geom_consistency_params(False, False)
        o.run(SEED)
    (pa, ca), (pb, cb) = pm.result(), ref.result()
    np.testing.assert_array_equal(pa, pb)
    np.testing.assert_array_equal(ca, cb)
    for o in (pm, ref):
        o.set_planar_prior_params()
        o.set_geom_consistency_params(False, True)
        o.set_prior(*prior_planes(c))
        o.run(SEED + 1)
    (pa, ca), (pb, cb) = pm.result(), ref.result()
    np.testing.assert_array_equal(pa, pb)
    np.testing.assert_array_equal(ca, cb)
    pm.destroy(); ref.destroy()
    # geometric consistency
    pm = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
    ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
    for o in (pm, ref):
        o.set_geom_consistency_params(True, False)
        o.set_src_depths(depths)
        o.set_state(*world_state_from_gt(c))
        o.run(SEED + 2)
    ga, gb = pm.result(geom=True), ref.result(geom=True)
    for a, b in zip(ga, gb):
        np.testing.assert_array_equal(a, b)
    pm.destroy(); ref.destroy()
