"""Golden vectors of the two host stages next to the hot path, tests/golden/prior_stage.npz and
fusion.npz, made by tests/golden/make_prior_golden.py with OpenCV's own Subdiv2D and SVD (cv2 is the library the reference
links): the restatements in oracle/ must keep reproducing them, and the product -- the triangulation on the host,
the planar-prior and fusion kernels on the GPU -- is checked against the frozen outputs, not only against a live oracle."""
import os

import numpy as np
import pytest

from conftest import ROOT

from mpmvs_b200 import capi, io_formats

GOLD = os.path.join(ROOT, "tests", "golden")


def prior_cases():
    z = np.load(os.path.join(GOLD, "prior_stage.npz"))
    names = sorted({k.split("/")[0] for k in z.files if "/" in k})
    return {n: {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names}


@pytest.mark.parametrize("name", ["room_a", "room_geom"])
def test_prior_restatement_reproduces_the_golden_vectors(name):
    import prior_oracle

    c = prior_cases()[name]
    geom = c["geom"] if c["geom"].size else None
    prior, mask, verts, tris = prior_oracle.build_prior(c["planes"], c["costs"], c["K"], float(c["depth_range"][0]), float(c["depth_range"][1]), geom)
    np.testing.assert_array_equal(verts, c["vertices"])
    np.testing.assert_array_equal(tris, c["triangles"])                 # cv2.Subdiv2D: the reference's triangulator
    np.testing.assert_array_equal(mask, c["mask"])
    np.testing.assert_allclose(prior, c["prior"], atol=1e-6)            # cv2.SVDecomp (cv::SVD::solveZ), float32


@pytest.mark.parametrize("name", ["room_a", "room_geom"])
def test_host_triangulation_reproduces_the_golden_triangle_list(name):
    """mpmvs_delaunay (pm_subdiv.h, host only) on the golden vertices: OpenCV's triangle list, order included."""
    c = prior_cases()[name]
    h, w = c["costs"].shape
    verts = c["vertices"]
    tris = capi.delaunay(verts, w, h)
    np.testing.assert_array_equal(np.asarray(verts)[tris], c["triangles"])


def test_fusion_restatement_reproduces_the_golden_vectors():
    import fusion_oracle

    z = np.load(os.path.join(GOLD, "fusion.npz"))
    n = len(z["lists"])
    depths, normals, bgr = [z[f"depth{i}"] for i in range(n)], [z[f"normal{i}"] for i in range(n)], [z[f"bgr{i}"] for i in range(n)]
    for dyn in (0, 1):
        got = fusion_oracle.fuse(z["cams"], depths, normals, bgr, [list(r) for r in z["lists"]], dynamic=bool(dyn))
        np.testing.assert_allclose(got, z[f"points_dyn{dyn}"], rtol=0, atol=1e-6)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["room_a", "room_geom"])
def test_gpu_prior_stage_against_the_golden_vectors(name):
    """Vertex picking (bit-exact), then rasterisation + plane fit + range check on the GOLDEN triangulation (so the comparison
    isolates them): triangle-id mask exact up to range-check ties, planes to 2e-3 (closed-form plane through three points
    against OpenCV's float32 SVD)."""
    from conftest import PKG, problem_arrays

    c = prior_cases()[name]
    h, w = c["costs"].shape
    sc = PKG.synth.make_eth3d_scene(width=w, height=h, n_views=3, n_src=2, jpeg=False, seed=3)
    ids, imgs, cams = problem_arrays(sc, 1)
    pm = capi.PatchMatch(0).set_problem(imgs, cams)
    with_geom = c["geom"].size > 0
    pm.set_geom_consistency_params(with_geom, with_geom)
    st = {"planes": c["planes"], "costs": c["costs"], "views": None, "rng": None, "geom": c["geom"] if with_geom else None}
    pm.set_dev_state(st)
    verts = pm.pick_vertices(with_geom)
    np.testing.assert_array_equal(verts, c["vertices"])
    lut = {(int(x), int(y)): i for i, (x, y) in enumerate(verts)}
    tris_idx = np.array([[lut[(int(x), int(y))] for x, y in t] for t in c["triangles"]], np.int32)
    pm.prior_from_triangles(verts, tris_idx)
    prior, mask = pm.get_prior()
    # the product's range check uses the handle's depth range (0.6 dmin .. 1.2 dmax of the camera), the golden one [1, 6]: compare
    # where both kept the pixel
    both = (mask > 0) & (c["mask"] > 0)
    assert both.mean() > 0.5
    assert (mask[both] == c["mask"][both]).mean() > 0.999
    same = both & (mask == c["mask"])
    assert np.abs(prior[same] - c["prior"][same]).max() < 2e-3
    pm.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("dyn", [0, 1])
def test_gpu_fusion_against_the_golden_vectors(dyn):
    z = np.load(os.path.join(GOLD, "fusion.npz"))
    n = len(z["lists"])
    f = capi.Fusion(0, n)
    for i in range(n):
        bgr = z[f"bgr{i}"]
        f.set_view(i, z["cams"][i:i + 1], z[f"depth{i}"], z[f"normal{i}"], np.ascontiguousarray(bgr[..., 0]))
        f.set_color(i, bgr)
    got, _ = f.run([list(r) for r in z["lists"]], bool(dyn))
    f.destroy()
    want = z[f"points_dyn{dyn}"]
    assert abs(len(got) - len(want)) <= max(3, 0.003 * len(want)), (len(got), len(want))
    key = lambda p: set(map(tuple, np.round(np.concatenate([p[:, :3] * 1e4, p[:, 6:9]], 1)).astype(np.int64)))  # noqa: E731
    a, b = key(got), key(want)
    assert len(a & b) > 0.98 * len(b), (len(a & b), len(b))
