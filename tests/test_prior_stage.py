"""Planar-prior host stage (ProcessProblem, /root/reference/src/PatchMatch.cpp:532-609).

CPU part: the product's triangulation (pm_subdiv.h through the C ABI, pure host code) against cv2.Subdiv2D -- the
reference's own triangulator: identical triangle lists, order included. GPU part: vertex picking, rasterisation, plane
fit and range check (pm_prior.cu) against the numpy/OpenCV restatement in oracle/prior_oracle.py, which itself reproduces
the prior the reference's own ProcessProblem builds bit for bit (tests/test_reference_program.py).
"""
import numpy as np
import pytest

from conftest import PKG, gt_planes_cam, problem_arrays

from mpmvs_b200 import capi


def cv_triangle_list(pts, w, h):
    """cv::Subdiv2D as the reference drives it (PatchMatch.cpp:763-771): one insert per vertex, getTriangleList."""
    import cv2

    sd = cv2.Subdiv2D((0, 0, w, h))
    for x, y in pts:
        sd.insert((float(x), float(y)))
    t = sd.getTriangleList()
    return np.zeros((0, 6), np.float32) if t is None else np.asarray(t, np.float32).reshape(-1, 6)


def our_triangle_list(pts, w, h):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    tris = capi.delaunay(pts, w, h)
    return pts[tris].reshape(len(tris), 6).astype(np.float32)


def cell_points(rng, w, h, keep=0.8, per_cell=1):
    pts = []
    for r in range(0, h, 5):
        for c in range(0, w, 5):
            if rng.random() < keep:
                for q in rng.permutation(25)[: rng.integers(1, per_cell + 1)]:
                    pts.append((min(w - 1, c + q % 5), min(h - 1, r + q // 5)))
    return np.array(pts, np.int32)


@pytest.mark.parametrize("kind", ["random", "cells", "three_per_cell", "grid", "collinear", "on_edge", "duplicates", "tiny"])
def test_triangle_list_is_opencvs(kind):
    """mpmvs_delaunay (pm_subdiv.h, pure host code) returns cv::Subdiv2D's triangle list: the same triangles, the same corner
    order, the same LIST order (the rasterisation that follows lets later triangles overwrite earlier ones, so a fifth of
    the prior depends on it) -- on generic sets, on the vertex picker's one / up-to-three points per 5x5 cell in scan order,
    and on the degenerate inputs integer pixels produce all the time: co-circular grids, collinear runs, points that fall on
    an existing edge, repeated points."""
    rng = np.random.default_rng(7)
    if kind == "random":
        w, h = 320, 240
        pts = np.unique(np.stack([rng.integers(0, w, 2500), rng.integers(0, h, 2500)], 1), axis=0).astype(np.int32)
        rng.shuffle(pts)
    elif kind == "cells":
        w, h = 403, 271
        pts = cell_points(rng, w, h)
    elif kind == "three_per_cell":
        w, h = 320, 240
        pts = cell_points(rng, w, h, keep=0.9, per_cell=3)
    elif kind == "grid":
        w, h = 150, 100
        pts = np.array([(c, r) for r in range(0, 100, 5) for c in range(0, 150, 5)], np.int32)
    elif kind == "collinear":
        w, h = 100, 80
        pts = np.array([(c, 10) for c in range(5, 95, 3)] + [(50, r) for r in range(12, 78, 3)] + [(7, 70), (93, 71)] + [(c, 20) for c in range(1, 100, 7)], np.int32)
    elif kind == "on_edge":
        w, h = 64, 48
        pts = np.array([(5, 5), (50, 5), (27, 5), (27, 30), (27, 17), (10, 17), (38, 17), (27, 11)], np.int32)
    elif kind == "duplicates":
        w, h = 64, 48
        pts = np.array([(5, 5), (50, 5), (5, 5), (27, 30), (50, 5), (10, 17), (27, 30)], np.int32)
    else:
        w, h = 20, 20
        pts = np.array([(3, 3), (15, 4), (8, 16)], np.int32)
    ours, want = our_triangle_list(pts, w, h), cv_triangle_list(pts, w, h)
    assert len(want) > 0
    np.testing.assert_array_equal(ours, want)


def test_triangle_list_fuzz_against_opencv():
    """400 small random sets, three quarters of them degenerate on purpose (lattices: collinear, co-circular and repeated points;
    almost everything on one line; a tiny dense cluster): the list is OpenCV's every time (3 000 such cases were run once)."""
    rng = np.random.default_rng(12345)
    for it in range(400):
        w, h, n = int(rng.integers(8, 60)), int(rng.integers(8, 60)), int(rng.integers(3, 80))
        kind = it % 4
        if kind == 0:
            pts = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1)
        elif kind == 1:
            pts = np.stack([rng.integers(0, w, n) // 3 * 3, rng.integers(0, h, n) // 3 * 3], 1)
        elif kind == 2:
            pts = np.stack([rng.integers(0, w, n), np.full(n, rng.integers(0, h))], 1)
            pts[::5, 1] = rng.integers(0, h, len(pts[::5]))
        else:
            pts = np.stack([rng.integers(0, min(w, 6), n), rng.integers(0, min(h, 6), n)], 1)
        np.testing.assert_array_equal(our_triangle_list(pts, w, h), cv_triangle_list(pts, w, h), err_msg=f"case {it}")


def test_delaunay_edge_cases():
    assert len(capi.delaunay(np.zeros((0, 2), np.int32), 10, 10)) == 0
    assert len(capi.delaunay(np.array([(1, 1), (5, 5)], np.int32), 10, 10)) == 0
    dup = np.array([(1, 1), (8, 2), (4, 8), (8, 2), (1, 1)], np.int32)        # a repeated point is the vertex it repeats
    assert len(capi.delaunay(dup, 10, 10)) == 1
    with pytest.raises(capi.MpmvsError):
        capi.delaunay(np.array([(1, 1), (50, 2), (4, 8)], np.int32), 10, 10)  # outside the image rectangle


def test_triangle_list_full_size():
    """The metric's size: one vertex per 5x5 cell of a 3200x2130 image (272 640 points, 545 k triangles), identical to OpenCV's
    list and well under a second of host time (it overlaps the other images' kernels)."""
    import time

    rng = np.random.default_rng(1)
    pts = cell_points(rng, 3200, 2130, keep=1.0)
    t = time.time()
    ours = our_triangle_list(pts, 3200, 2130)
    dt = time.time() - t
    assert abs(len(ours) - 2 * len(pts)) < 0.01 * len(pts)    # Euler: ~2n triangles
    assert dt < 5.0, dt
    np.testing.assert_array_equal(ours, cv_triangle_list(pts, 3200, 2130))


def test_prior_oracle_matches_opencv_semantics():
    """The restated rasterisation covers exactly the closed triangle it samples, later triangles win shared pixels."""
    import prior_oracle

    tris = np.array([[(2, 2), (12, 3), (5, 11)], [(12, 3), (5, 11), (14, 13)]], np.int32)
    m = prior_oracle.rasterise(tris, 20, 20)
    assert m[2, 2] == 1 and m[13, 14] == 2 and m[0, 0] == 0
    assert m[3, 12] == 2 and m[11, 5] == 2       # shared vertices belong to the later triangle
    assert (m == 1).sum() > 20 and (m == 2).sum() > 20


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("geom_variant", [False, True])
def test_gpu_prior_stage_vs_restatement(geom_variant):
    import prior_oracle

    sc = PKG.synth.make_eth3d_scene(width=160, height=107, n_views=5, n_src=4, jpeg=False)
    ids, imgs, cams = problem_arrays(sc, 2)
    pm = capi.PatchMatch(0).set_problem(imgs, cams)
    if geom_variant:
        pm.set_geom_consistency_params(True, True)
        pm.set_src_depths([sc.gt_depth[i] for i in ids[1:]])
        planes0 = gt_planes_cam(sc, 2).copy()
        planes0[..., :3] = sc.gt_normal[2]
        planes0[..., 3] = np.where(sc.gt_depth[2] > 0, sc.gt_depth[2], 3.0) * np.random.default_rng(3).uniform(0.97, 1.03, sc.gt_depth[2].shape)
        pm.set_state(planes0, np.random.default_rng(4).uniform(0, 0.6, sc.gt_depth[2].shape).astype(np.float32))
        pm.run(3)
        planes, costs, geom = pm.result(geom=True)
    else:
        pm.set_geom_consistency_params(False, False)
        pm.run(3)
        (planes, costs), geom = pm.result(), None
    verts = pm.pick_vertices(geom_variant)
    want_verts = prior_oracle.triangulate_vertices(costs, geom)
    np.testing.assert_array_equal(verts, want_verts)                       # integer work: bit-exact, same order
    assert len(verts) > 50
    # same triangulation into both (OpenCV's), so the comparison isolates rasterisation + plane fit + range check
    K = sc.cams[2].K
    dmin, dmax = pm.depth_range
    prior_w, mask_w, _, tris_px = prior_oracle.build_prior(planes, costs, K, dmin, dmax, geom)
    lut = {(int(x), int(y)): i for i, (x, y) in enumerate(verts)}
    tris_idx = np.array([[lut[(int(x), int(y))] for x, y in t] for t in tris_px], np.int32)
    bad = tris_idx.copy()
    bad[0, 0] = len(verts)                                                 # one past the last vertex: must be refused on the host
    with pytest.raises(capi.MpmvsError):
        pm.prior_from_triangles(verts, bad)
    n = pm.prior_from_triangles(verts, tris_idx)
    prior, mask = pm.get_prior()
    assert n == int((mask > 0).sum())
    assert (mask == mask_w).mean() > 0.999                                 # id mask: exact except range-check ties
    both = (mask > 0) & (mask == mask_w)
    assert np.abs(prior[both] - prior_w[both]).max() < 2e-3                # closed-form plane vs float32 SVD
    # full stage with the product's own triangulation: OpenCV's triangle list, so the same triangle ids per pixel
    st = pm.build_prior()
    prior2, mask2 = pm.get_prior()
    assert st["n_vertices"] == len(verts) and st["n_triangles"] == len(tris_px), (st, len(tris_px))
    same = (mask2 == mask_w).mean()
    print("prior stage:", st, "pixels with the restatement's triangle id", same)
    assert same > 0.999                                                    # exact except range-check ties
    both2 = (mask2 > 0) & (mask2 == mask_w)
    assert np.abs(prior2[both2] - prior_w[both2]).max() < 2e-3
    # and the prior run goes through
    pm.set_planar_prior_params()
    pm.set_geom_consistency_params(False, True)
    pm.run(4)
    p2, c2 = pm.result()
    assert np.isfinite(p2).all()
    acc = PKG.synth.accuracy_at(p2[..., 3], sc.gt_depth[2])
    acc0 = PKG.synth.accuracy_at(planes[..., 3], sc.gt_depth[2])
    print("accuracy before/after the prior run", acc0, acc)
    pm.destroy()


def test_c_restatement_matches_python_restatement():
    """oracle/pm_oracle.c's host prior stage (used at full size by bench.py's reference arm) against the line-by-line
    numpy restatement, on a rendered scene with a synthetic converged state."""
    import prior_oracle

    sc = PKG.synth.make_eth3d_scene(width=120, height=81, n_views=3, n_src=2, jpeg=False)
    rng = np.random.default_rng(5)
    gt = sc.gt_depth[1]
    planes = np.concatenate([sc.gt_normal[1], (np.where(gt > 0, gt, 3.0) * rng.uniform(0.99, 1.01, gt.shape))[..., None]], -1).astype(np.float32)
    costs = rng.uniform(0.0, 0.3, gt.shape).astype(np.float32)
    geom = rng.uniform(0.0, 0.8, gt.shape).astype(np.float32)
    K = sc.cams[1].K
    for g in (None, geom):
        pw, mw, vw, tw = prior_oracle.build_prior(planes, costs, K, 1.0, 6.0, g)
        pf, mf, vf, tf, cnt = prior_oracle.build_prior_fast(planes, costs, K, 1.0, 6.0, g)
        np.testing.assert_array_equal(vw, vf)
        np.testing.assert_array_equal(tw, tf)
        assert (mw == mf).mean() > 0.9995 and cnt == int((mf > 0).sum())
        both = (mw > 0) & (mw == mf)
        assert np.abs(pw[both] - pf[both]).max() < 2e-3
