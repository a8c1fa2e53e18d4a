"""Golden vectors for the planar-prior host stage and for depth-map fusion (row 3).

The reference's host program cannot be built here (it needs OpenCV's C++ headers and libraries; only the Python wheel is
in the image). What CAN be pinned is its third-party arithmetic: cv2 is the same library (cv::Subdiv2D for the
triangulation, PatchMatch.cpp:766-771; cv::SVD::solveZ for the plane fit, :743), so the restatement in
oracle/prior_oracle.py is DRIVEN by it and its outputs are frozen here. The loops around it (vertex picking :782-853,
rasterisation :554-579, range check :583-595, RunFusion :287-504) are restated line by line with the source's arithmetic
types; for them the golden files pin the restatement against drift, not against reference outputs -- DESIGN.md says so.

    python tests/golden/make_prior_golden.py          # CPU only (cv2 4.13, numpy); writes tests/golden/prior_stage.npz, fusion.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

from conftest import PKG  # noqa: E402
import fusion_oracle  # noqa: E402
import prior_oracle  # noqa: E402


def prior_case(seed, w, h, with_geom):
    """A converged-looking state on a rendered weak-texture room: GT planes slightly off, costs low on walls, high at edges."""
    sc = PKG.synth.make_eth3d_scene(width=w, height=h, n_views=3, n_src=2, jpeg=False, seed=3)
    rng = np.random.default_rng(seed)
    gt = sc.gt_depth[1]
    planes = np.concatenate([sc.gt_normal[1], (np.where(gt > 0, gt, 3.0) * rng.uniform(0.99, 1.01, gt.shape))[..., None]], -1).astype(np.float32)
    costs = rng.uniform(0.0, 0.3, gt.shape).astype(np.float32)
    geom = rng.uniform(0.0, 0.8, gt.shape).astype(np.float32) if with_geom else None
    K = sc.cams[1].K.astype(np.float32)
    prior, mask, verts, tris = prior_oracle.build_prior(planes, costs, K, 1.0, 6.0, geom)
    return dict(planes=planes, costs=costs, geom=np.zeros((0,), np.float32) if geom is None else geom, K=K, depth_range=np.array([1.0, 6.0], np.float32),
                vertices=verts, triangles=tris, mask=mask, prior=prior)


def fusion_case():
    sc = PKG.synth.make_dtu_scene(width=80, height=60, grid=2, n_src=3, seed=2, jpeg=False)
    rng = np.random.default_rng(11)
    depths = [(d * rng.uniform(0.998, 1.002, d.shape)).astype(np.float32) for d in sc.gt_depth]
    normals = [n.astype(np.float32) for n in sc.gt_normal]
    lists = [[i] + [j for j, _ in sc.pairs[i]] for i in range(sc.num_views)]
    cams = PKG.io_formats.pack_cameras(sc.cams)
    bgr = [np.stack([g, 255 - g, g // 2], -1).astype(np.uint8) for g in sc.images]
    out = {"cams": cams, "lists": np.array(lists, np.int32)}
    for i in range(sc.num_views):
        out[f"depth{i}"], out[f"normal{i}"], out[f"bgr{i}"] = depths[i], normals[i], bgr[i]
    for dyn in (0, 1):
        out[f"points_dyn{dyn}"] = fusion_oracle.fuse(cams, depths, normals, bgr, lists, dynamic=bool(dyn))
    return out


if __name__ == "__main__":
    import cv2

    cases = {"room_a": prior_case(5, 120, 81, False), "room_geom": prior_case(6, 100, 67, True)}
    flat = {f"{name}/{k}": v for name, c in cases.items() for k, v in c.items()}
    flat["opencv_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "prior_stage.npz"), **flat)
    np.savez_compressed(os.path.join(HERE, "fusion.npz"), **fusion_case(), opencv_version=np.array(cv2.__version__))
    for f in ("prior_stage.npz", "fusion.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KB")
