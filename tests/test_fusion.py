"""Depth-map fusion on the GPU (mp-mvs_b200/csrc/pm_fusion.cu) against the numpy restatement with the kernel's order
(oracle/fusion_oracle.py) point for point, and against the reference's sequential order (RunFusion restated in
mp-mvs_b200/csrc/mpmvs_main.cpp) statistically."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT

from mpmvs_b200 import io_formats


def noisy_scene(width=160, height=120, seed=0):
    """GT depth/normal maps with small noise and holes, as a converged PatchMatch run would leave them."""
    rng = np.random.default_rng(seed)
    sc = PKG.synth.make_dtu_scene(width=width, height=height, grid=3, n_src=4, seed=2, jpeg=False)
    depths, normals = [], []
    for i in range(sc.num_views):
        d = sc.gt_depth[i].astype(np.float32) * rng.uniform(0.997, 1.003, sc.gt_depth[i].shape).astype(np.float32)
        d[rng.random(d.shape) < 0.03] *= 1.2          # outliers that must fail the consistency test
        nrm = sc.gt_normal[i].astype(np.float32) + rng.normal(0, 0.03, sc.gt_normal[i].shape).astype(np.float32)
        nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-6)
        depths.append(d)
        normals.append(nrm.astype(np.float32))
    lists = [[i] + [j for j, _ in sc.pairs[i]] for i in range(sc.num_views)]
    return sc, depths, normals, lists


def test_fusion_oracle_basic_properties():
    import fusion_oracle

    sc, depths, normals, lists = noisy_scene(96, 72)
    cams = io_formats.pack_cameras(sc.cams)
    pts = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, dynamic=True)
    total = sum(int((d > 0).sum()) for d in depths)
    assert 0.1 * total < len(pts) < total            # every surface point is fused once, not once per view
    assert np.isfinite(pts).all()
    assert np.abs(np.linalg.norm(pts[:, 3:6], axis=1) - 1).mean() < 0.05
    # fused points lie on the scene: the table plane z = 0 or box faces -> z within the scene's extent
    assert np.quantile(pts[:, 2], 0.01) > -0.05 and np.quantile(pts[:, 2], 0.99) < 0.6   # a few chance-consistent outliers remain
    pts2 = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, dynamic=False)
    assert len(pts2) > 0 and len(pts2) != len(pts)


@pytest.mark.gpu
@pytest.mark.parametrize("dynamic", [True, False])
def test_gpu_fusion_matches_restatement(dynamic):
    import fusion_oracle

    from mpmvs_b200 import capi

    sc, depths, normals, lists = noisy_scene()
    cams = io_formats.pack_cameras(sc.cams)
    want = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, dynamic=dynamic)
    f = capi.Fusion(0, sc.num_views)
    for i in range(sc.num_views):
        f.set_view(i, cams[i:i + 1], depths[i], normals[i], sc.images[i])
    got, ms = f.run(lists, dynamic)
    f.destroy()
    print(f"fusion dynamic={dynamic}: {len(got)} points on the GPU in {ms:.2f} ms, {len(want)} in the restatement")
    # float rounding (fast-math division, acosf) moves a few borderline pixels across the thresholds
    assert abs(len(got) - len(want)) <= 0.003 * len(want)
    # same points (to 0.1 mm and 1e-3 in the normal) up to those borderline pixels; rows cannot be compared one to one
    # because a pixel that flips shifts every row after it
    key = lambda p: set(map(tuple, np.round(np.concatenate([p[:, :3] * 1e4, p[:, 3:6] * 1e3], 1)).astype(np.int64)))  # noqa: E731
    a, b = key(got), key(want)
    assert len(a & b) > 0.99 * len(b), (len(a & b), len(b))


@pytest.mark.gpu
def test_gpu_fusion_averages_the_colour_image():
    """RunFusion reads the images with IMREAD_COLOR and averages B, G, R per point (PatchMatch.cpp:322,399,443-445):
    mpmvs_fusion_set_color gives the GPU fusion the colour image; without it the grey level fills all three channels."""
    import fusion_oracle

    from mpmvs_b200 import capi

    sc, depths, normals, lists = noisy_scene()
    cams = io_formats.pack_cameras(sc.cams)
    bgr = [np.stack([g, 255 - g, g // 2], -1).astype(np.uint8) for g in sc.images]
    want = fusion_oracle.fuse(cams, depths, normals, bgr, lists, dynamic=True)
    f = capi.Fusion(0, sc.num_views)
    for i in range(sc.num_views):
        f.set_view(i, cams[i:i + 1], depths[i], normals[i], sc.images[i])
        f.set_color(i, bgr[i])
    got, _ = f.run(lists, True)
    f.destroy()
    assert abs(len(got) - len(want)) <= 0.003 * len(want)
    # colour statistics: the three channels differ from one another and agree with the restatement's
    assert np.abs(got[:, 6:9].mean(0) - want[:, 6:9].mean(0)).max() < 0.5
    assert abs(got[:, 6].mean() - got[:, 7].mean()) > 5.0
    key = lambda p: set(map(tuple, np.round(np.concatenate([p[:, :3] * 1e4, p[:, 6:9]], 1)).astype(np.int64)))  # noqa: E731
    a, b = key(got), key(want)
    assert len(a & b) > 0.98 * len(b), (len(a & b), len(b))


@pytest.mark.gpu
def test_gpu_fusion_vs_sequential_host(tmp_path):
    """Against the reference's order (C++ RunFusion in mpmvs_main, run on the PatchMatch results it produced itself)."""
    from test_cpp_host import MAIN, build_main, write_scene

    from mpmvs_b200 import capi

    build_main()
    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 1, "Planer prior": 0, "Geometric consistency planer prior": 0})
    r = subprocess.run([MAIN, yaml, "--seed", "5", "--tex", "u8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:]
    with open(os.path.join(root, "MPMVS", "MPMVS_model.ply"), "rb") as fh:
        head = fh.read(400).decode("latin1")
    n_host = int(head.split("element vertex ")[1].split("\n")[0])
    cams = io_formats.pack_cameras(sc.cams)
    f = capi.Fusion(0, sc.num_views)
    lists = []
    for i in range(sc.num_views):
        d = io_formats.read_dmb(os.path.join(root, "MPMVS", f"2333_{i:08d}", "depths.dmb"))
        n = io_formats.read_dmb(os.path.join(root, "MPMVS", f"2333_{i:08d}", "normals.dmb"))
        f.set_view(i, cams[i:i + 1], d, n, sc.images[i])
        lists.append([i] + [j for j, _ in sc.pairs[i]])
    got, ms = f.run(lists, True)
    f.destroy()
    print(f"GPU fusion {len(got)} points in {ms:.2f} ms; sequential host order {n_host} points")
    assert abs(len(got) - n_host) < 0.05 * n_host      # the orders differ only in which duplicate survives


def test_host_fusion_matches_restatement_without_gpu(tmp_path):
    """RunFusion restated in the C++ host (`mpmvs_main --fusion-only`: the reference's sequential order, no GPU involved)
    against the numpy restatement (per-image snapshot order): same surface, point counts within the few percent by which the
    two orders differ in which duplicate survives; the sky gate removes the gated pixels."""
    from test_cpp_host import MAIN, build_main

    import fusion_oracle

    build_main()
    sc, depths, normals, lists = noisy_scene(96, 72)
    root = str(tmp_path / "dense")
    PKG.synth.write_dense_folder(sc, root)                      # images/ (jpg + pgm), cams/, pair.txt
    for i in range(sc.num_views):
        d = io_formats.result_dir(root, i)
        os.makedirs(d, exist_ok=True)
        io_formats.write_dmb(os.path.join(d, "depths.dmb"), depths[i])
        io_formats.write_dmb(os.path.join(d, "normals.dmb"), normals[i])
    cams = io_formats.pack_cameras(sc.cams)

    def run(**cfg):
        yaml = str(tmp_path / "config.yaml")
        io_formats.write_config(yaml, **dict({"Input-folder": root, "Output-folder": root, "Max source images num": 4}, **cfg))
        r = subprocess.run([MAIN, yaml, "--fusion-only"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        n = int(r.stdout.split("store 3D points to ply file: ")[1].split(" points")[0])
        with open(os.path.join(root, "MPMVS", "MPMVS_model.ply"), "rb") as fh:
            raw = fh.read()
        body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
        assert len(body) == n * 27                              # 6 floats + 3 colour bytes per vertex (PatchMatch.cpp:145-198)
        pts = np.frombuffer(body, dtype=np.dtype([("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)]))
        return n, pts

    for dyn in (1, 0):
        n, pts = run(**{"Use dynamic_consistency to fuse": dyn})
        want = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, dynamic=bool(dyn))
        assert abs(n - len(want)) < 0.08 * len(want), (n, len(want))      # measured 0.3 % (dynamic) and 5.8 % (fixed threshold)
        key = lambda p: set(map(tuple, np.round(p * 1e3).astype(np.int64)))   # noqa: E731  (1 mm cells)
        a, b = key(pts["p"]), key(want[:, :3])
        assert len(a & b) > 0.75 * min(len(a), len(b))     # measured 0.83: the orders differ in which pixels of a surface patch get fused
        assert np.isfinite(pts["p"]).all() and abs(float(np.linalg.norm(pts["n"], axis=1).mean()) - 1) < 0.05
    # sky gate through the files the reference reads (skymask_refine.*): the top rows of every image are sky
    n_open = n
    for i in range(sc.num_views):
        m = np.zeros(depths[i].shape, np.uint8)
        m[:30] = 255
        PKG.synth.write_pgm(os.path.join(io_formats.result_dir(root, i), "skymask_refine.pgm"), m)
    n_gated, _ = run(**{"Use dynamic_consistency to fuse": 0, "Sky segment": 1})
    sky = [np.where(np.arange(d.shape[0])[:, None] < 30, 255, 0).astype(np.uint8) * np.ones(d.shape, np.uint8) for d in depths]
    want = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, dynamic=False, sky=sky)
    assert n_gated < n_open and abs(n_gated - len(want)) < 0.08 * len(want), (n_gated, n_open, len(want))
