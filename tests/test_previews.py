"""JPEG previews of the result maps (mp-mvs_b200/previews.py; reference src/utility.cpp:310-520)."""
import os

import numpy as np

from conftest import PKG

from mpmvs_b200 import io_formats, previews


def test_cost_and_normal_previews(tmp_path):
    import cv2

    cost = np.array([[0.0, 1.0, 2.0, 2.5], [0.5, 0.004, 1.996, -1.0]], np.float32)
    p = str(tmp_path / "costs.png")             # lossless container, same conversion
    assert previews.save_cost(cost, p)
    np.testing.assert_array_equal(cv2.imread(p, 0), [[0, 128, 255, 255], [64, 1, 254, 0]])    # x 127.5, round half to even, saturate
    assert not previews.save_cost(np.zeros((0, 0), np.float32), p)
    nrm = np.zeros((2, 2, 3), np.float32)
    nrm[..., 0], nrm[..., 1], nrm[..., 2] = 0.5, -0.25, 1.0
    q = str(tmp_path / "normals.png")
    assert previews.save_normal(nrm, q)
    got = cv2.imread(q, cv2.IMREAD_COLOR)
    assert tuple(got[0, 0]) == (128, 64, 255)   # x 255, y negated, saturated


def test_depth_preview_rules():
    import cv2

    rng = np.random.default_rng(0)
    d = rng.uniform(2.0, 4.0, (60, 80)).astype(np.float32)
    d[:5] = 0; d[5, :10] = -1                   # invalid pixels -> black
    d[10, 10] = 50.0                            # an outlier the 3 % clipping must absorb
    img = previews.depth_preview(d, hist_enhance=True)
    assert img.shape == (60, 80, 3) and img.dtype == np.uint8
    assert (img[:5] == 0).all() and (img[5, :10] == 0).all() and (img[6:] != 0).any()
    # with clipping the bulk of the range spreads over the colour map; without it the outlier squeezes it into the blue end
    flat = previews.depth_preview(d, hist_enhance=False)
    jet = cv2.applyColorMap(np.arange(256, dtype=np.uint8)[None], cv2.COLORMAP_JET)[0]
    assert len({tuple(c) for c in img[6:].reshape(-1, 3)}) > 5 * len({tuple(c) for c in flat[6:].reshape(-1, 3)})
    assert tuple(flat[10, 10]) == tuple(jet[255])
    top, down = previews._clip_levels(np.repeat(np.arange(256, dtype=np.uint8), 100))
    assert (top, down) == (7, 249)              # first level where > 3 % of the pixels lie below / above (bin 0 skipped)


def test_save_dmb_as_jpg_follows_the_yaml_keys(tmp_path):
    root = str(tmp_path)
    folder = io_formats.result_dir(root, 3)
    os.makedirs(folder)
    rng = np.random.default_rng(1)
    io_formats.write_dmb(os.path.join(folder, "depths.dmb"), rng.uniform(1, 2, (30, 40)).astype(np.float32))
    io_formats.write_dmb(os.path.join(folder, "costs.dmb"), rng.uniform(0, 2, (30, 40)).astype(np.float32))
    n = rng.normal(size=(30, 40, 3)).astype(np.float32)
    io_formats.write_dmb(os.path.join(folder, "normals.dmb"), n / np.linalg.norm(n, axis=-1, keepdims=True))
    cfg = dict(io_formats.DEFAULT_CONFIG, **{"Input-folder": root, "Output-folder": root})
    assert previews.save_dmb_as_jpg(cfg, [3]) == 0
    cfg.update({"Save Dmb as JPG": 1, "Save Cost Map": 1, "Save Normal Map": 1, "Save Prior Dmb as JPG": 1})
    assert previews.save_dmb_as_jpg(cfg, [3]) == 3          # no depths_prior.dmb was written: that preview is skipped
    assert sorted(f for f in os.listdir(folder) if f.endswith(".jpg")) == ["costs.jpg", "depths.jpg", "normals.jpg"]


def test_cli_loaders_follow_the_reference_rules(tmp_path):
    """run.py: PatchMatchInit's image rule (uint8 kept below `Max image size`, float resize above) and the sky files."""
    import importlib.util

    import cv2

    spec = importlib.util.spec_from_file_location("mpmvs_run", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mp-mvs_b200", "run.py"))
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 255, (60, 90), dtype=np.uint8)
    folder = str(tmp_path / "images")
    os.makedirs(folder)
    cv2.imwrite(os.path.join(folder, "00000007.pgm"), img)
    a, sx, sy = run.load_image(folder, 7, 3200)
    assert a.dtype == np.uint8 and (sx, sy) == (1.0, 1.0)
    np.testing.assert_array_equal(a, img)
    b, sx, sy = run.load_image(folder, 7, 45)                    # 90 x 60 -> 45 x 30 (PatchMatch.cpp:893-925)
    assert b.dtype == np.float32 and b.shape == (30, 45) and abs(sx - 0.5) < 1e-6 and abs(sy - 0.5) < 1e-6
    np.testing.assert_allclose(b, cv2.resize(img.astype(np.float32), (45, 30), interpolation=cv2.INTER_LINEAR))
    with np.testing.assert_raises(FileNotFoundError):
        run.load_image(folder, 8, 3200)
    assert run.sky_image_size(90, 60, 3200) == (90, 60) and run.sky_image_size(90, 60, 45) == (45, 30)
    # the sky gate RunFusion reads: skymask_refine.jpg brought to the depth map's size; absent file = no gate
    out = str(tmp_path)
    assert run.load_sky_mask(out, 7, (30, 45)) is None
    d = io_formats.result_dir(out, 7)
    os.makedirs(d)
    m = np.zeros((60, 90), np.uint8); m[:20] = 255
    cv2.imwrite(os.path.join(d, "skymask_refine.jpg"), m)
    got = run.load_sky_mask(out, 7, (30, 45))
    assert got.shape == (30, 45) and (got[:8] > 0).all() and (got[14:] > 0).mean() < 0.05
