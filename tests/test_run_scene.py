"""Host logic of mp-mvs_b200/run.py (the multi-GPU entry point over a dense folder) that needs no GPU."""
import importlib.util
import os

import numpy as np

from conftest import PKG, ROOT


def load_run():
    spec = importlib.util.spec_from_file_location("mpmvs_run", os.path.join(ROOT, "mp-mvs_b200", "run.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_every_camera_is_scaled_to_the_working_size_on_every_rank(tmp_path):
    """`Max image size` below the image size: PatchMatchInit resizes the views and scales K (PatchMatch.cpp:893-925). Rank 0
    fuses ALL depth maps, also those of images other ranks estimated and it never decoded -- their cameras must carry the same
    scaled intrinsics and the working size (a rank used to scale K only for the views of its own shard)."""
    run = load_run()
    scene = PKG.synth.make_dtu_scene(width=160, height=120, grid=3, n_src=2, seed=2, jpeg=True)
    root = str(tmp_path / "dense")
    PKG.synth.write_dense_folder(scene, root)
    cfg = {"Input-folder": root, "Output-folder": root, "Max source images num": 2, "Max image size": 100}
    _, cams1, images1 = run.load_scene(cfg, 0, 1)
    assert all(im.shape == (75, 100) for im in images1.values())
    seen_undecoded = False
    for world in (2, 4):
        for rank in range(world):
            _, cams, images = run.load_scene(cfg, rank, world)
            assert set(cams) == set(cams1)
            assert set(images) <= set(images1)                    # a rank decodes its shard's views only
            undecoded = set(cams) - set(images)
            seen_undecoded = seen_undecoded or bool(undecoded)
            for i, c in cams.items():
                assert (c.height, c.width) == (75, 100), (world, rank, i)
                np.testing.assert_array_equal(c.K, cams1[i].K)
    assert seen_undecoded          # the case the test is about did occur


def test_image_size_without_decoding(tmp_path):
    import cv2

    run = load_run()
    img = (np.random.default_rng(0).random((37, 91)) * 255).astype(np.uint8)
    cv2.imwrite(str(tmp_path / "00000007.jpg"), img)
    assert run.image_size(str(tmp_path), 7) == (91, 37)
    cv2.imwrite(str(tmp_path / "00000008.jpg"), img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert run.image_size(str(tmp_path), 8) == (91, 37)
    PKG.synth.write_pgm(str(tmp_path / "00000009.pgm"), img)
    assert run.image_size(str(tmp_path), 9) == (91, 37)
    assert run.resized_size(91, 37, 50)[:2] == (50, 20)
