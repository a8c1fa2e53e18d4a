"""File formats of the drop-in boundary: .dmb, *_cam.txt, pair.txt, config.yaml keys."""
import os

import numpy as np
import pytest

from conftest import PKG

io = PKG.io_formats


def test_dmb_roundtrip(tmp_path):
    # writeDepthDmb/readDepthDmb, /root/reference/src/utility.cpp:193-308: int32 type=1,h,w,nb + float32 row-major
    rng = np.random.default_rng(0)
    for shape in ((7, 5), (4, 9, 3)):
        a = rng.normal(size=shape).astype(np.float32)
        p = str(tmp_path / "x.dmb")
        io.write_dmb(p, a)
        raw = np.fromfile(p, dtype=np.int32, count=4)
        assert raw[0] == 1 and raw[1] == shape[0] and raw[2] == shape[1] and raw[3] == (1 if len(shape) == 2 else 3)
        assert os.path.getsize(p) == 16 + a.size * 4
        np.testing.assert_array_equal(io.read_dmb(p), a)


def test_dmb_truncated(tmp_path):
    p = str(tmp_path / "t.dmb")
    io.write_dmb(p, np.zeros((4, 4), np.float32))
    with open(p, "r+b") as f:
        f.truncate(40)
    with pytest.raises(ValueError):
        io.read_dmb(p)


def test_cam_roundtrip_and_center(tmp_path):
    # ReadCamera, /root/reference/src/PatchMatch.cpp:111-143; C = -R^T t (:134-136)
    K = np.array([[500, 0, 320], [0, 510, 240], [0, 0, 1]], np.float32)
    a = 0.3
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    t = np.array([0.1, -0.2, 2.0], np.float32)
    cam = io.Camera(K=K, R=R, t=t, height=480, width=640, depth_min=1.5, depth_max=7.25)
    p = str(tmp_path / "00000000_cam.txt")
    io.write_cam(p, cam)
    back = io.read_cam(p)
    np.testing.assert_array_equal(back.K, K)
    np.testing.assert_array_equal(back.R, R)
    np.testing.assert_array_equal(back.t, t)
    assert back.depth_min == 1.5 and back.depth_max == 7.25
    np.testing.assert_allclose(R @ cam.C + t, 0, atol=1e-6)
    rec = cam.pack()
    assert rec.itemsize == 112 and rec["width"][0] == 640 and rec["height"][0] == 480


def test_pairs_rules(tmp_path):
    # GenerateSampleList, /root/reference/src/PatchMatch.cpp:67-109: score <= 0 dropped, first `max_src` kept, n_src == 0 skipped
    p = str(tmp_path / "pair.txt")
    with open(p, "w") as f:
        f.write("4\n0\n3 1 5.0 2 0.0 3 2.0\n1\n0\n2\n2 0 1.0 1 2.0\n3\n0\n")
    entries = io.read_pairs(p, max_src=1)
    by_ref = {e.ref_id: e for e in entries}
    assert by_ref[0].src_ids == [0, 1]            # srcID[0] is the reference itself (:84)
    assert by_ref[2].src_ids == [2, 0]
    full = {e.ref_id: e for e in io.read_pairs(p, max_src=20)}
    assert full[0].src_ids == [0, 1, 3]           # score 0.0 dropped
    assert full[1].estimate is False and full[0].estimate is True


@pytest.mark.parametrize("text", [
    "3\n0\n1 1 1.0\n",                          # truncated
    "2\n1\n1 0 1.0\n0\n1 1 1.0\n",            # ids out of order: scenes[id] would be another image
    "1\n0\n2 1 1.0 x\n",                       # not a number
    "2\n0\n1 7 1.0\n1\n1 0 1.0\n",            # a source without an entry (Scenes[7] in the reference, unchecked)
    "2\n0\n1 1 1.0\n2147483647\n1 0 1.0\n",   # an id that would pad two billion empty scenes
])
def test_pairs_malformed(tmp_path, text):
    p = str(tmp_path / "pair.txt")
    with open(p, "w") as f:
        f.write(text)
    with pytest.raises(ValueError):
        io.read_pairs(p)


def test_config_keys(tmp_path):
    # readConfig, /root/reference/src/utility.cpp:8-35 (13 keys, "Planer" spelling is normative)
    p = str(tmp_path / "config.yaml")
    io.write_config(p, **{"Input-folder": "/a", "Output-folder": "/a", "Planer prior": 1})
    cfg = io.read_config(p)
    for k in ("Input-folder", "Output-folder", "Geometric consistency iterations", "Planer prior",
              "Geometric consistency planer prior", "Sky segment", "Use dynamic_consistency to fuse", "Save Dmb as JPG",
              "Save Prior Dmb as JPG", "Save Cost Map", "Save Normal Map", "Max source images num", "Max image size"):
        assert k in cfg, k
    assert cfg["Planer prior"] == 1 and cfg["Input-folder"] == "/a"


def test_dense_folder_layout(tmp_path):
    sc = PKG.synth.make_plane_scene(width=64, height=48, n_views=3, jpeg=True)
    root = str(tmp_path / "dense")
    PKG.synth.write_dense_folder(sc, root)
    for i in range(3):
        assert os.path.exists(os.path.join(root, "images", f"{i:08d}.jpg"))
        assert os.path.exists(os.path.join(root, "cams", f"{i:08d}_cam.txt"))
    entries = io.read_pairs(os.path.join(root, "pair.txt"))
    assert len(entries) == 3 and all(len(e.src_ids) == 3 and e.estimate for e in entries)
    import cv2

    dec = cv2.imread(os.path.join(root, "images", "00000001.jpg"), cv2.IMREAD_GRAYSCALE)
    np.testing.assert_array_equal(dec, sc.images[1])  # scene images are what the reference's imread returns
    np.testing.assert_array_equal(PKG.synth.read_pgm(os.path.join(root, "images", "00000001.pgm")), sc.images[1])
