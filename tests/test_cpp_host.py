"""The C++ host mirror (mp-mvs_b200/csrc/PatchMatchCUDA.h, mpmvs_host.cpp, mpmvs_main.cpp): the reference's entry point
(/root/reference/src/main.cpp) over the dense-folder layout, driven through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, has_gpu

MAIN = os.path.join(ROOT, "mp-mvs_b200", "mpmvs_main")


def build_main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mp-mvs_b200", "csrc")])
    assert os.path.exists(MAIN)


def write_scene(tmp_path, width=320, height=240, **cfg):
    sc = PKG.synth.make_dtu_scene(width=width, height=height, grid=3, n_src=4, seed=2, jpeg=True)
    root = str(tmp_path / "dense")
    PKG.synth.write_dense_folder(sc, root)
    yaml = str(tmp_path / "config.yaml")
    base = {"Input-folder": root, "Output-folder": root, "Max source images num": 4}
    base.update(cfg)
    PKG.io_formats.write_config(yaml, **base)
    return sc, root, yaml


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_main_fails_loudly_without_gpu(tmp_path):
    build_main()
    sc, root, yaml = write_scene(tmp_path, 64, 48)
    r = subprocess.run([MAIN, yaml], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stdout          # no CPU fallback: the first mpmvs_create fails
    assert "There are 9 depthmaps" in r.stdout   # config.yaml, pair.txt were parsed before that


def test_main_reports_missing_inputs(tmp_path):
    build_main()
    r = subprocess.run([MAIN, str(tmp_path / "nope.yaml")], capture_output=True, text=True)
    assert r.returncode != 0 and "can not open config file" in r.stdout
    yaml = str(tmp_path / "c.yaml")
    PKG.io_formats.write_config(yaml, **{"Input-folder": str(tmp_path), "Output-folder": str(tmp_path)})
    r = subprocess.run([MAIN, yaml], capture_output=True, text=True)
    assert r.returncode != 0 and "pair.txt" in r.stdout


@pytest.mark.parametrize("pairs, message", [
    ("3\n0\n1 1 1.0\n", "malformed pair.txt"),                       # truncated: three images announced, one present
    ("2\n1\n1 0 1.0\n0\n1 1 1.0\n", "malformed pair.txt"),         # ids out of order: Scenes[id] would be another image
    ("1\n0\n2 1 1.0 x\n", "malformed pair.txt"),                    # a source entry that is not a number
    ("2\n0\n1 7 1.0\n1\n1 0 1.0\n", "no entry of its own"),        # a source id beyond the list (Scenes[7] in the reference)
    ("2\n0\n1 1 1.0\n2147483647\n1 0 1.0\n", "malformed pair.txt"),  # an id that would pad two billion empty scenes
])
def test_main_refuses_malformed_inputs(tmp_path, pairs, message):
    """The reference indexes Scenes[srcID[i]] and reads pair.txt / cams unchecked (PatchMatch.cpp:67-143,871-890); the host
    mirror must end such inputs with a message and a non-zero exit code, never with an out-of-bounds access."""
    build_main()
    sc, root, yaml = write_scene(tmp_path, 64, 48)
    with open(os.path.join(root, "pair.txt"), "w") as f:
        f.write(pairs)
    dump = tmp_path / "dump"
    dump.mkdir()
    r = subprocess.run([MAIN, yaml, "--check-inputs", str(dump)], capture_output=True, text=True)
    assert r.returncode == 1 and message in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-500:])


def test_main_refuses_truncated_camera_and_dmb(tmp_path):
    build_main()
    sc, root, yaml = write_scene(tmp_path, 64, 48)
    cam = os.path.join(root, "cams", "00000000_cam.txt")
    text = open(cam).read().split()
    with open(cam, "w") as f:
        f.write(" ".join(text[:-3]))                                   # depth range cut off
    dump = tmp_path / "dump"
    dump.mkdir()
    r = subprocess.run([MAIN, yaml, "--check-inputs", str(dump)], capture_output=True, text=True)
    assert r.returncode == 1 and "malformed camera file" in r.stdout, (r.returncode, r.stdout[-500:])


def test_host_parsers_survive_mutated_inputs(tmp_path):
    """Fuzz: truncated, shuffled and garbled pair.txt / camera files / .pgm images must end `mpmvs_main --check-inputs` with
    exit code 0 or 1 (a message), never with a signal (segmentation fault, abort, bad_alloc escaping main)."""
    import random

    build_main()
    sc, root, yaml = write_scene(tmp_path, 64, 48)
    targets = [os.path.join(root, "pair.txt"), os.path.join(root, "cams", "00000000_cam.txt"), os.path.join(root, "images", "00000000.pgm")]
    originals = {t: open(t, "rb").read() for t in targets}
    rng = random.Random(7)
    dump = tmp_path / "dump"
    dump.mkdir()
    codes = set()
    for trial in range(40):
        t = targets[trial % len(targets)]
        data = bytearray(originals[t])
        kind = rng.randrange(4)
        if kind == 0:                                   # truncate
            data = data[:rng.randrange(0, len(data))]
        elif kind == 1:                                 # overwrite a stretch with random bytes
            a = rng.randrange(0, len(data)); b = min(len(data), a + rng.randrange(1, 24))
            data[a:b] = bytes(rng.randrange(256) for _ in range(b - a))
        elif kind == 2:                                 # blow up a number
            toks = data.split()
            if toks:
                toks[rng.randrange(len(toks))] = rng.choice([b"-1", b"99999999999", b"1e40", b"nan", b"2147483647"])
                data = bytearray(b" ".join(toks))
        else:                                           # duplicate a stretch
            a = rng.randrange(0, len(data)); data[a:a] = data[a:a + rng.randrange(1, 64)]
        with open(t, "wb") as f:
            f.write(bytes(data))
        r = subprocess.run([MAIN, yaml, "--check-inputs", str(dump)], capture_output=True, timeout=60)
        codes.add(r.returncode)
        assert r.returncode in (0, 1), (trial, t, kind, r.returncode, r.stdout[-300:], r.stderr[-300:])
        with open(t, "wb") as f:
            f.write(originals[t])
    assert 1 in codes           # some of the mutations were refused with a message


@pytest.mark.parametrize("max_size", [3200, 200])
def test_host_parsers_match_the_python_readers(tmp_path, max_size):
    """readConfig, GenerateSampleList, ReadCamera, the image loader with PatchMatchInit's resize rule and writeDmb of the
    C++ host (`mpmvs_main --check-inputs`, no GPU involved) against io_formats and cv2.resize."""
    import json

    import cv2

    build_main()
    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 1, "Planer prior": 0, "Sky segment": 1, "Save Cost Map": 1,
                                              "Use dynamic_consistency to fuse": 0, "Max source images num": 3, "Max image size": max_size})
    # a pair.txt with the quirks: a gap in the ids, a zero score, more sources than `Max source images num`, an image without sources
    with open(os.path.join(root, "pair.txt"), "w") as f:
        f.write("4\n0\n4 1 5.0 2 0.0 3 2.0 4 1.0\n1\n0\n2\n2 0 1.0 1 2.0\n5\n1 0 3.5\n")
    dump = tmp_path / "dump"
    dump.mkdir()
    r = subprocess.run([MAIN, yaml, "--check-inputs", str(dump)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    got = json.loads(r.stdout)
    cfg = PKG.io_formats.read_config(yaml)
    c = got["config"]
    assert c["input_folder"] == root and c["output_folder"] == root + "/MPMVS"          # utility.cpp:30
    assert (c["geom_iterations"], c["planar_prior"], c["geomPlanarPrior"], c["sky_seg"], c["use_dynamic_consistency"]) == (1, 0, 1, 1, 0)
    assert (c["saveDmb"], c["saveProirDmb"], c["saveCostDmb"], c["saveNormalDmb"]) == (0, 0, 1, 0)
    assert c["MaxSourceImageNum"] == int(cfg["Max source images num"]) == 3 and c["MaxImageSize"] == max_size
    want = PKG.io_formats.read_pairs(os.path.join(root, "pair.txt"), 3)
    assert len(got["scenes"]) == len(want) == 6
    for g, w in zip(got["scenes"], want):
        assert bool(g["estimate"]) == w.estimate
        if not w.estimate:
            continue
        assert g["refID"] == w.ref_id and g["srcID"] == w.src_ids
        cam = PKG.io_formats.read_cam(os.path.join(root, "cams", f"{w.ref_id:08d}_cam.txt"))
        img = cv2.imread(os.path.join(root, "images", f"{w.ref_id:08d}.pgm"), cv2.IMREAD_GRAYSCALE)
        h, wd = img.shape
        K = cam.K.astype(np.float32).copy()
        ref_img = img.astype(np.float32)
        if max(h, wd) > max_size:                                                        # PatchMatch.cpp:893-925
            factor = min(np.float32(max_size) / np.float32(wd), np.float32(max_size) / np.float32(h))
            nw, nh = int(round(float(np.float32(wd) * factor))), int(round(float(np.float32(h) * factor)))
            ipp = cv2.ipp.useIPP()
            cv2.ipp.setUseIPP(False)      # OpenCV's own resize; Intel IPP's (where the build has it) differs by up to 3e-3 grey levels
            ref_img = cv2.resize(ref_img, (nw, nh), interpolation=cv2.INTER_LINEAR)
            cv2.ipp.setUseIPP(ipp)
            sx, sy = np.float32(nw) / np.float32(wd), np.float32(nh) / np.float32(h)
            K[0, 0] *= sx; K[0, 2] *= sx; K[1, 1] *= sy; K[1, 2] *= sy
        assert (g["width"], g["height"], g["orig_width"], g["orig_height"]) == (ref_img.shape[1], ref_img.shape[0], wd, h)
        np.testing.assert_allclose(np.array(g["K"]).reshape(3, 3), K, rtol=1e-6)
        np.testing.assert_allclose(np.array(g["R"]).reshape(3, 3), cam.R, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(g["t"], cam.t, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(g["C"], cam.C, rtol=1e-5, atol=1e-6)
        assert abs(g["depth_min"] - cam.depth_min) < 1e-6 and abs(g["depth_max"] - cam.depth_max) < 1e-6
        np.testing.assert_array_equal(PKG.io_formats.read_dmb(str(dump / f"{w.ref_id:08d}.dmb")), ref_img)      # resizeLinear = cv::resize, bit for bit


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default_schedule", "photometric_only_resized"])
def test_main_end_to_end(tmp_path, mode):
    build_main()
    if mode == "default_schedule":       # the shipped config: planar prior inside the first geom iteration, 2 geom iterations
        sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 2, "Planer prior": 1,
                                                  "Geometric consistency planer prior": 1})
        scale = 1
    else:                                # Max image size below the image size: PatchMatchInit's resize + K scaling path
        sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 0, "Planer prior": 0,
                                                  "Geometric consistency planer prior": 0, "Max image size": 160})
        scale = 2
    r = subprocess.run([MAIN, yaml, "--seed", "5", "--tex", "u8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "cost time is" in r.stdout
    out = os.path.join(root, "MPMVS")
    assert os.path.getsize(os.path.join(out, "MPMVS_model.ply")) > 1000
    accs = []
    for i in range(sc.num_views):
        d = PKG.io_formats.read_dmb(os.path.join(out, f"2333_{i:08d}", "depths.dmb"))
        n = PKG.io_formats.read_dmb(os.path.join(out, f"2333_{i:08d}", "normals.dmb"))
        c = PKG.io_formats.read_dmb(os.path.join(out, f"2333_{i:08d}", "costs.dmb"))
        gt = sc.gt_depth[i][::scale, ::scale] if scale > 1 else sc.gt_depth[i]
        assert d.shape == (240 // scale, 320 // scale) and n.shape == d.shape + (3,) and c.shape == d.shape
        assert np.isfinite(d).all() and abs(float(np.linalg.norm(n, axis=-1).mean()) - 1) < 1e-3
        if scale == 1:
            accs.append(PKG.synth.accuracy_at(d, gt)[2])
        else:   # pixel (x, y) of the half-size image sees the scene along the ray of full-size pixel ~(2x+0.5, 2y+0.5)
            accs.append(float(100 * (np.abs(d - gt) < 0.05 * gt)[gt > 0].mean()))
    print(mode, "accuracy per view", [round(a, 1) for a in accs])
    assert np.median(accs) > (95 if scale == 1 else 85)
    with open(os.path.join(out, "MPMVS_model.ply"), "rb") as f:
        head = f.read(400).decode("latin1")
    nv = int(head.split("element vertex ")[1].split("\n")[0])
    assert nv > 1000


@pytest.mark.gpu
def test_main_resident_mode(tmp_path):
    """--resident: image cache, one resident handle per image, depth maps exchanged in device memory (Jacobi), images in flight.
    Same files, same quality as the sequential (reference-order) mode; bit-reproducible for any --in-flight."""
    build_main()
    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 2, "Planer prior": 1,
                                              "Geometric consistency planer prior": 1})
    out = os.path.join(root, "MPMVS")

    def run(*flags):
        r = subprocess.run([MAIN, yaml, "--seed", "5", "--tex", "u8", "--gpu-fusion", *flags], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        d = [PKG.io_formats.read_dmb(os.path.join(out, f"2333_{i:08d}", "depths.dmb")) for i in range(sc.num_views)]
        npts = int(r.stdout.split("store 3D points to ply file: ")[1].split(" points")[0])
        return d, npts, r.stdout

    d_seq, n_seq, _ = run()
    d_a, n_a, log = run("--resident", "--in-flight", "3")
    d_b, n_b, _ = run("--resident", "--in-flight", "1")
    assert "resident set-up" in log and "cost time is" in log
    for a, b in zip(d_a, d_b):
        np.testing.assert_array_equal(a, b)                     # the order of the host threads does not matter
    acc_seq = np.median([PKG.synth.accuracy_at(d, sc.gt_depth[i])[2] for i, d in enumerate(d_seq)])
    acc_res = np.median([PKG.synth.accuracy_at(d, sc.gt_depth[i])[2] for i, d in enumerate(d_a)])
    print("sequential / resident: accuracy", acc_seq, acc_res, "fused points", n_seq, n_a)
    assert acc_res > 95 and abs(acc_res - acc_seq) < 0.5 and n_a == n_b and abs(n_a - n_seq) < 0.03 * n_seq


@pytest.mark.gpu
def test_main_resident_two_gpus(tmp_path):
    """--resident --gpus 2: reference images sharded over two GPUs by one process, depth maps all-gathered with peer copies.
    Bit-identical to the one-GPU run."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    build_main()
    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 2, "Planer prior": 0,
                                              "Geometric consistency planer prior": 0})
    out = os.path.join(root, "MPMVS")
    res = {}
    for g in (1, 2):
        r = subprocess.run([MAIN, yaml, "--seed", "5", "--tex", "u8", "--no-fusion", "--resident", "--gpus", str(g)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert f"resident set-up on {g} GPU(s)" in r.stdout
        res[g] = [PKG.io_formats.read_dmb(os.path.join(out, f"2333_{i:08d}", name)) for i in range(sc.num_views) for name in ("depths.dmb", "normals.dmb", "costs.dmb")]
    for a, b in zip(res[1], res[2]):
        np.testing.assert_array_equal(a, b)
    accs = [PKG.synth.accuracy_at(res[2][3 * i], sc.gt_depth[i])[2] for i in range(sc.num_views)]
    assert np.median(accs) > 95


@pytest.mark.gpu
@pytest.mark.parametrize("max_size", [3200, 200])
def test_python_cli_matches_layout(tmp_path, max_size):
    """mp-mvs_b200/run.py: the sharded pipeline as a drop-in for main() over the same dense folder and YAML keys."""
    import sys

    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 1, "Planer prior": 1,
                                              "Geometric consistency planer prior": 0, "Max image size": max_size})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "mp-mvs_b200", "run.py"), yaml, "--seed", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "cost time is" in r.stdout
    scale = 1 if max_size >= 320 else 320 / 200
    accs = []
    for i in range(sc.num_views):
        d = PKG.io_formats.read_dmb(os.path.join(root, "MPMVS", f"2333_{i:08d}", "depths.dmb"))
        n = PKG.io_formats.read_dmb(os.path.join(root, "MPMVS", f"2333_{i:08d}", "normals.dmb"))
        assert d.shape == (round(240 / scale), round(320 / scale)) and n.shape == d.shape + (3,)
        if scale == 1:
            accs.append(PKG.synth.accuracy_at(d, sc.gt_depth[i])[2])
        else:
            import cv2

            gt = cv2.resize(sc.gt_depth[i], (d.shape[1], d.shape[0]), interpolation=cv2.INTER_NEAREST)
            accs.append(float(100 * (np.abs(d - gt) < 0.05 * gt)[gt > 0].mean()))
    print("cli accuracy per view", [round(a, 1) for a in accs])
    assert np.median(accs) > (95 if scale == 1 else 85)
    with open(os.path.join(root, "MPMVS", "MPMVS_model.ply"), "rb") as f:       # fused on the GPU by the driver
        head = f.read(400).decode("latin1")
    assert int(head.split("element vertex ")[1].split("\n")[0]) > 1000 and "fusion:" in r.stdout
