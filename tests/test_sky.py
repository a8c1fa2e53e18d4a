"""Sky-mask joint-bilateral upsampling (mp-mvs_b200/csrc/pm_sky.cu; reference SkySegment/src/SkyRegionDetect.cu:3-66) and
the sky gate of depth fusion (PatchMatch.cpp:358-388).

CPU: the numpy restatement against cv2.resize and against tests/golden/sky.npz (outputs of the reference's own kernel).
GPU: the CUDA kernel through the C ABI against the same golden file, the live reference kernel and the restatement."""
import os

import numpy as np
import pytest

from conftest import ROOT

import sky_oracle

GOLDEN = os.path.join(ROOT, "tests", "golden", "sky.npz")
CASES = ("a", "b", "c")


def golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/sky.npz not generated yet")
    return np.load(GOLDEN)


def near_threshold(prob, eps=2e-3):
    return np.abs(prob.astype(np.float64) - 0.6) < eps


@pytest.mark.parametrize("size", [(160, 120), (163, 77), (37, 200), (39, 29), (20, 15), (40, 30)])
def test_resize_restatement_matches_cv2(size):
    import cv2

    _, lo, _ = sky_oracle.make_sky_case()
    w, h = size
    np.testing.assert_allclose(sky_oracle.resize_linear(lo, w, h), cv2.resize(lo, (w, h), interpolation=cv2.INTER_LINEAR), atol=3e-6)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_kernel_outputs(name):
    g = golden()
    res, prob = sky_oracle.sky_mask_refine(g[f"{name}_bgr"], g[f"{name}_mask_lo"])
    np.testing.assert_array_equal(sky_oracle.resize_linear(g[f"{name}_mask_lo"], *g[f"{name}_bgr"].shape[1::-1]), g[f"{name}_mask_full"])
    diff = (res > 0) != (g[f"{name}_result"] > 0)
    # libm exp/sqrt against the kernel's fast-math ones: only pixels sitting on the 0.6 threshold may flip
    assert not (diff & ~near_threshold(prob)).any()
    assert diff.mean() < 2e-3


def test_oracle_edge_cases():
    bgr, lo, _ = sky_oracle.make_sky_case(40, 30, 2, 3)
    r1, p1 = sky_oracle.sky_mask_refine(bgr, np.ones_like(lo))
    r0, p0 = sky_oracle.sky_mask_refine(bgr, np.zeros_like(lo))
    assert (r1 == 255).all() and (r0 == 0).all()
    np.testing.assert_allclose(p1, 1.0, atol=1e-5)
    # a constant image reduces the filter to a spatial one: the result is a smoothed mask, symmetric under a flip
    flat = np.full((30, 40, 3), 77, np.uint8)
    m = np.zeros((30, 40), np.float32); m[:, :20] = 1
    _, p = sky_oracle.sky_mask_refine(flat, m)
    np.testing.assert_allclose(p + p[:, ::-1], 1.0, atol=1e-5)
    assert (np.diff(p[15]) <= 1e-6).all()


def test_fusion_oracle_sky_gate():
    import fusion_oracle
    from test_fusion import noisy_scene

    from mpmvs_b200 import io_formats

    sc, depths, normals, lists = noisy_scene(96, 72)
    cams = io_formats.pack_cameras(sc.cams)
    base = fusion_oracle.fuse(cams, depths, normals, sc.images, lists)
    sky = [np.zeros(d.shape, np.uint8) for d in depths]
    for s in sky:
        s[:30] = 255
    gated = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, sky=sky)
    assert 0 < len(gated) < len(base)
    none = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, sky=[np.full(d.shape, 255, np.uint8) for d in depths])
    assert len(none) == 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_kernel_outputs(name):
    from mpmvs_b200 import capi

    g = golden()
    res, prob, ms = capi.sky_mask_refine(g[f"{name}_bgr"], g[f"{name}_mask_lo"], want_prob=True)
    diff = (res > 0) != (g[f"{name}_result"] > 0)
    print(f"sky {name}: {int(diff.sum())} of {diff.size} pixels differ from the reference kernel; {ms:.3f} ms")
    assert set(np.unique(res)) <= {0.0, 255.0}
    assert not (diff & ~near_threshold(prob, 1e-4)).any()
    assert diff.mean() < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(640, 480, 4), (333, 217, 3), (50, 20, 1), (31, 17, 2)])
def test_gpu_vs_live_reference_and_restatement(size):
    from mpmvs_b200 import capi

    w, h, div = size
    bgr, lo, _ = sky_oracle.make_sky_case(w, h, div, seed=w)
    res, prob, _ = capi.sky_mask_refine(bgr, lo, want_prob=True)
    full = sky_oracle.resize_linear(lo, w, h)
    if sky_oracle.ref_available():
        ref = sky_oracle.ref_sky_filter(bgr, full)
        diff = (res > 0) != (ref > 0)
        print(f"{w}x{h}: {int(diff.sum())} pixels differ from the live reference kernel")
        assert not (diff & ~near_threshold(prob, 1e-4)).any()
    if w * h <= 333 * 217:
        want, wprob = sky_oracle.sky_mask_refine(bgr, lo)
        np.testing.assert_allclose(prob, wprob, atol=2e-4)
        assert not (((res > 0) != (want > 0)) & ~near_threshold(wprob)).any()


@pytest.mark.gpu
def test_gpu_full_size_properties():
    """BASELINE ETH3D-shaped frame: size-independent properties + device time next to the reference kernel."""
    from mpmvs_b200 import capi

    w, h = 3200, 2130
    bgr, lo, truth = sky_oracle.make_sky_case(w, h, 8, seed=5)
    res, prob, ms = capi.sky_mask_refine(bgr, lo, want_prob=True)
    assert np.isfinite(prob).all() and prob.min() >= -1e-6 and prob.max() <= 1 + 1e-5
    assert ((res > 0) == truth).mean() > 0.995
    ones, _, _ = capi.sky_mask_refine(bgr, np.ones_like(lo))
    zeros, _, _ = capi.sky_mask_refine(bgr, np.zeros_like(lo))
    assert (ones == 255).all() and (zeros == 0).all()
    line = f"sky filter 3200x2130: ours {ms:.2f} ms"
    if sky_oracle.ref_available():
        full = sky_oracle.resize_linear(lo, w, h)
        ref, ref_ms = sky_oracle.ref_sky_filter(bgr, full, reps=2)
        diff = (res > 0) != (ref > 0)
        assert not (diff & ~near_threshold(prob, 1e-4)).any()
        line += f", reference kernel {ref_ms:.2f} ms, {int(diff.sum())} of {diff.size} pixels differ"
    print(line)


@pytest.mark.gpu
def test_gpu_fusion_sky_gate():
    import fusion_oracle
    from test_fusion import noisy_scene

    from mpmvs_b200 import capi, io_formats

    sc, depths, normals, lists = noisy_scene()
    cams = io_formats.pack_cameras(sc.cams)
    sky = [np.zeros(d.shape, np.uint8) for d in depths]
    for k, s in enumerate(sky):
        s[: 20 + 5 * k] = 200
    sky[3] = None                                          # a view without a mask
    want = fusion_oracle.fuse(cams, depths, normals, sc.images, lists, sky=sky)
    f = capi.Fusion(0, sc.num_views)
    for i in range(sc.num_views):
        f.set_view(i, cams[i:i + 1], depths[i], normals[i], sc.images[i])
        if sky[i] is not None:
            f.set_sky_mask(i, sky[i])
    got, _ = f.run(lists, True)
    f.destroy()
    assert abs(len(got) - len(want)) <= 0.003 * len(want)
    key = lambda p: set(map(tuple, np.round(np.concatenate([p[:, :3] * 1e4, p[:, 3:6] * 1e3], 1)).astype(np.int64)))  # noqa: E731
    a, b = key(got), key(want)
    assert len(a & b) > 0.99 * len(b), (len(a & b), len(b))


def test_sky_error_paths():
    """No GPU here: the entry point must refuse, never fall back (argument errors are checked before the device)."""
    from mpmvs_b200 import capi

    with pytest.raises(capi.MpmvsError):
        capi.sky_mask_refine(np.zeros((4, 4, 3), np.uint8), np.zeros((0, 0), np.float32))


def place_coarse_masks(sc, root, div=4):
    """skymask.jpg per image, as the reference's segmenter leaves them (PatchMatch.cpp:42-44): here the top of every frame."""
    import cv2

    h, w = sc.images[0].shape
    for i in range(sc.num_views):
        folder = os.path.join(root, "MPMVS", f"2333_{i:08d}")
        os.makedirs(folder, exist_ok=True)
        m = np.zeros((h // div, w // div), np.float32)
        m[: (h // div) // 3] = 1.0
        cv2.imwrite(os.path.join(folder, "skymask.jpg"), (255 * m).astype(np.uint8))


@pytest.mark.gpu
def test_cli_sky_segment(tmp_path):
    """`Sky segment: 1` through mp-mvs_b200/run.py: masks refined on the GPU, fusion gated by them."""
    import subprocess
    import sys

    import cv2
    from test_cpp_host import write_scene

    from mpmvs_b200 import capi, io_formats

    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 0, "Planer prior": 0,
                                              "Geometric consistency planer prior": 0, "Sky segment": 1,
                                              "Save Dmb as JPG": 1, "Save Normal Map": 1})
    place_coarse_masks(sc, root)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "mp-mvs_b200", "run.py"), yaml, "--seed", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert f"refined {sc.num_views} sky masks" in r.stdout
    n_cli = int(r.stdout.split("fusion: ")[1].split(" points")[0])
    for name in ("costs.jpg", "depths.jpg", "normals.jpg"):          # previews (PatchMatch.cpp:624-629, main.cpp:51-53)
        assert os.path.getsize(os.path.join(root, "MPMVS", "2333_00000004", name)) > 500, name
    cams = io_formats.pack_cameras(sc.cams)
    counts = {}
    for gate in (False, True):
        f = capi.Fusion(0, sc.num_views)
        lists = []
        for i in range(sc.num_views):
            folder = os.path.join(root, "MPMVS", f"2333_{i:08d}")
            d = io_formats.read_dmb(os.path.join(folder, "depths.dmb"))
            n = io_formats.read_dmb(os.path.join(folder, "normals.dmb"))
            f.set_view(i, cams[i:i + 1], d, n, sc.images[i])
            if gate:
                m = cv2.imread(os.path.join(folder, "skymask_refine.jpg"), cv2.IMREAD_GRAYSCALE)
                assert m is not None and m.shape == d.shape
                assert (m[: d.shape[0] // 4] > 0).mean() > 0.9 and (m[d.shape[0] // 2:] > 0).mean() < 0.02
                assert os.path.exists(os.path.join(folder, "skymask_fuse.jpg"))
                f.set_sky_mask(i, m)
            lists.append([i] + [j for j, _ in sc.pairs[i]])
        counts[gate] = len(f.run(lists, True)[0])
        f.destroy()
    print("fused points without / with the sky gate:", counts, "cli:", n_cli)
    assert counts[True] == n_cli and counts[True] < 0.9 * counts[False]


@pytest.mark.gpu
def test_cpp_host_sky_segment(tmp_path):
    """The same through the C++ host (mpmvs_main): nvJPEG colour decode, skymask_refine.pgm, gated host and GPU fusion."""
    import subprocess

    import cv2
    from test_cpp_host import MAIN, build_main, write_scene

    from mpmvs_b200 import capi

    build_main()
    sc, root, yaml = write_scene(tmp_path, **{"Geometric consistency iterations": 0, "Planer prior": 0,
                                              "Geometric consistency planer prior": 0, "Sky segment": 1})
    place_coarse_masks(sc, root)
    n = {}
    for flag in ("--no-gate", "", "--gpu-fusion"):
        if flag == "--no-gate":
            from mpmvs_b200 import io_formats

            y2 = str(tmp_path / "nogate.yaml")
            io_formats.write_config(y2, **{"Input-folder": root, "Output-folder": root, "Max source images num": 4,
                                           "Geometric consistency iterations": 0, "Planer prior": 0,
                                           "Geometric consistency planer prior": 0, "Sky segment": 0})
            cmd = [MAIN, y2, "--seed", "5", "--tex", "u8"]
        else:
            cmd = [MAIN, yaml, "--seed", "5", "--tex", "u8"] + ([flag] if flag else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        n[flag] = int(r.stdout.split("store 3D points to ply file: ")[1].split(" points")[0])
        if flag:
            continue
        assert f"refined {sc.num_views} sky masks" in r.stdout, [ln for ln in r.stdout.splitlines() if "sky" in ln or "Can not" in ln]
        # the C++ path (nvJPEG colour decode) and the ctypes path (cv2 decode) refine to the same mask up to decoder rounding
        for i in (0, sc.num_views - 1):
            folder = os.path.join(root, "MPMVS", f"2333_{i:08d}")
            got = cv2.imread(os.path.join(folder, "skymask_refine.pgm"), cv2.IMREAD_GRAYSCALE)
            bgr = cv2.imread(os.path.join(root, "images", f"{i:08d}.jpg"), cv2.IMREAD_COLOR)
            coarse = cv2.imread(os.path.join(folder, "skymask.jpg"), cv2.IMREAD_GRAYSCALE).astype(np.float32) / 255
            want, _, _ = capi.sky_mask_refine(bgr, coarse)
            assert got.shape == want.shape and ((got > 0) == (want > 0)).mean() > 0.995
    print("C++ host fused points:", n)
    assert n[""] < 0.9 * n["--no-gate"] and abs(n["--gpu-fusion"] - n[""]) < 0.05 * n[""]
