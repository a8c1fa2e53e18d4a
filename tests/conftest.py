"""pytest configuration: markers, import path, shared fixtures (small synthetic scenes)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pkgload  # noqa: E402

PKG = pkgload.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def build_emul(tag: str = "", flags=()) -> str:
    """tests/emul/libpm_emul<tag>.so: host instantiation of the product's pm_core.cuh (debug aid, tests only)."""
    d = os.path.join(ROOT, "tests", "emul")
    out = os.path.join(d, f"libpm_emul{tag}.so")
    src = os.path.join(d, "pm_emul.cpp")
    deps = [src] + [os.path.join(ROOT, "mp-mvs_b200", "csrc", f) for f in ("pm_core.cuh", "pm_views.h")]
    if not os.path.exists(out) or any(os.path.getmtime(p) > os.path.getmtime(out) for p in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-ffp-contract=off",
                               *flags, "-o", out, src])
    return out


@pytest.fixture(scope="session")
def pkg():
    return PKG


@pytest.fixture(scope="session")
def oracle_cpu():
    import oracle_py

    oracle_py.build("cpu")
    return oracle_py


@pytest.fixture(scope="session")
def small_plane():
    """3-view 96x64 textured plane (config 1 shrunk so the CPU oracle finishes in seconds)."""
    return PKG.synth.make_plane_scene(width=96, height=64, n_views=3, seed=1, jpeg=False)


@pytest.fixture(scope="session")
def small_dtu():
    """5 of the DTU-shaped views at 128x96, 4 sources."""
    return PKG.synth.make_dtu_scene(width=128, height=96, grid=3, n_src=4, seed=2, jpeg=False)


def problem_arrays(scene, ref, max_src=20):
    ids, imgs, cams = scene.problem(ref, max_src)
    return ids, imgs, PKG.io_formats.pack_cameras(cams)


def gt_planes_cam(scene, ref):
    """Ground-truth (camera-frame normal, plane distance) per pixel of view `ref`; invalid pixels get a fronto plane."""
    cam = scene.cams[ref]
    depth = scene.gt_depth[ref].astype(np.float64)
    nw = scene.gt_normal[ref].astype(np.float64)
    h, w = depth.shape
    R = cam.R.astype(np.float64)
    nc = nw @ R.T
    K = cam.K.astype(np.float64)
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    X = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs, dtype=np.float64)], -1)
    bad = depth <= 0
    depth = np.where(bad, 0.5 * (cam.depth_min + cam.depth_max), depth)
    nc[bad] = np.array([0, 0, -1.0])
    flip = (nc * X).sum(-1) > 0
    nc[flip] *= -1
    d = -(nc * X * depth[..., None]).sum(-1)
    return np.concatenate([nc, d[..., None]], -1).astype(np.float32)
