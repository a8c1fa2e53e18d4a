"""Golden vectors of depth-map fusion made by THE REFERENCE ITSELF: RunFusion of
/root/reference/src/PatchMatch.cpp:287-504, compiled where it lies into oracle/_ref/libmpmvs_ref_host.so (oracle/ref_host_harness.cu),
reading its images through the real cv::imread (tests/ref_host.py). CPU only.

    python tests/golden/make_ref_fusion_golden.py        # writes tests/golden/ref_fusion.npz

The inputs (camera files, pair.txt, JPEG bytes, depth and normal maps) are stored with the outputs (the vertex records of
MPMVS_model.ply for `Use dynamic_consistency to fuse` = 0 and 1), so the test can rebuild the folder anywhere."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

from conftest import PKG  # noqa: E402
import ref_host  # noqa: E402


def half_size(a, h, w):
    """Samples a full-size map at the pixel centres of the (h, w) image (nearest full-size pixel)."""
    ys = np.clip(np.rint((np.arange(h) + 0.5) * a.shape[0] / h - 0.5).astype(int), 0, a.shape[0] - 1)
    xs = np.clip(np.rint((np.arange(w) + 0.5) * a.shape[1] / w - 0.5).astype(int), 0, a.shape[1] - 1)
    return np.ascontiguousarray(a[ys][:, xs])


def build_folder(dense, n_noise=0.03, map_size=None):
    """map_size (h, w): depth and normal maps smaller than the images, as `Max image size` produces them -- RunFusion then
    resizes the colour image and scales K (RescaleImageAndCamera, PatchMatch.cpp:264-285)."""
    import cv2

    sc = PKG.synth.make_dtu_scene(width=96, height=72, grid=2, n_src=3, seed=4, jpeg=True)
    PKG.synth.write_dense_folder(sc, dense)
    rng = np.random.default_rng(3)
    files = {"pair.txt": open(os.path.join(dense, "pair.txt"), "rb").read()}
    for i in range(sc.num_views):
        g = cv2.imread(os.path.join(dense, "images", f"{i:08d}.jpg"), cv2.IMREAD_GRAYSCALE).astype(np.float32)
        col = np.stack([np.clip(g * 0.6 + 20, 0, 255), np.clip(g * 0.9, 0, 255), np.clip(255 - g * 0.5, 0, 255)], -1).astype(np.uint8)
        jpg = os.path.join(dense, "images", f"{i:08d}.jpg")
        cv2.imwrite(jpg, col, [int(cv2.IMWRITE_JPEG_QUALITY), 95])            # a colour JPEG: B, G, R differ
        d = os.path.join(dense, "MPMVS", f"2333_{i:08d}")
        os.makedirs(d, exist_ok=True)
        gt = sc.gt_depth[i].astype(np.float32)
        gn = sc.gt_normal[i].astype(np.float32)
        if map_size:
            gt, gn = half_size(gt, *map_size), half_size(gn, *map_size)
        dep = (gt * (1 + 0.002 * rng.standard_normal(gt.shape))).astype(np.float32)
        dep[rng.random(gt.shape) < 0.05] = 0
        nrm = gn + n_noise * rng.standard_normal(gn.shape).astype(np.float32)
        nrm = (nrm / np.linalg.norm(nrm, axis=-1, keepdims=True)).astype(np.float32)
        PKG.io_formats.write_dmb(os.path.join(d, "depths.dmb"), dep)
        PKG.io_formats.write_dmb(os.path.join(d, "normals.dmb"), nrm)
        files[f"jpg{i}"] = open(jpg, "rb").read()
        files[f"cam{i}"] = open(os.path.join(dense, "cams", f"{i:08d}_cam.txt"), "rb").read()
        files[f"depth{i}"], files[f"normal{i}"] = dep, nrm
    return sc.num_views, files


def ply_vertices(path):
    b = open(path, "rb").read()
    return np.frombuffer(b[b.index(b"end_header\n") + 11:], np.uint8).copy()


if __name__ == "__main__":
    import cv2

    assert ref_host.available(), "build oracle/_ref first (make -C oracle ref)"
    work = tempfile.mkdtemp(prefix="reffusion_")
    dense = os.path.join(work, "dense")
    n, files = build_folder(dense)
    out = {"n": n, "opencv_version": np.array(cv2.__version__)}
    for k, v in files.items():
        out[k] = np.frombuffer(v, np.uint8) if isinstance(v, bytes) else v
    for dyn in (0, 1):
        ref_host.write_project(work, dense, **{"Use dynamic_consistency to fuse": dyn, "Max source images num": 3})
        ref_host.run_fusion(work)
        out[f"ply_dyn{dyn}"] = ply_vertices(os.path.join(dense, "MPMVS", "MPMVS_model.ply"))
        print("dynamic consistency", dyn, ":", len(out[f"ply_dyn{dyn}"]) // 27, "points")
    # the same images with 64 x 48 maps: the colour images are resized by cv::resize (OpenCV's own 8-bit path: IPP off)
    dense2 = os.path.join(work, "dense_small_maps")
    _, files2 = build_folder(dense2, map_size=(48, 64))
    for k, v in files2.items():
        if k.startswith(("depth", "normal")):
            out["small_" + k] = v
    ref_host.write_project(work, dense2, **{"Use dynamic_consistency to fuse": 1, "Max source images num": 3})
    ref_host.run_fusion(work, ipp=False)
    out["ply_small_maps_dyn1"] = ply_vertices(os.path.join(dense2, "MPMVS", "MPMVS_model.ply"))
    print("64 x 48 maps of 96 x 72 images:", len(out["ply_small_maps_dyn1"]) // 27, "points")
    np.savez_compressed(os.path.join(HERE, "ref_fusion.npz"), **out)
    shutil.rmtree(work, ignore_errors=True)
    print("ref_fusion.npz", os.path.getsize(os.path.join(HERE, "ref_fusion.npz")) // 1024, "KB")
