"""Multi-GPU drop-in for the reference's entry point (main(), /root/reference/src/main.cpp) over a dense folder:

    python mp-mvs_b200/run.py config.yaml [--seed N]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 mp-mvs_b200/run.py config.yaml

Reads the same config.yaml keys, `images/%08d.jpg` (or the decoded `.pgm` sidecar), `cams/%08d_cam.txt` and `pair.txt`, runs
stage 1 and the geometric-consistency iterations sharded by reference image (pipeline.DensePipeline) and writes
`<Output-folder>/MPMVS/2333_%08d/{depths,normals,costs}.dmb` with background writer threads, then (optionally) refines the sky
masks (`Sky segment: 1`) and fuses the depth maps on rank 0's GPU into `MPMVS_model.ply`.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkgload  # noqa: E402

pkgload.load_package()
from mpmvs_b200 import io_formats, pipeline, previews  # noqa: E402


def cv_resize(img, size):
    """cv::resize(INTER_LINEAR) as OpenCV itself computes it. Where the cv2 build has Intel IPP, IPP's resize answers up to 3e-3
    grey levels (float) / one level (8 bit) differently; the C++ host's resizes are OpenCV's own paths bit for bit
    (mpmvs_host.cpp: resizeLinear, resizeLinearBGR), so IPP is switched off for the call and both hosts see the same pixels."""
    import cv2

    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        return cv2.resize(img, size, interpolation=cv2.INTER_LINEAR)
    finally:
        cv2.ipp.setUseIPP(ipp)


def load_image(folder: str, image_id: int, max_size: int):
    """PatchMatchInit's loading rule (PatchMatch.cpp:871-925): grey uint8 -> float32, cv::resize(INTER_LINEAR) above
    `max_size`. Returns (image, scale_x, scale_y); the image stays uint8 when it was not resized."""
    import cv2

    pgm = os.path.join(folder, f"{image_id:08d}.pgm")
    img = cv2.imread(pgm if os.path.exists(pgm) else os.path.join(folder, f"{image_id:08d}.jpg"), cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(f"Can not read this image ! {image_id:08d}")
    h, w = img.shape
    if w <= max_size and h <= max_size:
        return img, 1.0, 1.0
    factor = min(np.float32(max_size) / np.float32(w), np.float32(max_size) / np.float32(h))
    nw, nh = int(round(float(np.float32(w) * factor))), int(round(float(np.float32(h) * factor)))
    out = cv_resize(img.astype(np.float32), (nw, nh))
    return out, nw / np.float32(w), nh / np.float32(h)


def load_color(folder: str, image_id: int, shape):
    """The colour image RunFusion reads (cv::imread(IMREAD_COLOR), PatchMatch.cpp:322) at the depth map's size (its
    RescaleImageAndCamera resizes with INTER_LINEAR, cpp:341-345). None when the folder only has the grey .pgm sidecar."""
    import cv2

    bgr = cv2.imread(os.path.join(folder, f"{image_id:08d}.jpg"), cv2.IMREAD_COLOR)
    if bgr is None:
        return None
    if bgr.shape[:2] != tuple(shape):
        bgr = cv_resize(bgr, (shape[1], shape[0]))
    return bgr


def image_size(folder: str, image_id: int):
    """(width, height) of an image of the dense folder without decoding it: the PGM header of the sidecar, else the JPEG's
    SOF segment. Falls back to decoding."""
    pgm = os.path.join(folder, f"{image_id:08d}.pgm")
    if os.path.exists(pgm):
        with open(pgm, "rb") as f:
            tok = f.read(64).split()
        if tok and tok[0] == b"P5":
            return int(tok[1]), int(tok[2])
    jpg = os.path.join(folder, f"{image_id:08d}.jpg")
    try:
        with open(jpg, "rb") as f:
            if f.read(2) == b"\xff\xd8":
                while True:
                    b = f.read(1)
                    while b and b != b"\xff":
                        b = f.read(1)
                    m = f.read(1)
                    while m == b"\xff":
                        m = f.read(1)
                    if not m:
                        break
                    k = m[0]
                    if k in (0xD8, 0x01) or 0xD0 <= k <= 0xD7:
                        continue
                    n = int.from_bytes(f.read(2), "big")
                    if 0xC0 <= k <= 0xCF and k not in (0xC4, 0xC8, 0xCC):      # SOFn: precision, height, width
                        seg = f.read(5)
                        return int.from_bytes(seg[3:5], "big"), int.from_bytes(seg[1:3], "big")
                    f.seek(n - 2, 1)
    except OSError:
        pass
    import cv2

    img = cv2.imread(pgm if os.path.exists(pgm) else jpg, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(f"Can not read this image ! {image_id:08d}")
    return img.shape[1], img.shape[0]


def resized_size(w: int, h: int, max_size: int):
    """PatchMatchInit's resize rule (PatchMatch.cpp:893-925): (new width, new height, scale x, scale y)."""
    if w <= max_size and h <= max_size:
        return w, h, 1.0, 1.0
    factor = min(np.float32(max_size) / np.float32(w), np.float32(max_size) / np.float32(h))
    nw, nh = int(round(float(np.float32(w) * factor))), int(round(float(np.float32(h) * factor)))
    return nw, nh, nw / np.float32(w), nh / np.float32(h)


def sky_image_size(w: int, h: int, max_size: int):
    """Size rule of GenerateSkyRegionMask (PatchMatch.cpp:21-33): the image size PatchMatch worked at."""
    if w <= max_size and h <= max_size:
        return w, h
    factor = min(np.float32(max_size) / np.float32(w), np.float32(max_size) / np.float32(h))
    return int(round(float(np.float32(w) * factor))), int(round(float(np.float32(h) * factor)))


def generate_sky_masks(cfg: dict, refs, rank: int, world: int, device: int):
    """GenerateSkyRegionMask (PatchMatch.cpp:4-57) without the network: `<out>/MPMVS/2333_%08d/skymask.jpg` (255 x the sky
    probability of any segmenter, any size; the reference's own ncnn model writes exactly this file, cpp:42-44) is refined by
    joint-bilateral upsampling on the GPU (mpmvs_sky_mask_refine) into skymask_refine.jpg, which gates fusion, plus the
    skymask_fuse.jpg preview (image_mask_fuse, SkyRegionDetect.cpp:462-478). Images are sharded over the ranks.
    Returns the number of masks this rank refined."""
    import cv2

    from mpmvs_b200 import capi

    inp, out = cfg["Input-folder"].rstrip("/"), cfg["Output-folder"].rstrip("/")
    done = 0
    for k, i in enumerate(refs):
        if k % world != rank:
            continue
        folder = io_formats.result_dir(out, i)
        coarse = cv2.imread(os.path.join(folder, "skymask.jpg"), cv2.IMREAD_GRAYSCALE)
        if coarse is None:
            continue
        bgr = cv2.imread(os.path.join(inp, "images", f"{i:08d}.jpg"), cv2.IMREAD_COLOR)
        if bgr is None:
            raise FileNotFoundError(f"Can not read this image ! {i:08d}")
        w, h = sky_image_size(bgr.shape[1], bgr.shape[0], int(cfg["Max image size"]))
        if (w, h) != (bgr.shape[1], bgr.shape[0]):
            bgr = cv2.resize(bgr, (w, h), interpolation=cv2.INTER_LINEAR)          # cpp:36
        refined, _, _ = capi.sky_mask_refine(bgr, coarse.astype(np.float32) / np.float32(255), device=device)
        cv2.imwrite(os.path.join(folder, "skymask_refine.jpg"), refined)           # cpp:47-50
        fuse = bgr.copy()
        fuse[refined > 0] = 0
        cv2.imwrite(os.path.join(folder, "skymask_fuse.jpg"), fuse)
        done += 1
    return done


def load_sky_mask(out: str, image_id: int, shape):
    """The sky gate RunFusion reads (PatchMatch.cpp:358-373): skymask_refine.jpg, brought to the depth map's size."""
    import cv2

    m = cv2.imread(os.path.join(io_formats.result_dir(out, image_id), "skymask_refine.jpg"), cv2.IMREAD_GRAYSCALE)
    if m is None:
        return None
    if m.shape != tuple(shape):
        m = cv2.resize(m, (shape[1], shape[0]), interpolation=cv2.INTER_LINEAR)
    return m


def load_scene(cfg: dict, rank: int, world: int):
    inp = cfg["Input-folder"].rstrip("/")
    entries = io_formats.read_pairs(os.path.join(inp, "pair.txt"), int(cfg["Max source images num"]))
    refs = sorted(e.ref_id for e in entries if e.estimate)
    mine = pipeline.shard_refs(refs, rank, world)
    by_ref = {e.ref_id: e for e in entries if e.estimate}
    need = sorted({i for r in mine for i in by_ref[r].src_ids[: 1 + int(cfg["Max source images num"])]} | set(refs[:1]))
    cams, images = {}, {}
    # decode on host threads (cv2 releases the GIL): the reference decodes every view again for every problem that uses it
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        decoded = dict(zip(need, ex.map(lambda i: load_image(os.path.join(inp, "images"), i, int(cfg["Max image size"])), need)))
    max_size = int(cfg["Max image size"])
    for i in sorted(set(refs) | set(need)):
        cam = io_formats.read_cam(os.path.join(inp, "cams", f"{i:08d}_cam.txt"))
        if i in need:
            img, sx, sy = decoded[i]
            images[i] = img
            h, w = img.shape
        else:
            # an image another rank estimates: this rank (rank 0: fusion) still needs its camera AT THE SIZE PatchMatch worked
            # at -- the same K scaling as RescaleImageAndCamera (PatchMatch.cpp:906-921), from the header of the file
            w, h, sx, sy = resized_size(*image_size(os.path.join(inp, "images"), i), max_size)
        cam.K = cam.K.copy()
        cam.K[0, 0] *= sx; cam.K[0, 2] *= sx; cam.K[1, 1] *= sy; cam.K[1, 2] *= sy
        cam.height, cam.width = h, w
        cams[i] = cam
    return entries, cams, images


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--seed", type=int, default=0x2333)
    ap.add_argument("--in-flight", type=int, default=8, help="reference images in flight per GPU (host threads: the planar-prior triangulation is host work)")
    ap.add_argument("--fusion", type=int, default=1, choices=[0, 1, 2],
                    help="1 (default): fuse the depth maps on rank 0's GPU (image-parallel kernels) and write MPMVS_model.ply; 2: the reference's "
                         "sequential host order through the C++ host (mpmvs_main --fusion-only): the reference's .ply byte for byte, "
                         "single-threaded; 0: no fusion")
    ap.add_argument("--order", default="jacobi", choices=["jacobi", "gauss_seidel"],
                    help="jacobi (default): a geometric pass reads the previous pass's depth maps, results independent of the number of "
                         "GPUs; gauss_seidel (one GPU): the reference's in-place order (src/PatchMatch.cpp:620-633) -- with the exact "
                         "arithmetic the .dmb files are then byte-identical to the reference's and to mpmvs_main's for the same --seed")
    ap.add_argument("--arithmetic", "--fidelity", dest="arithmetic", default="exact", choices=["exact", "fast"],
                    help="exact (default): the reference's kernel results bit for bit, float32 view storage; fast: the library's "
                         "hoisted arithmetic with 8-bit view storage, statistically equal results (mpmvs_set_arithmetic)")
    args = ap.parse_args()
    import torch

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = io_formats.read_config(args.config)
    t0 = time.time()
    entries, cams, images = load_scene(cfg, rank, world)
    from mpmvs_b200 import capi

    resized = any(im.dtype != np.uint8 for im in images.values())
    pcfg = pipeline.PipelineConfig(geom_iterations=int(cfg["Geometric consistency iterations"]), max_src=int(cfg["Max source images num"]),
                                   seed=args.seed, planar_prior=bool(int(cfg["Planer prior"])),
                                   geom_planar_prior=bool(int(cfg["Geometric consistency planer prior"])),
                                   tex_format=capi.TEX_F32 if (resized or args.arithmetic == "exact") else capi.TEX_U8, in_flight=args.in_flight,
                                   arithmetic=args.arithmetic, order=args.order)
    p = pipeline.DensePipeline(entries, cams, images, pcfg, rank=rank, world=world, device=local, dist=dist)
    p.setup()
    t1 = time.time()
    stats = p.run()
    torch.cuda.synchronize()
    t2 = time.time()
    out = cfg["Output-folder"].rstrip("/")
    p.write_results(out)
    for ref in p.my_refs:                                  # costs.jpg next to the .dmb files (PatchMatch.cpp:624-629)
        folder = io_formats.result_dir(out, ref)
        previews.save_cost(io_formats.read_dmb(os.path.join(folder, "costs.dmb")), os.path.join(folder, "costs.jpg"))
    if any(int(cfg[k]) for k in ("Save Dmb as JPG", "Save Prior Dmb as JPG", "Save Cost Map", "Save Normal Map")):   # main.cpp:51-53
        print("Start save data as jpg")
        previews.save_dmb_as_jpg(cfg, p.my_refs)
        print("save data success")
    t3 = time.time()
    sky_seg = bool(int(cfg["Sky segment"]))
    if sky_seg:                                            # main.cpp:44-46
        n_sky = generate_sky_masks(cfg, sorted(e.ref_id for e in entries if e.estimate), rank, world, local)
        print(f"rank {rank}: refined {n_sky} sky masks in {time.time() - t3:.2f} s")
    if dist is not None:
        dist.barrier()
    t4 = t3
    n_points = None
    p.destroy()            # the resident PatchMatch state (76 B/px per reference image) is not needed any more: free it before fusion
    if rank == 0 and args.fusion == 2:
        # RunFusion in the reference's own pixel order (a later pixel sees the masks earlier ones set, PatchMatch.cpp:374-499): host code,
        # which lives in the C++ host; it reads the .dmb files every rank has written
        import subprocess

        exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpmvs_main")
        if not os.path.exists(exe):
            raise FileNotFoundError(f"{exe}: build it with `make -C mp-mvs_b200/csrc`")
        r = subprocess.run([exe, args.config, "--fusion-only"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("mpmvs_main --fusion-only failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
        n_points = int(r.stdout.split("store 3D points to ply file: ")[1].split(" points")[0])
        t4 = time.time()
    elif rank == 0 and args.fusion:
        # RunFusion (PatchMatch.cpp:287-504) on the GPU: every rank has written its maps, rank 0 reads them back
        est = sorted(e.ref_id for e in entries if e.estimate)
        by_ref = {e.ref_id: e for e in entries if e.estimate}
        fz = capi.Fusion(local, max(est) + 1)
        lists = [None] * (max(est) + 1)
        for i in est:
            d = io_formats.read_dmb(os.path.join(io_formats.result_dir(out, i), "depths.dmb"))
            nrm = io_formats.read_dmb(os.path.join(io_formats.result_dir(out, i), "normals.dmb"))
            img, _, _ = load_image(os.path.join(cfg["Input-folder"].rstrip("/"), "images"), i, int(cfg["Max image size"]))
            cam = cams[i]
            cam.height, cam.width = d.shape
            fz.set_view(i, io_formats.pack_cameras([cam]), d, nrm, np.clip(np.rint(img), 0, 255).astype(np.uint8))
            bgr = load_color(os.path.join(cfg["Input-folder"].rstrip("/"), "images"), i, d.shape)
            if bgr is not None:
                fz.set_color(i, bgr)                       # RunFusion averages B, G, R of the colour image (cpp:322,443-445)
            sky = load_sky_mask(out, i, d.shape) if sky_seg else None
            if sky is not None:
                fz.set_sky_mask(i, sky)
            lists[i] = [i] + [j if j in by_ref else -1 for j in by_ref[i].src_ids[1:]]
        pts, fus_ms = fz.run(lists, bool(int(cfg["Use dynamic_consistency to fuse"])))
        fz.destroy()
        io_formats.write_ply(os.path.join(out, "MPMVS", "MPMVS_model.ply"), pts)
        n_points = len(pts)
        t4 = time.time()
    if rank == 0:
        n = len([e for e in entries if e.estimate])
        print(f"There are {n} depthmaps; load {t1 - t0:.2f} s, PatchMatch stages {t2 - t1:.2f} s "
              f"({[(s.name, round(s.device_ms)) for s in stats]}), write {t3 - t2:.2f} s on {world} GPU(s)")
        print(f"cost time is {(t2 - t1) * 1e6:.10f} us")
        if n_points is not None:
            print(f"fusion: {n_points} points in {t4 - t3:.2f} s ({'host, reference order' if args.fusion == 2 else 'load + GPU fusion'} + ply)")
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
