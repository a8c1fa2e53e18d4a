"""GPU tool (TEST INFRASTRUCTURE): the reference's WHOLE program -- main(), ProcessProblem, the planar-prior host stage,
RunFusion, Run() and every kernel, compiled where they lie into oracle/_ref/libmpmvs_ref_host.so with OpenCV's numerics
served by the real OpenCV (tests/ref_host.py) -- against the product's host program (mp-mvs_b200/mpmvs_main, reference
order, exact arithmetic, float32 views) on the same dense folder with the same seeds. Compares the files both write
(depths / normals / costs .dmb per image, MPMVS_model.ply) byte for byte, and, where the schedule has a planar prior, the
prior the reference built (captured at its upload, PatchMatch.cpp:994-995) with mpmvs_build_prior on the same state.

    python tests/tools/reference_program.py [--scene small|dtu|eth3d] [--width W --height H] [--schedules photo_geom,planar,geom_planar] [--golden out.npz]

Prints one JSON line."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)

from conftest import PKG  # noqa: E402
import ref_host  # noqa: E402

MAIN = os.path.join(ROOT, "mp-mvs_b200", "mpmvs_main")

SCHEDULES = {   # config.yaml keys (utility.cpp:8-35); main.cpp:19-41 decides which ProcessProblem gets a planar prior
    "photo_geom": {"Geometric consistency iterations": 2, "Planer prior": 0, "Geometric consistency planer prior": 0},
    "planar": {"Geometric consistency iterations": 1, "Planer prior": 1, "Geometric consistency planer prior": 0},
    "geom_planar": {"Geometric consistency iterations": 2, "Planer prior": 1, "Geometric consistency planer prior": 1},
    # GenerateSampleList's rules (PatchMatch.cpp:67-109) on a pair.txt that exercises them: zero-score entries are dropped but
    # still count towards `Max source images num` (the test is on the listed POSITION j), sources beyond it are ignored, the
    # images are listed in another order than their ids, and the fusion rule without dynamic consistency (>= 2 views)
    "quirks": {"Geometric consistency iterations": 1, "Planer prior": 0, "Geometric consistency planer prior": 0, "Max source images num": 3,
               "Use dynamic_consistency to fuse": 0},
    # images above `Max image size`: PatchMatchInit resizes them with cv::resize (PatchMatch.cpp:893-925) and RunFusion resizes the
    # colour image (RescaleImageAndCamera, :264-285). The product's resizes are OpenCV's own float and 8-bit paths bit for bit; the
    # reference side of this schedule therefore runs with Intel IPP switched off (IPP's resize differs by up to 3e-3 grey levels)
    "resized": {"Geometric consistency iterations": 1, "Planer prior": 0, "Geometric consistency planer prior": 0, "Max image size": 200},
}


def quirky_pairs(path, sc):
    order = list(range(sc.num_views))
    with open(path, "w") as f:
        f.write(f"{sc.num_views}\n")
        for ref in order:
            row = list(sc.pairs[ref])
            row.insert(1, (row[-1][0], 0.0))            # a zero-score entry in second position
            if ref % 2:
                row = row[:1] + row[1:][::-1]            # and another order of the rest
            f.write(f"{ref}\n{len(row)} " + " ".join(f"{i} {s:.4f}" for i, s in row) + "\n")


def render(kind, width, height, views):
    """small: 9 views, 4 sources (the test's size); dtu / eth3d: the shapes of BASELINE.json configs[1] / [2] (49 views
    1600x1200 and 11 views 3200x2130 by default, 10 sources)."""
    workers = min(32, os.cpu_count() or 1)
    if kind == "small":
        return PKG.synth.make_dtu_scene(width=width, height=height, grid=3, n_src=4, seed=2, jpeg=True), 4
    if kind == "dtu":
        return PKG.synth.make_dtu_scene(width=width, height=height, workers=workers, jpeg=True), 10
    return PKG.synth.make_eth3d_scene(width=width, height=height, n_views=views or 11, workers=workers, jpeg=True), 10


def write_inputs(dense, sc):
    """A rendered scene as a dense folder: JPEGs (what the reference decodes through cv::imread) plus the decoded .pgm / .ppm
    sidecars the OpenCV-free product host reads, so both sides see the same pixels."""
    import cv2

    PKG.synth.write_dense_folder(sc, dense)
    for i in range(sc.num_views):
        bgr = cv2.imread(os.path.join(dense, "images", f"{i:08d}.jpg"), cv2.IMREAD_COLOR)
        with open(os.path.join(dense, "images", f"{i:08d}.ppm"), "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (bgr.shape[1], bgr.shape[0]))
            f.write(np.ascontiguousarray(bgr[:, :, ::-1]).tobytes())
    return sc


def read_results(dense, n):
    out = []
    for i in range(n):
        d = os.path.join(dense, "MPMVS", f"2333_{i:08d}")
        out.append({k: open(os.path.join(d, k + ".dmb"), "rb").read() for k in ("depths", "normals", "costs")})
    path = os.path.join(dense, "MPMVS", "MPMVS_model.ply")
    ply = open(path, "rb").read() if os.path.exists(path) else b""
    return out, ply


def ply_points(b):
    h = b.index(b"end_header\n") + 11
    return np.frombuffer(b[h:], np.dtype([("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)]))


def compare_prior(sc, cap, width, height, max_src=4):
    """mpmvs_build_prior on the state the reference built its prior from, against the prior it uploaded."""
    from mpmvs_b200 import capi

    ref = int(cap["scene_index"])
    ids, imgs, cams = sc.problem(ref, max_src)
    pm = capi.PatchMatch(0).set_problem(imgs, PKG.io_formats.pack_cameras(cams))
    with_geom = cap["in_geom"] is not None
    pm.set_geom_consistency_params(with_geom, with_geom)
    pm.set_dev_state({"planes": cap["in_planes"].reshape(height, width, 4), "costs": cap["in_costs"].reshape(height, width), "views": None,
                      "rng": None, "geom": cap["in_geom"].reshape(height, width) if with_geom else None})
    verts = pm.pick_vertices(with_geom)                  # the three steps of mpmvs_build_prior (pm_capi.cu), which itself wants a Run() first
    tris = capi.delaunay(verts, width, height)
    pm.prior_from_triangles(verts, tris)
    stats = {"n_vertices": len(verts), "n_triangles": len(tris)}
    prior, mask = pm.get_prior()
    pm.destroy()
    if os.environ.get("REFPROG_DUMP"):
        np.savez_compressed(os.environ["REFPROG_DUMP"] + f"_{ref}_{int(cap['stage'])}.npz", prior=prior, mask=mask, verts=verts, tris=tris)
    rmask = cap["mask"].reshape(height, width)
    rprior = cap["planes"].reshape(height, width, 4)
    same = (mask == rmask)
    both = same & (mask > 0)
    d = np.abs(prior[both] - rprior[both])
    return {"scene_index": int(ref), "stage": int(cap["stage"]), "geom_variant": bool(with_geom), "vertices": int(stats["n_vertices"]),
            "triangles": int(stats["n_triangles"]), "reference_prior_pixels": float((rmask > 0).mean()), "our_prior_pixels": float((mask > 0).mean()),
            "same_triangle_id": float(same.mean()), "planes_bit_identical_same_id": float((d == 0).all(-1).mean()) if both.any() else None,
            "max_abs_plane_diff_same_id": float(d.max()) if both.any() else None, "mean_abs_plane_diff_same_id": float(d.mean()) if both.any() else None}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="small", choices=["small", "dtu", "eth3d"])
    ap.add_argument("--views", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--svd-noise", action="store_true", help="also run the reference with another SVD behind cv::SVD::solveZ (numpy, float64) "
                    "and report reference-against-reference agreement: how much of a difference is the plane fit's rounding noise")
    ap.add_argument("--python-host", action="store_true", help="also run the Python host (mp-mvs_b200/run.py --order gauss_seidel --fusion 2) and "
                    "compare its files with the reference's")
    ap.add_argument("--priors", type=int, default=3, help="how many of the captured priors to compare with the product's stage")
    ap.add_argument("--schedules", default="photo_geom,planar,geom_planar")
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--golden", help="write the reference's captured priors (inputs and outputs) of the first image of every schedule here")
    ap.add_argument("--keep", help="keep the working folders under this directory")
    a = ap.parse_args()
    assert ref_host.available(), "oracle/_ref/libmpmvs_ref_host.so is not built"
    dw, dh = {"small": (320, 240), "dtu": (1600, 1200), "eth3d": (3200, 2130)}[a.scene]
    a.width, a.height = a.width or dw, a.height or dh
    t0 = time.time()
    scene0, max_src = render(a.scene, a.width, a.height, a.views)
    print(f"rendered {scene0.num_views} views {a.width}x{a.height} in {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mp-mvs_b200", "csrc")])
    work = a.keep or tempfile.mkdtemp(prefix="refprog_")
    res = {"scene": a.scene, "views": scene0.num_views, "size": [a.width, a.height], "max_src": max_src, "seed": a.seed, "schedules": {}}
    golden = {}
    for name in a.schedules.split(","):
        cfg = dict({"Max source images num": max_src}, **SCHEDULES[name])
        sides = {}
        for side in ("reference", "ours") + (("reference_other_svd",) if a.svd_noise and (cfg["Planer prior"] or cfg["Geometric consistency planer prior"]) else ()) + (("python_host",) if a.python_host else ()):   # noqa: E501
            proj = os.path.join(work, name, side)
            shutil.rmtree(proj, ignore_errors=True)
            dense = os.path.join(proj, "dense")
            import copy

            sc = write_inputs(dense, copy.deepcopy(scene0))
            if name == "quirks":
                quirky_pairs(os.path.join(dense, "pair.txt"), sc)
            yaml = ref_host.write_project(proj, dense, **cfg)
            t0 = time.time()
            if side.startswith("reference"):
                cap = os.path.join(proj, "priors.npz")
                log = ref_host.run_main(proj, seed=a.seed, capture=cap, timeout=6000, solvez="cv2" if side == "reference" else "numpy64",
                                        ipp=name != "resized")
            elif side == "python_host":
                r = subprocess.run([sys.executable, os.path.join(ROOT, "mp-mvs_b200", "run.py"), yaml, "--seed", str(a.seed), "--order", "gauss_seidel",
                                    "--fusion", "2", "--arithmetic", "exact"], capture_output=True, text=True, timeout=6000)
                assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
                log = r.stdout
            else:
                r = subprocess.run([MAIN, yaml, "--seed", str(a.seed), "--tex", "f32", "--arithmetic", "exact"], capture_output=True, text=True, timeout=6000)
                assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
                log = r.stdout
            sides[side] = {"wall_s": round(time.time() - t0, 2), "results": read_results(dense, sc.num_views), "log_tail": log[-300:]}
        (ra, rply), (oa, oply) = sides["reference"]["results"], sides["ours"]["results"]
        ident = {k: [bool(x[k] == y[k]) for x, y in zip(ra, oa)] for k in ("depths", "normals", "costs")}
        entry = {"images": len(ra), "depth_maps_byte_identical": int(sum(ident["depths"])), "normal_maps_byte_identical": int(sum(ident["normals"])),
                 "cost_maps_byte_identical": int(sum(ident["costs"])), "ply_byte_identical": bool(rply == oply),
                 "ply_points": [int(len(ply_points(rply))), int(len(ply_points(oply)))],
                 "wall_s": {k: v["wall_s"] for k, v in sides.items()}, "reference_opencv_calls": sides["reference"]["log_tail"].strip().splitlines()[-1]}
        def closeness(xa, ya):
            agree, accs = [], []
            for i, (x, y) in enumerate(zip(xa, ya)):
                hh, ww = (int(v) for v in np.frombuffer(x["depths"][4:12], np.int32))
                if y["depths"][:16] != x["depths"][:16]:
                    return [0.0, 0.0], [0.0, 0.0]
                dr = np.frombuffer(x["depths"][16:], np.float32).reshape(hh, ww)
                do = np.frombuffer(y["depths"][16:], np.float32).reshape(hh, ww)
                nr = np.frombuffer(x["normals"][16:], np.float32).reshape(hh, ww, 3)
                no = np.frombuffer(y["normals"][16:], np.float32).reshape(hh, ww, 3)
                gt = sc.gt_depth[i]
                if (hh, ww) != gt.shape:      # ground truth at the resized pixel centres (nearest full-size pixel)
                    ys = np.clip(np.rint((np.arange(hh) + 0.5) * gt.shape[0] / hh - 0.5).astype(int), 0, gt.shape[0] - 1)
                    xs = np.clip(np.rint((np.arange(ww) + 0.5) * gt.shape[1] / ww - 0.5).astype(int), 0, gt.shape[1] - 1)
                    gt = gt[ys][:, xs]
                agree.append(PKG.synth.depth_normal_agreement(do, no, dr, nr, gt > 0))
                accs.append([PKG.synth.accuracy_at(dr, gt)[0], PKG.synth.accuracy_at(do, gt)[0]])
            return [float(np.median(agree)), float(np.min(agree))], [float(np.mean([x[0] for x in accs])), float(np.mean([x[1] for x in accs]))]

        if not all(ident["depths"]):      # how close, where not identical
            entry["agreement_median_min"], entry["accuracy_2cm_reference_ours"] = closeness(ra, oa)
        if "python_host" in sides:           # the Python host in the reference's order: same seeds, same files
            pa, _ = sides["python_host"]["results"]
            entry["python_host"] = {k + "_maps_byte_identical": int(sum(x[k + "s"] == y[k + "s"] for x, y in zip(ra, pa))) for k in ("depth", "normal", "cost")}
            entry["python_host"]["ply_byte_identical"] = bool(sides["python_host"]["results"][1] == rply)
            entry["python_host"]["wall_s"] = sides["python_host"]["wall_s"]
        if "reference_other_svd" in sides:   # the reference against itself with another SVD behind cv::SVD::solveZ
            rb, _ = sides["reference_other_svd"]["results"]
            entry["reference_vs_reference_other_svd"] = {"depth_maps_byte_identical": int(sum(x["depths"] == y["depths"] for x, y in zip(ra, rb)))}
            entry["reference_vs_reference_other_svd"]["agreement_median_min"], entry["reference_vs_reference_other_svd"]["accuracy_2cm"] = closeness(ra, rb)
        z = np.load(os.path.join(work, name, "reference", "priors.npz"))
        n_pri = int(z["n"])
        entry["reference_priors_captured"] = n_pri
        if n_pri:
            caps = [{k: (z[f"p{j}_{k}"] if f"p{j}_{k}" in z.files else None) for k in ("scene_index", "stage", "planes", "mask", "in_planes", "in_costs", "in_geom")}
                    for j in range(n_pri)]
            entry["prior_stage"] = [compare_prior(sc, c, a.width, a.height, max_src) for c in caps[:a.priors]]
            if a.golden:
                c = caps[0]
                ids, imgs, cams = sc.problem(int(c["scene_index"]), max_src)
                c = dict(c)
                c["planes"] = np.where((c["mask"] > 0)[:, None], c["planes"].reshape(-1, 4), 0).astype(np.float32)   # undefined (new[]) where there is no prior
                golden.update({f"{name}/{k}": v for k, v in c.items() if v is not None})
                golden[f"{name}/cams"] = PKG.io_formats.pack_cameras(cams)[:1]
                golden[f"{name}/size"] = np.array([a.width, a.height], np.int32)
        res["schedules"][name] = entry
        print(name, json.dumps(entry), file=sys.stderr, flush=True)
    if a.golden and golden:
        import cv2

        np.savez_compressed(a.golden, opencv_version=np.array(cv2.__version__), **golden)
    if not a.keep:
        shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(res))
