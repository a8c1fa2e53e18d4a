"""The host stages either side of the hot path and the whole program, against THE REFERENCE
ITSELF: /root/reference/src/main.cpp, PatchMatch.cpp, utility.cpp and PatchMatch.cu compiled where they lie into
oracle/_ref/libmpmvs_ref_host.so (oracle/ref_host_harness.cu), with OpenCV's numerics -- cv::imread, cv::resize,
cv::Subdiv2D, cv::SVD::solveZ -- served by the real OpenCV of this image (tests/ref_host.py).

Frozen outputs of that library (tests/golden/ref_fusion.npz: RunFusion, made on CPU by make_ref_fusion_golden.py;
tests/golden/ref_program_priors.npz: the planar prior ProcessProblem uploaded, with the state it was built from, captured
on a B200 by tests/tools/reference_program.py --golden) are checked everywhere; the live comparisons run where the
library is present (it travels to the GPU box with the snapshot; it can only be BUILT where /root/reference is mounted)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

import ref_host
from mpmvs_b200 import capi, io_formats

GOLD = os.path.join(ROOT, "tests", "golden")
MAIN = os.path.join(ROOT, "mp-mvs_b200", "mpmvs_main")


def build_main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mp-mvs_b200", "csrc")])
    assert os.path.exists(MAIN)


def ply_vertices(path):
    b = open(path, "rb").read()
    return np.frombuffer(b[b.index(b"end_header\n") + 11:], np.uint8)


def golden_fusion_folder(tmp_path, maps=""):
    """Rebuilds the dense folder the golden file was made from; the product host reads decoded sidecars (it has no JPEG
    decoder of OpenCV's): the pixels cv2.imread returns for the stored JPEG bytes. maps="small_": the 64 x 48 maps."""
    import cv2

    z = np.load(os.path.join(GOLD, "ref_fusion.npz"))
    dense = str(tmp_path / ("dense" + maps))
    for sub in ("images", "cams", "MPMVS"):
        os.makedirs(os.path.join(dense, sub), exist_ok=True)
    open(os.path.join(dense, "pair.txt"), "wb").write(z["pair.txt"].tobytes())
    for i in range(int(z["n"])):
        jpg = os.path.join(dense, "images", f"{i:08d}.jpg")
        open(jpg, "wb").write(z[f"jpg{i}"].tobytes())
        open(os.path.join(dense, "cams", f"{i:08d}_cam.txt"), "wb").write(z[f"cam{i}"].tobytes())
        grey, bgr = cv2.imread(jpg, cv2.IMREAD_GRAYSCALE), cv2.imread(jpg, cv2.IMREAD_COLOR)
        PKG.synth.write_pgm(os.path.join(dense, "images", f"{i:08d}.pgm"), grey)
        with open(os.path.join(dense, "images", f"{i:08d}.ppm"), "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (bgr.shape[1], bgr.shape[0]))
            f.write(np.ascontiguousarray(bgr[:, :, ::-1]).tobytes())
        d = os.path.join(dense, "MPMVS", f"2333_{i:08d}")
        os.makedirs(d, exist_ok=True)
        io_formats.write_dmb(os.path.join(d, "depths.dmb"), z[f"{maps}depth{i}"])
        io_formats.write_dmb(os.path.join(d, "normals.dmb"), z[f"{maps}normal{i}"])
    return z, dense


@pytest.mark.parametrize("dyn", [0, 1])
def test_host_fusion_reproduces_the_references_ply(tmp_path, dyn):
    """RunFusion of the product's host program (mpmvs_main.cpp) on the golden folder: every vertex record of MPMVS_model.ply
    -- position, normal, colour, order -- byte-identical to what the reference's own RunFusion wrote (PatchMatch.cpp:287-504;
    noisy normals and a colour JPEG, so the cv::Vec3f `/=` of the averaged normal and the B, G, R averages are exercised)."""
    build_main()
    z, dense = golden_fusion_folder(tmp_path)
    yaml = ref_host.write_project(str(tmp_path), dense, **{"Use dynamic_consistency to fuse": dyn, "Max source images num": 3})
    r = subprocess.run([MAIN, yaml, "--fusion-only"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = ply_vertices(os.path.join(dense, "MPMVS", "MPMVS_model.ply"))
    want = z[f"ply_dyn{dyn}"]
    assert len(want) // 27 > 1000
    np.testing.assert_array_equal(got, want)


def test_host_fusion_with_smaller_maps_reproduces_the_references_ply(tmp_path):
    """Depth maps smaller than the images (what `Max image size` produces): RunFusion resizes the colour image and scales K
    (RescaleImageAndCamera, PatchMatch.cpp:264-285). The host's resizeLinearBGR is cv::resize's 8-bit fixed-point path, so the
    .ply is byte-identical to the reference's with OpenCV's own resize (Intel IPP, where OpenCV has it, differs by a grey level)."""
    build_main()
    z, dense = golden_fusion_folder(tmp_path, "small_")
    yaml = ref_host.write_project(str(tmp_path), dense, **{"Use dynamic_consistency to fuse": 1, "Max source images num": 3})
    r = subprocess.run([MAIN, yaml, "--fusion-only"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = ply_vertices(os.path.join(dense, "MPMVS", "MPMVS_model.ply"))
    assert len(z["ply_small_maps_dyn1"]) // 27 > 500
    np.testing.assert_array_equal(got, z["ply_small_maps_dyn1"])


@pytest.mark.skipif(not ref_host.available(), reason="oracle/_ref/libmpmvs_ref_host.so not built here")
def test_golden_fusion_is_what_the_reference_writes_today(tmp_path):
    """The frozen file against the live library: the reference's RunFusion on the rebuilt folder."""
    z, dense = golden_fusion_folder(tmp_path)
    for dyn in (0, 1):
        ref_host.write_project(str(tmp_path), dense, **{"Use dynamic_consistency to fuse": dyn, "Max source images num": 3})
        ref_host.run_fusion(str(tmp_path))
        np.testing.assert_array_equal(ply_vertices(os.path.join(dense, "MPMVS", "MPMVS_model.ply")), z[f"ply_dyn{dyn}"])
    z, dense = golden_fusion_folder(tmp_path, "small_")
    ref_host.write_project(str(tmp_path), dense, **{"Use dynamic_consistency to fuse": 1, "Max source images num": 3})
    ref_host.run_fusion(str(tmp_path), ipp=False)
    np.testing.assert_array_equal(ply_vertices(os.path.join(dense, "MPMVS", "MPMVS_model.ply")), z["ply_small_maps_dyn1"])


def prior_cases():
    z = np.load(os.path.join(GOLD, "ref_program_priors.npz"))
    out = {}
    for name in ("planar", "geom_planar"):
        w, h = (int(v) for v in z[f"{name}/size"])
        c = {"w": w, "h": h, "cams": z[f"{name}/cams"], "planes": z[f"{name}/planes"].reshape(h, w, 4), "mask": z[f"{name}/mask"].reshape(h, w),
             "in_planes": z[f"{name}/in_planes"].reshape(h, w, 4), "in_costs": z[f"{name}/in_costs"].reshape(h, w),
             "in_geom": z[f"{name}/in_geom"].reshape(h, w) if f"{name}/in_geom" in z.files else None}
        out[name] = c
    return out


@pytest.mark.parametrize("name", ["planar", "geom_planar"])
def test_prior_restatement_reproduces_the_reference_programs_prior(name):
    """oracle/prior_oracle.py (vertex picking, cv2.Subdiv2D, rasterisation, cv2.SVDecomp plane fit, range check) on the state
    the reference's ProcessProblem built its prior from: the triangle-id mask and every prior plane it uploaded
    (PatchMatch.cpp:978-996), bit for bit -- both vertex-picking variants. This is what pins the restatement."""
    import prior_oracle

    c = prior_cases()[name]
    cam = c["cams"][0]
    K = np.asarray(cam["K"], np.float32).reshape(3, 3)
    prior, mask, verts, tris = prior_oracle.build_prior(c["in_planes"], c["in_costs"], K, float(np.float32(cam["depth_min"]) * np.float32(0.6)),
                                                        float(np.float32(cam["depth_max"]) * np.float32(1.2)), c["in_geom"])
    assert (c["mask"] > 0).mean() > 0.5
    np.testing.assert_array_equal(mask, c["mask"])
    on = mask > 0
    np.testing.assert_array_equal(prior[on], c["planes"][on])


@pytest.mark.parametrize("name", ["planar", "geom_planar"])
def test_c_restatement_reproduces_the_reference_programs_prior(name):
    """The same for the C loops that whole-scene comparisons and bench.py's reference arm run (oracle/pm_oracle.c through
    prior_oracle.build_prior_fast): with OpenCV's SVD per triangle (plane_fit="svd", what tests/tools/scene_parity.py uses) the
    reference's prior bit for bit; with the closed-form plane (what the bench arm times) the same triangle ids and planes
    within the SVD's float32 noise."""
    import prior_oracle

    c = prior_cases()[name]
    cam = c["cams"][0]
    K = np.asarray(cam["K"], np.float32).reshape(3, 3)
    dmin, dmax = float(np.float32(cam["depth_min"]) * np.float32(0.6)), float(np.float32(cam["depth_max"]) * np.float32(1.2))
    on = c["mask"] > 0
    prior, mask, _, _, cnt = prior_oracle.build_prior_fast(c["in_planes"], c["in_costs"], K, dmin, dmax, c["in_geom"], plane_fit="svd")
    np.testing.assert_array_equal(mask, c["mask"])
    np.testing.assert_array_equal(prior[on], c["planes"][on])
    assert cnt == int(on.sum())
    prior, mask, _, _, _ = prior_oracle.build_prior_fast(c["in_planes"], c["in_costs"], K, dmin, dmax, c["in_geom"], plane_fit="closed")
    assert (mask == c["mask"]).mean() > 0.9995
    both = (mask == c["mask"]) & on
    assert np.abs(prior[both] - c["planes"][both]).max() < 2e-3


@pytest.mark.parametrize("name", ["planar", "geom_planar"])
def test_product_triangulation_on_the_reference_programs_state(name):
    """Host part of the product's prior stage on the same state: the restated vertex picking feeds mpmvs_delaunay; every
    triangle id the reference rasterised must be one of its list (same order => same ids)."""
    import prior_oracle

    c = prior_cases()[name]
    verts = np.array(prior_oracle.triangulate_vertices(c["in_costs"], c["in_geom"]), np.int32)
    tris = capi.delaunay(verts, c["w"], c["h"])
    want = prior_oracle.delaunay_cv(verts, c["w"], c["h"])
    np.testing.assert_array_equal(verts[tris], want)
    assert int(c["mask"].max()) <= len(tris)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["planar", "geom_planar"])
def test_gpu_prior_stage_against_the_reference_programs_prior(name):
    """The product's whole prior stage (vertex picking and rasterisation kernels, pm_subdiv.h, closed-form plane fit, range
    check) on the state the reference built its prior from: the same triangle id on every pixel but range-check ties, and
    planes within float32-SVD noise of the ones the reference uploaded (mean 2..5e-6, up to a few 1e-4 on the thinnest
    triangles: closed-form plane through three points in double against cv::SVD::solveZ in float32, which is LAPACK's sgesdd
    in this OpenCV build and OpenCV's own Jacobi SVD in a build without LAPACK -- bits no restatement can follow)."""
    c = prior_cases()[name]
    w, h = c["w"], c["h"]
    sc = PKG.synth.make_dtu_scene(width=w, height=h, grid=3, n_src=4, seed=2, jpeg=False)
    ids, imgs, _ = sc.problem(0, 4)
    cams = np.repeat(c["cams"], len(imgs), 0)          # the reference camera is what the stage reads
    pm = capi.PatchMatch(0).set_problem(imgs, cams)
    with_geom = c["in_geom"] is not None
    pm.set_geom_consistency_params(with_geom, with_geom)
    pm.set_dev_state({"planes": c["in_planes"], "costs": c["in_costs"], "views": None, "rng": None, "geom": c["in_geom"]})
    verts = pm.pick_vertices(with_geom)
    tris = capi.delaunay(verts, w, h)
    pm.prior_from_triangles(verts, tris)
    prior, mask = pm.get_prior()
    pm.destroy()
    same = mask == c["mask"]
    print(name, "pixels with the reference's triangle id:", float(same.mean()), "prior pixels ours / reference:", float((mask > 0).mean()), float((c["mask"] > 0).mean()))
    assert same.mean() > 0.9995
    both = same & (mask > 0)
    d = np.abs(prior[both] - c["planes"][both])
    print(name, "plane difference: max", float(d.max()), "mean", float(d.mean()))
    assert d.mean() < 2e-5 and d.max() < 2e-3          # float32-SVD noise; the maximum sits on the thinnest triangles


@pytest.mark.gpu
@pytest.mark.skipif(not ref_host.available(), reason="oracle/_ref/libmpmvs_ref_host.so not built on this box")
def test_whole_program_against_the_references_main():
    """The reference's main() and the product's hosts -- mpmvs_main (C++) and run.py (Python, --order gauss_seidel) -- on the same
    dense folder, same seeds (9 views, 320x240):
    * photometric + 2 geometric passes + fusion; a pair.txt that exercises GenerateSampleList's rules; images above `Max image
      size`: every depths / normals / costs .dmb and MPMVS_model.ply BYTE-IDENTICAL;
    * the two schedules with a planar prior: the product's prior has the reference's triangle id on every pixel and its planes
      within float32-SVD noise; the depth maps are then no longer bit-identical (that noise, amplified by the propagation)
      but agree on > 90 % of the pixels at 1 % depth / 5 degrees (measured 97.8 / 94.8 % median), with the same accuracy."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "reference_program.py"), "--schedules", "photo_geom,quirks,resized,planar,geom_planar", "--python-host"],
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])["schedules"]
    print(json.dumps(res))
    n = res["photo_geom"]["images"]
    # quirks: a pair.txt with zero scores, too many sources, shuffled order, the >= 2 views fusion rule; resized: `Max image size`
    # below the image size (PatchMatchInit's cv::resize + K scaling, RunFusion's colour resize; OpenCV's own resize, IPP off)
    for name in ("photo_geom", "quirks", "resized"):
        a = res[name]
        assert a["depth_maps_byte_identical"] == n and a["normal_maps_byte_identical"] == n and a["cost_maps_byte_identical"] == n and a["ply_byte_identical"], a
        b = a["python_host"]       # mp-mvs_b200/run.py --order gauss_seidel --fusion 2: the Python host writes the same files
        assert b["depth_maps_byte_identical"] == n and b["normal_maps_byte_identical"] == n and b["cost_maps_byte_identical"] == n and b["ply_byte_identical"], b
    for name in ("planar", "geom_planar"):
        b = res[name]
        assert b["reference_priors_captured"] == n
        for p in b["prior_stage"]:
            assert p["same_triangle_id"] > 0.999 and p["mean_abs_plane_diff_same_id"] < 2e-5 and p["max_abs_plane_diff_same_id"] < 2e-3, p
        assert b["agreement_median_min"][0] > 0.9, b
        assert abs(b["accuracy_2cm_reference_ours"][0] - b["accuracy_2cm_reference_ours"][1]) < 0.5, b
