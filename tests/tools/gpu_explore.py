"""Exploratory parity/timing report on a GPU box (prints, asserts nothing). Usage: python tests/tests/tools/gpu_explore.py [quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import PKG, gt_planes_cam, problem_arrays  # noqa: E402
from cases import CASES, SEED, make_case, prior_planes, random_planes, src_depths, world_state_from_gt  # noqa: E402
import oracle_py  # noqa: E402
from mpmvs_b200 import capi, synth  # noqa: E402


def stat(tag, a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    print(f"   {tag}: max {d.max():.3e} mean {d.mean():.3e} p99.9 {np.quantile(d, 0.999):.3e} frac>2e-3 {(d > 2e-3).mean():.5f}")


def cmp_state(tag, sa, sb, colour=None):
    h, w = sa["costs"].shape
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.ones((h, w), bool) if colour is None else (((xx + yy) & 1) == colour)
    samep = np.all(sa["planes"] == sb["planes"], -1)
    closep = np.all(np.abs(sa["planes"] - sb["planes"]) <= 1e-4 * (1 + np.abs(sb["planes"])), -1)
    print(f"   {tag}: planes bit-identical {samep[m].mean():.4f} close {closep[m].mean():.4f} "
          f"|dcost| mean {np.abs(sa['costs'] - sb['costs'])[m].mean():.2e} "
          f"cost<1e-3 {(np.abs(sa['costs'] - sb['costs'])[m] < 1e-3).mean():.4f} "
          f"views same {(sa['views'] == sb['views'])[m].mean():.4f} rng same {np.all(sa['rng'] == sb['rng'], -1)[m].mean():.4f}")


def small_cases():
    for name in CASES:
        print("== case", name)
        c = make_case(name)
        ours = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
        ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        cpu = oracle_py.Oracle("cpu").set_problem(c["images"], c["cams"])
        u = [capi.PatchMatch.uniform_stream(SEED, 3, 4, 64), ref.uniform_stream(SEED, 3, 4, 64), cpu.uniform_stream(SEED, 3, 4, 64)]
        print("   uniform stream ours==ref", np.array_equal(u[0], u[1]), "cpu==ref", np.array_equal(u[2], u[1]))
        gt, rnd = gt_planes_cam(c["scene"], c["ref"]), random_planes(c)
        for s in (0, 1, 2):
            for tag, pl in (("gt", gt), ("rnd", rnd)):
                r = ref.ncc_map(pl, s)
                stat(f"ncc {tag} s{s} ours-ref", ours.ncc_map(pl, s), r)
                stat(f"ncc {tag} s{s} cpu-ref ", cpu.ncc_map(pl, s), r)
        for o in (ours, ref, cpu):
            o.set_geom_consistency_params(False, False)
            o.init_only(SEED)
        sr = ref.get_state()
        cmp_state("init ours-ref", ours.get_state(), sr)
        cmp_state("init cpu-ref ", cpu.get_state(), sr)
        for scale in (2, 0):
            for red in (0, 1):
                ours.set_dev_state(sr); cpu.set_dev_state(sr)
                for o in (ours, ref, cpu):
                    o.half_sweep(red, 0, scale)
                sr = ref.get_state()
                cmp_state(f"sweep s{scale} red{red} ours-ref", ours.get_state(), sr, red)
                cmp_state(f"sweep s{scale} red{red} cpu-ref ", cpu.get_state(), sr, red)
        for o in (ours, ref, cpu):
            o.finalize()
        sr = ref.get_state()
        so = ours.get_state()
        print("   finalize: ours-ref depth maxdiff where init same", np.abs(so["planes"] - sr["planes"]).max())
        # geom
        for o in (ours, ref, cpu):
            o.set_geom_consistency_params(True, False)
            o.set_src_depths(src_depths(c, 0.002))
            o.set_state(*world_state_from_gt(c))
        r = ref.geom_map(rnd)
        stat("geom map ours-ref", ours.geom_map(rnd), r)
        stat("geom map cpu-ref ", cpu.geom_map(rnd), r)
        for o in (ours, ref, cpu):
            o.init_only(SEED + 2)
        sr = ref.get_state()
        cmp_state("ginit ours-ref", ours.get_state(), sr)
        for red in (0, 1):
            ours.set_dev_state(sr); cpu.set_dev_state(sr)
            for o in (ours, ref, cpu):
                o.half_sweep(red, 0, 0)
            sr = ref.get_state()
            so = ours.get_state()
            cmp_state(f"gsweep red{red} ours-ref", so, sr, red)
            cmp_state(f"gsweep red{red} cpu-ref ", cpu.get_state(), sr, red)
            print("      geom cost |d| mean", np.abs(so["geom"] - sr["geom"]).mean())
        ours.destroy(); ref.destroy(); cpu.destroy()
        # planar prior
        ours = capi.PatchMatch(0).set_problem(c["images"], c["cams"])
        ref = oracle_py.Oracle("ref").set_problem(c["images"], c["cams"])
        for o in (ours, ref):
            o.set_geom_consistency_params(False, False)
        ref.run(SEED)
        ours.init_only(SEED)
        ours.set_dev_state(ref.get_state())
        for o in (ours, ref):
            o.set_planar_prior_params()
            o.set_geom_consistency_params(False, True)
            o.set_prior(*prior_planes(c))
            o.init_only(SEED + 1)
        sr = ref.get_state()
        cmp_state("pinit ours-ref", ours.get_state(), sr)
        for red in (0, 1):
            ours.set_dev_state(sr)
            for o in (ours, ref):
                o.half_sweep(red, 0, 0)
            sr = ref.get_state()
            cmp_state(f"psweep red{red} ours-ref", ours.get_state(), sr, red)
        ours.destroy(); ref.destroy()


def full_run(scene, refv, tag, reps=2, geom=False):
    ids, imgs, cams = problem_arrays(scene, refv)
    gt = scene.gt_depth[refv]
    gtn = scene.gt_normal[refv]
    valid = gt > 0
    ours = capi.PatchMatch(0).set_problem(imgs, cams)
    ref = oracle_py.Oracle("ref").set_problem(imgs, cams)
    ours.set_geom_consistency_params(False, False)
    ref.set_geom_consistency_params(False, False)
    res = {}
    for seed in (1, 2):
        t = time.time(); ms_o = ours.run(seed); wall_o = time.time() - t
        res["ours", seed] = ours.result()
        t = time.time(); ms_r = ref.run(seed); wall_r = time.time() - t
        res["ref", seed] = ref.result()
        print(f"   [{tag}] seed {seed}: ours {ms_o:.1f} ms device ({wall_o*1e3:.1f} wall), ref {ms_r:.1f} ms ({wall_r*1e3:.1f} wall), speed-up {ms_r/ms_o:.2f}x,"
              f" {scene.width*scene.height/ms_o/1e3:.2f} vs {scene.width*scene.height/ms_r/1e3:.2f} Mpix/s")
    def agree(a, b, m):
        return synth.depth_normal_agreement(a[0][..., 3], a[0][..., :3], b[0][..., 3], b[0][..., :3], m)
    for k in (("ours", 1), ("ours", 2), ("ref", 1), ("ref", 2)):
        p = res[k]
        print(f"   [{tag}] {k}: acc@2/5/10cm {['%.2f' % a for a in synth.accuracy_at(p[0][..., 3], gt)]} mean cost {p[1].mean():.4f}"
              f" agreement with GT(1%,5deg) {synth.depth_normal_agreement(p[0][..., 3], p[0][..., :3], gt, gtn, valid):.4f}")
    rv = valid & (res["ref", 1][1] < 0.5)
    print(f"   [{tag}] agreement ours1-ref1 {agree(res['ours',1], res['ref',1], rv):.4f}  ref1-ref2 (noise floor) {agree(res['ref',1], res['ref',2], rv):.4f}"
          f"  ours1-ours2 {agree(res['ours',1], res['ours',2], rv):.4f}  valid frac {rv.mean():.3f}")
    ours.destroy(); ref.destroy()


if __name__ == "__main__":
    quick = "quick" in sys.argv
    small_cases()
    print("== full runs")
    full_run(synth.make_plane_scene(), 1, "plane 640x480 n=3")
    sc = synth.make_dtu_scene(views=list(range(17, 32)), workers=16)
    ids = sc.problem(24)[0]
    print("dtu ref 24 sources", ids)
    sc = synth.make_dtu_scene(views=ids, workers=16)
    full_run(sc, 24, "dtu 1600x1200 n=11")
    if not quick:
        sc = synth.make_eth3d_scene(workers=16)
        full_run(sc, 5, "eth3d 3200x2130 n=11")
