"""Find where two builds of the library first diverge: step both through init + half-sweeps from identical states.

usage: python tests/tests/tools/variant_diverge.py [variantA] [variantB] [case] [mode: photo|geom]
Each build runs in a child process and saves the state after every half-sweep (always restarted from variant A's previous
state, so a difference is local to one half-sweep); the parent reports the differing pixels.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = "/tmp/mpmvs_diverge"      # states can be hundreds of MB: not under gpurun_out/


def schedule(mode):
    if mode == "geom":
        return [(red, it, 0) for it in range(3) for red in (0, 1)]
    return [(red, it, sc) for sc in (2, 1, 0) for it in range(3) for red in (0, 1)]


def child(case_name, mode, tag, follow):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
    from pkgload import load_package

    load_package()
    from cases import make_case, src_depths, world_state_from_gt
    from mpmvs_b200 import capi

    if case_name.startswith("eth:"):          # eth:W:H:nsrc -- a larger weak-texture room than the test cases
        from conftest import problem_arrays
        from mpmvs_b200 import synth

        w, h, ns = (int(t) for t in case_name.split(":")[1:])
        sc = synth.make_eth3d_scene(width=w, height=h, n_views=ns + 1, n_src=ns, seed=3, jpeg=False)
        ids, imgs, cams = problem_arrays(sc, 2)
        c = dict(scene=sc, ref=2, ids=ids, images=imgs, cams=cams)
    else:
        c = make_case(case_name)
    pm = capi.PatchMatch(0)
    pm.set_profiling(2)
    pm.set_problem(c["images"], c["cams"])
    if mode == "geom":
        pm.set_geom_consistency_params(True, False)
        pm.set_src_depths(src_depths(c, 0.002))
        pm.set_state(*world_state_from_gt(c))
    else:
        pm.set_geom_consistency_params(False, False)
    pm.init_only(77)
    states = {"init": pm.get_state()}
    lead = np.load(follow, allow_pickle=True)["states"].item() if follow else None
    prev = "init"
    for red, it, sc in schedule(mode):
        if lead is not None:
            pm.set_dev_state(lead[prev])
        pm.half_sweep(red, it, sc)
        prev = f"s{sc}i{it}r{red}"
        states[prev] = pm.get_state()
    np.savez(os.path.join(OUT, f"diverge_{tag}.npz"), states=np.array(states, dtype=object))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
        sys.exit(0)
    a = sys.argv[1] if len(sys.argv) > 1 else "default"
    b = sys.argv[2] if len(sys.argv) > 2 else "nocompact"
    case = sys.argv[3] if len(sys.argv) > 3 else "room6"
    mode = sys.argv[4] if len(sys.argv) > 4 else "geom"
    os.makedirs(OUT, exist_ok=True)
    for v, follow in ((a, None), (b, os.path.join(OUT, "diverge_A.npz"))):
        env = dict(os.environ)
        env.pop("MPMVS_LIB_VARIANT", None)
        if v != "default":
            env["MPMVS_LIB_VARIANT"] = v
        tag = "A" if follow is None else "B"
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", case, mode, tag] + ([follow] if follow else []),
                           env=env, capture_output=True, text=True)
        if r.returncode:
            print(v, "FAILED", r.stderr[-2000:])
            sys.exit(1)
    A = np.load(os.path.join(OUT, "diverge_A.npz"), allow_pickle=True)["states"].item()
    B = np.load(os.path.join(OUT, "diverge_B.npz"), allow_pickle=True)["states"].item()
    shown = 0
    for k in A:
        diff = np.zeros(A[k]["costs"].shape, bool)
        for f in ("planes", "costs", "views", "rng", "geom"):
            d = A[k][f] != B[k][f]
            if f in ("planes", "costs", "geom"):
                d &= ~(np.isnan(A[k][f]) & np.isnan(B[k][f]))
            diff |= d.reshape(diff.shape + (-1,)).any(-1)
        print(f"{k:8s} differing pixels {int(diff.sum())} of {diff.size}")
        if diff.any() and shown < 3:
            shown += 1
            ys, xs = np.nonzero(diff)
            for y, x in list(zip(ys, xs))[:8]:
                print(f"   ({x},{y}) A cost {A[k]['costs'][y, x]!r} geom {A[k]['geom'][y, x]!r} plane {A[k]['planes'][y, x]} views {A[k]['views'][y, x]:#x}"
                      f" rng_same {bool((A[k]['rng'][y, x] == B[k]['rng'][y, x]).all())}")
                print(f"   {'':9s} B cost {B[k]['costs'][y, x]!r} geom {B[k]['geom'][y, x]!r} plane {B[k]['planes'][y, x]} views {B[k]['views'][y, x]:#x}")
