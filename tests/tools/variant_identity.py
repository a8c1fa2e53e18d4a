"""Check on the GPU that two builds of the library give bit-identical results (photometric, planar-prior and geometric runs).

usage: python tests/tests/tools/variant_identity.py [variantA] [variantB] [case]    ("default" = the in-tree library)
Each build runs in its own process (the library is chosen at import time by MPMVS_LIB_VARIANT) and prints one digest per
run; the parent compares them.
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(case_name: str) -> None:
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
    import numpy as np
    from pkgload import load_package

    pkg = load_package()
    from cases import make_case
    from mpmvs_b200 import capi
    from cases import prior_planes, src_depths, world_state_from_gt

    c = make_case(case_name)

    def digest(tag, *arrs):
        h = hashlib.sha256()
        for a in arrs:
            h.update(np.ascontiguousarray(a).tobytes())
        print("DIGEST", tag, h.hexdigest(), flush=True)

    for fmt in (capi.TEX_F32, capi.TEX_U8):
        pm = capi.PatchMatch(0)
        pm.set_tex_format(fmt)
        pm.set_problem(c["images"], c["cams"])
        pm.set_geom_consistency_params(False, False)
        pm.run(1234)
        digest(f"photo{fmt}", *pm.result())
        pm.build_prior()                                  # GPU prior stage on the resident state
        pm.set_planar_prior_params()
        pm.set_geom_consistency_params(False, True)
        pm.run(1235)
        digest(f"gprior{fmt}", *pm.result())
        pm.set_prior(*prior_planes(c))                    # host-supplied prior with holes
        pm.run(1236)
        digest(f"hprior{fmt}", *pm.result())
        pm.destroy()
        for planar in (False,):
            pm = capi.PatchMatch(0)
            pm.set_tex_format(fmt)
            pm.set_problem(c["images"], c["cams"])
            pm.set_geom_consistency_params(True, planar)
            pm.set_src_depths(src_depths(c, 0.002))
            pm.set_state(*world_state_from_gt(c))
            if planar:
                pm.set_planar_prior_params()
                pm.set_prior(*prior_planes(c))
            pm.run(1237)
            digest(f"geom{int(planar)}{fmt}", *pm.result(geom=True))
            pm.destroy()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    a = sys.argv[1] if len(sys.argv) > 1 else "default"
    b = sys.argv[2] if len(sys.argv) > 2 else "noeo"
    case = sys.argv[3] if len(sys.argv) > 3 else "room6"
    out = {}
    for v in (a, b):
        env = dict(os.environ)
        env.pop("MPMVS_LIB_VARIANT", None)
        if v != "default":
            env["MPMVS_LIB_VARIANT"] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", case], env=env, capture_output=True, text=True)
        if r.returncode:
            print(v, "FAILED", r.stderr[-2000:])
            sys.exit(1)
        out[v] = [ln.split()[1:] for ln in r.stdout.splitlines() if ln.startswith("DIGEST")]
    ok = True
    for (ta, da), (tb, db) in zip(out[a], out[b]):
        same = da == db
        ok &= same
        print(f"{ta:8s} {a}={da[:16]} {b}={db[:16]} {'identical' if same else 'DIFFERENT'}")
    print("variant identity:", "OK" if ok and out[a] else "MISMATCH")
    sys.exit(0 if ok and out[a] else 1)
