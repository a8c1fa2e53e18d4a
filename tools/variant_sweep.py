"""Run bench.py for every (library variant, view storage format) on this GPU box and print one line per run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload = sys.argv[1] if len(sys.argv) > 1 else "dtu"
variants = sys.argv[2].split(",") if len(sys.argv) > 2 and sys.argv[2] not in ("", "all") else sorted(
    f[len("libmpmvs_b200_"):-3] for f in os.listdir(os.path.join(ROOT, "mp-mvs_b200", "variants")) if f.endswith(".so"))
fmts = sys.argv[3].split(",") if len(sys.argv) > 3 else ["f32", "f16", "u8"]
for v in variants:
    for f in fmts:
        env = dict(os.environ, MPMVS_LIB_VARIANT=v)
        if v == "default" and not os.path.exists(os.path.join(ROOT, "mp-mvs_b200", "variants", "libmpmvs_b200_default.so")):
            env.pop("MPMVS_LIB_VARIANT")          # the in-tree library
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "2", "--warmup", "1",
                            "--no-cpu-baseline", "--tex", f], env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            k = j["kernel_ms_per_step"]
            print(f"{v:10s} {f:4s} ms/step {j['ms_per_step']:9.3f} sweeps {k['sweeps']:9.3f} init {k['init']:7.3f} value {j['value']:8.3f} e2e {j['e2e']['value']:8.3f} "
                  f"frac {j['roofline']['frac']:.3f} cost {j['checksum_mean_cost']:.5f} acc {j['accuracy_2_5_10cm']}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{v:10s} {f:4s} FAILED rc={r.returncode} {e} {r.stderr[-300:]}", flush=True)
