"""Run the C++ host (mp-mvs_b200/mpmvs_main: the reference's main() over the C ABI, strictly sequential like the reference)
on a rendered dense folder and report the stage times it prints: PatchMatch stages (GPU) and fusion (host, one thread).

    python tools/cpp_host_scene.py --scene dtu --scale 1.0 [--planar 1 --geom-planar 1]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkgload  # noqa: E402

pkgload.load_package()
from mpmvs_b200 import io_formats, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="dtu")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--planar", type=int, default=0)
ap.add_argument("--geom-planar", type=int, default=0)
ap.add_argument("--gpu-fusion", type=int, default=0)
ap.add_argument("--extra", default="", help="extra flags for mpmvs_main")
args = ap.parse_args()
workers = min(32, os.cpu_count() or 1)
if args.scene == "dtu":
    sc = synth.make_dtu_scene(width=int(1600 * args.scale), height=int(1200 * args.scale), workers=workers)
else:
    sc = synth.make_eth3d_scene(width=int(3200 * args.scale), height=int(2130 * args.scale), workers=workers)
tmp = tempfile.mkdtemp(prefix="mpmvs_dense_", dir="/dev/shm")
root = os.path.join(tmp, "dense")
synth.write_dense_folder(sc, root)
yaml = os.path.join(tmp, "config.yaml")
io_formats.write_config(yaml, **{"Input-folder": root, "Output-folder": root, "Geometric consistency iterations": 2, "Planer prior": args.planar,
                                 "Geometric consistency planer prior": args.geom_planar, "Max source images num": 10})
t = time.time()
r = subprocess.run([os.path.join(ROOT, "mp-mvs_b200", "mpmvs_main"), yaml, "--tex", "u8", "--seed", "7", "--profile"] + (["--gpu-fusion"] if args.gpu_fusion else []) + args.extra.split(), capture_output=True, text=True)
wall = time.time() - t
if r.returncode != 0:
    print(r.stdout[-2000:], r.stderr[-2000:])
    sys.exit(1)
pm_us = float(re.search(r"cost time is ([0-9.]+) us", r.stdout).group(1))
fu_us = float(re.search(r"fusion time is ([0-9.]+) us", r.stdout).group(1))
npts = int(re.search(r"ply file: (\d+) points", r.stdout).group(1))
disk_us = float(re.search(r"results on disk after ([0-9.]+) us", r.stdout).group(1))
phases = re.search(r"host phases \(s\): (.*)", r.stdout).group(1)
extra_lines = [ln for ln in r.stdout.splitlines() if "resident set-up" in ln or "done after" in ln or "results collected" in ln or "GPU fusion kernels" in ln]
mpix = sc.num_views * sc.width * sc.height / 1e6
import numpy as np  # noqa: E402

acc = np.mean([synth.accuracy_at(io_formats.read_dmb(os.path.join(io_formats.result_dir(root, i), "depths.dmb")), sc.gt_depth[i]) for i in range(sc.num_views)], 0)
out = {"scene": f"{args.scene}-shaped {sc.num_views} views {sc.width}x{sc.height}", "planar": args.planar, "geom_planar": args.geom_planar,
       "patchmatch_stages_s": round(pm_us / 1e6, 3), "results_on_disk_s": round(disk_us / 1e6, 3), "host_phases_s": phases, "log": extra_lines, "flags": args.extra, "patchmatch_mpix_per_s": round(mpix / (pm_us / 1e6), 3),
       "fusion_s": round(fu_us / 1e6, 3), "fusion": "gpu" if args.gpu_fusion else "host, 1 thread", "host_cores": os.cpu_count(), "fused_points": npts,
       "total_wall_s": round(wall, 3), "accuracy_2_5_10cm": [round(float(a), 3) for a in acc],
       "note": "C++ host (mpmvs_main): default = the reference's sequential order; --resident = image cache + resident handles + device depth exchange"}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"cpp_host_{args.scene}_p{args.planar}g{args.geom_planar}{'_gpufusion' if args.gpu_fusion else ''}{args.extra.replace(' ', '').replace('--', '_')}.json"), "w"), indent=1)
subprocess.run(["rm", "-rf", tmp])
