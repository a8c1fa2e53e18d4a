"""Whole-scene run of the sharded pipeline (BASELINE.json config 4 shape: a Tanks-and-Temples-like ring of 1920x1080 views,
10 sources each, photometric + geometric-consistency stages with the depth maps exchanged over NVLink between stages).

    python tools/scene_bench.py --views 96
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/scene_bench.py --views 96

Strong scaling: the scene is fixed, reference images are sharded over the ranks. Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkgload  # noqa: E402

pkgload.load_package()
from mpmvs_b200 import io_formats, pipeline, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=96)
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--geom-iters", type=int, default=2)
ap.add_argument("--planar", type=int, default=0)
ap.add_argument("--geom-planar", type=int, default=0)
ap.add_argument("--in-flight", type=int, default=8)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = (int(a) for a in args.size.split("x"))
cache = f"/dev/shm/mpmvs_scene_tnt_{args.views}_{W}x{H}_{os.getuid()}.npz"
if rank == 0 and not os.path.exists(cache):
    t = time.time()
    sc = synth.make_tnt_scene(width=W, height=H, n_views=args.views, n_src=10, workers=min(32, os.cpu_count() or 1))
    np.savez(cache + ".tmp.npz", images=np.stack(sc.images), cams=io_formats.pack_cameras(sc.cams),
             pairs=np.array([[j for j, _ in sc.pairs[i]] for i in range(sc.num_views)]), gt=np.stack(sc.gt_depth).astype(np.float16))
    os.replace(cache + ".tmp.npz", cache)
    print(f"[scene] rendered {args.views} views {W}x{H} in {time.time() - t:.1f} s on {os.cpu_count()} cores", file=sys.stderr, flush=True)
if dist is not None:
    dist.barrier()
z = np.load(cache)
cams_packed = z["cams"]
cams = {}
for i in range(len(cams_packed)):
    r = cams_packed[i]
    cams[i] = io_formats.Camera(K=r["K"].reshape(3, 3), R=r["R"].reshape(3, 3), t=r["t"], height=int(r["height"]), width=int(r["width"]),
                                depth_min=float(r["depth_min"]), depth_max=float(r["depth_max"]))
images = {i: z["images"][i] for i in range(len(cams_packed))}
entries = [io_formats.SceneEntry(ref_id=i, src_ids=[i] + [int(j) for j in z["pairs"][i]], estimate=True) for i in range(len(cams_packed))]
cfg = pipeline.PipelineConfig(geom_iterations=args.geom_iters, max_src=10, seed=9, planar_prior=bool(args.planar),
                              geom_planar_prior=bool(args.geom_planar), in_flight=args.in_flight)
p = pipeline.DensePipeline(entries, cams, images, cfg, rank=rank, world=world, device=local, dist=dist)
t = time.time()
n_cached = p.setup()
torch.cuda.synchronize()
setup_s = time.time() - t
if dist is not None:     # NCCL communicator set-up outside the timed region
    w = torch.zeros(1024, device="cuda")
    o = torch.zeros(1024 * world, device="cuda")
    dist.all_gather_into_tensor(o, w)
    torch.cuda.synchronize()
    dist.barrier()
t = time.time()
stats = p.run()
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
dt = time.time() - t
res = p.results()
gt = z["gt"]
acc = [synth.accuracy_at(res[r_][0][..., 3], gt[r_].astype(np.float32))[2] for r_ in list(res)[:4]]
tt = torch.tensor([dt] + [s.device_ms for s in stats] + [s.exchange_ms for s in stats], dtype=torch.float64, device="cuda")
if dist is not None:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
tt = tt.tolist()
if rank == 0:
    n = len(stats)
    out = {"scene": f"tnt-shaped ring, {args.views} views {W}x{H}, 10 src", "n_gpus": world, "refs": len(entries), "refs_per_rank": p.block,
           "views_cached_rank0": n_cached, "setup_s_rank0": round(setup_s, 2), "total_s": round(tt[0], 3),
           "mpix_per_s": round(len(entries) * W * H / 1e6 / tt[0], 3), "schedule": [s.name for s in stats],
           "pass_ms_max": [round(v, 1) for v in tt[1:1 + n]], "exchange_ms_max": [round(v, 2) for v in tt[1 + n:]],
           "exchange_bytes_per_stage": len(entries) * W * H * 4, "accuracy_10cm_first_refs_rank0": [round(a, 2) for a in acc],
           "planar": args.planar, "geom_planar": args.geom_planar, "in_flight": args.in_flight}
    print(json.dumps(out), flush=True)
    os.makedirs(args.out, exist_ok=True)
    json.dump(out, open(os.path.join(args.out, f"scene_tnt_{args.views}_w{world}.json"), "w"), indent=1)
p.destroy()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
