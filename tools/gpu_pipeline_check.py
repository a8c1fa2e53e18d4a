"""Sharded pipeline on real GPUs. Run plain (1 GPU) or under torchrun (N GPUs); writes per-ref checksums to gpurun_out/.

    python tools/gpu_pipeline_check.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_pipeline_check.py
"""
import hashlib
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

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
dist = None
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

W, H = (int(a) for a in os.environ.get("PIPE_SIZE", "640x480").split("x"))
sc = synth.make_dtu_scene(width=W, height=H, grid=3, n_src=4, seed=2, jpeg=False, workers=8)
entries = [io_formats.SceneEntry(ref_id=i, src_ids=[i] + [j for j, _ in sc.pairs[i]], estimate=True) for i in range(sc.num_views)]
cams = {i: c for i, c in enumerate(sc.cams)}
images = {i: im for i, im in enumerate(sc.images)}
cfg = pipeline.PipelineConfig(geom_iterations=2, max_src=4, seed=5, planar_prior=os.environ.get('PIPE_PLANAR', '0') == '1',
                              geom_planar_prior=os.environ.get('PIPE_GEOMPLANAR', '0') == '1')
tag = f"p{int(cfg.planar_prior)}g{int(cfg.geom_planar_prior)}"
p = pipeline.DensePipeline(entries, cams, images, cfg, rank=rank, world=world, device=local, dist=dist)
t = time.time()
stats = p.run()
torch.cuda.synchronize()
dt = time.time() - t
res = p.results()
out = {}
for ref, (planes, costs) in res.items():
    acc = synth.accuracy_at(planes[..., 3], sc.gt_depth[ref])
    out[str(ref)] = {"sha": hashlib.sha1(planes.tobytes() + costs.tobytes()).hexdigest()[:16], "acc": [round(a, 2) for a in acc],
                     "mean_cost": float(costs.mean())}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"world": world, "rank": rank, "wall_s": dt, "stats": [s.__dict__ for s in stats], "refs": out},
          open(os.path.join(ROOT, "gpurun_out", f"pipeline_{tag}_w{world}_r{rank}.json"), "w"), indent=1)
print(f"rank {rank}/{world}: {len(res)} refs in {dt:.2f} s;", [(s.name, round(s.device_ms, 1), round(s.exchange_ms, 2)) for s in stats], flush=True)
p.destroy()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
