#!/usr/bin/env python
"""bench.py -- depth-map throughput of the PatchMatch hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload eth3d|dtu|plane]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): the configuration the metric is quoted on -- synthetic ETH3D-shaped indoor scene,
3200x2130, 1 reference + 10 source views. One STEP = the complete PatchMatchCUDA::Run() of one reference image
(/root/reference/src/PatchMatch.cu:1188-1254: InitializeScore, 3 scales x 3 iterations x {black, red} half-sweeps,
GetDepthandNormal, black/red median filter) = 22 kernel launches; metric = W*H*steps / seconds, summed over ranks.

  value  : device-resident throughput (views already in the per-GPU layered texture, results left in HBM), CUDA events.
  e2e    : the same step through the reference-facing C ABI with HOST buffers: mpmvs_set_views (H2D of the 11 float
           images from pinned memory) + mpmvs_run_into (22 launches + D2H of planes/costs) inside the timed region.
  roofline : dominant kernel = pm_sweep_kernel (18 of 22 launches). Bound = SM L1/TEX path (not HBM,
           not tensor cores): algorithmic cost 16 B of L1/TEX traffic per executed tap; achieved = executed taps of the
           18 sweep launches x 16 B / their summed device time (CUDA events between launches on the launching stream);
           peak = 4 bilinear fetches/clk/SM x 148 SMs x sm_max_mhz x 16 B (nominal TMU rate at the MEASURED max SM
           clock of MEASURED_PEAKS.json). hbm_frac is reported beside it from the same timing.
  cpu_baseline : the plain-C oracle port (oracle/pm_oracle.c, all host cores) on a centre crop of the same views.
  --impl reference : the reference's own CUDA path (oracle/_ref/libmpmvs_ref.so = /root/reference/src/PatchMatch.cu
           compiled in place for sm_100) on the same workload; the reference has no CPU implementation of this path.
           Rank 0 only (the reference is single-GPU, src/PatchMatch.cpp:509).

Multi-GPU: reference images are independent within a pass; each rank runs its own reference images,
no data-path collective in the photometric pass -> "scaling": "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import pkgload  # noqa: E402

pkgload.load_package()
from mpmvs_b200 import io_formats, synth  # noqa: E402

TAPS_PER_NCC = 36
BYTES_PER_TAP = 16          # one bilinear source sample = 4 texels x 4 B of L1/TEX traffic
TEX_PER_CLK_SM = 4
N_SM = 148


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------- workload
def make_workload(name: str, workers: int):
    """Returns (scene, ref view, label). Rendered once per box and cached in /dev/shm for the other ranks."""
    if name == "eth3d":
        shape, ref, label = (3200, 2130), 5, "eth3d-shaped 3200x2130, 1 ref + 10 src views, photometric Run()"
        mk = lambda: synth.make_eth3d_scene(width=3200, height=2130, n_views=11, n_src=10, workers=workers)  # noqa: E731
    elif name == "dtu":
        shape, ref, label = (1600, 1200), 24, "dtu-shaped 1600x1200, 1 ref + 10 src views, photometric Run()"

        def mk():
            probe = synth.make_dtu_scene(views=[])
            ids = [24] + [i for i, _ in probe.pairs[24]][:10]
            return synth.make_dtu_scene(views=ids, workers=workers)
    elif name == "plane":
        shape, ref, label = (640, 480), 1, "textured plane 640x480, 1 ref + 2 src views, photometric Run()"
        mk = lambda: synth.make_plane_scene()  # noqa: E731
    else:
        raise SystemExit(f"unknown workload {name}")
    return mk, ref, label, shape


def load_problem(name: str, rank: int, world: int, barrier):
    mk, ref, label, shape = make_workload(name, workers=min(16, os.cpu_count() or 1))
    cache = f"/dev/shm/mpmvs_bench_{name}_{os.getuid()}.npz"
    if rank == 0 and not os.path.exists(cache):
        t = time.time()
        scene = mk()
        ids, imgs, cams = scene.problem(ref, 10)
        np.savez(cache + ".tmp.npz", images=np.stack(imgs).astype(np.uint8), cams=io_formats.pack_cameras(cams), ids=np.array(ids),
                 gt_depth=scene.gt_depth[ref], gt_normal=scene.gt_normal[ref])
        os.replace(cache + ".tmp.npz", cache)
        log(f"[bench] rendered {label} in {time.time() - t:.1f} s")
    barrier()
    z = np.load(cache)
    imgs = [np.ascontiguousarray(i, dtype=np.float32) for i in z["images"]]
    return dict(images=imgs, cams=z["cams"], ids=z["ids"], gt_depth=z["gt_depth"], gt_normal=z["gt_normal"], label=label,
                width=shape[0], height=shape[1])


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------- arms
def run_ours(args, rank, world, local_rank, prob, barrier, allmax):
    import torch
    from mpmvs_b200 import capi

    torch.cuda.set_device(local_rank)
    W, H, n = prob["width"], prob["height"], len(prob["images"])
    fmt = {"f32": capi.TEX_F32, "f16": capi.TEX_F16, "u8": capi.TEX_U8}[args.tex]
    pm = capi.PatchMatch(device=local_rank).set_tex_format(fmt)
    # resident arm: views uploaded once (the per-GPU image cache), runs leave results in HBM
    pm.set_problem(prob["images"], prob["cams"])
    pm.set_geom_consistency_params(False, False)
    pm.synchronize()
    # pinned host buffers for the e2e arm
    # host images for the e2e arm: uint8 grey levels as decoded from the JPEGs when the storage is 8-bit, else float32
    host_imgs = [i.astype(np.uint8) for i in prob["images"]] if args.tex == "u8" else prob["images"]
    pin_imgs = [torch.from_numpy(i).pin_memory() for i in host_imgs]
    h_planes = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    h_costs = torch.empty((H, W), dtype=torch.float32).pin_memory()
    pin_np = [t.numpy() for t in pin_imgs]

    seed = 1000 * (rank + 1)
    for i in range(args.warmup):
        pm.run_async(seed + i)
    pm.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region 1: device-resident throughput, per-launch events on the launching stream
    pm.set_profiling(1)
    barrier(); torch.cuda.synchronize()
    tot_ms = sweep_ms = init_ms = fin_ms = 0.0
    n_sweeps = launches = 0
    t0 = time.time()
    for i in range(args.steps):
        pm.run_async(seed + 100 + i)
        pm.synchronize()
        tot_ms += pm.last_run_ms()
        pr = pm.last_run_profile()
        sweep_ms += pr["sweep_ms"]; init_ms += pr["init_ms"]; fin_ms += pr["finalize_ms"]; n_sweeps += pr["n_sweeps"]
        launches += pm.last_run_launches()
    torch.cuda.synchronize(); barrier()
    wall_resident = time.time() - t0
    clocks = sampler.stop()
    pm.set_profiling(0)
    tot_ms = allmax(tot_ms)

    # ---- timed region 2: end to end through the C ABI with host buffers
    pm2 = capi.PatchMatch(device=local_rank).set_tex_format(fmt)
    pm2.set_geom_consistency_params(False, False)
    for i in range(max(1, min(args.warmup, 2))):
        pm2.set_problem(pin_np, prob["cams"])
        pm2.run_into(seed + i, h_planes.numpy(), h_costs.numpy())
    barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    for i in range(args.steps):
        pm2.set_problem(pin_np, prob["cams"])
        pm2.run_into(seed + 200 + i, h_planes.numpy(), h_costs.numpy())
    torch.cuda.synchronize()
    e2e_s = allmax(time.time() - t0)   # host wall clock: includes H2D, launches, D2H and the final synchronisation
    barrier()
    checksum = float(h_costs.double().mean())
    acc = synth.accuracy_at(h_planes.numpy()[..., 3], prob["gt_depth"])

    # ---- untimed: count executed NCC evaluations of one run (roofline numerator)
    pm.set_profiling(2)
    pm.run_async(seed + 100)
    pm.synchronize()
    ncc_exec = pm.last_run_profile(timing=False, count=True)["ncc_evaluations"]
    pm.set_profiling(0)

    mpix = W * H / 1e6
    value = world * args.steps * mpix / (tot_ms / 1e3)
    e2e = world * args.steps * mpix / e2e_s
    peaks = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    # executed taps of the sweep launches: total minus the init launch's (N-1) evaluations per pixel
    nsrc = n - 1
    taps_sweeps = max(0, ncc_exec - W * H * nsrc) * TAPS_PER_NCC
    sweep_s = sweep_ms / 1e3 / args.steps
    achieved = taps_sweeps * BYTES_PER_TAP / sweep_s / 1e9
    peak = TEX_PER_CLK_SM * N_SM * sm_max * 1e6 * BYTES_PER_TAP / 1e9
    ref_equiv_taps = W * H * nsrc * (1 + 14 * 9) * TAPS_PER_NCC
    out = {
        "metric": "depth-map Mpix/s (3200x2130, 10 src views)" if args.workload == "eth3d" else f"depth-map Mpix/s ({W}x{H}, {nsrc} src views)",
        "value": round(value, 4), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(tot_ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": prob["label"], "width": W, "height": H, "src_views": nsrc, "passes": "photometric (scales 2,1,0 x 3 iterations)",
                   "view_storage": args.tex, "lib_variant": os.environ.get("MPMVS_LIB_VARIANT", "default"),
                   "l2": "inputs_larger_than_l2 (11 float views = %.0f MB + %.0f MB state vs 126 MB L2)" % (n * W * H * 4 / 1e6, W * H * 56 / 1e6),
                   "parallelism": f"refs sharded over {world} gpu(s), no data-path collective in this pass"},
        "e2e": {"value": round(e2e, 4), "unit": "Mpix/s", "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in pin_imgs) + 112 * n),
                "d2h_bytes_per_step": int(W * H * 20), "ms_per_step": round(e2e_s * 1e3 / args.steps, 3)},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {"init": round(init_ms / args.steps, 3), "sweeps": round(sweep_ms / args.steps, 3),
                               "finalize": round(fin_ms / args.steps, 3), "n_sweep_launches": n_sweeps // max(1, args.steps)},
        "roofline": {"bound": "l1tex", "kernel": "pm_sweep_kernel", "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None,
                     "peak_source": "4 bilinear/clk/SM x 148 SM x measured sm_max_mhz x 16 B",
                     "executed_taps_per_step": int(taps_sweeps), "reference_equivalent_taps_per_step": int(ref_equiv_taps),
                     "gtaps_per_s": round(taps_sweeps / sweep_s / 1e9, 2),
                     "hbm_peak_gbs": peaks.get("hbm_gbs")},
        "clocks": clocks,
        "checksum_mean_cost": round(checksum, 6), "accuracy_2_5_10cm": [round(a, 3) for a in acc],
        "wall_s_resident": round(wall_resident, 3),
    }
    pm.destroy(); pm2.destroy()
    return out


def cpu_baseline(prob, budget_s=20.0):
    """The plain-C oracle port on a centre crop of the same views (principal point shifted), all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    W, H = prob["width"], prob["height"]
    cw, ch = min(W, 256), min(H, 160)
    x0, y0 = (W - cw) // 2, (H - ch) // 2
    # keep the sources whole (their warped windows must stay inside), crop only the reference: the oracle needs all
    # views at the size its camera says, so crop every view with the same window instead and shift cx, cy
    imgs = [np.ascontiguousarray(i[y0:y0 + ch, x0:x0 + cw]) for i in prob["images"]]
    cams = prob["cams"].copy()
    for c in cams:
        c["K"][2] -= x0
        c["K"][5] -= y0
        c["width"], c["height"] = cw, ch
    o = oracle_py.Oracle("cpu").set_problem(imgs, cams)
    o.set_geom_consistency_params(False, False)
    cores = oracle_py.Oracle("cpu").lib.pmo_get_threads()
    t = time.time()
    o.run(1)
    dt = time.time() - t
    o.destroy()
    return {"value": round(cw * ch / 1e6 / dt, 6), "unit": "Mpix/s", "cores": int(cores), "kind": "port",
            "sample": f"centre crop {cw}x{ch} of all {len(imgs)} views, one full photometric Run() in {dt:.1f} s (oracle/pm_oracle.c, pthreads)"}


def run_reference(args, prob):
    """The reference's own CUDA path (oracle/_ref) on GPU 0, same workload, same step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    if not oracle_py.available("ref"):
        return {"impl": "reference", "unavailable": "oracle/_ref/libmpmvs_ref.so not built (needs /root/reference at build time)"}
    W, H, n = prob["width"], prob["height"], len(prob["images"])
    R = oracle_py.Oracle("ref")
    sampler = ClockSampler(0)
    R.set_problem(prob["images"], prob["cams"])
    R.set_geom_consistency_params(False, False)
    for i in range(args.warmup):
        R.run(10 + i)
    sampler.start()
    ms = 0.0
    for i in range(args.steps):
        ms += R.run(100 + i)           # CUDA events around PatchMatchCUDA::Run(), which ends with its blocking D2H copies
    clocks = sampler.stop()
    # e2e: what ProcessProblem does per reference image (minus file I/O): allocate + upload + Run + Release
    R.destroy()
    t0 = time.time()
    for i in range(args.steps):
        R = oracle_py.Oracle("ref")
        R.set_problem(prob["images"], prob["cams"])
        R.set_geom_consistency_params(False, False)
        R.run(200 + i)
        R.result()
        R.destroy()
    e2e_s = time.time() - t0
    mpix = W * H / 1e6
    v = args.steps * mpix / (ms / 1e3)
    return {
        "impl": "reference", "metric": "depth-map Mpix/s (3200x2130, 10 src views)" if args.workload == "eth3d" else f"depth-map Mpix/s ({W}x{H}, {n-1} src views)",
        "value": round(v, 4), "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": prob["label"], "width": W, "height": H, "src_views": n - 1,
                                        "passes": "photometric (scales 2,1,0 x 3 iterations)"},
        "cpu_baseline": {"value": round(v, 4), "unit": "Mpix/s", "cores": 0, "kind": "reference",
                         "sample": "the reference has no CPU path: its own CUDA kernels (PatchMatch.cu rebuilt for sm_100, -O3 --use_fast_math "
                                   "--maxrregcount=128) on one B200, full workload"},
        "e2e": {"value": round(args.steps * mpix / e2e_s, 4), "unit": "Mpix/s", "h2d_bytes_per_step": int(n * W * H * 4 + 112 * n),
                "d2h_bytes_per_step": int(W * H * 20), "ms_per_step": round(e2e_s * 1e3 / args.steps, 3)},
        "gpu_launches": 22 * args.steps, "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="eth3d")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tex", default="u8", choices=["f32", "f16", "u8"], help="storage format of the views in HBM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        if args.impl == "reference" and rank != 0:
            return 0
        if args.impl == "ours":
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prob = load_problem(args.workload, rank if args.impl == "ours" else 0, world, barrier)
    if args.impl == "reference":
        out = run_reference(args, prob)
        print(json.dumps(out), flush=True)
        return 0
    out = run_ours(args, rank, world, local_rank, prob, barrier, allmax)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(prob)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
