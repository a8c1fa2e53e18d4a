#!/usr/bin/env python
"""bench.py -- depth-map throughput of the PatchMatch hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload eth3d|dtu|plane] [--arithmetic exact|fast] [--tex u8|f32|f16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): the configuration the metric is quoted on (BASELINE.json configs[2]) -- synthetic
ETH3D-shaped indoor weak-texture scene, 3200x2130, 1 reference + 10 source views, planar prior + multi-scale windows on.
One STEP = the compute of one ProcessProblem(geom=false, planar=true) call (/root/reference/src/PatchMatch.cpp:506-638),
i.e. what stage 1 of main() does per reference image with `Planer prior: 1`:
    photometric Run()  (PatchMatch.cu:1188-1254: InitializeScore + 3 scales x 3 iterations x {black, red} + finalize, 22 launches)
    planar-prior stage (PatchMatch.cpp:532-609: vertex picking, Delaunay, rasterisation, plane fit, range check)
    planar-prior Run() (InitializeScore + 3 iterations x {black, red} at scale 0 + finalize, 10 launches)
metric = W*H*steps / seconds, summed over ranks.

  value  : device-resident throughput (views already in the per-GPU layered texture, results left in HBM); CUDA events on
           the stream the kernels are launched on bracket the K steps (host triangulation time is inside).
  e2e    : the same step through the reference-facing C ABI with HOST buffers: mpmvs_set_views_u8 (H2D of the 11 grey
           images from pinned memory) + the step + D2H of planes/costs, host wall clock.
  roofline : dominant kernel = pm_sweep_kernel (24 of 32 launches, ~97 % of the device time). Bound = the SM's L1/TEX pipe
           (not HBM -- ncu shows 2 % DRAM utilisation -- and not tensor cores): algorithmic cost 16 B of
           L1/TEX traffic per executed tap (one bilinear source sample = 4 texels x 4 B); achieved = executed taps of the
           sweep launches x 16 B / their summed device time (CUDA events between launches on the launching stream);
           peak = 4 bilinear fetches/clk/SM x 148 SMs x sm_max_mhz x 16 B -- the nominal TMU rate, which
           tools/tex_microbench.cu reaches on this GPU (3.98/clk/SM, profiles/r01_tex_microbench.log).
  arithmetic : the headline is the EXACT arithmetic of the library (mpmvs_set_arithmetic, include/mpmvs_b200.h): the
           reference's operations one for one, float32 view storage, results bit-identical to the reference's kernels
           (tests/test_zz_fidelity_build_gpu.py). `arithmetic_fast` (N=1) reports the same step through the library's fast
           arithmetic with 8-bit view storage (statistically equal results) beside it; --no-fast-arm skips it.
  cpu_baseline : the plain-C oracle port (oracle/pm_oracle.c, all host cores) on a centre crop of the same views, one whole
           step (photometric Run + planar-prior stage + prior Run).
  --impl reference : the reference's own CUDA path (oracle/_ref/libmpmvs_ref.so = /root/reference/src/PatchMatch.cu
           compiled in place for sm_100) on the same workload and step; its host planar-prior stage is the restatement in
           oracle/ (OpenCV Subdiv2D + plain C, one core). The reference has no CPU implementation of the kernels.
           Rank 0 only (the reference is single-GPU, src/PatchMatch.cpp:509).

Multi-GPU: two things in one line.
  value / e2e   : the metric's configuration (configs[2]) has no geometric-consistency pass, so its reference images are
           independent: each rank runs its own reference images, no data-path collective -> "scaling": "weak".
  sharded_scene : the north star's multi-GPU configuration (configs[3]: 300 views 1920x1080 sharded by reference image,
           photometric + 2 geometric-consistency passes, the depth all-gather over NVLink INSIDE the timed region) through
           mp-mvs_b200/pipeline.py at every N, strong scaling -- `value` there is the whole scene's Mpix/s; see
           run_sharded_scene. --scene-views 0 skips it.
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
        shape, ref, label = (3200, 2130), 5, "eth3d-shaped 3200x2130 indoor weak-texture, 1 ref + 10 src views, planar prior + multi-scale windows"
        mk = lambda: synth.make_eth3d_scene(width=3200, height=2130, n_views=11, n_src=10, workers=workers)  # noqa: E731
    elif name == "dtu":
        shape, ref, label = (1600, 1200), 24, "dtu-shaped 1600x1200, 1 ref + 10 src views, planar prior + multi-scale windows"

        def mk():
            probe = synth.make_dtu_scene(views=[])
            ids = [24] + [i for i, _ in probe.pairs[24]][:10]
            return synth.make_dtu_scene(views=ids, workers=workers)
    elif name == "plane":
        shape, ref, label = (640, 480), 1, "textured plane 640x480, 1 ref + 2 src views, planar prior + multi-scale windows"
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
    """SM clock, power and throttle reasons of one GPU, sampled every 100 ms DURING a timed region. In-process NVML
    (pynvml) when available: a polling nvidia-smi process per rank costs the kernels a few percent once four or more ranks
    do it at the same time (profiles/r01_bench_eth3d_w{4,8}_v2.json: 430 ms/step with, 415 without). Falls back to
    `nvidia-smi --query-gpu ... -lms 100`."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        while not self.stop_flag:
            try:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)),
                                  str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)] + ["Active" if mask & bit else "Not Active" for _, bit in self.REASONS])
            except Exception:
                pass
            time.sleep(0.1)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"]}
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for (name, _), v in zip(self.REASONS, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def ncu_traffic(workload: str, arithmetic: str = "exact", tex: str = "f32"):
    """DRAM bytes per pm_sweep_kernel launch from the committed ncu capture of this workload, arithmetic and view storage
    (profiles/r02_ncu_traffic.json <- r02_ncu_sweep_exact_v2.txt, r01_ncu_traffic.json), else None."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        return j["dram_bytes_per_launch"].get(f"{arithmetic}_{tex}") if j.get("workload") == workload else None
    except Exception:
        return None


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def workload_config(prob, tex: str) -> dict:
    """The keys that DEFINE the measured configuration -- identical for this arm and for `--impl reference`, so the driver can
    tell that both ran the same thing; what is specific to an implementation goes to the line's `implementation` key."""
    W, H, n = prob["width"], prob["height"], len(prob["images"])
    return {"workload": prob["label"], "width": W, "height": H, "src_views": n - 1,
            "passes": "ProcessProblem(geom=0, planar=1): photometric Run (scales 2,1,0 x 3 it) + planar-prior stage + prior Run (3 it)",
            "view_storage": tex,
            "l2": "inputs_larger_than_l2 (%d views = %.0f MB + %.0f MB state vs 126 MB L2)" % (n, n * W * H * (1 if tex == "u8" else 4) / 1e6, W * H * 76 / 1e6)}


# --------------------------------------------------------------------------------------------- arms
TAKE_TURNS = os.environ.get("MPMVS_BENCH_TURNS", "1") != "0"


def start_problem(pm, seed):
    """First half of ProcessProblem(geom=false, planar=true), /root/reference/src/PatchMatch.cpp:516-531: photometric Run()."""
    pm.reset_params()
    pm.set_geom_consistency_params(False, True)
    pm.run_async(seed)


def finish_problem(pm, seed, prof=None):
    """Second half, PatchMatch.cpp:532-609: planar-prior stage on the state in HBM, then the planar-prior Run()."""
    pm.set_planar_prior_params()
    pm.set_geom_consistency_params(False, True)
    st = pm.build_prior()               # waits for this handle's stream: the vertex cells come back for the host triangulation
    if prof is not None:
        prof.append(pm.last_run_profile(timing=prof.timing, count=prof.count))
    pm.run_async(seed + 1)
    return st


def process_problems(handles, n_steps, seed0, upload=None, download=None, prof=None, streams=None, trace=None):
    """n_steps reference images through ProcessProblem. Every handle is driven by its own host thread (the C ABI blocks
    only the calling thread and releases the GIL), image i goes to handle i mod len(handles). The handles' streams have
    different priorities, so the GPU serves the most urgent handle first and fills the gaps it leaves -- its host
    triangulation, uploads, downloads -- with the other handles' kernels: the device never waits for the host.
    With one handle this is the reference's strictly sequential order (used for the per-kernel profile)."""
    from concurrent.futures import ThreadPoolExecutor

    stats = [None] * n_steps
    from mpmvs_b200.pipeline import GpuTurn     # whole Run()s take turns on the device (see its docstring)

    turn = GpuTurn() if (len(handles) > 1 and prof is None and TAKE_TURNS) else None

    def timed(name, k, i, fn, *a):
        if trace is None:
            return fn(*a)
        t = time.time()
        r = fn(*a)
        trace.append((k, i, name, t, time.time()))      # host-side duration of the call: shows which C-ABI calls block
        return r

    def worker(k):
        pm = handles[k]
        for i in range(k, n_steps, len(handles)):
            if upload is not None:
                timed("upload", k, i, upload, pm)
            if turn is None:
                timed("start", k, i, start_problem, pm, seed0 + 2 * i)
                stats[i] = timed("finish", k, i, finish_problem, pm, seed0 + 2 * i, prof)
            else:
                with turn:                                   # photometric Run()
                    timed("start", k, i, start_problem, pm, seed0 + 2 * i)
                    timed("wait", k, i, pm.synchronize)
                pm.set_planar_prior_params()
                pm.set_geom_consistency_params(False, True)
                stats[i] = timed("prior", k, i, pm.build_prior)      # vertex pick, host triangulation, rasterisation
                with turn:                                   # planar-prior Run()
                    timed("run2", k, i, pm.run_async, seed0 + 2 * i + 1)
                    timed("wait2", k, i, pm.synchronize)
            if download is not None:
                timed("download", k, i, download, pm)
            if prof is not None:        # sequential profiling pass: read the prior run's events too
                pm.synchronize()
                prof.append(pm.last_run_profile(timing=prof.timing, count=prof.count))

    if len(handles) == 1:
        worker(0)
    else:
        with ThreadPoolExecutor(len(handles)) as ex:
            list(ex.map(worker, range(len(handles))))
    return stats


class Prof(list):
    def __init__(self, timing=True, count=False):
        super().__init__()
        self.timing, self.count = timing, count


def run_ours(args, rank, world, local_rank, prob, barrier, allmax, arithmetic=None, tex=None):
    import torch
    from mpmvs_b200 import capi

    arithmetic = arithmetic or args.arithmetic
    tex = tex or args.tex
    torch.cuda.set_device(local_rank)
    W, H, n = prob["width"], prob["height"], len(prob["images"])
    fmt = {"f32": capi.TEX_F32, "f16": capi.TEX_F16, "u8": capi.TEX_U8}[tex]
    depth = max(1, args.in_flight)
    # each handle works on its own torch stream, so torch events bracket its work
    lo, hi = torch.cuda.Stream.priority_range()      # (0, -5) on B200: lower number = more urgent
    streams = [torch.cuda.Stream(device=local_rank, priority=max(hi, lo - (depth - 1 - k))) for k in range(depth)]
    handles = [capi.PatchMatch(device=local_rank, stream=s.cuda_stream).set_arithmetic(arithmetic).set_tex_format(fmt) for s in streams]
    # host images: uint8 grey levels as decoded from the JPEGs when the storage is 8-bit, else float32; pinned
    host_imgs = [i.astype(np.uint8) for i in prob["images"]] if tex == "u8" else prob["images"]
    pin_imgs = [torch.from_numpy(i).pin_memory() for i in host_imgs]
    pin_np = [t.numpy() for t in pin_imgs]
    h_planes = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in handles]
    h_costs = [torch.empty((H, W), dtype=torch.float32).pin_memory() for _ in handles]
    # resident arm: views uploaded once per handle (the per-GPU image cache), results stay in HBM
    for pm in handles:
        pm.set_problem(pin_np, prob["cams"])
        pm.synchronize()

    def sync_all():
        for pm in handles:
            pm.synchronize()

    def end_events():
        evs = []
        for st_ in streams:
            e = torch.cuda.Event(enable_timing=True)
            e.record(st_)
            evs.append(e)
        for e in evs:
            e.synchronize()
        return evs

    seed = 1000 * (rank + 1)
    process_problems(handles, max(args.warmup, len(handles)), seed, streams=streams)     # every handle runs at least once untimed
    sync_all()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region 1: device-resident throughput. e0 is recorded while the GPU is idle; the region ends when the last
    # of the handles' streams has drained (max over the per-stream end events).
    e0 = torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t0 = time.time()
    e0.record(streams[0])
    pstats = process_problems(handles, args.steps, seed + 100, streams=streams)
    tot_ms = max(float(e0.elapsed_time(e)) for e in end_events())
    torch.cuda.synchronize(); barrier()
    wall_resident = time.time() - t0
    tot_ms = allmax(tot_ms)
    clocks = sampler.stop()
    launches = args.steps * (22 + 10 + 5)   # photometric Run, prior Run (init + 6 sweeps + 3), pick + triangle + expand + 2 memsets
    st = pstats[-1]
    dl_ms = sum(s_["delaunay_ms"] for s_ in pstats)

    # ---- timed region 2: end to end through the C ABI with host buffers (H2D of the views, D2H of the results), pipelined
    # the same way; host wall clock from the first upload to the last byte of the last result
    idx = {id(pm): k for k, pm in enumerate(handles)}

    def upload(pm):
        pm.set_problem(pin_np, prob["cams"])

    def download(pm):
        k = idx[id(pm)]
        pm.get_results_async(h_planes[k].numpy(), h_costs[k].numpy())

    process_problems(handles, max(len(handles), min(args.warmup, 2)), seed, upload, download, streams=streams)
    sync_all()
    barrier(); torch.cuda.synchronize()
    t0 = time.time()
    trace = [] if args.trace else None
    skip = os.environ.get("MPMVS_E2E_SKIP", "") if args.trace else ""      # diagnostic only: isolate the cost of either copy
    if skip:
        print(f"[diagnostic] e2e region without: {skip} -- not a valid e2e number", file=sys.stderr)
    process_problems(handles, args.steps, seed + 200, None if "upload" in skip else upload, None if "download" in skip else download,
                     streams=streams, trace=trace)
    sync_all()
    if trace:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"e2e_trace_rank{rank}.json"), "w") as fh:
            json.dump({"t0": t0, "t_end": time.time(), "calls": sorted(trace, key=lambda c: c[3])}, fh)
    torch.cuda.synchronize()
    e2e_s = allmax(time.time() - t0)
    barrier()
    last = (args.steps - 1) % len(handles)
    checksum = float(h_costs[last].double().mean())
    acc = synth.accuracy_at(h_planes[last].numpy()[..., 3], prob["gt_depth"])

    # ---- untimed for `value`: per-kernel device times (events between launches) and executed NCC counts, strictly
    # sequential on one handle so a kernel is timed alone
    pm = handles[0]
    pm.set_profiling(1)
    prof = Prof(timing=True)
    nprof = min(args.steps, 2)
    process_problems([pm], nprof, seed + 100, prof=prof)
    pm.synchronize()
    sweep_ms = sum(p_["sweep_ms"] for p_ in prof) / nprof
    init_ms = sum(p_["init_ms"] for p_ in prof) / nprof
    fin_ms = sum(p_["finalize_ms"] for p_ in prof) / nprof
    n_sweeps = sum(p_["n_sweeps"] for p_ in prof) // nprof
    pm.set_profiling(2)
    cnt = Prof(timing=False, count=True)
    process_problems([pm], 1, seed + 100, prof=cnt)
    pm.synchronize()
    ncc_exec = sum(c["ncc_evaluations"] for c in cnt)
    pm.set_profiling(0)

    mpix = W * H / 1e6
    value = world * args.steps * mpix / (tot_ms / 1e3)
    e2e = world * args.steps * mpix / e2e_s
    peaks = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    nsrc = n - 1
    # executed taps of the sweep launches: total minus the two init launches' (N-1) evaluations per pixel (upper bound)
    taps_sweeps = max(0, ncc_exec - 2 * W * H * nsrc) * TAPS_PER_NCC
    sweep_s = sweep_ms / 1e3
    achieved = taps_sweeps * BYTES_PER_TAP / sweep_s / 1e9
    peak = TEX_PER_CLK_SM * N_SM * sm_max * 1e6 * BYTES_PER_TAP / 1e9
    ref_equiv_taps = W * H * nsrc * ((1 + 14 * 9) + (1 + 14 * 3)) * TAPS_PER_NCC
    out = {
        "metric": "depth-map Mpix/s (3200x2130, 10 src views)" if args.workload == "eth3d" else f"depth-map Mpix/s ({W}x{H}, {nsrc} src views)",
        "value": round(value, 4), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(tot_ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(prob, tex),
        "implementation": {"arithmetic": handles[0].arithmetic, "library": capi.build_flavor(), "lib_variant": os.environ.get("MPMVS_LIB_VARIANT", "default"),
                           "parity": ("bit-identical to the reference's kernels in all three modes (tests/test_zz_fidelity_build_gpu.py)"
                                      if arithmetic == "exact" and tex == "f32" else "statistically equal to the reference's kernels (DESIGN.md section 5)"),
                           "parallelism": f"reference images sharded over {world} gpu(s), no data-path collective in this pass (see sharded_scene)"},
        "in_flight": depth,
        "e2e": {"value": round(e2e, 4), "unit": "Mpix/s", "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in pin_imgs) + 112 * n),
                "d2h_bytes_per_step": int(W * H * 20), "ms_per_step": round(e2e_s * 1e3 / args.steps, 3)},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {"init": round(init_ms, 3), "sweeps": round(sweep_ms, 3), "finalize": round(fin_ms, 3),
                               "n_sweep_launches": n_sweeps, "host_delaunay": round(dl_ms / args.steps, 3),
                               "note": "kernels timed alone in a sequential pass; value/e2e keep `in_flight` images in flight"},
        "roofline": {"bound": "l1tex", "kernel": "pm_sweep_kernel", "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": ncu_traffic(args.workload, arithmetic, tex),
                     "peak_source": "4 bilinear/clk/SM x 148 SM x measured sm_max_mhz x 16 B (profiles/r01_tex_microbench.log reaches 3.98/clk/SM)",
                     "executed_taps_per_step": int(taps_sweeps), "reference_equivalent_taps_per_step": int(ref_equiv_taps),
                     "gtaps_per_s": round(taps_sweeps / sweep_s / 1e9, 2),
                     "achieved_per_launch_bytes": int(taps_sweeps * BYTES_PER_TAP / max(1, n_sweeps)),
                     "traffic_note": "DRAM bytes per sweep launch (ncu, profiles/r02_ncu_traffic.json): the source views once (300 MB as float32) + the per-pixel state a half-sweep reads and writes in 32-byte sectors that hold both checkerboard colours; < 1 % of the HBM peak -- the kernel is bound by the L1 data stage, which texture and shared-memory wavefronts share: tex 72 % + lsu 23 % = 95 % busy (profiles/r02_ncu_sweep_exact_v2.txt, r02_l1_data_stage.md), not by DRAM",
                     "hbm_peak_gbs": peaks.get("hbm_gbs")},
        "clocks": clocks,
        "checksum_mean_cost": round(checksum, 6), "accuracy_2_5_10cm": [round(a, 3) for a in acc],
        "prior_stage": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
        "wall_s_resident": round(wall_resident, 3),
    }
    for pm in handles:
        pm.destroy()
    return out


def run_sharded_scene(args, rank, world, local_rank, dist, barrier):
    """BASELINE.json configs[3], the north star's multi-GPU configuration, through the product's own pipeline
    (mp-mvs_b200/pipeline.py: DensePipeline) with the exchange INSIDE the timed region: a Tanks-and-Temples-shaped ring of
    `--scene-views` (300) views 1920x1080, 10 sources each, photometric pass + 2 geometric-consistency passes
    (main(), /root/reference/src/main.cpp:20-41), reference images sharded over the ranks in contiguous blocks, depth maps
    all-gathered over NVLink with NCCL between the passes instead of the reference's depths.dmb files
    (PatchMatch.cpp:620-633 writes, :934-950 reads). STRONG scaling: the scene is fixed, so `value` at N GPUs over `value` at
    one GPU is the speed-up. Timing: a barrier + device synchronisation on both sides, CUDA events per pass and per exchange,
    maximum over the ranks."""
    import torch

    from mpmvs_b200 import pipeline

    W, H, V = 1920, 1080, args.scene_views
    cache = f"/dev/shm/mpmvs_scene_tnt_{V}_{W}x{H}_{os.getuid()}.npz"
    if rank == 0 and not os.path.exists(cache):
        t = time.time()
        sc = synth.make_tnt_scene(width=W, height=H, n_views=V, n_src=10, workers=min(32, os.cpu_count() or 1))
        np.savez(cache + ".tmp.npz", images=np.stack(sc.images), cams=io_formats.pack_cameras(sc.cams),
                 pairs=np.array([[j for j, _ in sc.pairs[i]] for i in range(sc.num_views)]), gt=np.stack(sc.gt_depth).astype(np.float16))
        os.replace(cache + ".tmp.npz", cache)
        log(f"[bench] rendered the {V}-view ring {W}x{H} in {time.time() - t:.1f} s on {os.cpu_count()} cores")
    barrier()
    z = np.load(cache)
    cams_packed = z["cams"]
    cams = {}
    for i in range(len(cams_packed)):
        r = cams_packed[i]
        cams[i] = io_formats.Camera(K=r["K"].reshape(3, 3), R=r["R"].reshape(3, 3), t=r["t"], height=int(r["height"]), width=int(r["width"]),
                                    depth_min=float(r["depth_min"]), depth_max=float(r["depth_max"]))
    entries = [io_formats.SceneEntry(ref_id=i, src_ids=[i] + [int(j) for j in z["pairs"][i]], estimate=True) for i in range(len(cams_packed))]
    mine = pipeline.shard_refs(list(range(V)), rank, world)
    need = sorted({j for i in mine for j in entries[i].src_ids})
    all_imgs = z["images"]
    images = {i: all_imgs[i] for i in need}           # a rank loads only its shard's reference and source views
    cfg = pipeline.PipelineConfig(geom_iterations=2, max_src=10, seed=9, planar_prior=False, geom_planar_prior=False, in_flight=8,
                                  arithmetic=args.arithmetic)
    p = pipeline.DensePipeline(entries, cams, images, cfg, rank=rank, world=world, device=local_rank, dist=dist)
    t = time.time()
    n_cached = p.setup()
    torch.cuda.synchronize()
    setup_s = time.time() - t
    if dist is not None:     # NCCL communicator and its buffers: outside the timed region
        w = torch.zeros(1024, device="cuda")
        o = torch.zeros(1024 * world, device="cuda")
        dist.all_gather_into_tensor(o, w)
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t = time.time()
    stats = p.run()
    torch.cuda.synchronize()
    barrier()
    dt = time.time() - t
    clocks = sampler.stop()
    gt = z["gt"]
    acc, comp = [], []
    for r_ in mine[:3]:
        planes, costs = p.engines[r_].result()
        a_, c_ = synth.accuracy_completeness_at(planes[..., 3], costs, gt[r_].astype(np.float32))
        acc.append(a_); comp.append(c_)
    tt = torch.tensor([dt] + [s_.device_ms for s_ in stats] + [s_.exchange_ms for s_ in stats], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tt = tt.tolist()
    n = len(stats)
    p.destroy()
    torch.cuda.empty_cache()
    limiting = max(range(n), key=lambda k: tt[1 + k])
    return {"workload": f"tnt-shaped ring, {V} views {W}x{H}, 10 src views, photometric + 2 geometric-consistency passes (BASELINE.json configs[3])",
            "scaling": "strong", "n_gpus": world, "refs": V, "refs_per_rank_max": p.block, "arithmetic": args.arithmetic,
            "view_storage": "f32" if args.arithmetic == "exact" else "u8", "value": round(V * W * H / 1e6 / tt[0], 3), "unit": "Mpix/s",
            "total_s": round(tt[0], 3), "schedule": [s_.name for s_ in stats], "pass_ms_max_over_ranks": [round(v, 1) for v in tt[1:1 + n]],
            "exchange_ms_max_over_ranks": [round(v, 2) for v in tt[1 + n:]], "exchange_bytes_per_pass": V * W * H * 4,
            "collective": "NCCL all_gather_into_tensor of the depth maps (in place, buffers preallocated) after the photometric pass and after geometric pass 0"
                          if world > 1 else "none (one GPU: the maps are exported into the same buffer the next pass reads)",
            "limiting_pass": stats[limiting].name, "ideal_speedup": round(V / p.block, 2), "views_cached_this_rank": n_cached,
            "setup_s_rank0": round(setup_s, 2), "clocks": clocks,
            "accuracy_2_5_10cm_first_refs_rank0": [[round(x, 2) for x in a_] for a_ in acc],
            "completeness_2_5_10cm_first_refs_rank0": [[round(x, 2) for x in c_] for c_ in comp]}


def cpu_baseline(prob, budget_s=20.0):
    """The plain-C oracle port on a centre crop of the same views (principal point shifted), all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    W, H = prob["width"], prob["height"]
    cw, ch = min(W, 352), min(H, 240)        # ~12 s of CPU work on the GPU box's 16 host cores
    x0, y0 = (W - cw) // 2, (H - ch) // 2
    # keep the sources whole (their warped windows must stay inside), crop only the reference: the oracle needs all
    # views at the size its camera says, so crop every view with the same window instead and shift cx, cy
    imgs = [np.ascontiguousarray(i[y0:y0 + ch, x0:x0 + cw]) for i in prob["images"]]
    cams = prob["cams"].copy()
    for c in cams:
        c["K"][2] -= x0
        c["K"][5] -= y0
        c["width"], c["height"] = cw, ch
    import prior_oracle

    o = oracle_py.Oracle("cpu").set_problem(imgs, cams)
    o.set_geom_consistency_params(False, True)
    cores = oracle_py.Oracle("cpu").lib.pmo_get_threads()
    t = time.time()
    o.run(1)                                                   # photometric Run()
    planes, costs = o.result()
    dmin, dmax = o.depth_range
    prior, mask, _, _, _ = prior_oracle.build_prior_fast(planes, costs, cams["K"][0].reshape(3, 3), dmin, dmax)   # host prior stage
    o.set_planar_prior_params()
    o.set_geom_consistency_params(False, True)
    o.set_prior(prior, mask)
    o.run(2)                                                   # planar-prior Run()
    dt = time.time() - t
    o.destroy()
    return {"value": round(cw * ch / 1e6 / dt, 6), "unit": "Mpix/s", "cores": int(cores), "kind": "port",
            "sample": f"centre crop {cw}x{ch} of all {len(imgs)} views, one whole step (photometric Run + planar-prior stage + prior Run) in {dt:.1f} s "
                      "(oracle/pm_oracle.c, pthreads; the reference has no CPU implementation of this path)"}


def run_reference(args, prob):
    """The reference's own CUDA path (oracle/_ref) on GPU 0, same workload, same step: Run() + the host planar-prior
    stage (restated: oracle/prior_oracle.py, OpenCV's Subdiv2D + plain-C loops on one core, as the reference's host code
    is single-threaded) + CudaPlanarPriorInitialization + Run()."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import prior_oracle

    if not oracle_py.available("ref"):
        return {"impl": "reference", "unavailable": "oracle/_ref/libmpmvs_ref.so not built (needs /root/reference at build time)"}
    oracle_py.build("cpu")
    W, H, n = prob["width"], prob["height"], len(prob["images"])
    K = prob["cams"]["K"][0].reshape(3, 3)

    def step(R, seed):
        R.set_geom_consistency_params(False, True)
        ms = R.run(seed)                       # CUDA events around PatchMatchCUDA::Run(), which ends with its blocking D2H copies
        planes, costs = R.result()
        t = time.time()
        dmin, dmax = R.depth_range
        prior, mask, verts, tris, cnt = prior_oracle.build_prior_fast(planes, costs, K, dmin, dmax)
        host_ms = (time.time() - t) * 1e3
        R.set_planar_prior_params()
        R.set_geom_consistency_params(False, True)
        t = time.time()
        R.set_prior(prior, mask)
        up_ms = (time.time() - t) * 1e3
        ms2 = R.run(seed + 1)
        return ms + host_ms + up_ms + ms2, host_ms, R.result()

    sampler = ClockSampler(0)
    for i in range(args.warmup):
        R = oracle_py.Oracle("ref").set_problem(prob["images"], prob["cams"])
        step(R, 10 + 2 * i)
        R.destroy()
    sampler.start()
    ms = host = 0.0
    t0 = time.time()
    for i in range(args.steps):
        # a fresh object per reference image, as ProcessProblem does (PatchMatch.cpp:516); its allocation and the upload
        # of the views are in `e2e` but not in `value`
        R = oracle_py.Oracle("ref").set_problem(prob["images"], prob["cams"])
        a, b, res = step(R, 100 + 2 * i)
        ms += a; host += b
        R.destroy()
    e2e_s = time.time() - t0
    clocks = sampler.stop()
    mpix = W * H / 1e6
    v = args.steps * mpix / (ms / 1e3)
    acc = synth.accuracy_at(res[0][..., 3], prob["gt_depth"])
    return {
        "impl": "reference", "metric": "depth-map Mpix/s (3200x2130, 10 src views)" if args.workload == "eth3d" else f"depth-map Mpix/s ({W}x{H}, {n-1} src views)",
        "value": round(v, 4), "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(prob, "f32"),
        "implementation": {"arithmetic": "the reference's own kernels and launcher (oracle/_ref = /root/reference/src/PatchMatch.cu compiled in place for sm_100)",
                           "host_prior_stage": "restated in oracle/ (OpenCV's Subdiv2D + plain C, one core)"},
        "cpu_baseline": {"value": round(v, 4), "unit": "Mpix/s", "cores": 1, "kind": "reference",
                         "sample": "the reference has no CPU implementation of the kernels: its own CUDA path (PatchMatch.cu rebuilt for sm_100, -O3 "
                                   "--use_fast_math --maxrregcount=128) on one B200, full workload; host prior stage %.0f ms/step on 1 core" % (host / args.steps)},
        "e2e": {"value": round(args.steps * mpix / e2e_s, 4), "unit": "Mpix/s", "h2d_bytes_per_step": int(n * W * H * 4 + 112 * n + W * H * 20),
                "d2h_bytes_per_step": int(2 * W * H * 20), "ms_per_step": round(e2e_s * 1e3 / args.steps, 3)},
        "gpu_launches": 44 * args.steps, "clocks": clocks, "host_prior_stage_ms": round(host / args.steps, 1),
        "accuracy_2_5_10cm": [round(a, 3) for a in acc],
    }


def fast_arm(args, rank, world, local_rank, prob, barrier, allmax) -> dict:
    """The same step through the library's FAST arithmetic with 8-bit view storage (mpmvs_set_arithmetic; statistically equal
    results), in this process after the headline. Reported beside the headline, never instead of it: any failure ends up in
    `error`."""
    try:
        j = run_ours(args, rank, world, local_rank, prob, barrier, allmax, arithmetic="fast", tex="u8")
        keep = ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "kernel_ms_per_step", "clocks", "checksum_mean_cost", "accuracy_2_5_10cm")
        res = {k: j[k] for k in keep if k in j}
        res["roofline_frac"] = j.get("roofline", {}).get("frac")
        res["executed_taps_per_step"] = j.get("roofline", {}).get("executed_taps_per_step")
        res["config"] = {"view_storage": j["config"]["view_storage"], **{k: j["implementation"][k] for k in ("arithmetic", "parity")}}
        res["note"] = "same workload, steps and timing rules as the headline"
        return res
    except Exception as e:          # noqa: BLE001 -- a reported extra must never take the headline down
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="eth3d")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace", action="store_true", help="write the host-side call timeline of the e2e region to gpurun_out/")
    ap.add_argument("--in-flight", type=int, default=2, help="reference images kept in flight per GPU (1 = the reference's sequential order)")
    ap.add_argument("--arithmetic", "--fidelity", dest="arithmetic", default="exact", choices=["exact", "fast"],
                    help="exact (default): the reference's operations one for one, bit-identical results; fast: hoisted arithmetic, "
                         "statistically equal results (mpmvs_set_arithmetic)")
    ap.add_argument("--tex", default=None, choices=["f32", "f16", "u8"],
                    help="storage format of the views in HBM (default: f32 with the exact arithmetic -- the reference's, and what its "
                         "bit-identity needs -- and u8 with the fast one)")
    ap.add_argument("--scene-views", type=int, default=300, help="views of the sharded config-4 scene (sharded_scene); 0 skips it")
    ap.add_argument("--no-fast-arm", "--no-exact-arm", dest="no_fast_arm", action="store_true",
                    help="at N=1, skip the extra measurement of the fast arithmetic (arithmetic_fast)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 1)
    if args.tex is None:
        args.tex = "f32" if args.arithmetic == "exact" else "u8"

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
    scene = None
    if args.scene_views > 0:
        try:
            scene = run_sharded_scene(args, rank, world, local_rank, dist, barrier)
        except Exception as e:          # noqa: BLE001 -- every rank reaches the barrier below; the headline line survives
            scene = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if rank == 0:
        if scene is not None:
            out["sharded_scene"] = scene
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(prob)
        if world == 1 and args.arithmetic == "exact" and not args.no_fast_arm:
            out["arithmetic_fast"] = fast_arm(args, rank, world, local_rank, prob, barrier, allmax)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
