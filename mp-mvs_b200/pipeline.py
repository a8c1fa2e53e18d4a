"""Multi-GPU stage schedule of MP-MVS's depth estimation (main(), /root/reference/src/main.cpp:20-41), sharded by
reference image: one process per GPU, `torch.distributed` for the plumbing, depth maps exchanged between passes with an
all-gather over NVLink instead of through `depths.dmb` files (PatchMatch.cpp:620-633 writes, :934-950 reads).

    stage 1   photometric Run() of every reference image            (main.cpp:20-26)
    exchange  all-gather of the fresh depth maps (4 B/px per image)
    stage 2   `geom_iterations` x { geometric-consistency Run(); exchange }   (main.cpp:29-41)

Differences from the reference's schedule, both deliberate:
  * Jacobi order: every image of pass k reads the depth maps of pass k-1 (the reference overwrites the files in place
    image by image, so image i sees already-updated maps of images < i). Results are therefore identical for any number
    of GPUs given the same seed;
  * every reference image keeps its per-pixel state (planes, costs, view masks, RNG) resident in HBM between passes
    (one `mpmvs_problem` per image), so a geom pass restarts from device memory, not from normals.dmb/costs.dmb.

The compute engine is the CUDA library (CudaEngine -> capi.PatchMatch). The engine is injectable only so that the
sharding/exchange logic can be exercised on CPU with the gloo backend by the tests; there is no CPU engine in the product.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def shard_refs(ref_ids: Sequence[int], rank: int, world: int) -> List[int]:
    """Contiguous blocks (rank r gets refs [r*B, (r+1)*B)), B = ceil(n / world): neighbouring images share sources,
    so a rank's working set of views stays small. Equal image sizes -> equal cost per image."""
    n = len(ref_ids)
    b = (n + world - 1) // world
    return list(ref_ids[rank * b:(rank + 1) * b])


def block_size(n_refs: int, world: int) -> int:
    return (n_refs + world - 1) // world


def stage_seed(seed: int, ref_id: int, stage: int) -> int:
    """Per (image, pass) seed, independent of the rank that runs it -- the C++ host's convention (mpmvs_main.cpp: seed +
    1000003 i + 7919 stage, the planar-prior Run() of a pass XORs 0x5DEECE66D), so both hosts draw the same random numbers
    for the same --seed (the library mixes the seed with the pixel position, pm_rng_init)."""
    return (seed + 1000003 * ref_id + 7919 * stage) & 0xFFFFFFFFFFFFFFFF


@dataclass
class PipelineConfig:
    geom_iterations: int = 2          # "Geometric consistency iterations"
    planar_prior: bool = False        # "Planer prior"
    geom_planar_prior: bool = False   # "Geometric consistency planer prior"
    in_flight: int = 8                # reference images a rank keeps in flight (host threads; the C ABI releases the GIL)
    max_src: int = 20                 # "Max source images num"
    seed: int = 0
    arithmetic: str = "exact"         # mpmvs_set_arithmetic: "exact" = the reference's kernel results bit for bit, "fast"
    tex_format: Optional[int] = None  # capi.TEX_*; None = float32 with the exact arithmetic (what its bit-identity needs),
                                      # 8-bit with the fast one (views are 8-bit grey levels as decoded from the JPEGs)
    keep_normals: bool = True
    order: str = "jacobi"             # "jacobi": every image of a geometric pass reads the maps of the previous pass (any number of
                                      # GPUs, results independent of the sharding); "gauss_seidel": the reference's own order -- images
                                      # one after the other, each reading the maps images before it already updated in this pass
                                      # (PatchMatch.cpp:620-633 overwrites depths.dmb in place) -- one GPU only
    keep_priors: bool = False         # tests: keep a host copy of every prior (planes, mask) the planar-prior stage builds


@dataclass
class PassStats:
    name: str
    device_ms: float = 0.0           # wall/device time of the pass on this rank (all its reference images)
    exchange_ms: float = 0.0
    n_refs: int = 0


class GpuTurn:
    """FIFO ticket for whole Run()s on one GPU. Several reference images are in flight so that the device never waits for
    the host (triangulation, copies); but Run()s that share the SMs finish together, send their images into the host stage
    together and leave the device idle together, and that lock-step is stable. Taking turns keeps the images out of phase
    by construction: while one image's Run() has the device, the others do their host work."""

    def __init__(self):
        import threading

        self.cv = threading.Condition()
        self.next_ticket = 0
        self.serving = 0

    def __enter__(self):
        with self.cv:
            t = self.next_ticket
            self.next_ticket += 1
            self.cv.wait_for(lambda: self.serving == t)
        return self

    def __exit__(self, *exc):
        with self.cv:
            self.serving += 1
            self.cv.notify_all()
        return False


class _NoTurn:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class CudaEngine:
    """One reference image on one GPU: a resident `mpmvs_problem` (capi.PatchMatch)."""

    def __init__(self, device: int, cache, ids: Sequence[int], cams_packed: np.ndarray, arithmetic: str = "exact", stream: Optional[int] = None):
        from . import capi

        self.pm = capi.PatchMatch(device, stream=stream).set_arithmetic(arithmetic)
        self.pm.set_problem_cached(cache, list(ids), cams_packed)
        self.prior_stats = None
        self.keep_prior = False
        self.saved_priors = []      # (planes, mask) of every planar-prior stage of this image, in pass order (keep_prior)

    def process(self, seed: int, geom: bool, planar: bool, src_depth_ptrs: Optional[Sequence[int]] = None, turn=None):
        """The compute of one ProcessProblem(geom, planar) call (PatchMatch.cpp:516-609) on the resident state:
        Run(); if planar: planar-prior stage + planar-prior Run(). With a `turn` (GpuTurn) every Run() waits for its turn on
        the device and is complete on return; without, the last run is still in flight on return."""
        pm = self.pm
        pm.reset_params()                                   # a fresh PatchMatchCUDA object per call (PatchMatch.cpp:516)
        pm.set_geom_consistency_params(geom, planar)
        if geom:
            pm.set_src_depths_device(list(src_depth_ptrs))
        with (turn or _NoTurn()):
            pm.run_async(seed)
            if turn is not None:
                pm.synchronize()
        if planar:
            pm.set_planar_prior_params()
            pm.set_geom_consistency_params(False, True)
            self.prior_stats = pm.build_prior()             # blocks this host thread only
            if self.keep_prior:
                self.saved_priors.append(pm.get_prior())
            with (turn or _NoTurn()):
                pm.run_async(seed ^ 0x5DEECE66D)
                if turn is not None:
                    pm.synchronize()

    def export_depth(self, dst_tensor):
        self.pm.export_depth_device(dst_tensor.data_ptr(), dst_tensor.stride(0) * 4)

    def device_ms(self) -> float:
        return self.pm.last_run_ms()

    def synchronize(self):
        self.pm.synchronize()

    def result(self):
        import ctypes as C

        from . import capi

        planes = np.empty((self.pm.hgt, self.pm.w, 4), np.float32)
        costs = np.empty((self.pm.hgt, self.pm.w), np.float32)
        capi._ck(capi.lib().mpmvs_get_device_state(self.pm.h, planes.ctypes.data, costs.ctypes.data, None, None, None), "get_device_state")
        return planes, costs

    def destroy(self):
        self.pm.destroy()


class DensePipeline:
    """entries: io_formats.SceneEntry list (pair.txt); cams: id -> io_formats.Camera; images: id -> uint8/float32 (H, W)."""

    def __init__(self, entries, cams: Dict[int, object], images: Dict[int, np.ndarray], cfg: PipelineConfig, rank: int = 0,
                 world: int = 1, device: int = 0, dist=None, engine_factory: Optional[Callable] = None, torch_device=None):
        import torch

        from . import io_formats

        self.torch = torch
        self.io = io_formats
        self.cfg, self.rank, self.world, self.device, self.dist = cfg, rank, world, device, dist
        self.entries = {e.ref_id: e for e in entries if e.estimate}
        self.ref_ids = sorted(self.entries)
        self.my_refs = shard_refs(self.ref_ids, rank, world)
        self.block = block_size(len(self.ref_ids), world)
        self.cams, self.images = cams, images
        self.tdev = torch_device if torch_device is not None else torch.device("cuda", device)
        self.engine_factory = engine_factory
        self.engines: Dict[int, object] = {}
        self.stats: List[PassStats] = []
        any_cam = cams[self.ref_ids[0]]
        self.H, self.W = int(any_cam.height), int(any_cam.width)
        for i in self.ref_ids:
            if (int(cams[i].height), int(cams[i].width)) != (self.H, self.W):
                raise ValueError("the sharded pipeline requires equally sized views (one all-gather buffer)")
        # slot of an image in the gathered buffer: rank-major, then position inside the rank's block
        self.slot = {}
        for r in range(world):
            for k, ref in enumerate(shard_refs(self.ref_ids, r, world)):
                self.slot[ref] = r * self.block + k
        self.gathered = None          # [world * block, H, W] float32: depth maps of the previous pass
        self.gather_bufs = None       # the two buffers `gathered` alternates between (Jacobi double buffering), allocated once
        self.gather_turn = 0
        self.zero_map = None          # depth map of a source that is no reference image of the scene (no estimate): all 0
        self.streams = []             # CUDA streams the engines work on (in_flight of them, shared round-robin)
        self.cache = None

    # ------------------------------------------------------------------------------------------ set-up
    def problem_ids(self, ref: int) -> List[int]:
        ids = self.entries[ref].src_ids          # src_ids[0] == ref (GenerateSampleList, PatchMatch.cpp:84)
        return [ids[0]] + list(ids[1:1 + self.cfg.max_src])

    def setup(self):
        need = sorted({i for ref in self.my_refs for i in self.problem_ids(ref)})
        if self.engine_factory is None:
            from . import capi

            fmt = self.cfg.tex_format
            if fmt is None:
                fmt = capi.TEX_F32 if self.cfg.arithmetic == "exact" else capi.TEX_U8
            self.cache = capi.ImageCache(self.device, self.W, self.H, max(1, len(need)), fmt)
            for i in need:
                self.cache.put(i, self.images[i])
            # the engines work on `in_flight` torch streams, so that the exchange can be ordered against them with stream
            # events instead of device-wide synchronisation
            self.streams = [self.torch.cuda.Stream(device=self.tdev) for _ in range(max(1, self.cfg.in_flight))]
        for k, ref in enumerate(self.my_refs):
            ids = self.problem_ids(ref)
            packed = self.io.pack_cameras([self.cams[i] for i in ids])
            if self.engine_factory is None:
                self.engines[ref] = CudaEngine(self.device, self.cache, ids, packed, self.cfg.arithmetic,
                                               self.streams[k % len(self.streams)].cuda_stream)
                self.engines[ref].keep_prior = self.cfg.keep_priors
            else:
                self.engines[ref] = self.engine_factory(ids, [self.images[i] for i in ids], packed)
        if self.cfg.geom_iterations > 0:
            # both gather buffers once, zero-filled (slots past the last image of a short block, and `zero_map`, stay 0):
            # no allocation and no zero-fill inside a pass
            n = self.world * self.block
            self.gather_bufs = [self.torch.zeros((n, self.H, self.W), dtype=self.torch.float32, device=self.tdev)
                                for _ in range(2 if self.cfg.geom_iterations > 1 else 1)]
            self.zero_map = self.torch.zeros((self.H, self.W), dtype=self.torch.float32, device=self.tdev)
            self._sync_torch()
        return len(need)

    # ------------------------------------------------------------------------------------------ exchange
    def _exchange(self, st: PassStats):
        """All-gather of this pass's depth maps: every engine writes its map straight into this rank's block of the gather
        buffer (mpmvs_export_depth_device), NCCL gathers the blocks in place. The buffers are allocated once (setup) and
        alternate (Jacobi double buffering); ordering against the engines' streams is by stream events, not by device-wide
        synchronisation."""
        torch = self.torch
        out = self.gather_bufs[self.gather_turn % len(self.gather_bufs)]
        self.gather_turn += 1
        mine = out[self.rank * self.block:(self.rank + 1) * self.block]
        cuda = self.tdev.type == "cuda"
        cur = torch.cuda.current_stream(self.tdev) if cuda else None
        if cuda and self.streams:
            for s_ in self.streams:          # the previous pass may still be reading this buffer's other twin only; but its
                s_.wait_stream(cur)          # exports must not overtake an all-gather still in flight on `cur`
        t0 = self._tick()
        for k, ref in enumerate(self.my_refs):
            self.engines[ref].export_depth(mine[k])
        if cuda and self.streams:
            for s_ in self.streams:
                cur.wait_stream(s_)          # the gather starts when every engine stream has written its maps
        else:
            for ref in self.my_refs:
                self.engines[ref].synchronize()
        if self.world > 1:
            self.dist.all_gather_into_tensor(out, mine)       # in place: `mine` is this rank's block of `out`
        if cuda and self.streams:
            for s_ in self.streams:
                s_.wait_stream(cur)          # ... and the next pass's kernels start when the gathered maps are complete
        self.gathered = out
        st.exchange_ms = self._tock(t0)

    def _sync_torch(self):
        if self.tdev.type == "cuda":
            self.torch.cuda.synchronize(self.tdev)

    def _tick(self):
        if self.tdev.type == "cuda":
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        import time

        return time.time()

    def _tock(self, t0) -> float:
        if self.tdev.type == "cuda":
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            e.synchronize()
            return float(t0.elapsed_time(e))
        import time

        return (time.time() - t0) * 1e3

    # ------------------------------------------------------------------------------------------ passes
    def _run_pass(self, st: PassStats, stage: int, geom: bool, planar: bool, prev=None):
        """One ProcessProblem per reference image of this rank, `in_flight` of them at a time on host threads."""
        from concurrent.futures import ThreadPoolExecutor

        # passes with a host stage between two Run()s (the planar prior): whole Run()s take turns on the device (GpuTurn).
        # Without a host stage nothing can leave the device idle, and runs sharing it fill each other's kernel tails
        # (measured: DTU-shaped 49 views 5.62 s shared, 5.88 s with turns). Test engines have no device.
        turn = GpuTurn() if (planar and self.engine_factory is None and self.cfg.in_flight > 1 and len(self.my_refs) > 1) else None
        extra = {"turn": turn} if turn is not None else {}

        def one(ref):
            seed = stage_seed(self.cfg.seed, ref, stage)
            if geom:
                # a source without an estimate of its own (not a reference image of the scene) has depth 0 everywhere, which
                # ComputeGeomConsistencyCost answers with the maximum cost (cu:627-629) -- what the C++ hosts do as well
                views = [prev[self.slot[i]] if i in self.slot else self.zero_map for i in self.problem_ids(ref)[1:]]
                self.engines[ref].process(seed, True, planar, [v.data_ptr() for v in views] if self.tdev.type == "cuda" else views, **extra)
            else:
                self.engines[ref].process(seed, False, planar, **extra)

        t0 = self._tick()
        if geom and self.cfg.order == "gauss_seidel":
            # the reference's order (main.cpp:35-40 + PatchMatch.cpp:620-633): one image after the other, its fresh depth map
            # replaces the old one at once, so the images after it in this pass already read it
            if self.world != 1:
                raise ValueError("the Gauss-Seidel order is sequential over ALL reference images: one GPU only")
            for ref in self.my_refs:
                one(ref)
                self.engines[ref].synchronize()
                self.engines[ref].export_depth(prev[self.slot[ref]])
                self.engines[ref].synchronize()
        elif self.cfg.in_flight > 1 and len(self.my_refs) > 1:
            with ThreadPoolExecutor(self.cfg.in_flight) as ex:
                list(ex.map(one, self.my_refs))
        else:
            for ref in self.my_refs:
                one(ref)
        for ref in self.my_refs:
            self.engines[ref].synchronize()
        st.device_ms = self._tock(t0)

    def run(self):
        """The stage schedule of main() (main.cpp:20-41) over this rank's reference images."""
        if not self.engines:
            self.setup()
        cfg = self.cfg
        st = PassStats("photometric" + (" + planar prior" if cfg.planar_prior and not cfg.geom_planar_prior else ""), n_refs=len(self.my_refs))
        self._run_pass(st, 0, False, cfg.planar_prior and not cfg.geom_planar_prior)
        if cfg.geom_iterations > 0:
            self._exchange(st)
        self.stats.append(st)
        for g in range(cfg.geom_iterations):
            planar = cfg.geom_planar_prior and g != cfg.geom_iterations - 1
            st = PassStats(f"geometric {g}" + (" + planar prior" if planar else ""), n_refs=len(self.my_refs))
            self._run_pass(st, 1 + g, True, planar, self.gathered)
            if g + 1 < cfg.geom_iterations and cfg.order != "gauss_seidel":      # Gauss-Seidel updated `gathered` in place
                self._exchange(st)
            self.stats.append(st)
        return self.stats

    def results(self) -> Dict[int, tuple]:
        return {ref: self.engines[ref].result() for ref in self.my_refs}

    def write_results(self, out_folder: str, writers: int = 4):
        """<out>/MPMVS/2333_%08d/{depths,normals,costs}.dmb (PatchMatch.cpp:510-513,620-633), written by background threads
        while the next image's results are still coming off the GPU."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        def write(ref, planes, costs):
            d = self.io.result_dir(out_folder, ref)
            os.makedirs(d, exist_ok=True)
            self.io.write_dmb(os.path.join(d, "depths.dmb"), np.ascontiguousarray(planes[..., 3]))
            self.io.write_dmb(os.path.join(d, "normals.dmb"), np.ascontiguousarray(planes[..., :3]))
            self.io.write_dmb(os.path.join(d, "costs.dmb"), costs)

        with ThreadPoolExecutor(max(1, writers)) as ex:
            futs = []
            for ref in self.my_refs:
                planes, costs = self.engines[ref].result()
                futs.append(ex.submit(write, ref, planes, costs))
            for f in futs:
                f.result()

    def destroy(self):
        for e in self.engines.values():
            e.destroy()
        self.engines.clear()
        if self.cache is not None:
            self.cache.destroy()
            self.cache = None
