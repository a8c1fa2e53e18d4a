"""The sharded stage schedule (mp-mvs_b200/pipeline.py) on CPU with the gloo backend, world_size 2.

The compute engine here is the CPU oracle (a test stand-in injected through `engine_factory`; the product default is the
CUDA engine). What is tested is the host logic the multi-GPU path adds: sharding by reference image, slot bookkeeping of
the all-gather buffer, Jacobi double buffering, per-(image, pass) seeds -- so that N ranks produce bit-identical results
to 1 rank.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT

from mpmvs_b200 import io_formats, pipeline


class OracleEngine:
    def __init__(self, ids, images, cams_packed):
        import oracle_py

        self.o = oracle_py.Oracle("cpu").set_problem([np.asarray(i, np.float32) for i in images], cams_packed)
        self.ms = 0.0

    def process(self, seed, geom, planar, src_depths=None):
        assert not planar, "the CPU stand-in has no planar-prior stage"
        self.o.set_geom_consistency_params(geom, False)
        if geom:
            self.o.set_src_depths([d.numpy() for d in src_depths])
        self.o.run(seed)

    def export_depth(self, dst):
        dst.copy_(torch.from_numpy(self.o.result()[0][..., 3].copy()))

    def device_ms(self):
        return 0.0

    def synchronize(self):
        pass

    def result(self):
        return self.o.result()

    def destroy(self):
        self.o.destroy()


def make_inputs():
    sc = PKG.synth.make_dtu_scene(width=48, height=36, grid=2, n_src=2, seed=2, jpeg=False)     # 4 views
    entries = [io_formats.SceneEntry(ref_id=i, src_ids=[i] + [j for j, _ in sc.pairs[i]], estimate=True) for i in range(sc.num_views)]
    cams = {i: c for i, c in enumerate(sc.cams)}
    images = {i: im for i, im in enumerate(sc.images)}
    return entries, cams, images


def run_pipeline(rank, world, d=None):
    entries, cams, images = make_inputs()
    cfg = pipeline.PipelineConfig(geom_iterations=2, max_src=2, seed=77)
    p = pipeline.DensePipeline(entries, cams, images, cfg, rank=rank, world=world, dist=d,
                               engine_factory=lambda ids, imgs, packed: OracleEngine(ids, imgs, packed), torch_device=torch.device("cpu"))
    p.run()
    res = p.results()
    p.destroy()
    return res


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PMO_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = run_pipeline(rank, world, dist)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **{f"p{r}": v[0] for r, v in res.items()}, **{f"c{r}": v[1] for r, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_helpers():
    refs = list(range(10))
    blocks = [pipeline.shard_refs(refs, r, 4) for r in range(4)]
    assert sum(blocks, []) == refs and [len(b) for b in blocks] == [3, 3, 3, 1]
    assert pipeline.block_size(300, 8) == 38
    assert pipeline.stage_seed(1, 5, 0) != pipeline.stage_seed(1, 5, 1) != pipeline.stage_seed(1, 6, 0)
    assert pipeline.stage_seed(1, 5, 0) == pipeline.stage_seed(1, 5, 0) < 2 ** 64


def test_two_ranks_match_one_rank(tmp_path, oracle_cpu):
    single = run_pipeline(0, 1)
    assert sorted(single) == [0, 1, 2, 3]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    merged = {}
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        for k in z.files:
            merged[k] = z[k]
    assert sorted(k for k in merged if k.startswith("p")) == ["p0", "p1", "p2", "p3"]
    for ref, (planes, costs) in single.items():
        np.testing.assert_array_equal(merged[f"p{ref}"], planes)      # bit-identical to the 1-rank schedule
        np.testing.assert_array_equal(merged[f"c{ref}"], costs)
    # the geom passes really consumed exchanged maps: results differ from a photometric-only run
    entries, cams, images = make_inputs()
    p0 = pipeline.DensePipeline(entries, cams, images, pipeline.PipelineConfig(geom_iterations=0, max_src=2, seed=77),
                                engine_factory=lambda ids, imgs, packed: OracleEngine(ids, imgs, packed), torch_device=torch.device("cpu"))
    p0.run()
    photo = p0.results()
    p0.destroy()
    assert any(not np.array_equal(photo[r][0], single[r][0]) for r in single)


def test_gauss_seidel_order_is_the_references_in_place_exchange(oracle_cpu):
    """PipelineConfig.order = "gauss_seidel" (one GPU): inside a geometric pass the images run one after the other and each
    reads the depth maps the images before it have ALREADY written in this pass -- what the reference does through its
    depths.dmb files (main.cpp:35-40, PatchMatch.cpp:620-633, 934-950). Checked against a plain sequential loop."""
    entries, cams, images = make_inputs()
    cfg = pipeline.PipelineConfig(geom_iterations=2, max_src=2, seed=77, order="gauss_seidel")
    p = pipeline.DensePipeline(entries, cams, images, cfg, engine_factory=lambda ids, imgs, packed: OracleEngine(ids, imgs, packed),
                               torch_device=torch.device("cpu"))
    p.run()
    got = p.results()
    p.destroy()
    # the same schedule by hand
    eng = {e.ref_id: OracleEngine(e.src_ids[:3], [images[i] for i in e.src_ids[:3]], io_formats.pack_cameras([cams[i] for i in e.src_ids[:3]]))
           for e in entries}
    for r in sorted(eng):
        eng[r].process(pipeline.stage_seed(77, r, 0), False, False)
    maps = {r: torch.from_numpy(eng[r].result()[0][..., 3].copy()) for r in eng}
    for g in range(2):
        for r in sorted(eng):
            srcs = [e for e in entries if e.ref_id == r][0].src_ids[1:3]
            eng[r].process(pipeline.stage_seed(77, r, 1 + g), True, False, [maps[j] for j in srcs])
            maps[r] = torch.from_numpy(eng[r].result()[0][..., 3].copy())        # in place: the next image sees it
    for r in sorted(eng):
        np.testing.assert_array_equal(got[r][0], eng[r].result()[0])
        np.testing.assert_array_equal(got[r][1], eng[r].result()[1])
        eng[r].destroy()
    jac = run_pipeline(0, 1)
    assert any(not np.array_equal(jac[r][0], got[r][0]) for r in got)          # and it is not the Jacobi result
    with pytest.raises(ValueError):
        pipeline.DensePipeline(entries, cams, images, cfg, rank=0, world=2, engine_factory=lambda *a: None,
                               torch_device=torch.device("cpu"))._run_pass(pipeline.PassStats("x"), 1, True, False, None)


def test_gpu_turn_is_fifo():
    """pipeline.GpuTurn: whole Run()s take turns on a GPU in the order the host threads asked for them."""
    import threading
    import time

    from mpmvs_b200.pipeline import GpuTurn

    turn, order, inside = GpuTurn(), [], []
    gate = threading.Event()

    def worker(k):
        gate.wait()
        time.sleep(0.01 * k)                 # ask in a known order
        with turn:
            inside.append(k)
            assert len(inside) == 1          # exclusive
            time.sleep(0.02)
            order.append(k)
            inside.pop()

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in ts:
        t.start()
    gate.set()
    for t in ts:
        t.join()
    assert order == list(range(6))
    # an exception inside a turn releases it
    try:
        with turn:
            raise ValueError("boom")
    except ValueError:
        pass
    with turn:
        pass


def test_write_results_layout(tmp_path, oracle_cpu):
    """DensePipeline.write_results: <out>/MPMVS/2333_%08d/{depths,normals,costs}.dmb written by background threads, holding
    exactly what results() returns (the layout of PatchMatch.cpp:510-513,620-633)."""
    entries, cams, images = make_inputs()
    p = pipeline.DensePipeline(entries, cams, images, pipeline.PipelineConfig(geom_iterations=0, max_src=2, seed=5),
                               engine_factory=lambda ids, imgs, packed: OracleEngine(ids, imgs, packed), torch_device=torch.device("cpu"))
    p.run()
    res = p.results()
    p.write_results(str(tmp_path), writers=3)
    p.destroy()
    for ref, (planes, costs) in res.items():
        d = io_formats.result_dir(str(tmp_path), ref)
        assert d.endswith(f"MPMVS/2333_{ref:08d}")
        np.testing.assert_array_equal(io_formats.read_dmb(os.path.join(d, "depths.dmb")), planes[..., 3])
        np.testing.assert_array_equal(io_formats.read_dmb(os.path.join(d, "normals.dmb")), planes[..., :3])
        np.testing.assert_array_equal(io_formats.read_dmb(os.path.join(d, "costs.dmb")), costs)
