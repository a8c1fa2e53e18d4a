"""Host-side logic of bench.py that does not need a GPU: the reference arm's `unavailable` line, the software pipeline over
handles, the clock sampler's degraded mode, the roofline constants."""
import argparse
import sys
import threading

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


class FakeHandle:
    def __init__(self, name, log):
        self.name, self.log = name, log

    def reset_params(self): self.log.append((self.name, "reset"))
    def set_geom_consistency_params(self, g, p): pass
    def set_planar_prior_params(self): pass
    def run_async(self, seed): self.log.append((self.name, "run", seed))
    def build_prior(self): self.log.append((self.name, "prior")); return {"delaunay_ms": 1.0, "total_ms": 2.0}
    def synchronize(self): pass
    def last_run_profile(self, timing=True, count=False): return {"sweep_ms": 1.0, "init_ms": 0.1, "finalize_ms": 0.1, "n_sweeps": 18}


def test_process_problems_runs_every_step_once_per_handle_in_order():
    log = []
    lock = threading.Lock()

    class L(list):
        def append(self, x):
            with lock:
                super().append(x)

    log = L()
    handles = [FakeHandle("A", log), FakeHandle("B", log)]
    stats = bench.process_problems(handles, 5, 100)
    assert len(stats) == 5 and all(s is not None for s in stats)
    for name, steps in (("A", [0, 2, 4]), ("B", [1, 3])):
        mine = [e for e in log if e[0] == name]
        runs = [e[2] for e in mine if e[1] == "run"]
        # per image: photometric run with seed s, prior stage, prior run with seed s + 1 -- strictly in that order per handle
        assert runs == [v for i in steps for v in (100 + 2 * i, 100 + 2 * i + 1)]
        kinds = [e[1] for e in mine]
        assert kinds == ["reset", "run", "prior", "run"] * len(steps)
    # one handle = the reference's sequential order, with the per-run profile collected
    log2 = L()
    prof = bench.Prof(timing=True)
    bench.process_problems([FakeHandle("S", log2)], 2, 7, prof=prof)
    assert [e[1] for e in log2] == ["reset", "run", "prior", "run"] * 2 and len(prof) == 4


def test_reference_arm_reports_unavailable(monkeypatch, tmp_path):
    sys.path.insert(0, ROOT + "/oracle")
    import oracle_py

    monkeypatch.setitem(oracle_py.LIBS, "ref", (str(tmp_path / "missing.so"), "ref_"))
    out = bench.run_reference(argparse.Namespace(steps=1, warmup=1, workload="eth3d"), {"width": 8, "height": 8, "images": [None] * 3})
    assert out["impl"] == "reference" and "unavailable" in out


def test_clock_sampler_without_nvidia_smi(monkeypatch):
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert "reasons" in out and out.get("sm_mhz") is None


def test_roofline_constants():
    assert bench.TAPS_PER_NCC == 36 and bench.BYTES_PER_TAP == 16 and bench.N_SM == 148 and bench.TEX_PER_CLK_SM == 4
    assert bench.ncu_traffic("eth3d") == pytest.approx(920.5e6, rel=1e-3)          # exact arithmetic, float32 views
    assert bench.ncu_traffic("eth3d", "fast", "u8") == pytest.approx(684e6)
    assert bench.ncu_traffic("dtu") is None


def test_fast_arm_never_takes_the_headline_down():
    """`arithmetic_fast` is measured after the headline; whatever happens to it (here: no GPU) must come back as an `error`
    entry, not as an exception in the process that prints the headline line."""
    import types

    from conftest import has_gpu

    if has_gpu():
        pytest.skip("only meaningful on a box without a GPU")
    sys.path.insert(0, ROOT)
    import bench

    prob = {"width": 8, "height": 8, "images": [None] * 3}
    res = bench.fast_arm(types.SimpleNamespace(steps=1, warmup=1, workload="plane", in_flight=2, arithmetic="exact", tex="f32"), 0, 1, 0, prob,
                         lambda: None, lambda x: x)
    assert set(res) == {"error"} and res["error"]
