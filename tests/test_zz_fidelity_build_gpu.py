"""GPU tests of the EXACT arithmetic of the library (MPMVS_ARITH_EXACT, the default of every handle; namespace pm_exact of
mp-mvs_b200/csrc/pm_core.cuh, compiled into libmpmvs_b200.so next to the fast arithmetic): the reference's own operations,
BIT-IDENTICAL to the reference's kernels (/root/reference/src/PatchMatch.cu compiled in place into oracle/_ref) in all
three modes -- photometric, planar prior, geometric consistency -- after every launch and over whole runs, up to the
metric's full size. Kept in a file of its own that sorts last: whatever happens here cannot hide the other GPU tests under
`pytest -x`. Measured on B200: profiles/r02_fidelity_exact.json, r02_modes_bisect_exact.json, r02_fullsize_exact.json."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    import oracle_py

    return oracle_py


def need_ref(oracle):
    if not oracle.available("ref"):
        pytest.skip("oracle/_ref/libmpmvs_ref.so not built on this box")


def tool(name, *args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", name), *args], env=dict(os.environ, **(env or {})),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_ex2_is_monotone():
    """The early-out of refinement proposals under the planar prior (PM_PRIOR_EARLY_OUT, pm_core.cuh) argues that the
    acceptance test is monotone in the running cost; the one step of that chain that is hardware, MUFU.EX2, is checked here
    for every float in [-160, -0] (1.13e9 adjacent pairs)."""
    from mpmvs_b200 import capi

    assert capi.selftest_ex2_monotone(0) == 0


def test_exact_arithmetic_is_bit_identical_to_the_reference_photometric(oracle):
    """Started from the reference's state, EVERY half-sweep of a photometric Run() reproduces the reference's planes, costs and
    view masks bit for bit, and so does a whole same-seed Run() -- on the three parity cases."""
    need_ref(oracle)
    data = tool("variant_fidelity.py", env={"MPMVS_ARITHMETIC": "exact"})
    assert data["arithmetic"] == "exact", data
    print("exact arithmetic, photometric:", json.dumps(data["cases"]))
    for name, c in data["cases"].items():
        assert c["half_sweep_planes_min"] == 1.0 and c["half_sweep_costs_min"] == 1.0 and c["half_sweep_views_min"] == 1.0, (name, c)
        assert c["run_planes_identical"] == 1.0 and c["run_costs_identical"] == 1.0, (name, c)


def test_exact_arithmetic_is_bit_identical_in_the_prior_and_geom_modes(oracle):
    """The planar-prior mode (InitializeScore's prior branch cu:552-562, the prior selection cu:924-978, refinement under the
    prior cu:656-711) and the geometric-consistency mode (cu:617-640 inside candidate scoring cu:880-913, the current plane
    and refinement cu:684-694): InitializeScore, every half-sweep from the reference's state, the finalize kernels and whole
    same-seed Run()s -- planes, costs, view masks, RNG states and geometric costs all bit-identical, on the three cases."""
    need_ref(oracle)
    data = tool("modes_bisect.py", env={"MPMVS_ARITHMETIC": "exact"})
    assert data["arithmetic"] == "exact", data
    slim = {m: {n: {k: {a: b for a, b in v.items() if a != "examples"} for k, v in c.items()} for n, c in data[m].items()} for m in ("prior", "geom")}
    print("exact arithmetic, other modes:", json.dumps(slim))
    for mode in ("prior", "geom"):
        for name, stages in data[mode].items():
            for stage, c in stages.items():
                for key, v in c.items():
                    if key in ("examples", "n_diff", "max_abs_diff"):
                        continue
                    assert v == 1.0, (mode, name, stage, key, v)
                assert c.get("n_diff", 0) == 0 and c.get("max_abs_diff", 0.0) == 0.0, (mode, name, stage, c)


def test_exact_arithmetic_at_full_size_whole_bench_step(oracle):
    """The metric's full size (3200x2130, 10 sources): a same-seed photometric Run() -- all 6.8 M planes and costs -- and then
    the planar-prior Run() of the bench step with the same prior, both bit-identical to the reference's, at less than 0.6 of
    the reference's device time."""
    need_ref(oracle)
    full = tool("variant_fullsize.py", "--check-ref", env={"MPMVS_ARITHMETIC": "exact"})
    print("exact arithmetic at full size:", json.dumps(full))
    assert full["arithmetic"] == "exact"
    assert full["run_planes_identical"] == 1.0 and full["run_costs_identical"] == 1.0 and full["max_abs_depth_diff"] == 0.0, full
    assert full["prior_run_planes_identical"] == 1.0 and full["prior_run_costs_identical"] == 1.0, full
    assert full["photometric_run_ms"] < 0.6 * full["reference_run_ms"], full      # r01: 398 ms against 1 062 ms
