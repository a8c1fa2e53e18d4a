"""GPU tests of the FIDELITY BUILD (mp-mvs_b200/variants/libmpmvs_b200_literal2.so = the same sources with
-DPM_LITERAL_NCC=2, built by `make exact` / __graft_entry__.build()): the reference's own operation order, bit-identical to
the reference's kernels (/root/reference/src/PatchMatch.cu compiled in place into oracle/_ref). The library is chosen at
import time, so the build under test runs in child processes (tests/tools/variant_fidelity.py, variant_fullsize.py).
Kept in a file of its own that sorts last: whatever happens here cannot hide the other GPU tests under `pytest -x`.
Measured on B200: profiles/r01_fidelity_literal2.json, r01_fullsize_literal2.json, r01_literal_variant.md."""
import os

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


def test_literal2_build_is_bit_identical_to_the_reference(oracle):
    """-DPM_LITERAL_NCC=2 (variant library literal2, built by __graft_entry__.build()): the reference's own arithmetic with
    unrolled taps and pinned roundings. Started from the reference's state, EVERY half-sweep reproduces the reference's
    planes, costs and view masks bit for bit, and so does a whole same-seed Run() -- on the three parity cases and at the
    metric's full size (3200x2130, 10 sources: all 6.8 M planes and costs). Measured on B200: profiles/r01_fidelity_literal2.json,
    r01_fullsize_literal2.json (the shipped kernels: 89-98 % of the planes per half-sweep, r01_fidelity_shipped.json).
    The library is chosen at import time, so the build under test runs in its own process."""
    import json
    import subprocess
    import sys

    need_ref(oracle)
    lib = os.path.join(ROOT, "mp-mvs_b200", "variants", "libmpmvs_b200_literal2.so")
    if not os.path.exists(lib):
        pytest.skip("variant library not built (python tools/build_variants.py literal2)")
    env = dict(os.environ, MPMVS_LIB_VARIANT="literal2")

    def tool(name, *args):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", name), *args], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])

    data = tool("variant_fidelity.py")
    assert data["arithmetic"] == "literal2", data["library"]        # the child really loaded the fidelity build
    print("literal2 fidelity:", json.dumps(data["cases"]))
    for name, c in data["cases"].items():
        assert c["half_sweep_planes_min"] == 1.0 and c["half_sweep_costs_min"] == 1.0 and c["half_sweep_views_min"] == 1.0, (name, c)
        assert c["run_planes_identical"] == 1.0 and c["run_costs_identical"] == 1.0, (name, c)
    full = tool("variant_fullsize.py", "--check-ref")
    print("literal2 at full size:", json.dumps(full))
    assert full["run_planes_identical"] == 1.0 and full["run_costs_identical"] == 1.0 and full["max_abs_depth_diff"] == 0.0, full
    assert full["photometric_run_ms"] < 0.6 * full["reference_run_ms"], full      # measured 398 ms against 1 062 ms


@pytest.mark.xfail(strict=False, reason="the planar-prior and geometric-consistency modes of the fidelity build are bit-identical to the "
                                        "oracle on the CPU and compile to the reference's instruction mix, but the round's GPU budget ended "
                                        "before they could be run: this is their first run; the suite does not depend on it")
def test_literal2_prior_and_geom_runs_bit_identical(oracle):
    import json
    import subprocess
    import sys

    need_ref(oracle)
    lib = os.path.join(ROOT, "mp-mvs_b200", "variants", "libmpmvs_b200_literal2.so")
    if not os.path.exists(lib):
        pytest.skip("variant library not built (make -C mp-mvs_b200/csrc exact)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "variant_fidelity.py"), "--modes"],
                       env=dict(os.environ, MPMVS_LIB_VARIANT="literal2"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    modes = json.loads(r.stdout.strip().splitlines()[-1])["modes"]
    print("literal2, other modes:", json.dumps(modes))
    for name, c in modes.items():
        assert all(v == 1.0 for v in c.values()), (name, c)
