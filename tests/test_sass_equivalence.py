"""The exact arithmetic of the library (pm_kernels.cu compiled with -DPM_EXACT=1: the pm_exact:: kernels of
mp-mvs_b200/libmpmvs_b200.so) must keep COMPILING to the reference's arithmetic: under --use_fast_math its bit-identity with the reference on the GPU
(tests/test_zz_fidelity_build_gpu.py) depends on which products end up fused into FMAs, and that is only visible in the
SASS. This CPU test disassembles both builds (cuobjdump, no GPU) and compares the floating-point expression trees with
tests/tools/sass_expr.py, so an edit of pm_core.cuh that changes a rounding is caught where there is no GPU."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))

REF = os.path.join(ROOT, "oracle", "_ref", "libmpmvs_ref.so")
LIT = os.path.join(ROOT, "mp-mvs_b200", "libmpmvs_b200.so")
SWEEP = "pm_exact15pm_sweep_kernelILi%dELb0ELb0"      # pm_exact::pm_sweep_kernel<scale, false, false>
NCC_MAP = "pm_exact17pm_ncc_map_kernelILi0ELb0ELb0"
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass(tmp_path_factory):
    if not (os.path.exists(REF) and os.path.exists(LIT) and os.path.exists(CUOBJDUMP)):
        pytest.skip("needs oracle/_ref/libmpmvs_ref.so, the library (build()) and cuobjdump")
    d = tmp_path_factory.mktemp("sass")
    out = {}
    for tag, lib in (("ref", REF), ("lit", LIT)):
        out[tag] = str(d / f"{tag}.sass")
        with open(out[tag], "w") as f:
            subprocess.check_call([CUOBJDUMP, "-sass", lib], stdout=f)
    return out


@pytest.mark.parametrize("scale", [0, 1, 2])
def test_source_coordinates_are_the_reference_expression(sass, scale):
    """ComputeHomography + ComputeCorrespondingPoint + 0.5 (cu:228-288,375-378): all 36 unrolled source fetches of our sweep
    kernel carry the expression tree of the reference's BlackPixelUpdate (identical, or identical once the plane distance
    that is still in registers at that point is taken as the leaf it is in the reference's first two inlined copies)."""
    import sass_expr

    ours = sass_expr.source_coordinate_trees(sass["lit"], SWEEP % scale)
    assert len(ours) == 36
    # the relative pose R_s R_r^T, R_s (C_r - C_s) is hoisted in the product (formed once per problem with the same
    # operations, pm_views.h: pm_view_prep) and arrives as loaded values: those subtrees of the reference count as leaves
    refs = sass_expr.source_coordinate_trees(sass["ref"], "BlackPixelUpdate", fold=True)
    assert len(refs) >= 3                                    # cost vector, current plane, refinement: three inlined copies
    assert set(refs) == set(sass_expr.source_coordinate_trees(sass["ref"], "RedPixelUpdate", fold=True))
    assert sass_expr.compare(sass["ref"], "BlackPixelUpdate", sass["lit"], SWEEP % scale)


def test_ncc_test_hook_matches_too(sass):
    import sass_expr

    ref = sass_expr.source_coordinate_trees(sass["ref"], "RefNccMap", fold=True)
    ours = sass_expr.source_coordinate_trees(sass["lit"], NCC_MAP)
    assert len(ours) == 36 and len(set(ref)) == 1 and set(ours) == set(ref)


def test_no_rsqrt_in_the_ncc_tail(sass):
    """1 / sqrt(var_r var_s) must stay MUFU.SQRT followed by MUFU.RCP as in the reference (cu:411-412 compiles to that),
    not one MUFU.RSQ: the test hook kernel contains the NCC and nothing else that takes a square root."""
    import sass_expr

    body = [ins for _, ins in sass_expr.function_body(sass["lit"], NCC_MAP)]
    assert not any("MUFU.RSQ" in i for i in body)
    assert any("MUFU.SQRT" in i for i in body) and any("MUFU.RCP" in i for i in body)
    ref = [ins for _, ins in sass_expr.function_body(sass["ref"], "RefNccMap")]
    assert not any("MUFU.RSQ" in i for i in ref)


def test_tap_loop_uses_blackwell_packed_fp32(sass):
    """PM_PACKED (pm_core.cuh): two taps share one FFMA2 / FADD2 / FMUL2 per numerator, coordinate and moment -- every lane of a
    packed instruction is the reference's own rounded operation (the coordinate trees above are read THROUGH the packed
    instructions), so the results stay bit-identical with ~16 % fewer issue slots in the NCC."""
    import sass_expr

    body = [ins for _, ins in sass_expr.function_body(sass["lit"], SWEEP % 0)]
    n = {k: sum(1 for i in body if i.split()[0].startswith(k)) for k in ("FFMA2", "FADD2", "FMUL2")}
    assert n["FFMA2"] >= 100 and n["FADD2"] >= 36 and n["FMUL2"] >= 18, n


# md5 of the instruction text of the exact arithmetic's pipeline kernels (pm_exact::*) (sass_expr.pipeline_checksum) as they were when
# tests/test_zz_fidelity_build_gpu.py measured them bit-identical to the reference on a B200 (nvcc 12.9.86, sm_100a).
VERIFIED_ON_GPU = "4f3b05a0ddfbac5c69cee218152f8927"    # profiles/r02_fullsize_exact_packed.json, r02_gpu_tests_fid_v3.log, r02_fullsize_geom_bisect_v2.json


def test_fidelity_kernels_are_the_ones_verified_on_the_gpu(sass):
    """Any edit that changes an instruction of the exact arithmetic's sweep / init / finalize kernels lands here first: run
    tests/test_zz_fidelity_build_gpu.py on a B200 again and, if it is still bit-identical, record the new checksum
    (`python tests/tools/sass_expr.py md5 <cuobjdump -sass of the library>`). Comments and host code do not change it."""
    import sass_expr

    nvcc = subprocess.run([os.path.join(os.path.dirname(CUOBJDUMP), "nvcc"), "--version"], capture_output=True, text=True).stdout
    if "V12.9.86" not in nvcc:
        pytest.skip("recorded for nvcc 12.9.86; another compiler schedules differently")
    assert sass_expr.pipeline_checksum(sass["lit"]) == VERIFIED_ON_GPU
