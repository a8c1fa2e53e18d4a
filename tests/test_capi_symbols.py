"""The C-ABI shared library loads and exports every symbol include/mpmvs_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu

from mpmvs_b200 import capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpmvs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpmvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    capi.build()
    L = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mpmvs_b200.h but not exported"


def test_binding_covers_header():
    L = capi.lib()
    for n in declared_symbols():
        fn = getattr(L, n)
        assert fn.argtypes is not None or n in ("mpmvs_version",), f"capi.py does not bind {n}"


def test_camera_struct_is_binary_compatible():
    from mpmvs_b200 import io_formats

    assert io_formats.CAMERA_DTYPE.itemsize == 112  # sizeof(struct Camera), /root/reference/include/PatchMatch.h:35-46


def test_error_strings_and_version():
    L = capi.lib()
    assert L.mpmvs_version() >= 100
    assert capi.build_flavor() == "exact+fast"   # ONE library with both arithmetics; a handle picks at run time
    assert L.mpmvs_arithmetic_name(0) == b"exact" and L.mpmvs_arithmetic_name(1) == b"fast"
    assert capi.default_arithmetic() == os.environ.get("MPMVS_ARITHMETIC", "exact")
    assert b"no CUDA device" in L.mpmvs_error_string(-2)
    assert L.mpmvs_error_string(0) == b"ok"


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_gpu():
    """There is no CPU fallback: on a GPU-less box create() must fail with MPMVS_E_NO_DEVICE, not compute something."""
    with pytest.raises(capi.MpmvsError) as e:
        capi.PatchMatch(device=0)
    assert "no CUDA device" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product sources must not include, link or load anything under oracle/ (or the test emulation)."""
    words = ("pm_oracle", "oracle_py", "prior_oracle", "fusion_oracle", "sky_oracle", "libmpmvs_ref", "pm_emul")
    for top in ("mp-mvs_b200", "tools", "include"):          # the package, the C ABI and the product-side tools
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    for w in words:
                        assert w not in text, (os.path.join(dirpath, f), w)
