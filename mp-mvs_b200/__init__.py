"""mpmvs_b200 -- B200-native PatchMatch depth/normal estimation behind MP-MVS's surface.

The directory is named ``mp-mvs_b200`` (not importable as written); load it with
``__graft_entry__.load_package()`` which registers it as ``mpmvs_b200``.

Only the hot path lives here: ``csrc/`` holds the sm_100a kernels,
the C-ABI (``include/mpmvs_b200.h``) and the C++ ``PatchMatchCUDA`` mirror;
the Python modules are the ctypes binding, the file formats and the synthetic
scene generator used by the tests and bench.py.
"""
from . import io_formats, synth  # noqa: F401

__all__ = ["io_formats", "synth"]
