"""ctypes front-end for the two test oracles (TEST INFRASTRUCTURE -- never imported by the product).

* ``Oracle("cpu")``  -> oracle/libpm_oracle.so      (pm_oracle.c, plain-C restatement, runs anywhere)
* ``Oracle("ref")``  -> oracle/_ref/libmpmvs_ref.so (the reference's own CUDA path, needs a GPU)

Both libraries export the same entry points with a ``pmo_`` / ``ref_`` prefix, so one class drives
either. The calling sequence mirrors ProcessProblem (/root/reference/src/PatchMatch.cpp:506-638).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBS = {
    "cpu": (os.path.join(HERE, "libpm_oracle.so"), "pmo_"),
    "ref": (os.path.join(HERE, "_ref", "libmpmvs_ref.so"), "ref_"),
    # host instantiation of the product's pm_core.cuh templates, built by tests/emul (debug aid, not an oracle)
    "emul": (os.path.join(os.path.dirname(HERE), "tests", "emul", "libpm_emul.so"), "emu_"),
}


def build(which: str = "cpu") -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, which])


def available(which: str) -> bool:
    return os.path.exists(LIBS[which][0])


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class Oracle:
    def __init__(self, which: str = "cpu"):
        path, self.pfx = LIBS[which]
        if which == "cpu" and not os.path.exists(path):
            build("cpu")
        self.lib = C.CDLL(path)
        self.which = which
        f = self._f
        f("create").restype = C.c_void_p
        f("run").restype = C.c_float
        f("run").argtypes = [C.c_void_p, C.c_uint64]
        f("depth_min").restype = C.c_float
        f("depth_max").restype = C.c_float
        for name in ("depth_min", "depth_max", "destroy", "set_planar_prior_params", "finalize"):
            f(name).argtypes = [C.c_void_p]
        f("set_geom_consistency_params").argtypes = [C.c_void_p, C.c_int, C.c_int]
        f("set_problem").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]
        f("set_src_depths").argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        f("init_only").argtypes = [C.c_void_p, C.c_uint64]
        f("half_sweep").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        f("set_state").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("set_prior").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("get_result").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f("get_device_state").argtypes = [C.c_void_p] + [C.c_void_p] * 5
        f("set_device_state").argtypes = [C.c_void_p] + [C.c_void_p] * 5
        f("ncc_map").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        f("geom_map").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("uniform_stream").argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        assert f("sizeof_camera")() == 112
        self.h = None
        self.n = self.w = self.hgt = 0
        self._keep = []

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    # ------------------------------------------------------------------ problem set-up
    def set_problem(self, images, cams_packed: np.ndarray):
        """images: list of float32 (h,w) arrays, [0] = reference; cams_packed: (n,) CAMERA_DTYPE."""
        if self.h:
            self.destroy()
        self.h = C.c_void_p(self._f("create")())
        imgs = [np.ascontiguousarray(i, dtype=np.float32) for i in images]
        ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
        cams = np.ascontiguousarray(cams_packed)
        assert cams.itemsize == 112 and len(cams) == len(imgs)
        self._keep = [imgs, cams]
        self.n = len(imgs)
        self.hgt, self.w = imgs[0].shape
        rc = self._f("set_problem")(self.h, self.n, ptrs, cams.ctypes.data)
        assert rc == 0
        return self

    def set_geom_consistency_params(self, geom: bool, planar: bool):
        self._f("set_geom_consistency_params")(self.h, int(geom), int(planar))

    def set_planar_prior_params(self):
        self._f("set_planar_prior_params")(self.h)

    def set_src_depths(self, depths):
        d = [np.ascontiguousarray(x, dtype=np.float32) for x in depths]
        assert len(d) == self.n - 1
        ptrs = (C.c_void_p * len(d))(*[x.ctypes.data for x in d])
        self._keep.append(d)
        assert self._f("set_src_depths")(self.h, ptrs) == 0

    def set_state(self, planes4, costs):
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        c = np.ascontiguousarray(costs, dtype=np.float32)
        assert p.shape == (self.hgt, self.w, 4) and c.shape == (self.hgt, self.w)
        self._f("set_state")(self.h, p.ctypes.data, c.ctypes.data)

    def set_prior(self, prior4, mask):
        p = np.ascontiguousarray(prior4, dtype=np.float32)
        m = np.ascontiguousarray(mask, dtype=np.uint32)
        self._f("set_prior")(self.h, p.ctypes.data, m.ctypes.data)

    # ------------------------------------------------------------------ running
    def run(self, seed: int) -> float:
        return float(self._f("run")(self.h, C.c_uint64(seed)))

    def init_only(self, seed: int):
        self._f("init_only")(self.h, C.c_uint64(seed))

    def half_sweep(self, red: int, it: int, scale: int):
        self._f("half_sweep")(self.h, int(red), int(it), int(scale))

    def finalize(self):
        self._f("finalize")(self.h)

    def result(self, geom=False):
        planes = np.empty((self.hgt, self.w, 4), np.float32)
        costs = np.empty((self.hgt, self.w), np.float32)
        g = np.empty((self.hgt, self.w), np.float32) if geom else None
        self._f("get_result")(self.h, planes.ctypes.data, costs.ctypes.data, g.ctypes.data if geom else None)
        return (planes, costs, g) if geom else (planes, costs)

    def get_state(self):
        s = dict(
            planes=np.empty((self.hgt, self.w, 4), np.float32),
            costs=np.empty((self.hgt, self.w), np.float32),
            views=np.empty((self.hgt, self.w), np.uint32),
            rng=np.empty((self.hgt, self.w, 6), np.uint32),
            geom=np.empty((self.hgt, self.w), np.float32),
        )
        self._f("get_device_state")(self.h, *[s[k].ctypes.data for k in ("planes", "costs", "views", "rng", "geom")])
        return s

    def set_dev_state(self, s):
        arrs = []
        for k, dt in (("planes", np.float32), ("costs", np.float32), ("views", np.uint32), ("rng", np.uint32), ("geom", np.float32)):
            a = s.get(k)
            arrs.append(None if a is None else np.ascontiguousarray(a, dtype=dt))
        self._f("set_device_state")(self.h, *[None if a is None else a.ctypes.data for a in arrs])

    def ncc_map(self, planes4, scale: int):
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        out = np.empty((self.n - 1, self.hgt, self.w), np.float32)
        self._f("ncc_map")(self.h, p.ctypes.data, int(scale), out.ctypes.data)
        return out

    def geom_map(self, planes4):
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        out = np.empty((self.n - 1, self.hgt, self.w), np.float32)
        assert self._f("geom_map")(self.h, p.ctypes.data, out.ctypes.data) == 0
        return out

    def uniform_stream(self, seed: int, x: int, y: int, n: int):
        out = np.empty(n, np.float32)
        self._f("uniform_stream")(C.c_uint64(seed), x, y, n, out.ctypes.data)
        return out

    @property
    def depth_range(self):
        return float(self._f("depth_min")(self.h)), float(self._f("depth_max")(self.h))

    def destroy(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
