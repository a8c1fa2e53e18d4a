"""TEST INFRASTRUCTURE: drives oracle/_ref/libmpmvs_ref_host.so -- the reference's whole program (main.cpp, PatchMatch.cpp,
utility.cpp, PatchMatch.cu compiled where they lie, oracle/ref_host_harness.cu) -- and supplies the OpenCV numerics its
stand-in headers forward: cv2.imread, cv2.resize, cv2.Subdiv2D, cv2.SVDecomp (= cv::SVD::solveZ), i.e. the real OpenCV of
this image (the reference links OpenCV and pins no version, README.md:5).

    python tests/ref_host.py main   <project_dir> [--seed S] [--capture out.npz]     # the reference's main(): needs a GPU
    python tests/ref_host.py fusion <project_dir>                                    # RunFusion alone: CPU only

<project_dir>/config/config.yaml is the reference's configuration file (main.cpp:8). The reference calls exit() on errors,
so callers run this as a child process (run_main / run_fusion below)."""
import argparse
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libmpmvs_ref_host.so")

# the product's seed convention (mp-mvs_b200/csrc/mpmvs_main.cpp: seed + 1000003 i + 7919 stage; prior Run() ^ 0x5DEECE66D)
PER_IMAGE, PER_STAGE, SECOND_RUN_XOR = 1000003, 7919, 0x5DEECE66D

IMREAD = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                          ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p))
RESIZE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p)
SUBDIV = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                          ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int))
SOLVEZ = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float))


def available() -> bool:
    return os.path.exists(LIB)


class RefHost:
    def __init__(self):
        import cv2

        self.cv2 = cv2
        self.lib = ctypes.CDLL(LIB)
        self.keep = {}
        self.cb = (IMREAD(self._guard(self._imread)), RESIZE(self._guard(self._resize)), SUBDIV(self._guard(self._subdiv)),
                   SOLVEZ(self._guard(self._solvez)))
        self.lib.ref_host_set_callbacks(*self.cb)
        self.lib.ref_host_set_seed_plan.argtypes = [ctypes.c_ulonglong] * 4
        self.lib.ref_host_main.argtypes = [ctypes.c_char_p]
        self.lib.ref_host_fusion.argtypes = [ctypes.c_char_p]
        self.lib.ref_host_prior_info.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)]
        self.lib.ref_host_get_prior.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.ref_host_get_prior_input.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        self.counts = {"imread": 0, "resize": 0, "subdiv": 0, "solvez": 0}
        self.solvez = os.environ.get("REF_HOST_SOLVEZ", "cv2")          # cv2 (the reference's own dependency) | numpy32 | numpy64
        if os.environ.get("REF_HOST_IPP", "1") == "0":                  # OpenCV's own resize instead of Intel IPP's (a build without IPP)
            cv2.ipp.setUseIPP(False)

    # ---- the real OpenCV behind the stand-in headers (a Python exception inside a ctypes callback would be swallowed and
    # leave the outputs unset: every callback reports failure through its return value instead, and the stand-in aborts)
    @staticmethod
    def _guard(fn):
        def wrapped(*a):
            try:
                return fn(*a)
            except Exception as e:      # noqa: BLE001
                print("ref_host callback failed:", repr(e), file=sys.stderr, flush=True)
                return 1
        return wrapped

    def _imread(self, path, flags, rows, cols, channels, data):
        self.counts["imread"] += 1
        img = self.cv2.imread(path.decode(), self.cv2.IMREAD_COLOR if flags == 1 else self.cv2.IMREAD_GRAYSCALE)
        if img is None:
            return 1
        img = np.ascontiguousarray(img)
        self.keep["imread"] = img
        rows[0], cols[0] = img.shape[0], img.shape[1]
        channels[0] = 1 if img.ndim == 2 else img.shape[2]
        data[0] = img.ctypes.data
        return 0

    def _resize(self, src, rows, cols, typ, new_rows, new_cols, dst):
        self.counts["resize"] += 1
        depth, cn = typ & 7, (typ >> 3) + 1
        dt = np.uint8 if depth == 0 else np.float32
        n = rows * cols * cn
        a = np.frombuffer((ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(src), dtype=dt).reshape(rows, cols, cn)
        out = self.cv2.resize(a if cn > 1 else a[:, :, 0], (new_cols, new_rows), interpolation=self.cv2.INTER_LINEAR)
        out = np.ascontiguousarray(out, dtype=dt)
        ctypes.memmove(dst, out.ctypes.data, out.nbytes)
        return 0

    def _subdiv(self, rx, ry, rw, rh, pts, npts, tris, ntris):
        self.counts["subdiv"] += 1
        p = np.ctypeslib.as_array(pts, shape=(npts, 2))
        sd = self.cv2.Subdiv2D((rx, ry, rw, rh))
        for x, y in p:                                   # one insert per point, in order (PatchMatch.cpp:767-769)
            sd.insert((float(x), float(y)))
        t = np.ascontiguousarray(sd.getTriangleList(), dtype=np.float32).reshape(-1, 6)
        self.keep["subdiv"] = t
        tris[0] = t.ctypes.data
        ntris[0] = t.shape[0]
        return 0

    def _solvez(self, A, rows, cols, z):
        self.counts["solvez"] += 1
        a = np.ctypeslib.as_array(A, shape=(rows, cols)).astype(np.float32)
        if self.solvez == "cv2":
            # cv::SVD::solveZ = SVD(m, m.rows >= m.cols ? 0 : FULL_UV), last row of vt (modules/core/src/lapack.cpp)
            _, _, vt = self.cv2.SVDecomp(a, flags=0 if rows >= cols else self.cv2.SVD_FULL_UV)
        else:
            # another, equally valid SVD (numpy's LAPACK in float32 or float64): what a reference linked against another OpenCV
            # build would compute up to rounding noise -- used to measure how much of a whole-program difference is that noise
            _, _, vt = np.linalg.svd(a.astype(np.float64 if self.solvez == "numpy64" else np.float32), full_matrices=True)
            if vt[-1] @ self.cv2.SVDecomp(a, flags=self.cv2.SVD_FULL_UV)[2][-1] < 0:      # same sign convention as the default
                vt = -vt
        out = np.ascontiguousarray(vt[-1], dtype=np.float32)
        ctypes.memmove(z, out.ctypes.data, 4 * cols)
        return 0

    # ---- entry points
    def main(self, project, seed=0, capture=False):
        self.lib.ref_host_set_seed_plan(seed, PER_IMAGE, PER_STAGE, SECOND_RUN_XOR)
        self.lib.ref_host_capture_priors(1 if capture else 0)
        return self.lib.ref_host_main(project.encode())

    def fusion(self, project):
        return self.lib.ref_host_fusion(project.encode())

    def priors(self):
        out = []
        for k in range(self.lib.ref_host_num_priors()):
            si, st, px = ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
            self.lib.ref_host_prior_info(k, ctypes.byref(si), ctypes.byref(st), ctypes.byref(px))
            n = px.value
            planes, mask = np.empty((n, 4), np.float32), np.empty(n, np.uint32)
            assert self.lib.ref_host_get_prior(k, planes.ctypes.data, mask.ctypes.data) == 0
            ip, ic, ig = np.empty((n, 4), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
            rc = self.lib.ref_host_get_prior_input(k, ip.ctypes.data, ic.ctypes.data, ig.ctypes.data)
            assert rc in (0, 3), rc
            out.append({"scene_index": si.value, "stage": st.value, "planes": planes, "mask": mask, "in_planes": ip, "in_costs": ic,
                        "in_geom": ig if rc == 0 else None})
        return out


def _child(*args, timeout=3600, env=None):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), *map(str, args)], capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


def run_main(project, seed=0, capture=None, timeout=3600, solvez="cv2", ipp=True):
    """The reference's main() over <project>/config/config.yaml in a child process; `capture`: .npz for the priors it built;
    `solvez`: which SVD serves cv::SVD::solveZ (cv2 = the reference's own dependency); `ipp`: False = OpenCV's own cv::resize
    (what a build without Intel IPP runs; IPP's resize answers up to 3e-3 grey levels / 1 level of 255 differently)."""
    return _child("main", project, "--seed", seed, *(["--capture", capture] if capture else []), timeout=timeout,
                  env={"REF_HOST_SOLVEZ": solvez, "REF_HOST_IPP": "1" if ipp else "0"})


def run_fusion(project, timeout=3600, ipp=True):
    return _child("fusion", project, timeout=timeout, env={"REF_HOST_IPP": "1" if ipp else "0"})


def write_project(project, dense, **cfg):
    """<project>/config/config.yaml for the reference (utility.cpp:8-35): every key it reads, previews off."""
    os.makedirs(os.path.join(project, "config"), exist_ok=True)
    keys = {"Input-folder": dense, "Output-folder": dense, "Geometric consistency iterations": 2, "Planer prior": 0,
            "Geometric consistency planer prior": 0, "Sky segment": 0, "Use dynamic_consistency to fuse": 1, "Save Dmb as JPG": 0,
            "Save Prior Dmb as JPG": 0, "Save Cost Map": 0, "Save Normal Map": 0, "Max source images num": 20, "Max image size": 3200}
    keys.update(cfg)
    path = os.path.join(project, "config", "config.yaml")
    with open(path, "w") as f:
        f.write("%YAML:1.0\n---\n")
        for k, v in keys.items():
            f.write(f'{k}: "{v}"\n' if isinstance(v, str) else f"{k}: {int(v)}\n")
    return path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["main", "fusion"])
    ap.add_argument("project")
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0)
    ap.add_argument("--capture")
    a = ap.parse_args()
    h = RefHost()
    rc = h.main(a.project, a.seed, capture=bool(a.capture)) if a.what == "main" else h.fusion(a.project)
    if a.capture:
        pri = h.priors()
        flat = {}
        for k, p in enumerate(pri):
            for key, v in p.items():
                if v is not None:
                    flat[f"p{k}_{key}"] = np.asarray(v)
        np.savez_compressed(a.capture, n=len(pri), **flat)
    print("ref_host", a.what, "rc", rc, "opencv calls", h.counts)
    sys.exit(rc)
