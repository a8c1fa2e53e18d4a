"""ctypes binding of the C ABI (include/mpmvs_b200.h) -- the Python host side of the product.

``PatchMatch`` mirrors the calling sequence ProcessProblem applies to a PatchMatchCUDA object
(/root/reference/src/PatchMatch.cpp:516-637): set_geom_consistency_params -> set_problem
(PatchMatchInit + AllocatePatchMatch + CudaMemInit) -> run -> [set_planar_prior_params,
set_geom_consistency_params(False, True), set_prior, run] -> result -> destroy (Release).

The compute path is the CUDA library only: loading fails loudly when libmpmvs_b200.so is missing,
and every call raises if the library reports an error (e.g. MPMVS_E_NO_DEVICE on a box without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmpmvs_b200.so")
# tuning experiments: tools/build_variants.py drops alternative builds of the same library next to the default one
if os.environ.get("MPMVS_LIB_VARIANT"):
    LIB_PATH = os.path.join(PKG_DIR, "variants", f"libmpmvs_b200_{os.environ['MPMVS_LIB_VARIANT']}.so")
TEX_F32, TEX_F16, TEX_U8 = 0, 1, 2
ARITH_EXACT, ARITH_FAST = 0, 1      # MPMVS_ARITH_*: both arithmetics are in the one library, a handle picks at run time

_lib = None


class MpmvsError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-s", "-j", str(min(6, os.cpu_count() or 1)), "-C", os.path.join(PKG_DIR, "csrc")]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpmvsError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (run __graft_entry__.build()). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, u64, i, fp = C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)
    sig = {
        "mpmvs_create": [i, vp, C.POINTER(vp)],
        "mpmvs_destroy": [vp],
        "mpmvs_set_views": [vp, i, C.POINTER(vp), vp],
        "mpmvs_set_views_device": [vp, i, C.POINTER(vp), vp, vp],
        "mpmvs_cache_create": [i, i, i, i, C.POINTER(vp)],
        "mpmvs_cache_create_fmt": [i, i, i, i, i, C.POINTER(vp)],
        "mpmvs_set_views_u8": [vp, i, C.POINTER(vp), vp],
        "mpmvs_set_tex_format": [vp, i],
        "mpmvs_cache_destroy": [vp],
        "mpmvs_cache_put": [vp, i, vp, i, i],
        "mpmvs_cache_put_u8": [vp, i, vp, i, i],
        "mpmvs_cache_has": [vp, i],
        "mpmvs_set_views_cached": [vp, vp, i, vp, vp],
        "mpmvs_set_geom_consistency_params": [vp, i, i],
        "mpmvs_set_planar_prior_params": [vp],
        "mpmvs_reset_params": [vp],
        "mpmvs_set_src_depths": [vp, C.POINTER(vp)],
        "mpmvs_set_src_depths_device": [vp, C.POINTER(vp), vp],
        "mpmvs_set_state": [vp, vp, vp],
        "mpmvs_set_prior": [vp, vp, vp],
        "mpmvs_run": [vp, u64],
        "mpmvs_run_async": [vp, u64],
        "mpmvs_run_into": [vp, u64, vp, vp, vp],
        "mpmvs_synchronize": [vp],
        "mpmvs_get_results_async": [vp, vp, vp, vp],
        "mpmvs_last_run_ms": [vp, fp],
        "mpmvs_last_run_launches": [vp, C.POINTER(i)],
        "mpmvs_set_profiling": [vp, i],
        "mpmvs_last_run_profile": [vp, fp, fp, C.POINTER(i), fp, C.POINTER(u64)],
        "mpmvs_get_size": [vp, C.POINTER(i), C.POINTER(i)],
        "mpmvs_get_depth_range": [vp, fp, fp],
        "mpmvs_get_planes": [vp, vp],
        "mpmvs_get_costs": [vp, vp],
        "mpmvs_get_geom_costs": [vp, vp],
        "mpmvs_device_planes": [vp, C.POINTER(vp)],
        "mpmvs_device_costs": [vp, C.POINTER(vp)],
        "mpmvs_export_depth_device": [vp, vp, C.c_size_t],
        "mpmvs_init_only": [vp, u64],
        "mpmvs_half_sweep": [vp, i, i, i],
        "mpmvs_finalize": [vp],
        "mpmvs_get_device_state": [vp, vp, vp, vp, vp, vp],
        "mpmvs_set_device_state": [vp, vp, vp, vp, vp, vp],
        "mpmvs_ncc_map": [vp, vp, i, vp],
        "mpmvs_geom_map": [vp, vp, vp],
        "mpmvs_ncc_bench": [vp, vp, i, i, i, i, fp, C.POINTER(u64)],
        "mpmvs_uniform_stream": [u64, i, i, i, vp],
        "mpmvs_delaunay": [vp, i, i, i, vp, i, C.POINTER(i)],
        "mpmvs_fusion_create": [i, i, C.POINTER(vp)],
        "mpmvs_fusion_destroy": [vp],
        "mpmvs_fusion_set_view": [vp, i, vp, vp, vp, vp],
        "mpmvs_fusion_run": [vp, vp, i, i, C.POINTER(u64), fp],
        "mpmvs_fusion_get_points": [vp, vp, u64],
        "mpmvs_fusion_set_sky_mask": [vp, i, vp],
        "mpmvs_fusion_set_color": [vp, i, vp],
        "mpmvs_sky_mask_refine": [i, vp, vp, i, i, vp, i, i, vp, vp, fp],
        "mpmvs_build_prior": [vp, vp],
        "mpmvs_pick_vertices": [vp, i, vp, i, C.POINTER(i)],
        "mpmvs_prior_from_triangles": [vp, vp, i, vp, i, C.POINTER(i)],
        "mpmvs_get_prior": [vp, vp, vp],
        "mpmvs_get_prior_pixels": [vp, C.POINTER(i)],
        "mpmvs_set_arithmetic": [vp, i],
        "mpmvs_get_arithmetic": [vp, C.POINTER(i)],
        "mpmvs_default_arithmetic": [],
        "mpmvs_selftest_ex2_monotone": [i, C.POINTER(u64)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.mpmvs_error_string.argtypes = [C.c_int]
    L.mpmvs_error_string.restype = C.c_char_p
    L.mpmvs_version.restype = C.c_int
    L.mpmvs_build_flavor.argtypes = []
    L.mpmvs_build_flavor.restype = C.c_char_p
    L.mpmvs_arithmetic_name.argtypes = [C.c_int]
    L.mpmvs_arithmetic_name.restype = C.c_char_p
    _lib = L
    return L


class PriorStats(C.Structure):
    _fields_ = [("n_vertices", C.c_int), ("n_triangles", C.c_int), ("n_prior_pixels", C.c_int), ("pick_ms", C.c_float),
                ("delaunay_ms", C.c_float), ("raster_ms", C.c_float), ("total_ms", C.c_float)]


def build_flavor() -> str:
    """What the loaded library holds: "exact+fast" (both arithmetics, chosen per handle at run time)."""
    return lib().mpmvs_build_flavor().decode()


def selftest_ex2_monotone(device: int = 0) -> int:
    """Adjacent float pairs in [-160, -0] for which MUFU.EX2 is not monotone (0 expected; see mpmvs_selftest_ex2_monotone)."""
    v = C.c_uint64()
    _ck(lib().mpmvs_selftest_ex2_monotone(int(device), C.byref(v)), "selftest_ex2_monotone")
    return int(v.value)


def default_arithmetic() -> str:
    """The arithmetic a new handle starts with: "exact" (bit-identical to the reference's kernels) unless the environment
    has MPMVS_ARITHMETIC=fast."""
    return lib().mpmvs_arithmetic_name(lib().mpmvs_default_arithmetic()).decode()


def _ck(rc: int, what: str):
    if rc != 0:
        raise MpmvsError(f"{what} failed: {lib().mpmvs_error_string(rc).decode()} ({rc})")


def _ptr(a):
    return None if a is None else a.ctypes.data


def delaunay(xy: np.ndarray, width: int, height: int) -> np.ndarray:
    """Delaunay triangulation of integer pixel positions (n, 2) -> (m, 3) vertex indices (host code, no GPU needed)."""
    pts = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
    out = np.empty((2 * len(pts) + 4, 3), np.int32)
    n = C.c_int()
    _ck(lib().mpmvs_delaunay(pts.ctypes.data, len(pts), width, height, out.ctypes.data, len(out), C.byref(n)), "delaunay")
    return out[: n.value].copy()


def sky_mask_refine(bgr, mask, device: int = 0, stream: int | None = None, want_prob: bool = False):
    """bilateral_filter() of GenerateSkyRegionMask (SkySegment/src/SkyRegionDetect.cu:3-66): joint-bilateral upsampling of a
    sky probability map `mask` (float, any size) guided by the colour image `bgr` ([h][w][3] uint8).
    Returns (result [h][w] float32 of 0 / 255, prob or None, device ms)."""
    img = np.ascontiguousarray(bgr, np.uint8)
    m = np.ascontiguousarray(mask, np.float32)
    assert img.ndim == 3 and img.shape[2] == 3 and m.ndim == 2
    h, w = img.shape[:2]
    out = np.empty((h, w), np.float32)
    prob = np.empty((h, w), np.float32) if want_prob else None
    ms = C.c_float()
    _ck(lib().mpmvs_sky_mask_refine(device, C.c_void_p(stream) if stream else None, img.ctypes.data, w, h, m.ctypes.data, m.shape[1],
                                    m.shape[0], out.ctypes.data, prob.ctypes.data if want_prob else None, C.byref(ms)), "sky_mask_refine")
    return out, prob, float(ms.value)


class Fusion:
    """Depth-map fusion of a whole scene on one GPU (mpmvs_fusion_*; RunFusion, PatchMatch.cpp:287-504)."""

    def __init__(self, device: int, n_images: int):
        self.h = C.c_void_p()
        self.n = n_images
        _ck(lib().mpmvs_fusion_create(device, n_images, C.byref(self.h)), "fusion_create")

    def set_view(self, index: int, cam_packed: np.ndarray, depth, normal, gray):
        d = np.ascontiguousarray(depth, np.float32)
        nrm = np.ascontiguousarray(normal, np.float32)
        g = np.ascontiguousarray(gray, np.uint8)
        cam = np.ascontiguousarray(cam_packed)
        assert nrm.shape == d.shape + (3,) and g.shape == d.shape
        _ck(lib().mpmvs_fusion_set_view(self.h, index, cam.ctypes.data, d.ctypes.data, nrm.ctypes.data, g.ctypes.data), "fusion_set_view")

    def set_color(self, index: int, bgr):
        """bgr: (h, w, 3) uint8 in cv2.imread(IMREAD_COLOR) order, at the depth map's size (RunFusion's colour, cpp:322)."""
        b = np.ascontiguousarray(bgr, dtype=np.uint8)
        assert b.ndim == 3 and b.shape[2] == 3
        _ck(lib().mpmvs_fusion_set_color(self.h, int(index), b.ctypes.data), "fusion_set_color")

    def set_sky_mask(self, index: int, sky):
        """sky [h][w] uint8, > 0 = sky (skymask_refine.jpg): masked when that view's turn comes (PatchMatch.cpp:385-388)."""
        m = np.ascontiguousarray(sky, np.uint8)
        _ck(lib().mpmvs_fusion_set_sky_mask(self.h, index, m.ctypes.data), "fusion_set_sky_mask")

    def run(self, src_lists, use_dynamic_consistency: bool = True):
        """src_lists: per image [ref, sources...] (None = not estimated). Returns (points (n, 9) float32, device ms)."""
        width = max(len(r) for r in src_lists if r is not None) + 1
        tab = np.full((self.n, width), -2, np.int32)
        for k, r in enumerate(src_lists):
            if r is None:
                tab[k, 0] = -1
            else:
                tab[k, : len(r)] = r
        n, ms = C.c_uint64(), C.c_float()
        _ck(lib().mpmvs_fusion_run(self.h, tab.ctypes.data, width, int(use_dynamic_consistency), C.byref(n), C.byref(ms)), "fusion_run")
        pts = np.empty((int(n.value), 9), np.float32)
        if len(pts):
            _ck(lib().mpmvs_fusion_get_points(self.h, pts.ctypes.data, len(pts)), "fusion_get_points")
        return pts, float(ms.value)

    def destroy(self):
        if self.h:
            lib().mpmvs_fusion_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class ImageCache:
    """All images resident on one GPU in one layered texture (mpmvs_image_cache)."""

    def __init__(self, device: int, width: int, height: int, capacity: int, tex_format: int = 0):
        self.h = C.c_void_p()
        _ck(lib().mpmvs_cache_create_fmt(device, width, height, capacity, tex_format, C.byref(self.h)), "cache_create")

    def put(self, image_id: int, img: np.ndarray):
        if img.dtype == np.uint8:
            a = np.ascontiguousarray(img)
            _ck(lib().mpmvs_cache_put_u8(self.h, image_id, a.ctypes.data, a.shape[1], a.shape[0]), "cache_put_u8")
        else:
            a = np.ascontiguousarray(img, dtype=np.float32)
            _ck(lib().mpmvs_cache_put(self.h, image_id, a.ctypes.data, a.shape[1], a.shape[0]), "cache_put")

    def has(self, image_id: int) -> bool:
        return bool(lib().mpmvs_cache_has(self.h, image_id))

    def destroy(self):
        if self.h:
            lib().mpmvs_cache_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class PatchMatch:
    """One reference-image problem on one GPU (wraps mpmvs_problem; replaces a PatchMatchCUDA object)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.h = C.c_void_p()
        _ck(lib().mpmvs_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h)), "create")
        self.n = self.w = self.hgt = 0
        self._keep = []

    # ------------------------------------------------------------------ arithmetic
    def set_arithmetic(self, arithmetic):
        """ARITH_EXACT / "exact" (default: the reference's operations bit for bit) or ARITH_FAST / "fast"."""
        a = {"exact": ARITH_EXACT, "fast": ARITH_FAST}.get(arithmetic, arithmetic)
        _ck(lib().mpmvs_set_arithmetic(self.h, int(a)), "set_arithmetic")
        return self

    @property
    def arithmetic(self) -> str:
        a = C.c_int()
        _ck(lib().mpmvs_get_arithmetic(self.h, C.byref(a)), "get_arithmetic")
        return lib().mpmvs_arithmetic_name(a.value).decode()

    # ------------------------------------------------------------------ inputs
    def set_tex_format(self, fmt: int):
        _ck(lib().mpmvs_set_tex_format(self.h, int(fmt)), "set_tex_format")
        return self

    def set_problem(self, images, cams_packed: np.ndarray):
        """images: float32 arrays (mpmvs_set_views) or uint8 arrays (mpmvs_set_views_u8), [0] = reference."""
        u8 = all(i.dtype == np.uint8 for i in images)
        imgs = [np.ascontiguousarray(i, dtype=np.uint8 if u8 else np.float32) for i in images]
        cams = np.ascontiguousarray(cams_packed)
        assert cams.itemsize == 112 and len(cams) == len(imgs)
        ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
        if u8:
            _ck(lib().mpmvs_set_views_u8(self.h, len(imgs), ptrs, cams.ctypes.data), "set_views_u8")
        else:
            _ck(lib().mpmvs_set_views(self.h, len(imgs), ptrs, cams.ctypes.data), "set_views")
        self.n = len(imgs)
        self.hgt, self.w = imgs[0].shape
        self._keep = [imgs, cams]
        return self

    def set_problem_device(self, dev_ptrs, pitches, cams_packed: np.ndarray, shape):
        cams = np.ascontiguousarray(cams_packed)
        n = len(dev_ptrs)
        ptrs = (C.c_void_p * n)(*dev_ptrs)
        pit = np.asarray(pitches, dtype=np.uint64) if pitches is not None else None
        _ck(lib().mpmvs_set_views_device(self.h, n, ptrs, _ptr(pit), cams.ctypes.data), "set_views_device")
        self.n = n
        self.hgt, self.w = shape
        self._keep = [cams]
        return self

    def set_problem_cached(self, cache: ImageCache, ids, cams_packed: np.ndarray):
        cams = np.ascontiguousarray(cams_packed)
        ida = np.asarray(ids, dtype=np.int32)
        _ck(lib().mpmvs_set_views_cached(self.h, cache.h, len(ida), ida.ctypes.data, cams.ctypes.data), "set_views_cached")
        self.n = len(ida)
        self.w, self.hgt = int(cams["width"][0]), int(cams["height"][0])
        self._keep = [cams, cache]
        return self

    def set_geom_consistency_params(self, geom: bool, planar: bool):
        _ck(lib().mpmvs_set_geom_consistency_params(self.h, int(geom), int(planar)), "set_geom_consistency_params")

    def set_planar_prior_params(self):
        _ck(lib().mpmvs_set_planar_prior_params(self.h), "set_planar_prior_params")

    def reset_params(self):
        _ck(lib().mpmvs_reset_params(self.h), "reset_params")

    def set_src_depths(self, depths):
        d = [np.ascontiguousarray(x, dtype=np.float32) for x in depths]
        assert len(d) == self.n - 1
        ptrs = (C.c_void_p * len(d))(*[x.ctypes.data for x in d])
        _ck(lib().mpmvs_set_src_depths(self.h, ptrs), "set_src_depths")
        lib().mpmvs_synchronize(self.h)

    def set_src_depths_device(self, dev_ptrs, pitches=None):
        ptrs = (C.c_void_p * len(dev_ptrs))(*dev_ptrs)
        pit = np.asarray(pitches, dtype=np.uint64) if pitches is not None else None
        _ck(lib().mpmvs_set_src_depths_device(self.h, ptrs, _ptr(pit)), "set_src_depths_device")

    def set_state(self, planes4, costs):
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        c = np.ascontiguousarray(costs, dtype=np.float32)
        assert p.shape == (self.hgt, self.w, 4) and c.shape == (self.hgt, self.w)
        _ck(lib().mpmvs_set_state(self.h, p.ctypes.data, c.ctypes.data), "set_state")

    def set_prior(self, prior4, mask):
        p = np.ascontiguousarray(prior4, dtype=np.float32)
        m = np.ascontiguousarray(mask, dtype=np.uint32)
        _ck(lib().mpmvs_set_prior(self.h, p.ctypes.data, m.ctypes.data), "set_prior")

    # ------------------------------------------------------------------ running
    def run(self, seed: int) -> float:
        _ck(lib().mpmvs_run(self.h, C.c_uint64(seed)), "run")
        return self.last_run_ms()

    def run_async(self, seed: int):
        _ck(lib().mpmvs_run_async(self.h, C.c_uint64(seed)), "run_async")

    def run_into(self, seed: int, planes=None, costs=None, geom=None):
        _ck(lib().mpmvs_run_into(self.h, C.c_uint64(seed), _ptr(planes), _ptr(costs), _ptr(geom)), "run_into")

    def synchronize(self):
        _ck(lib().mpmvs_synchronize(self.h), "synchronize")

    def get_results_async(self, planes=None, costs=None, geom=None):
        """Enqueue D2H copies into (pinned) numpy arrays; synchronize() before reading them."""
        _ck(lib().mpmvs_get_results_async(self.h, _ptr(planes), _ptr(costs), _ptr(geom)), "get_results_async")

    def last_run_ms(self) -> float:
        ms = C.c_float()
        _ck(lib().mpmvs_last_run_ms(self.h, C.byref(ms)), "last_run_ms")
        return float(ms.value)

    def last_run_launches(self) -> int:
        n = C.c_int()
        _ck(lib().mpmvs_last_run_launches(self.h, C.byref(n)), "last_run_launches")
        return int(n.value)

    # ------------------------------------------------------------------ planar-prior stage
    def build_prior(self) -> dict:
        """Vertex picking + Delaunay + rasterisation + plane fit on the state of the previous run (mpmvs_build_prior)."""
        st = PriorStats()
        _ck(lib().mpmvs_build_prior(self.h, C.byref(st)), "build_prior")
        return {k: getattr(st, k) for k, _ in PriorStats._fields_}

    def pick_vertices(self, geom_variant: bool = False) -> np.ndarray:
        cap = 3 * ((self.w + 4) // 5) * ((self.hgt + 4) // 5)
        out = np.empty((cap, 2), np.int32)
        n = C.c_int()
        _ck(lib().mpmvs_pick_vertices(self.h, int(geom_variant), out.ctypes.data, cap, C.byref(n)), "pick_vertices")
        return out[: n.value].copy()

    def prior_from_triangles(self, xy: np.ndarray, tris: np.ndarray) -> int:
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        n = C.c_int()
        _ck(lib().mpmvs_prior_from_triangles(self.h, xy.ctypes.data, len(xy), tris.ctypes.data, len(tris), C.byref(n)), "prior_from_triangles")
        return int(n.value)

    def prior_pixels(self) -> int:
        n = C.c_int()
        _ck(lib().mpmvs_get_prior_pixels(self.h, C.byref(n)), "get_prior_pixels")
        return int(n.value)

    def get_prior(self):
        prior = np.empty((self.hgt, self.w, 4), np.float32)
        mask = np.empty((self.hgt, self.w), np.uint32)
        _ck(lib().mpmvs_get_prior(self.h, prior.ctypes.data, mask.ctypes.data), "get_prior")
        return prior, mask

    def set_profiling(self, flags: int):
        _ck(lib().mpmvs_set_profiling(self.h, int(flags)), "set_profiling")

    def last_run_profile(self, timing=True, count=False) -> dict:
        a, b, d, n, e = C.c_float(), C.c_float(), C.c_float(), C.c_int(), C.c_uint64()
        _ck(lib().mpmvs_last_run_profile(self.h, C.byref(a) if timing else None, C.byref(b) if timing else None,
                                         C.byref(n) if timing else None, C.byref(d) if timing else None,
                                         C.byref(e) if count else None), "last_run_profile")
        out = {}
        if timing:
            out.update(init_ms=a.value, sweep_ms=b.value, n_sweeps=n.value, finalize_ms=d.value)
        if count:
            out["ncc_evaluations"] = int(e.value)
        return out

    def init_only(self, seed: int):
        _ck(lib().mpmvs_init_only(self.h, C.c_uint64(seed)), "init_only")

    def half_sweep(self, red: int, it: int, scale: int):
        _ck(lib().mpmvs_half_sweep(self.h, int(red), int(it), int(scale)), "half_sweep")

    def finalize(self):
        _ck(lib().mpmvs_finalize(self.h), "finalize")

    # ------------------------------------------------------------------ results
    def result(self, geom=False):
        planes = np.empty((self.hgt, self.w, 4), np.float32)
        costs = np.empty((self.hgt, self.w), np.float32)
        _ck(lib().mpmvs_get_planes(self.h, planes.ctypes.data), "get_planes")
        _ck(lib().mpmvs_get_costs(self.h, costs.ctypes.data), "get_costs")
        if geom:
            g = np.empty((self.hgt, self.w), np.float32)
            _ck(lib().mpmvs_get_geom_costs(self.h, g.ctypes.data), "get_geom_costs")
            return planes, costs, g
        return planes, costs

    @property
    def depth_range(self):
        a, b = C.c_float(), C.c_float()
        _ck(lib().mpmvs_get_depth_range(self.h, C.byref(a), C.byref(b)), "get_depth_range")
        return float(a.value), float(b.value)

    def device_planes_ptr(self) -> int:
        p = C.c_void_p()
        _ck(lib().mpmvs_device_planes(self.h, C.byref(p)), "device_planes")
        return int(p.value)

    def export_depth_device(self, dev_ptr: int, pitch_bytes: int = 0):
        _ck(lib().mpmvs_export_depth_device(self.h, C.c_void_p(dev_ptr), pitch_bytes), "export_depth_device")

    def get_state(self):
        s = dict(
            planes=np.empty((self.hgt, self.w, 4), np.float32),
            costs=np.empty((self.hgt, self.w), np.float32),
            views=np.empty((self.hgt, self.w), np.uint32),
            rng=np.empty((self.hgt, self.w, 6), np.uint32),
            geom=np.empty((self.hgt, self.w), np.float32),
        )
        _ck(lib().mpmvs_get_device_state(self.h, *[s[k].ctypes.data for k in ("planes", "costs", "views", "rng", "geom")]),
            "get_device_state")
        return s

    def set_dev_state(self, s):
        arrs = []
        for k, dt in (("planes", np.float32), ("costs", np.float32), ("views", np.uint32), ("rng", np.uint32), ("geom", np.float32)):
            a = s.get(k)
            arrs.append(None if a is None else np.ascontiguousarray(a, dtype=dt))
        _ck(lib().mpmvs_set_device_state(self.h, *[_ptr(a) for a in arrs]), "set_device_state")

    def ncc_map(self, planes4, scale: int):
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        out = np.empty((self.n - 1, self.hgt, self.w), np.float32)
        _ck(lib().mpmvs_ncc_map(self.h, p.ctypes.data, int(scale), out.ctypes.data), "ncc_map")
        return out

    def ncc_bench(self, planes4, scale: int, taps: int, n_views: int, reps: int):
        """(ms of one pass, executed NCC evaluations) of the NCC microbenchmark kernel."""
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        ms, cnt = C.c_float(), C.c_uint64()
        _ck(lib().mpmvs_ncc_bench(self.h, p.ctypes.data, scale, taps, n_views, reps, C.byref(ms), C.byref(cnt)), "ncc_bench")
        return float(ms.value), int(cnt.value)

    def geom_map(self, planes4):
        p = np.ascontiguousarray(planes4, dtype=np.float32)
        out = np.empty((self.n - 1, self.hgt, self.w), np.float32)
        _ck(lib().mpmvs_geom_map(self.h, p.ctypes.data, out.ctypes.data), "geom_map")
        return out

    @staticmethod
    def uniform_stream(seed: int, x: int, y: int, n: int):
        out = np.empty(n, np.float32)
        _ck(lib().mpmvs_uniform_stream(C.c_uint64(seed), x, y, n, out.ctypes.data), "uniform_stream")
        return out

    def destroy(self):
        if self.h:
            lib().mpmvs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
