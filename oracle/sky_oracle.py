"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's sky-mask joint-bilateral upsampling.

Follows /root/reference/SkySegment/src/SkyRegionDetect.cu:3-35 (Pixel_bilateral_filter) and :37-42 (cv::resize of the
network's mask to image size), called from GenerateSkyRegionMask, /root/reference/src/PatchMatch.cpp:4-57.
Pinned by tests/golden/sky.npz (outputs of the reference kernel itself, tests/golden/make_golden_sky.py) and, for the
resize, by cv2.resize in tests/test_sky.py. Only tests/, __graft_entry__.smoke() and bench.py's baseline leg may import it.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HALF_WINDOW = 18                                   # SkyRegionDetect.cu:13
SIGMA_SPATIAL = np.float32(2.0 * 6.0 * 6.0)        # :9
SIGMA_COLOR = np.float32(2.0 * 2.0 * 2.0)          # :10


def resize_linear(src: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv::resize(src, dst, Size(width, height), 0, 0, INTER_LINEAR) for one float32 channel (SkyRegionDetect.cu:41-42)."""
    src = np.ascontiguousarray(src, np.float32)
    sh, sw = src.shape
    if (sw, sh) == (width, height):
        return src.copy()

    def taps(n_dst, n_src):
        scale = n_src / n_dst
        f = ((np.arange(n_dst) + 0.5) * scale - 0.5).astype(np.float32)
        i = np.floor(f).astype(np.int64)
        f = (f - i.astype(np.float32)).astype(np.float32)
        lo = i < 0
        i[lo], f[lo] = 0, 0.0
        hi = i >= n_src - 1
        i[hi], f[hi] = n_src - 1, 0.0
        return i, np.minimum(i + 1, n_src - 1), f

    x0, x1, fx = taps(width, sw)
    y0, y1, fy = taps(height, sh)
    one = np.float32(1.0)
    rows = src[:, x0] * (one - fx)[None, :] + src[:, x1] * fx[None, :]          # horizontal pass
    return (rows[y0] * (one - fy)[:, None] + rows[y1] * fy[:, None]).astype(np.float32)


def sky_mask_refine(bgr: np.ndarray, mask: np.ndarray):
    """Returns (result float32 of 0 / 255, prob float32). bgr [h][w][3] uint8, mask float32 of any size."""
    img = np.ascontiguousarray(bgr, np.uint8).astype(np.float32)
    h, w = img.shape[:2]
    m = resize_linear(mask, w, h)
    R = HALF_WINDOW
    pad_img = np.zeros((h + 2 * R, w + 2 * R, 3), np.float32)
    pad_img[R:R + h, R:R + w] = img
    pad_m = np.zeros((h + 2 * R, w + 2 * R), np.float32)
    pad_m[R:R + h, R:R + w] = m
    valid = np.zeros((h + 2 * R, w + 2 * R), bool)
    valid[R:R + h, R:R + w] = True
    wsum = np.zeros((h, w), np.float32)
    prob = np.zeros((h, w), np.float32)
    for i in range(-R, R + 1):              # x offset, outer loop (:16)
        for j in range(-R, R + 1):          # y offset (:17)
            sl = (slice(R + j, R + j + h), slice(R + i, R + i + w))
            d = pad_img[sl] - img
            dis_color = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2], dtype=np.float32)
            distance = np.sqrt(np.float32(i * i + j * j), dtype=np.float32)
            wgt = np.exp(-distance / SIGMA_SPATIAL - dis_color / SIGMA_COLOR, dtype=np.float32)
            wgt = np.where(valid[sl], wgt, np.float32(0))      # out-of-image taps are skipped (:20-21)
            wsum += wgt
            prob += wgt * pad_m[sl]
    prob = prob / wsum
    return np.where(prob.astype(np.float64) > 0.6, np.float32(255), np.float32(0)), prob   # :34


# ------------------------------------------------------------------ the reference kernel itself (oracle/_ref)
REF_LIB = os.path.join(HERE, "_ref", "libmpmvs_ref_sky.so")


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_sky_filter(bgr: np.ndarray, mask_full: np.ndarray, reps: int = 0):
    """Pixel_bilateral_filter of the reference, launched with its own grid; mask already at image size.
    Returns result (and the mean kernel ms when reps > 0). Needs a GPU."""
    L = C.CDLL(REF_LIB)
    img = np.ascontiguousarray(bgr, np.uint8)
    m = np.ascontiguousarray(mask_full, np.float32)
    h, w = m.shape
    assert img.shape == (h, w, 3)
    out = np.empty((h, w), np.float32)
    if reps > 0:
        L.ref_sky_filter_timed.restype = C.c_float
        ms = L.ref_sky_filter_timed(img.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), h, w, reps)
        return out, float(ms)
    rc = L.ref_sky_filter(img.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), h, w)
    if rc:
        raise RuntimeError(f"reference sky kernel failed: CUDA error {rc}")
    return out


def make_sky_case(width=160, height=120, mask_div=4, seed=7):
    """Synthetic outdoor frame: smooth sky with soft clouds above a ragged, textured skyline, and the low-resolution
    probability map a segmentation network would give (blurred, noisy, 1/mask_div size). Returns (bgr, mask_lo, truth)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    skyline = 0.45 * height + 0.12 * height * np.sin(xx[0] / width * 7.0) + rng.normal(0, 2.0, width).cumsum() * 0.3
    for k in range(6):                                    # a few buildings
        a = int(rng.integers(0, max(1, width - 12))); b = a + int(rng.integers(6, 24))
        skyline[a:b] -= rng.uniform(5, max(6.0, 0.25 * height))
    truth = yy < skyline[None, :]
    sky = np.stack([200 + 40 * (1 - yy / height), 150 + 50 * (1 - yy / height), 90 + 60 * (yy / height)], -1)   # b g r
    cloud = 30 * np.clip(np.sin(xx / 17.0 + 1.3) * np.cos(yy / 11.0 + 0.4), 0, 1)
    sky += cloud[..., None]
    ground = rng.uniform(20, 140, (height, width, 3)).astype(np.float32)
    ground = 0.5 * ground + 0.5 * np.roll(ground, 1, 1)
    img = np.where(truth[..., None], sky, ground)
    bgr = np.clip(img + rng.normal(0, 2.0, img.shape), 0, 255).astype(np.uint8)
    mh, mw = max(2, height // mask_div), max(2, width // mask_div)
    lo = truth[:mh * mask_div, :mw * mask_div].reshape(mh, mask_div, mw, mask_div).mean((1, 3)).astype(np.float32)
    lo = np.clip(lo + rng.normal(0, 0.08, lo.shape), 0, 1).astype(np.float32)
    return bgr, lo, truth
