"""Debug-image dumps of the reference: the JPEG previews written next to the .dmb files.

    save_cost      SaveCost,     /root/reference/src/utility.cpp:465-477 (and costs.jpg of ProcessProblem, PatchMatch.cpp:624-629)
    save_normal    SaveNormal,   utility.cpp:310-320
    save_depth     SaveDmb,      utility.cpp:376-463 (JET colour map, optional 3 % histogram clipping, invalid pixels black)
    save_dmb_as_jpg  saveDmbAsJpg, utility.cpp:479-520, driven by the four `Save ...` YAML keys (main.cpp:51-53)

Host-side and off the hot path; OpenCV does the encoding exactly as in the reference (cv2 is the same library).
"""
from __future__ import annotations

import os

import numpy as np

from . import io_formats


def _u8(a: np.ndarray, scale: float) -> np.ndarray:
    """cv::Mat::convertTo(CV_8U, scale): round half to even, saturate."""
    return np.clip(np.rint(a.astype(np.float64) * scale), 0, 255).astype(np.uint8)


def save_cost(cost: np.ndarray, path: str) -> bool:
    import cv2

    if cost is None or cost.size == 0:
        print("Can not read this cost image !")
        return False
    return bool(cv2.imwrite(path, _u8(cost, 255.0 / 2.0)))


def save_normal(normal: np.ndarray, path: str, k: float = 255.0) -> bool:
    import cv2

    n = normal.astype(np.float32) * np.float32(k)
    n[..., 1] = -n[..., 1]
    return bool(cv2.imwrite(path, n))          # imwrite saturates the float image to 8 bits, channels in storage order


def _clip_levels(u8: np.ndarray):
    """getMax10 / getMin10, utility.cpp:349-374: grey levels that cut ~3 % of the pixels at either end (bin 0 is skipped)."""
    hist = np.bincount(u8.ravel(), minlength=256)
    total = u8.size
    top, down = 1, 1
    acc = 0
    for i in range(1, 256):
        acc += int(hist[i])
        if np.float32(acc) / np.float32(total) > 0.03:
            top = i - 1
            break
    acc = 0
    for i in range(255, 1, -1):
        acc += int(hist[i])
        if np.float32(acc) / np.float32(total) > 0.03:
            down = i + 1
            break
    return top, down


def depth_preview(depth: np.ndarray, hist_enhance: bool = True) -> np.ndarray:
    """The BGR image SaveDmb writes."""
    import cv2

    d = depth.astype(np.float32).copy()
    valid = d > 0
    d[~valid] = 0
    mn, mx = float(d.min()), float(d.max())
    norm = (d - np.float32(mn)) * np.float32(1 / (mx - mn + 1e-8))
    u8 = _u8(norm, 255.0)
    if hist_enhance:
        top, down = _clip_levels(u8)
        new_min = mn + (mx - mn) * (np.float32(top) / 256.0)
        new_max = mn + (mx - mn) * (np.float32(down) / 256.0)
        d = np.where(d < new_min, np.float32(new_min), np.where(d > new_max, np.float32(new_max), d)).astype(np.float32)
        norm = (d - np.float32(new_min) + np.float32(1e-8)) * np.float32(1 / (new_max - new_min + 1e-8))
        u8 = _u8(norm, 255.0)
    bgr = cv2.applyColorMap(u8, cv2.COLORMAP_JET)
    bgr[~valid] = 0
    return bgr


def save_depth(depth: np.ndarray, path: str, hist_enhance: bool = True) -> bool:
    import cv2

    if depth is None or depth.size == 0:
        print("Can not read this depth image !")
        return False
    return bool(cv2.imwrite(path, depth_preview(depth, hist_enhance)))


def save_dmb_as_jpg(cfg: dict, image_ids, hist_enhance: bool = True) -> int:
    """saveDmbAsJpg over the result folders of `image_ids`. Returns the number of previews written.
    (The reference also copies every depth map to a hard-coded /home/xuan/... folder, utility.cpp:484,493 -- not kept.)"""
    out = cfg["Output-folder"].rstrip("/")
    n = 0
    for i in image_ids:
        folder = io_formats.result_dir(out, i)
        if int(cfg["Save Dmb as JPG"]):
            n += save_depth(io_formats.read_dmb(os.path.join(folder, "depths.dmb")), os.path.join(folder, "depths.jpg"), hist_enhance)
        prior = os.path.join(folder, "depths_prior.dmb")
        if int(cfg["Save Prior Dmb as JPG"]) and (int(cfg["Planer prior"]) or int(cfg["Geometric consistency planer prior"])) and os.path.exists(prior):
            n += save_depth(io_formats.read_dmb(prior), os.path.join(folder, "depths_prior.jpg"), hist_enhance)
        if int(cfg["Save Cost Map"]):
            n += save_cost(io_formats.read_dmb(os.path.join(folder, "costs.dmb")), os.path.join(folder, "costs.jpg"))
        if int(cfg["Save Normal Map"]):
            n += save_normal(io_formats.read_dmb(os.path.join(folder, "normals.dmb")), os.path.join(folder, "normals.jpg"))
    return n
