"""File formats of the MP-MVS dense folder (the drop-in surface).

Readers/writers restate the reference's on-disk layouts so the synthetic
generator, the tests and the C++ host all speak the same bytes:

* ``.dmb``            -- /root/reference/src/utility.cpp:193-308
                         int32 type=1, h, w, nb, then h*w*nb float32 row-major.
* ``cams/%08d_cam.txt`` -- /root/reference/src/PatchMatch.cpp:111-143 (ReadCamera),
                         writer colmap2mvsnet_acm.py:423-441.
* ``pair.txt``        -- /root/reference/src/PatchMatch.cpp:67-109 (GenerateSampleList).
* ``config.yaml``     -- /root/reference/src/utility.cpp:8-35 (13 keys, "Planer" spelling normative).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

CONFIG_KEYS = (
    "Input-folder",
    "Output-folder",
    "Geometric consistency iterations",
    "Planer prior",
    "Geometric consistency planer prior",
    "Sky segment",
    "Use dynamic_consistency to fuse",
    "Save Dmb as JPG",
    "Save Prior Dmb as JPG",
    "Save Cost Map",
    "Save Normal Map",
    "Max source images num",
    "Max image size",
)


# ----------------------------------------------------------------------------- .dmb
def write_dmb(path: str, arr: np.ndarray) -> None:
    """Write a depth/cost (h,w) or normal (h,w,3) map. utility.cpp:225-247, 283-308."""
    a = np.ascontiguousarray(arr, dtype=np.float32)
    if a.ndim == 2:
        h, w = a.shape
        nb = 1
    elif a.ndim == 3:
        h, w, nb = a.shape
    else:
        raise ValueError("dmb arrays are (h,w) or (h,w,nb)")
    with open(path, "wb") as f:
        np.array([1, h, w, nb], dtype=np.int32).tofile(f)
        a.tofile(f)


def read_dmb(path: str) -> np.ndarray:
    """Read a .dmb; returns (h,w) for nb==1 else (h,w,nb). utility.cpp:193-223, 250-281."""
    with open(path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.int32, count=4)
        if hdr.size != 4 or hdr[0] != 1:
            raise ValueError(f"{path}: not a type-1 dmb")
        _, h, w, nb = (int(v) for v in hdr)
        data = np.fromfile(f, dtype=np.float32, count=h * w * nb)
    if data.size != h * w * nb:
        raise ValueError(f"{path}: truncated dmb")
    return data.reshape(h, w) if nb == 1 else data.reshape(h, w, nb)


# ----------------------------------------------------------------------------- cameras
@dataclass
class Camera:
    """Mirror of struct Camera, /root/reference/include/PatchMatch.h:35-46 (112 bytes)."""

    K: np.ndarray  # (3,3) float32
    R: np.ndarray  # (3,3) float32 world->cam
    t: np.ndarray  # (3,)  float32
    height: int = 0
    width: int = 0
    depth_min: float = 0.0
    depth_max: float = 0.0

    @property
    def C(self) -> np.ndarray:
        # PatchMatch.cpp:134-136: C[k] = -(R[k] t[0] + R[3+k] t[1] + R[6+k] t[2]) in float32, every product and sum rounded, left to
        # right (a BLAS matrix product may fuse or reorder them: the last bit of C reaches every homography).
        R = self.R.astype(np.float32)
        t = self.t.astype(np.float32)
        f = np.float32
        return np.array([-f(f(f(R[0, k] * t[0]) + f(R[1, k] * t[1])) + f(R[2, k] * t[2])) for k in range(3)], dtype=np.float32)

    def pack(self) -> np.ndarray:
        """112-byte record laid out exactly like the C struct."""
        rec = np.zeros(1, dtype=CAMERA_DTYPE)
        rec["K"] = self.K.astype(np.float32).reshape(9)
        rec["R"] = self.R.astype(np.float32).reshape(9)
        rec["t"] = self.t.astype(np.float32)
        rec["C"] = self.C
        rec["height"] = self.height
        rec["width"] = self.width
        rec["depth_min"] = self.depth_min
        rec["depth_max"] = self.depth_max
        return rec


CAMERA_DTYPE = np.dtype(
    [
        ("K", np.float32, 9),
        ("R", np.float32, 9),
        ("t", np.float32, 3),
        ("C", np.float32, 3),
        ("height", np.int32),
        ("width", np.int32),
        ("depth_min", np.float32),
        ("depth_max", np.float32),
    ]
)
assert CAMERA_DTYPE.itemsize == 112


def pack_cameras(cams: Sequence[Camera]) -> np.ndarray:
    return np.concatenate([c.pack() for c in cams])


def write_cam(path: str, cam: Camera, interval: float = 0.0, depth_num: int = 192) -> None:
    E = np.eye(4, dtype=np.float64)
    E[:3, :3] = cam.R
    E[:3, 3] = cam.t
    with open(path, "w") as f:
        f.write("extrinsic\n")
        for r in range(4):
            f.write(" ".join(repr(float(v)) for v in E[r]) + "\n")
        f.write("\nintrinsic\n")
        for r in range(3):
            f.write(" ".join(repr(float(v)) for v in cam.K[r]) + "\n")
        f.write(f"\n{cam.depth_min!r} {interval!r} {depth_num} {cam.depth_max!r}\n")


def read_cam(path: str) -> Camera:
    """ReadCamera, PatchMatch.cpp:111-143: token stream, values parsed as float32."""
    toks = open(path).read().split()
    if not toks or toks[0] != "extrinsic":
        raise ValueError(f"{path}: missing 'extrinsic'")
    v = toks[1:17]
    E = np.array([np.float32(x) for x in v], dtype=np.float32).reshape(4, 4)
    if toks[17] != "intrinsic":
        raise ValueError(f"{path}: missing 'intrinsic'")
    K = np.array([np.float32(x) for x in toks[18:27]], dtype=np.float32).reshape(3, 3)
    dmin, _interval, _num, dmax = (np.float32(x) for x in toks[27:31])
    return Camera(K=K, R=E[:3, :3].copy(), t=E[:3, 3].copy(), depth_min=float(dmin), depth_max=float(dmax))


# ----------------------------------------------------------------------------- pair.txt
@dataclass
class SceneEntry:
    """Mirror of struct Scene (utility.h:17-26) minus the image payloads."""

    ref_id: int = -1
    src_ids: List[int] = field(default_factory=list)  # src_ids[0] == ref_id
    estimate: bool = False


def write_pairs(path: str, pairs: Dict[int, Sequence[Tuple[int, float]]]) -> None:
    with open(path, "w") as f:
        f.write(f"{len(pairs)}\n")
        for ref in sorted(pairs):
            f.write(f"{ref}\n")
            row = pairs[ref]
            f.write(f"{len(row)} " + " ".join(f"{i} {s:.4f}" for i, s in row) + "\n")


def read_pairs(path: str, max_src: int = 20) -> List[SceneEntry]:
    """GenerateSampleList, PatchMatch.cpp:67-109, including its quirks:
    gaps in ref ids are padded with estimate=false scenes, score<=0 entries are
    dropped, only the first ``max_src`` *listed positions* may be kept (j < max)."""
    toks = open(path).read().split()
    it = iter(toks)
    try:
        return _read_pairs(it, max_src)
    except (StopIteration, ValueError) as e:       # same verdict as the C++ host (mpmvs_host.cpp, GenerateSampleList)
        raise ValueError(f"malformed pair.txt: {path} ({e or 'truncated'})") from None


def _read_pairs(it, max_src: int) -> List[SceneEntry]:
    n = int(next(it))
    scenes: List[SceneEntry] = []
    for _ in range(n):
        ref = int(next(it))
        if ref < len(scenes):                      # scenes are indexed by image id: ids must ascend
            raise ValueError(f"reference id {ref} out of order")
        if ref > 1 << 24:                          # the gap up to `ref` is padded with empty scenes (cpp:87-91)
            raise ValueError(f"reference id {ref} is not an image id")
        sc = SceneEntry(ref_id=ref, src_ids=[ref])
        while ref > len(scenes):
            scenes.append(SceneEntry())
        ns = int(next(it))
        for j in range(ns):
            sid = int(next(it))
            score = float(next(it))
            if score <= 0.0:
                continue
            if j < max_src:
                sc.src_ids.append(sid)
        sc.estimate = ns != 0
        scenes.append(sc)
    for sc in scenes:
        for sid in sc.src_ids:
            if not 0 <= sid < len(scenes):
                raise ValueError(f"image {sid} has no entry of its own")
    return scenes


# ----------------------------------------------------------------------------- config.yaml
DEFAULT_CONFIG = {
    "Input-folder": "",
    "Output-folder": "",
    "Geometric consistency iterations": 2,
    "Planer prior": 1,
    "Geometric consistency planer prior": 1,
    "Sky segment": 0,
    "Use dynamic_consistency to fuse": 1,
    "Save Dmb as JPG": 0,
    "Save Prior Dmb as JPG": 0,
    "Save Cost Map": 0,
    "Save Normal Map": 0,
    "Max source images num": 20,
    "Max image size": 3200,
}


def write_config(path: str, **overrides) -> dict:
    cfg = dict(DEFAULT_CONFIG)
    for k, v in overrides.items():
        if k not in cfg:
            raise KeyError(k)
        cfg[k] = v
    with open(path, "w") as f:
        f.write("%YAML:1.0\n---\n")
        for k in CONFIG_KEYS:
            v = cfg[k]
            f.write(f'{k}: "{v}"\n' if isinstance(v, str) else f"{k}: {int(v)}\n")
    return cfg


def read_config(path: str) -> dict:
    """Subset of cv::FileStorage YAML that config/config.yaml uses: 'key: value' lines."""
    cfg = dict(DEFAULT_CONFIG)
    for line in open(path):
        s = line.strip()
        if not s or s.startswith(("#", "%", "---")):
            continue
        k, _, v = s.partition(":")
        k = k.strip()
        v = v.strip()
        if k not in cfg:
            continue
        if v.startswith('"'):
            cfg[k] = v.strip('"')
        else:
            cfg[k] = int(float(v))
    for k in ("Input-folder", "Output-folder"):
        cfg[k] = cfg[k].rstrip("/")  # checkpath(), utility.cpp:3-6
    return cfg


def result_dir(out_folder: str, ref_id: int) -> str:
    """<Output-folder>/MPMVS/2333_%08d (PatchMatch.cpp:510-513, utility.cpp:30)."""
    return os.path.join(out_folder, "MPMVS", f"2333_{ref_id:08d}")


def write_ply(path: str, points9: np.ndarray) -> None:
    """StoreColorPlyFileBinaryPointCloud, PatchMatch.cpp:145-198: binary little-endian x y z nx ny nz + uchar r g b.
    points9: (n, 9) float32 = coord, normal, colour (b, g, r as in PointList::color)."""
    p = np.asarray(points9, dtype=np.float32).reshape(-1, 9)
    rec = np.zeros(len(p), dtype=[("xyz", "<f4", 3), ("n", "<f4", 3), ("rgb", "u1", 3)])
    xyz = p[:, :3].copy()
    xyz[~np.isfinite(xyz).all(1)] = 0.0                      # :175-179
    rec["xyz"], rec["n"] = xyz, p[:, 3:6]
    rec["rgb"] = np.clip(p[:, [8, 7, 6]], 0, 255).astype(np.int32).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
                 "end_header\n" % len(p)).encode())
        rec.tofile(f)
