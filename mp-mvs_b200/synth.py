"""Synthetic rendered scenes in the MP-MVS dense-folder layout.

There is no network and no dataset in this environment, so every input of the
tests and of bench.py is rendered here: textured planar faces (boxes, a room,
a ground plane) seen by pinhole cameras, with a procedural texture anchored in
world coordinates so that all views are photo-consistent, and exact ground
truth depth (camera z) and world normals per view.

Layout written by :func:`write_dense_folder` (what colmap2mvsnet_acm.py emits,
/root/reference/colmap2mvsnet_acm.py:423-451, and what PatchMatchInit reads,
/root/reference/src/PatchMatch.cpp:871-890):

    <root>/images/%08d.jpg      8-bit gray JPEG (q=97)
    <root>/images/%08d.pgm      the *decoded* gray image, byte-identical to what
                                cv2.imread(jpg, IMREAD_GRAYSCALE) returns -- the
                                OpenCV-free C++ hosts (ours and the oracle harness)
                                read this so both consume identical pixels
    <root>/cams/%08d_cam.txt
    <root>/pair.txt
    <root>/gt/%08d_depth.dmb, %08d_normal.dmb   ground truth (not part of the reference layout)

Pixel convention is the reference's: integer pixel (x, y) is the ray
K^-1 (x, y, 1) (GetPointI2C, /root/reference/src/PatchMatch.cu:163-168).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .io_formats import Camera, write_cam, write_dmb, write_pairs


# ----------------------------------------------------------------------------- texture
def _hash2(ix: np.ndarray, iy: np.ndarray, seed: int) -> np.ndarray:
    """Integer lattice hash -> float in [0,1). uint32 arithmetic, deterministic."""
    h = (ix.astype(np.uint32) * np.uint32(374761393)) ^ (iy.astype(np.uint32) * np.uint32(668265263))
    h = h + np.uint32((seed * 2246822519) & 0xFFFFFFFF)
    h = (h ^ (h >> np.uint32(13))) * np.uint32(1274126177)
    h = h ^ (h >> np.uint32(16))
    return (h & np.uint32(0xFFFFFF)).astype(np.float32) / np.float32(1 << 24)


def value_noise(u: np.ndarray, v: np.ndarray, seed: int, octaves: int = 4) -> np.ndarray:
    """Multi-octave value noise in [0,1], smoothstep-interpolated."""
    out = np.zeros(u.shape, dtype=np.float32)
    amp = 1.0
    tot = 0.0
    fu = u.astype(np.float32)
    fv = v.astype(np.float32)
    for o in range(octaves):
        iu = np.floor(fu)
        iv = np.floor(fv)
        a = fu - iu
        b = fv - iv
        a = a * a * (3.0 - 2.0 * a)
        b = b * b * (3.0 - 2.0 * b)
        iu = iu.astype(np.int64)
        iv = iv.astype(np.int64)
        s = seed * 31 + o
        n00 = _hash2(iu, iv, s)
        n10 = _hash2(iu + 1, iv, s)
        n01 = _hash2(iu, iv + 1, s)
        n11 = _hash2(iu + 1, iv + 1, s)
        out += amp * ((n00 * (1 - a) + n10 * a) * (1 - b) + (n01 * (1 - a) + n11 * a) * b)
        tot += amp
        amp *= 0.75
        fu = fu * 2.0 + 17.0
        fv = fv * 2.0 + 29.0
    out = out / np.float32(tot)
    return np.clip(0.5 + (out - 0.5) * 2.2, 0.0, 1.0)  # stretch: averaged octaves concentrate near 0.5


# ----------------------------------------------------------------------------- geometry
@dataclass
class Face:
    """A textured rectangle: origin o, orthogonal edges eu, ev; outward normal eu x ev."""

    o: np.ndarray
    eu: np.ndarray
    ev: np.ndarray
    base: float = 128.0  # mean intensity
    amp: float = 60.0  # +- amplitude of the texture
    freq: float = 8.0  # coarsest lattice cells per world unit
    seed: int = 0
    octaves: int = 4

    @property
    def normal(self) -> np.ndarray:
        n = np.cross(self.eu, self.ev)
        return n / np.linalg.norm(n)


def box_faces(lo, hi, inward=False, seed=0, **tex) -> List[Face]:
    """Six faces of an axis-aligned box; inward=True gives a room seen from inside."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    d = hi - lo
    ex, ey, ez = np.diag(d)
    faces = [
        (lo, ey, ex),  # z = lo (normal -z)
        (lo + ez, ex, ey),  # z = hi (normal +z)
        (lo, ex, ez),  # y = lo (normal -y)
        (lo + ey, ez, ex),  # y = hi (normal +y)
        (lo, ez, ey),  # x = lo (normal -x)
        (lo + ex, ey, ez),  # x = hi (normal +x)
    ]
    out = []
    for k, (o, eu, ev) in enumerate(faces):
        if inward:
            eu, ev = ev, eu
        out.append(Face(o=o.copy(), eu=eu.copy(), ev=ev.copy(), seed=seed * 8 + k, **tex))
    return out


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> Tuple[np.ndarray, np.ndarray]:
    """World->cam (R, t) with image x right, y down, z forward."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])
    t = -R @ eye
    return R, t


def render_view(faces: Sequence[Face], K: np.ndarray, R: np.ndarray, t: np.ndarray, width: int, height: int,
                rows: Optional[Tuple[int, int]] = None):
    """Ray-cast one view. Returns (gray float32 [0,255], depth z float32 (0 = no hit), normal world float32)."""
    r0, r1 = rows if rows is not None else (0, height)
    ys, xs = np.mgrid[r0:r1, 0:width].astype(np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    dc = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], axis=-1)  # cam-frame ray with z=1
    dw = dc @ R  # R^T applied to row vectors
    C = -R.T @ t
    best = np.full(xs.shape, np.inf)
    gray = np.zeros(xs.shape, dtype=np.float32)
    nrm = np.zeros(xs.shape + (3,), dtype=np.float32)
    for f in faces:
        n = f.normal
        denom = dw @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            z = ((f.o - C) @ n) / denom
        ok = (denom < -1e-9) & (z > 1e-6) & (z < best)
        if not ok.any():
            continue
        P = C + dw * z[..., None]
        rel = P - f.o
        lu = np.linalg.norm(f.eu)
        lv = np.linalg.norm(f.ev)
        u = rel @ (f.eu / lu)
        v = rel @ (f.ev / lv)
        ok &= (u >= 0) & (u <= lu) & (v >= 0) & (v <= lv)
        if not ok.any():
            continue
        uu = u[ok] * f.freq
        vv = v[ok] * f.freq
        tex = value_noise(uu, vv, f.seed, f.octaves)
        gray[ok] = f.base + f.amp * (2.0 * tex - 1.0)
        best[ok] = z[ok]
        nrm[ok] = n.astype(np.float32)
    depth = np.where(np.isfinite(best), best, 0.0).astype(np.float32)
    return np.clip(gray, 0, 255), depth, nrm


# ----------------------------------------------------------------------------- scenes
@dataclass
class SyntheticScene:
    name: str
    width: int
    height: int
    faces: List[Face]
    cams: List[Camera]
    pairs: Dict[int, List[Tuple[int, float]]]
    images: List[np.ndarray] = field(default_factory=list)  # uint8 gray, as decoded from the JPEG
    gt_depth: List[np.ndarray] = field(default_factory=list)
    gt_normal: List[np.ndarray] = field(default_factory=list)

    @property
    def num_views(self) -> int:
        return len(self.cams)

    def problem(self, ref: int, max_src: int = 20):
        """(ids, images float32 0..255, cameras) for one reference, ids[0] == ref."""
        ids = [ref] + [i for i, s in self.pairs[ref] if s > 0][:max_src]
        imgs = [self.images[i].astype(np.float32) for i in ids]
        cams = [self.cams[i] for i in ids]
        return ids, imgs, cams


def _intrinsics(f: float, width: int, height: int) -> np.ndarray:
    return np.array([[f, 0, (width - 1) / 2.0], [0, f, (height - 1) / 2.0], [0, 0, 1]], dtype=np.float64)


def _nearest_pairs(eyes: np.ndarray, target: np.ndarray, n_src: int) -> Dict[int, List[Tuple[int, float]]]:
    """Sources = views with the smallest viewing-angle difference (score = 1/(angle+eps))."""
    d = eyes - target
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cosang = np.clip(d @ d.T, -1, 1)
    ang = np.degrees(np.arccos(cosang))
    pairs = {}
    for i in range(len(eyes)):
        order = [j for j in np.argsort(ang[i], kind="stable") if j != i][:n_src]
        pairs[i] = [(int(j), float(100.0 / (1.0 + ang[i, j]))) for j in order]
    return pairs


def _jpeg_roundtrip(gray: np.ndarray, quality: int = 97) -> Tuple[np.ndarray, bytes]:
    import cv2

    img8 = np.clip(np.rint(gray), 0, 255).astype(np.uint8)
    ok, buf = cv2.imencode(".jpg", img8, [int(cv2.IMWRITE_JPEG_QUALITY), quality])
    assert ok
    dec = cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)  # the call PatchMatchInit makes (PatchMatch.cpp:877)
    return dec, buf.tobytes()


def _render_one(args):
    faces, K, R, t, width, height, jpeg = args
    gray, depth, nrm = render_view(faces, K, R, t, width, height)
    img = _jpeg_roundtrip(gray)[0] if jpeg else np.clip(np.rint(gray), 0, 255).astype(np.uint8)
    return img, depth, nrm


def _finish(name, width, height, faces, Ks, poses, pairs, jpeg=True, views: Optional[Sequence[int]] = None,
            workers: int = 0) -> SyntheticScene:
    """Render the requested views (all by default); unrendered views keep empty placeholders.

    workers > 1 renders views in a fork()ed process pool (full-size scenes take ~15 s per view on one core)."""
    cams, images, gtd, gtn = [], [], [], []
    want = sorted(set(range(len(poses))) if views is None else set(views))
    jobs = [(list(faces), Ks[i], poses[i][0], poses[i][1], width, height, jpeg) for i in want]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            rendered = dict(zip(want, pool.map(_render_one, jobs)))
    else:
        rendered = {i: _render_one(j) for i, j in zip(want, jobs)}
    for i, (K, (R, t)) in enumerate(zip(Ks, poses)):
        if i in rendered:
            img, depth, nrm = rendered[i]
            valid = depth > 0
            dmin = float(depth[valid].min()) if valid.any() else 1.0
            dmax = float(depth[valid].max()) if valid.any() else 2.0
        else:
            img = np.zeros((0, 0), np.uint8)
            depth = np.zeros((0, 0), np.float32)
            nrm = np.zeros((0, 0, 3), np.float32)
            dmin, dmax = 1.0, 2.0
        cams.append(Camera(K=K.astype(np.float32), R=R.astype(np.float32), t=t.astype(np.float32),
                           height=height, width=width, depth_min=dmin, depth_max=dmax))
        images.append(img)
        gtd.append(depth)
        gtn.append(nrm)
    return SyntheticScene(name, width, height, list(faces), cams, pairs, images, gtd, gtn)


def make_plane_scene(width=640, height=480, n_views=3, seed=1, jpeg=True, n_src=2) -> SyntheticScene:
    """Config 1: textured, mildly tilted plane at depth ~5, cameras on a short arc."""
    f = 560.0 * width / 640.0
    K = _intrinsics(f, width, height)
    # plane through the origin, tilted about x and y; cameras sit on the -y side looking along +y
    tilt = np.radians([12.0, -8.0])
    eu = np.array([np.cos(tilt[0]), np.sin(tilt[0]), 0.0]) * 14.0
    ev = np.array([0.0, np.sin(tilt[1]), np.cos(tilt[1])]) * 12.0
    ev -= (ev @ eu) / (eu @ eu) * eu
    o = -0.5 * (eu + ev)
    n = np.cross(eu, ev)
    if n[1] > 0:  # must face the cameras (-y side)
        eu, ev = ev, eu
        o = -0.5 * (eu + ev)
    faces = [Face(o=o, eu=eu, ev=ev, base=128, amp=70, freq=3.0, seed=seed, octaves=5)]
    target = np.zeros(3)
    angs = np.radians(np.linspace(-6.0, 6.0, n_views) if n_views > 1 else [0.0])
    eyes = np.stack([[5.0 * np.sin(a), -5.0 * np.cos(a), 0.15 * np.sin(3 * a)] for a in angs])
    poses = [look_at(e, target) for e in eyes]
    pairs = _nearest_pairs(eyes, target - np.array([0, 0, 0]), n_src)
    return _finish("plane", width, height, faces, [K] * n_views, poses, pairs, jpeg)


def _table_scene_faces(seed: int) -> List[Face]:
    faces = [Face(o=np.array([-1.6, -1.6, 0.0]), eu=np.array([3.2, 0, 0.0]), ev=np.array([0, 3.2, 0.0]),
                  base=120, amp=55, freq=6.0, seed=seed * 100 + 1, octaves=5)]
    faces += box_faces([-0.45, -0.30, 0.0], [0.25, 0.30, 0.42], seed=seed * 100 + 2, base=135, amp=60, freq=9.0, octaves=5)
    faces += box_faces([0.35, -0.55, 0.0], [0.75, -0.10, 0.25], seed=seed * 100 + 3, base=110, amp=65, freq=11.0, octaves=5)
    faces += box_faces([-0.20, 0.45, 0.0], [0.30, 0.80, 0.18], seed=seed * 100 + 4, base=140, amp=50, freq=10.0, octaves=5)
    return faces


def make_dtu_scene(width=1600, height=1200, grid=7, n_src=10, seed=2, jpeg=True, views=None, workers=0) -> SyntheticScene:
    """Config 2: DTU-shaped, grid x grid views on a spherical cap over boxes on a table."""
    f = 2890.0 * width / 1600.0
    K = _intrinsics(f, width, height)
    faces = _table_scene_faces(seed)
    target = np.array([0.1, 0.05, 0.15])
    az = np.radians(np.linspace(-24, 24, grid))
    el = np.radians(np.linspace(34, 64, grid))
    eyes = []
    for e in el:
        for a in az:
            eyes.append(target + 3.0 * np.array([np.cos(e) * np.sin(a), -np.cos(e) * np.cos(a), np.sin(e)]))
    eyes = np.array(eyes)
    poses = [look_at(e, target) for e in eyes]
    pairs = _nearest_pairs(eyes, target, n_src)
    return _finish("dtu", width, height, faces, [K] * len(eyes), poses, pairs, jpeg, views, workers)


def make_eth3d_scene(width=3200, height=2130, n_views=11, n_src=10, seed=3, jpeg=True, views=None, workers=0) -> SyntheticScene:
    """Config 3: ETH3D-shaped indoor room, large weakly-textured walls plus textured furniture blocks."""
    f = 1800.0 * width / 3200.0
    K = _intrinsics(f, width, height)
    faces = box_faces([-3.0, -2.5, 0.0], [3.0, 2.5, 3.0], inward=True, seed=seed * 100 + 1,
                      base=170, amp=5, freq=5.0, octaves=5)  # weak texture: +-5 grey levels
    faces[0].amp, faces[0].base, faces[0].freq = 25, 110, 7.0  # floor (z = lo) has a visible pattern
    faces += box_faces([-1.6, 1.2, 0.0], [-0.4, 2.2, 0.9], seed=seed * 100 + 2, base=120, amp=60, freq=8.0, octaves=5)
    faces += box_faces([0.3, 1.4, 0.0], [1.5, 2.3, 1.5], seed=seed * 100 + 3, base=100, amp=55, freq=7.0, octaves=5)
    faces += box_faces([-0.5, 0.2, 0.0], [0.4, 0.9, 0.5], seed=seed * 100 + 4, base=140, amp=60, freq=10.0, octaves=5)
    faces += box_faces([1.9, -0.5, 0.0], [2.8, 0.6, 2.0], seed=seed * 100 + 5, base=125, amp=50, freq=6.0, octaves=5)
    target = np.array([0.0, 1.6, 1.0])
    xs = np.linspace(-1.2, 1.2, n_views)
    eyes = np.stack([[x, -2.1 + 0.12 * np.cos(2.5 * x), 1.45 + 0.10 * np.sin(3.0 * x)] for x in xs])
    poses = [look_at(e, target + np.array([0.25 * x, 0, 0])) for e, x in zip(eyes, xs)]
    pairs = {}
    for i in range(n_views):
        order = sorted((j for j in range(n_views) if j != i), key=lambda j: (abs(j - i), j))[:n_src]
        pairs[i] = [(j, float(100.0 / (1 + abs(j - i)))) for j in order]
    return _finish("eth3d", width, height, faces, [K] * n_views, poses, pairs, jpeg, views, workers)


def make_tnt_scene(width=1920, height=1080, n_views=300, n_src=10, seed=4, jpeg=True, views=None, workers=0) -> SyntheticScene:
    """Config 4: Tanks-and-Temples-shaped ring of views around textured blocks on a textured ground."""
    f = 1160.0 * width / 1920.0
    K = _intrinsics(f, width, height)
    faces = [Face(o=np.array([-9.0, -9.0, 0.0]), eu=np.array([18.0, 0, 0.0]), ev=np.array([0, 18.0, 0.0]),
                  base=115, amp=55, freq=2.5, seed=seed * 100 + 1, octaves=6)]
    faces += box_faces([-1.0, -0.7, 0.0], [1.0, 0.7, 1.6], seed=seed * 100 + 2, base=135, amp=60, freq=5.0, octaves=5)
    faces += box_faces([-0.5, -0.4, 1.6], [0.5, 0.4, 2.3], seed=seed * 100 + 3, base=105, amp=60, freq=7.0, octaves=5)
    faces += box_faces([1.4, -0.3, 0.0], [2.0, 0.3, 0.6], seed=seed * 100 + 4, base=150, amp=50, freq=8.0, octaves=5)
    faces += box_faces([-2.2, 0.6, 0.0], [-1.5, 1.3, 0.9], seed=seed * 100 + 5, base=125, amp=55, freq=8.0, octaves=5)
    target = np.array([0.0, 0.0, 0.9])
    angs = np.arange(n_views) * (2 * np.pi / max(n_views, 300))  # same angular spacing as the 300-view ring
    eyes = np.stack([[5.5 * np.cos(a), 5.5 * np.sin(a), 1.7 + 0.25 * np.sin(5 * a)] for a in angs])
    poses = [look_at(e, target) for e in eyes]
    pairs = {}
    full_ring = n_views >= 300
    for i in range(n_views):
        cand = []
        for k in range(1, n_views):
            for j in (i + k, i - k):
                jj = j % n_views if full_ring else j
                if 0 <= jj < n_views and jj != i and jj not in [c for c, _ in cand]:
                    cand.append((jj, float(100.0 / (1 + k))))
            if len(cand) >= n_src:
                break
        pairs[i] = cand[:n_src]
    return _finish("tnt", width, height, faces, [K] * n_views, poses, pairs, jpeg, views, workers)


SCENES = {"plane": make_plane_scene, "dtu": make_dtu_scene, "eth3d": make_eth3d_scene, "tnt": make_tnt_scene}


# ----------------------------------------------------------------------------- dense folder
def write_pgm(path: str, img8: np.ndarray) -> None:
    h, w = img8.shape
    with open(path, "wb") as f:
        f.write(f"P5\n{w} {h}\n255\n".encode())
        f.write(np.ascontiguousarray(img8, dtype=np.uint8).tobytes())


def read_pgm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P5"
        w, h = (int(v) for v in f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h), dtype=np.uint8).reshape(h, w).copy()


def write_dense_folder(scene: SyntheticScene, root: str, jpeg_quality: int = 97) -> None:
    import cv2

    for sub in ("images", "cams", "gt"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    for i, cam in enumerate(scene.cams):
        write_cam(os.path.join(root, "cams", f"{i:08d}_cam.txt"), cam)
        if scene.images[i].size == 0:
            continue
        jpg = os.path.join(root, "images", f"{i:08d}.jpg")
        cv2.imwrite(jpg, scene.images[i], [int(cv2.IMWRITE_JPEG_QUALITY), jpeg_quality])
        dec = cv2.imread(jpg, cv2.IMREAD_GRAYSCALE)
        scene.images[i] = dec  # what a reader of the folder will see
        write_pgm(os.path.join(root, "images", f"{i:08d}.pgm"), dec)
        write_dmb(os.path.join(root, "gt", f"{i:08d}_depth.dmb"), scene.gt_depth[i])
        write_dmb(os.path.join(root, "gt", f"{i:08d}_normal.dmb"), scene.gt_normal[i])
    write_pairs(os.path.join(root, "pair.txt"), scene.pairs)


# ----------------------------------------------------------------------------- GT metrics
def depth_normal_agreement(depth_a, normal_a, depth_b, normal_b, valid, rel=0.01, deg=5.0) -> float:
    """Fraction of valid pixels within `rel` relative depth and `deg` degrees normal (north-star parity metric)."""
    da = np.asarray(depth_a, np.float64)
    db = np.asarray(depth_b, np.float64)
    ok_d = np.abs(da - db) <= rel * np.abs(db)
    cosang = np.clip(np.sum(np.asarray(normal_a, np.float64) * np.asarray(normal_b, np.float64), axis=-1), -1, 1)
    ok_n = np.degrees(np.arccos(cosang)) <= deg
    v = np.asarray(valid, bool)
    return float((ok_d & ok_n & v).sum() / max(1, v.sum()))


def accuracy_at(depth, gt_depth, thresholds=(0.02, 0.05, 0.10)) -> List[float]:
    """Per-view depth accuracy: % of GT-valid pixels whose |z - z_gt| is below each threshold (scene units = m)."""
    v = gt_depth > 0
    err = np.abs(np.asarray(depth, np.float64) - gt_depth)
    return [float(100.0 * ((err < t) & v).sum() / max(1, v.sum())) for t in thresholds]


def accuracy_completeness_at(depth, cost, gt_depth, thresholds=(0.02, 0.05, 0.10), max_cost=0.5):
    """Accuracy and completeness of one depth map against the rendered ground truth, the depth-map form of the two MVS
    measures: a pixel is RECONSTRUCTED when its matching cost is below `max_cost` (the reference never marks a pixel
    invalid, /root/reference/src/PatchMatch.cpp:610-618; the cost is what its fusion and its planar prior gate on);
    accuracy(t)     = % of the reconstructed pixels that have ground truth and lie within t of it      (precision)
    completeness(t) = % of the ground-truth pixels that are reconstructed and lie within t of the truth  (recall).
    Scene units are metres, so t = 0.02 / 0.05 / 0.10 are the north star's 2 / 5 / 10 cm. Returns (accuracy, completeness)."""
    v = np.asarray(gt_depth) > 0
    rec = (np.asarray(cost) < max_cost) & np.isfinite(depth)
    err = np.abs(np.asarray(depth, np.float64) - gt_depth)
    acc = [float(100.0 * ((err < t) & v & rec).sum() / max(1, (v & rec).sum())) for t in thresholds]
    comp = [float(100.0 * ((err < t) & v & rec).sum() / max(1, v.sum())) for t in thresholds]
    return acc, comp
