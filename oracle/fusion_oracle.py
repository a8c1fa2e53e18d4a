"""TEST INFRASTRUCTURE: numpy restatement of depth-map fusion, RunFusion (/root/reference/src/PatchMatch.cpp:287-504).

Two orders are offered:
  * fuse(..., per_image_snapshot=True): the pixels of one image are tested against the masks as of the start of that image
    and a point masks only the source pixels it used itself -- the semantics of the GPU kernel (mp-mvs_b200/csrc/pm_fusion.cu),
    vectorised per image; must match the kernel point for point;
  * the reference's strictly sequential order (pixel after pixel, stale `used_list` included) is restated in C++ in
    mp-mvs_b200/csrc/mpmvs_main.cpp (RunFusion); tests compare the two statistically.
Only tests/ may import this module.
"""
import numpy as np


def _world(x, y, depth, cam):
    K, R, C = cam["K"].reshape(3, 3), cam["R"].reshape(3, 3), cam["C"]
    f32 = np.float32
    px = depth * (x - K[0, 2]) / K[0, 0]
    py = depth * (y - K[1, 2]) / K[1, 1]
    pz = depth
    X = np.stack([R[0, 0] * px + R[1, 0] * py + R[2, 0] * pz + C[0],
                  R[0, 1] * px + R[1, 1] * py + R[2, 1] * pz + C[1],
                  R[0, 2] * px + R[1, 2] * py + R[2, 2] * pz + C[2]], -1)
    return X.astype(f32)


def _project(X, cam):
    K, R, t = cam["K"].reshape(3, 3), cam["R"].reshape(3, 3), cam["t"]
    tx = R[0, 0] * X[..., 0] + R[0, 1] * X[..., 1] + R[0, 2] * X[..., 2] + t[0]
    ty = R[1, 0] * X[..., 0] + R[1, 1] * X[..., 1] + R[1, 2] * X[..., 2] + t[1]
    tz = R[2, 0] * X[..., 0] + R[2, 1] * X[..., 1] + R[2, 2] * X[..., 2] + t[2]
    depth = K[2, 0] * tx + K[2, 1] * ty + K[2, 2] * tz
    with np.errstate(all="ignore"):
        u = (K[0, 0] * tx + K[0, 1] * ty + K[0, 2] * tz) / depth
        v = (K[1, 0] * tx + K[1, 1] * ty + K[1, 2] * tz) / depth
    return u.astype(np.float32), v.astype(np.float32), depth.astype(np.float32)


def fuse(cams, depths, normals, grays, src_lists, dynamic=True, sky=None):
    """cams: packed camera records; depths/normals/grays: per image arrays; src_lists[i] = [i, sources...] or None.
    sky: optional per-image uint8 masks (> 0 = sky, None = no mask): masked when the image's turn comes (cpp:385-388).
    Returns (n, 9) float32 points in the kernel's order (images in order, raster order inside an image)."""
    n = len(cams)
    # colour: (h, w, 3) B, G, R images as RunFusion reads them (cpp:322), or grey (h, w) images counted thrice
    grays = [g.astype(np.float32) if g.ndim == 3 else np.repeat(g.astype(np.float32)[..., None], 3, -1) for g in grays]
    masks = [np.zeros(d.shape, bool) for d in depths]
    out = []
    for i in range(n):
        if src_lists[i] is None:
            continue
        h, w = depths[i].shape
        ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
        ref_depth = depths[i].astype(np.float32)
        if sky is not None and sky[i] is not None:
            masks[i] |= np.asarray(sky[i]) > 0
        alive = (~masks[i]) & (ref_depth > 0)
        PX = _world(xs, ys, ref_depth, cams[i])
        rn = normals[i].astype(np.float32)
        sumP, sumN, sumC = PX.copy(), rn.copy(), grays[i].copy()
        numc = np.zeros((h, w), np.int32)
        dyn = np.zeros((h, w), np.float32)
        num_ngb = len(src_lists[i])
        used = {}
        new_masks = {}
        for j in range(1, num_ngb):
            act = alive.copy()
            if j == num_ngb - 1:
                act &= numc > 0
            s = src_lists[i][j]
            if s is None or s < 0 or src_lists[s] is None:
                continue
            sh, sw = depths[s].shape
            u, v, _ = _project(PX, cams[s])
            with np.errstate(all="ignore"):
                sr = np.where(np.isfinite(v), v + np.float32(0.5), -1e9).astype(np.int64)   # int() truncation
                sc = np.where(np.isfinite(u), u + np.float32(0.5), -1e9).astype(np.int64)
                sr = np.where(v + np.float32(0.5) < 0, np.ceil(v + np.float32(0.5)), sr).astype(np.int64)
                sc = np.where(u + np.float32(0.5) < 0, np.ceil(u + np.float32(0.5)), sc).astype(np.int64)
            inb = (sc >= 0) & (sc < sw) & (sr >= 0) & (sr < sh)
            act &= inb
            src_r, src_c = np.clip(sr, 0, sh - 1), np.clip(sc, 0, sw - 1)
            act &= ~masks[s][src_r, src_c]
            sd = depths[s][src_r, src_c].astype(np.float32)
            act &= sd > 0
            TX = _world(src_c.astype(np.float32), src_r.astype(np.float32), sd, cams[s])
            bu, bv, pd = _project(TX, cams[i])
            with np.errstate(all="ignore"):
                reproj = np.sqrt((xs - bu) ** 2 + (ys - bv) ** 2).astype(np.float32)
                rel = (np.abs(pd - ref_depth) / ref_depth).astype(np.float32)
                sn = normals[s][src_r, src_c].astype(np.float32)
                ang = np.arccos((rn * sn).sum(-1).astype(np.float32)).astype(np.float32)
            ang = np.where(np.isnan(ang), np.float32(0), ang)
            ok = act & (reproj < 2.0) & (rel < 0.01) & (ang < np.float32(0.174533))
            sumP[ok] += TX[ok]
            sumN[ok] += sn[ok]
            sumC[ok] += grays[s][src_r, src_c][ok]
            with np.errstate(all="ignore"):
                dyn[ok] += np.exp(-(reproj + 200 * rel + ang * 10)).astype(np.float32)[ok]
            numc[ok] += 1
            used[j] = (ok, src_r, src_c, s)
        keep = alive & ((numc >= 1) & (dyn > np.float32(0.3) * numc) if dynamic else (numc >= 2))
        # cpp:454-460: point and colour are divided by (n + 1); the normal is a cv::Vec3f, whose operator/=(float) multiplies
        # by the float reciprocal (OpenCV core/matx.hpp) -- pinned by the reference's own RunFusion, tests/test_reference_program.py
        n1 = (numc + 1.0).astype(np.float32)
        inv = (np.float32(1.0) / n1).astype(np.float32)
        pts = np.concatenate([(sumP / n1[..., None]).astype(np.float32), (sumN * inv[..., None]).astype(np.float32),
                              (sumC / n1[..., None]).astype(np.float32)], -1)
        out.append(pts[keep])
        for j, (ok, src_r, src_c, s) in used.items():
            m = ok & keep
            new_masks.setdefault(s, np.zeros(depths[s].shape, bool))[src_r[m], src_c[m]] = True
        for s, m in new_masks.items():
            masks[s] |= m
    return np.concatenate(out, 0).astype(np.float32) if out else np.zeros((0, 9), np.float32)
