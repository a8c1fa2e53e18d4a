"""TEST INFRASTRUCTURE: restatement of the host planar-prior stage of ProcessProblem
(/root/reference/src/PatchMatch.cpp:532-609 and the members it calls), in numpy + OpenCV.

cv2.Subdiv2D is the reference's own triangulator (cv::Subdiv2D, PatchMatch.cpp:766-771; OpenCV version unpinned by the
reference, README.md:5 -- 4.13 here), so the triangulation golden values come from the reference's dependency itself;
the rest is restated line by line with the arithmetic types of the C++ source (float32 where it uses float, float64 where
a double literal promotes). Pure-Python loops: small cases only. Only tests/ may import this module.
"""
import numpy as np


def triangulate_vertices(costs, geom_costs=None):
    """GetTriangulateVertices, PatchMatch.cpp:782-853. geom_costs given -> the geomPlanarPrior variant."""
    h, w = costs.shape
    out = []
    for row in range(0, h, 5):
        for col in range(0, w, 5):
            c_bound, r_bound = min(w, col + 5), min(h, row + 5)
            if geom_costs is None:
                min_cost, pt = np.float32(2.0), None
                for r in range(row, r_bound):
                    for c in range(col, c_bound):
                        cost = costs[r, c]
                        if cost < 2.0 and min_cost > cost:
                            pt, min_cost = (c, r), cost
                if min_cost < np.float32(0.1):
                    out.append(pt)
            else:
                mc = [np.float32(2.0)] * 3
                pts = [(0, 0)] * 3
                cost_sum = np.float32(0.0)
                for r in range(row, r_bound):
                    for c in range(col, c_bound):
                        cost = costs[r, c]
                        cost_sum = np.float32(cost_sum + cost)
                        if cost < 1.0 and geom_costs[r, c] < np.float32(0.4) and cost < mc[2]:
                            mc[2], pts[2] = cost, (c, r)
                            for i in (1, 0):
                                if mc[i] <= mc[i + 1]:
                                    break
                                mc[i], mc[i + 1] = mc[i + 1], mc[i]
                                pts[i], pts[i + 1] = pts[i + 1], pts[i]
                cost_sum = np.float32(np.float64(np.float32(cost_sum / np.float32(r_bound * c_bound))) * 0.85)   # :841
                thresh = max(cost_sum, np.float32(0.2))
                for i in range(3):
                    if mc[i] < thresh:
                        out.append(pts[i])
                    else:
                        break
    return np.array(out, dtype=np.int32).reshape(-1, 2)


def delaunay_cv(points, w, h):
    """DelaunayTriangulation, PatchMatch.cpp:757-780 + the imageRC.contains filter of :555. Returns (m, 3, 2) int pixels."""
    import cv2

    sd = cv2.Subdiv2D((0, 0, w, h))
    for x, y in points:
        sd.insert((float(x), float(y)))
    tris = []
    for t in sd.getTriangleList():
        p = [(int(t[0]), int(t[1])), (int(t[2]), int(t[3])), (int(t[4]), int(t[5]))]
        if all(0 <= x < w and 0 <= y < h for x, y in p):
            tris.append(p)
    return np.array(tris, dtype=np.int32).reshape(-1, 3, 2)


def solve_z(A):
    """cv::SVD::solveZ (modules/core/src/lapack.cpp: SVD svd(m, rows >= cols ? 0 : FULL_UV); dst = last row of vt) on a
    float32 matrix, through the SAME library the reference links (cv2 = OpenCV's Python binding; SVDecomp is cv::SVD::compute):
    the third-party arithmetic of GetPriorPlaneParams comes from the dependency itself, not from a restatement."""
    import cv2

    A = np.ascontiguousarray(A, np.float32)
    _, _, vt = cv2.SVDecomp(A, flags=0 if A.shape[0] >= A.shape[1] else cv2.SVD_FULL_UV)
    return vt[-1].astype(np.float32)


def plane_from_triangle(tri, depth, K):
    """GetPriorPlaneParams, PatchMatch.cpp:723-755: null vector of the 3x4 system [X 1] (cv::SVD::solveZ, float32)."""
    A = np.zeros((3, 4), np.float32)
    for k, (x, y) in enumerate(tri):
        d = np.float32(depth[y, x])
        A[k, 0] = d * (np.float32(x) - K[0, 2]) / K[0, 0]
        A[k, 1] = d * (np.float32(y) - K[1, 2]) / K[1, 1]
        A[k, 2] = d
        A[k, 3] = 1.0
    n4 = solve_z(A)
    norm2 = np.float32(np.sqrt(np.float64(n4[0]) ** 2 + np.float64(n4[1]) ** 2 + np.float64(n4[2]) ** 2))
    if n4[3] < 0:
        norm2 = -norm2
    return (n4 / norm2).astype(np.float32)


def rasterise(tris, w, h):
    """The sampling loop of ProcessProblem, PatchMatch.cpp:556-569 (float p, q, step; double (1.0 - p - q))."""
    mask = np.zeros((h, w), np.float32)
    for idx, ((x1, y1), (x2, y2), (x3, y3)) in enumerate(tris):
        L01 = np.float32(np.sqrt(float((x1 - x2) ** 2 + (y1 - y2) ** 2)))
        L02 = np.float32(np.sqrt(float((x1 - x3) ** 2 + (y1 - y3) ** 2)))
        L12 = np.float32(np.sqrt(float((x2 - x3) ** 2 + (y2 - y3) ** 2)))
        step = np.float32(1.0 / np.float64(max(L01, L02, L12)))
        p = np.float32(0)
        while p < 1.0:
            q = np.float32(0)
            while np.float64(q) < 1.0 - np.float64(p):
                r = 1.0 - np.float64(p) - np.float64(q)
                x = int(np.float64(np.float32(np.float32(p * np.float32(x1)) + np.float32(q * np.float32(x2)))) + r * x3)
                y = int(np.float64(np.float32(np.float32(p * np.float32(y1)) + np.float32(q * np.float32(y2)))) + r * y3)
                mask[y, x] = idx + 1.0
                q = np.float32(q + step)
            p = np.float32(p + step)
    return mask


def build_prior(planes_world, costs, K, depth_min, depth_max, geom_costs=None, tris=None):
    """The whole stage. planes_world: (h, w, 4) result of the first Run (world normal, depth). Returns
    (prior float32 (h, w, 4), mask uint32 (h, w), vertices, triangles)."""
    h, w = costs.shape
    K = np.asarray(K, np.float32).reshape(3, 3)
    verts = triangulate_vertices(costs, geom_costs)
    if tris is None:
        tris = delaunay_cv(verts, w, h)
    mask = rasterise(tris, w, h)
    depth = planes_world[..., 3]
    planes = np.array([plane_from_triangle(t, depth, K) for t in tris], np.float32).reshape(-1, 4)
    prior = np.zeros((h, w, 4), np.float32)
    out_mask = np.zeros((h, w), np.uint32)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    for j in range(h):
        for i in range(w):
            m = int(mask[j, i])
            if m > 0:
                n4 = planes[m - 1]
                with np.errstate(all="ignore"):
                    d = -n4[3] * fx / ((np.float32(i) - cx) * n4[0] + (fx / fy) * (np.float32(j) - cy) * n4[1] + fx * n4[2])   # :650-653
                if d <= depth_max and d >= depth_min:
                    out_mask[j, i] = m
                    prior[j, i] = n4
    return prior, out_mask, verts, tris


# ----------------------------------------------------------------------------- fast path (C loops in libpm_oracle.so)
def build_prior_fast(planes_world, costs, K, depth_min, depth_max, geom_costs=None, plane_fit="closed"):
    """Same stage with the loops in C (oracle/pm_oracle.c: pmo_pick_vertices, pmo_prior_from_triangles_fit) and OpenCV's
    Subdiv2D for the triangulation. plane_fit="closed": the plane through the three vertices in double -- what bench.py's
    reference arm runs as the reference pipeline's host stage (no 545 k SVD calls: in the reference's favour);
    plane_fit="svd": every triangle's plane from cv2.SVDecomp as in build_prior -- the reference's own fit, bit for bit
    (tests/test_reference_program.py) -- for whole-scene comparisons (tests/tools/scene_parity.py)."""
    import ctypes as C
    import os

    import cv2

    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpm_oracle.so"))
    h, w = costs.shape
    costs = np.ascontiguousarray(costs, np.float32)
    planes_world = np.ascontiguousarray(planes_world, np.float32)
    g = None if geom_costs is None else np.ascontiguousarray(geom_costs, np.float32)
    xy = np.empty((3 * ((w + 4) // 5) * ((h + 4) // 5), 2), np.int32)
    n = lib.pmo_pick_vertices(costs.ctypes.data_as(C.c_void_p), None if g is None else g.ctypes.data_as(C.c_void_p), w, h,
                              0 if g is None else 1, xy.ctypes.data_as(C.c_void_p))
    verts = xy[:n]
    sd = cv2.Subdiv2D((0, 0, w, h))
    if n:
        sd.insert(np.ascontiguousarray(verts, dtype=np.float32))
    tl = sd.getTriangleList() if n >= 3 else np.zeros((0, 6), np.float32)
    t = tl.astype(np.int32).reshape(-1, 3, 2)
    ok = np.all((t[..., 0] >= 0) & (t[..., 0] < w) & (t[..., 1] >= 0) & (t[..., 1] < h), axis=1)
    tris = np.ascontiguousarray(t[ok])
    prior = np.zeros((h, w, 4), np.float32)
    mask = np.zeros((h, w), np.uint32)
    Kf = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    tri_planes = None
    if plane_fit == "svd":
        K33 = np.asarray(K, np.float32).reshape(3, 3)
        depth = planes_world[..., 3]
        tri_planes = np.ascontiguousarray(np.array([plane_from_triangle(t, depth, K33) for t in tris], np.float32).reshape(-1, 4))
    lib.pmo_prior_from_triangles_fit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                                 C.c_void_p, C.c_void_p]
    cnt = lib.pmo_prior_from_triangles_fit(tris.ctypes.data, len(tris), None if tri_planes is None else tri_planes.ctypes.data,
                                           planes_world.ctypes.data, Kf.ctypes.data, w, h, C.c_float(depth_min), C.c_float(depth_max),
                                           prior.ctypes.data, mask.ctypes.data)
    return prior, mask, verts, tris, cnt
