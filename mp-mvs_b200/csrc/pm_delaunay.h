// pm_delaunay.h -- exact incremental Delaunay triangulation of integer pixel coordinates (host, header-only).
//
// Replaces cv::Subdiv2D in PatchMatchCUDA::DelaunayTriangulation (/root/reference/src/PatchMatch.cpp:757-780), which
// OpenCV-free builds cannot call. Vertices are pixel positions (one per 5x5 cell, GetTriangulateVertices :782-853), so
// all predicates are evaluated exactly in 64/128-bit integers: no epsilons, no robustness fallbacks. Points are
// inserted in the caller's order when that is a scan (as cv::Subdiv2D does), else band by band, with Lawson flips. The triangulation of co-circular
// point sets is not unique (common on a pixel grid); like OpenCV's, the result is then ONE valid Delaunay triangulation.
#ifndef MPMVS_PM_DELAUNAY_H
#define MPMVS_PM_DELAUNAY_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace pmd {

struct Tri {
    int v[3];  // counter-clockwise (y down: "CCW" in the mathematical sense of the orient() sign below)
    int n[3];  // n[i] = triangle across the edge opposite v[i], -1 = none
};

class Delaunay {
  public:
    // xy: n points (x0,y0,x1,y1,...) with 0 <= x < width, 0 <= y < height. Duplicates are ignored.
    Delaunay(const int* xy, int n, int width, int height) : n_(n) {
        px_.resize(n + 3);
        py_.resize(n + 3);
        for (int i = 0; i < n; ++i) { px_[i] = xy[2 * i]; py_[i] = xy[2 * i + 1]; }
        // The bounding triangle of cv::Subdiv2D::initDelaunay for the rect (0, 0, width, height): with these three outer
        // vertices the hull slivers that OpenCV's triangulation lacks are missing here too (same prior coverage at the borders).
        const int64_t m = 3 * (int64_t)(width > height ? width : height);
        small_ = m <= 3 * 4096;
        px_[n] = m;      py_[n] = 0;
        px_[n + 1] = 0;  py_[n + 1] = m;
        px_[n + 2] = -m; py_[n + 2] = -m;
        tris_.reserve(2 * (size_t)n + 8);
        Tri t;
        t.v[0] = n; t.v[1] = n + 1; t.v[2] = n + 2;
        if (orient(t.v[0], t.v[1], t.v[2]) < 0) { int s = t.v[1]; t.v[1] = t.v[2]; t.v[2] = s; }
        t.n[0] = t.n[1] = t.n[2] = -1;
        tris_.push_back(t);
        last_ = 0;
        // Insertion order. Co-circular vertex quadruples are common on a pixel grid and which of their two diagonals
        // survives depends on the order of insertion. cv::Subdiv2D inserts in the caller's order; GetTriangulateVertices
        // (and mpmvs_pick_vertices) emit the cells row by row, so inserting in that order reproduces OpenCV's choice: every
        // triangle built here is then one of OpenCV's (tests/test_prior_stage.py), against 99.3 % with a re-sorted order.
        // A caller's order is kept when consecutive points are close to each other (a row-major scan jumps once per row);
        // anything else (e.g. a shuffled set) would make every point-location walk O(sqrt n), so it is re-sorted into
        // bands of rows walked boustrophedon.
        std::vector<int> order((size_t)n);
        double path = 0.0;
        for (int i = 1; i < n; ++i) path += std::abs(xy[2 * i] - xy[2 * i - 2]) + std::abs(xy[2 * i + 1] - xy[2 * i - 1]);
        const double spacing = n > 0 ? std::sqrt((double)width * height / n) : 1.0;
        const bool keep_caller_order = n < 2 || path / (n - 1) <= 8.0 * spacing;
        if (keep_caller_order) {
            for (int i = 0; i < n; ++i) order[i] = i;
        } else {
            const int band = n > 0 ? (int)std::max(1.0, std::sqrt((double)width * height / n)) : 1;
            const int n_bands = height / band + 1;
            std::vector<int> start((size_t)n_bands + 1, 0);
            for (int i = 0; i < n; ++i) ++start[xy[2 * i + 1] / band + 1];
            for (int b = 0; b < n_bands; ++b) start[b + 1] += start[b];
            {
                std::vector<int> fill(start.begin(), start.end() - 1);
                for (int i = 0; i < n; ++i) order[fill[xy[2 * i + 1] / band]++] = i;     // counting sort by band (stable)
            }
            for (int b = 0; b < n_bands; ++b) {                                          // then by x, alternating direction
                const bool rev = b & 1;
                std::sort(order.begin() + start[b], order.begin() + start[b + 1], [&](int p, int q) {
                    const int xp = xy[2 * p], xq = xy[2 * q];
                    if (xp != xq) return rev ? xp > xq : xp < xq;
                    return xy[2 * p + 1] != xy[2 * q + 1] ? xy[2 * p + 1] < xy[2 * q + 1] : p < q;
                });
            }
        }
        for (int i = 0; i < n; ++i) insert(order[i]);
    }

    // triangles whose three vertices are input points, as vertex indices
    void triangles(std::vector<int>& out) const {
        out.clear();
        for (const Tri& t : tris_)
            if (t.v[0] < n_ && t.v[1] < n_ && t.v[2] < n_) { out.push_back(t.v[0]); out.push_back(t.v[1]); out.push_back(t.v[2]); }
    }
    size_t num_all_triangles() const { return tris_.size(); }

  private:
    int n_;
    std::vector<int64_t> px_, py_;
    std::vector<Tri> tris_;
    int last_;
    bool small_ = false;
    std::vector<int> stack_;

    int64_t orient(int a, int b, int c) const {
        return (px_[b] - px_[a]) * (py_[c] - py_[a]) - (py_[b] - py_[a]) * (px_[c] - px_[a]);
    }
    // > 0 iff d lies strictly inside the circumcircle of the positively oriented triangle (a, b, c).
    // |det| <= 12 D^4 for coordinate differences <= D: with the bounding triangle at 3 * max(w, h), 64-bit integers are
    // exact up to 4096-pixel images (D = 24576); larger images take the 128-bit path.
    int incircle_sign(int a, int b, int c, int d) const {
        if (small_) {
            const int64_t ax = px_[a] - px_[d], ay = py_[a] - py_[d];
            const int64_t bx = px_[b] - px_[d], by = py_[b] - py_[d];
            const int64_t cx = px_[c] - px_[d], cy = py_[c] - py_[d];
            const int64_t a2 = ax * ax + ay * ay, b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
            const int64_t det = ax * (by * c2 - b2 * cy) - ay * (bx * c2 - b2 * cx) + a2 * (bx * cy - by * cx);
            return det > 0 ? 1 : (det < 0 ? -1 : 0);
        }
        const __int128 det = incircle(a, b, c, d);
        return det > 0 ? 1 : (det < 0 ? -1 : 0);
    }
    __int128 incircle(int a, int b, int c, int d) const {
        const __int128 ax = px_[a] - px_[d], ay = py_[a] - py_[d];
        const __int128 bx = px_[b] - px_[d], by = py_[b] - py_[d];
        const __int128 cx = px_[c] - px_[d], cy = py_[c] - py_[d];
        const __int128 a2 = ax * ax + ay * ay, b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
        return ax * (by * c2 - b2 * cy) - ay * (bx * c2 - b2 * cx) + a2 * (bx * cy - by * cx);
    }
    static int nxt(int i) { return i == 2 ? 0 : i + 1; }
    static int prv(int i) { return i == 0 ? 2 : i - 1; }
    int index_of(const Tri& t, int vertex) const { return t.v[0] == vertex ? 0 : (t.v[1] == vertex ? 1 : 2); }
    int index_of_neighbour(const Tri& t, int tri) const { return t.n[0] == tri ? 0 : (t.n[1] == tri ? 1 : 2); }
    void relink(int tri, int old_nb, int new_nb) {
        if (tri < 0) return;
        Tri& t = tris_[tri];
        for (int i = 0; i < 3; ++i)
            if (t.n[i] == old_nb) { t.n[i] = new_nb; return; }
    }

    // visibility walk; returns the triangle containing p (closed) and, in `edge`, the index of an edge p lies on (or -1)
    int locate(int p, int& edge) const {
        int t = last_;
        for (size_t guard = 0; guard < 4 * tris_.size() + 16; ++guard) {
            const Tri& T = tris_[t];
            int moved = -1;
            for (int i = 0; i < 3; ++i) {
                if (orient(T.v[nxt(i)], T.v[prv(i)], p) < 0 && T.n[i] >= 0) { moved = T.n[i]; break; }
            }
            if (moved < 0) {
                edge = -1;
                for (int i = 0; i < 3; ++i)
                    if (orient(T.v[nxt(i)], T.v[prv(i)], p) == 0) edge = i;
                return t;
            }
            t = moved;
        }
        edge = -1;
        return t;
    }

    void insert(int p) {
        int edge;
        const int t0 = locate(p, edge);
        {
            const Tri& T = tris_[t0];
            for (int i = 0; i < 3; ++i)
                if (px_[T.v[i]] == px_[p] && py_[T.v[i]] == py_[p]) return;  // duplicate point
        }
        stack_.clear();
        if (edge < 0) {
            // split t0 = (a, b, c) into (a, b, p), (b, c, p), (c, a, p)
            const Tri T = tris_[t0];
            const int a = T.v[0], b = T.v[1], c = T.v[2];
            const int t1 = (int)tris_.size(), t2 = t1 + 1;
            tris_.resize(tris_.size() + 2);
            Tri& A = tris_[t0]; Tri& B = tris_[t1]; Tri& Cc = tris_[t2];
            A.v[0] = a; A.v[1] = b; A.v[2] = p; A.n[0] = t1; A.n[1] = t2; A.n[2] = T.n[2];
            B.v[0] = b; B.v[1] = c; B.v[2] = p; B.n[0] = t2; B.n[1] = t0; B.n[2] = T.n[0];
            Cc.v[0] = c; Cc.v[1] = a; Cc.v[2] = p; Cc.n[0] = t0; Cc.n[1] = t1; Cc.n[2] = T.n[1];
            relink(T.n[0], t0, t1);
            relink(T.n[1], t0, t2);
            stack_.push_back(t0); stack_.push_back(t1); stack_.push_back(t2);
        } else {
            // p lies on edge `edge` of t0 = (apex r, s, e) with the edge (s, e); split t0 and its neighbour u across it
            const Tri T = tris_[t0];
            const int r = T.v[edge], s = T.v[nxt(edge)], e = T.v[prv(edge)];
            const int u = T.n[edge];
            const int t1 = (int)tris_.size();
            tris_.resize(tris_.size() + 1);
            // t0 -> (r, s, p), t1 -> (r, p, e)
            const int n_rs = T.n[prv(edge)];  // across edge (r, s): opposite e
            const int n_er = T.n[nxt(edge)];  // across edge (e, r): opposite s
            if (u < 0) {
                Tri& A = tris_[t0]; Tri& B = tris_[t1];
                A.v[0] = r; A.v[1] = s; A.v[2] = p; A.n[0] = -1; A.n[1] = t1; A.n[2] = n_rs;
                B.v[0] = r; B.v[1] = p; B.v[2] = e; B.n[0] = -1; B.n[1] = n_er; B.n[2] = t0;
                relink(n_er, t0, t1);
                stack_.push_back(t0); stack_.push_back(t1);
            } else {
                const Tri U = tris_[u];
                const int ku = index_of_neighbour(U, t0);
                const int q = U.v[ku];                // apex of u; u = (q, e, s) positively oriented
                const int n_qe = U.n[prv(ku)];        // across (q, e): opposite s
                const int n_sq = U.n[nxt(ku)];        // across (s, q): opposite e
                const int t2 = (int)tris_.size();
                tris_.resize(tris_.size() + 1);
                Tri& A = tris_[t0]; Tri& B = tris_[t1]; Tri& Cc = tris_[u]; Tri& D = tris_[t2];
                A.v[0] = r; A.v[1] = s; A.v[2] = p; A.n[0] = t2; A.n[1] = t1; A.n[2] = n_rs;      // (r, s, p)
                B.v[0] = r; B.v[1] = p; B.v[2] = e; B.n[0] = u;  B.n[1] = n_er; B.n[2] = t0;      // (r, p, e)
                Cc.v[0] = q; Cc.v[1] = e; Cc.v[2] = p; Cc.n[0] = t1; Cc.n[1] = t2; Cc.n[2] = n_qe; // (q, e, p)
                D.v[0] = q; D.v[1] = p; D.v[2] = s; D.n[0] = t0; D.n[1] = n_sq; D.n[2] = u;       // (q, p, s)
                relink(n_er, t0, t1);
                relink(n_sq, u, t2);
                stack_.push_back(t0); stack_.push_back(t1); stack_.push_back(u); stack_.push_back(t2);
            }
        }
        // Lawson legalisation of the edges opposite p
        while (!stack_.empty()) {
            const int t = stack_.back();
            stack_.pop_back();
            Tri& T = tris_[t];
            const int kp = index_of(T, p);
            const int u = T.n[kp];
            if (u < 0) continue;
            Tri& U = tris_[u];
            const int ku = index_of_neighbour(U, t);
            const int q = U.v[ku];
            const int a = T.v[nxt(kp)], b = T.v[prv(kp)];   // shared edge (a, b); T = (p, a, b), U = (q, b, a)
            if (incircle_sign(p, a, b, q) <= 0) continue;
            // flip (a, b) -> (p, q):  T = (p, a, q), U = (p, q, b)
            const int n_pa = T.n[prv(kp)];   // across (p, a): opposite b
            const int n_bp = T.n[nxt(kp)];   // across (b, p): opposite a
            const int n_qb = U.n[prv(ku)];   // across (q, b): opposite a  [U = (q, b, a): nxt(ku) = b, prv(ku) = a]
            const int n_aq = U.n[nxt(ku)];   // across (a, q): opposite b
            T.v[0] = p; T.v[1] = a; T.v[2] = q; T.n[0] = n_aq; T.n[1] = u;    T.n[2] = n_pa;
            U.v[0] = p; U.v[1] = q; U.v[2] = b; U.n[0] = n_qb; U.n[1] = n_bp; U.n[2] = t;
            relink(n_aq, u, t);
            relink(n_bp, t, u);
            stack_.push_back(t);
            stack_.push_back(u);
        }
        last_ = (int)tris_.size() - 1;
    }
};

}  // namespace pmd
#endif
