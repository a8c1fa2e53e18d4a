// pm_subdiv.h -- the triangulation of the planar-prior host stage, in the ORDER the reference gets it.
//
// The reference triangulates its vertices with cv::Subdiv2D (insert every vertex, getTriangleList;
// /root/reference/src/PatchMatch.cpp:757-780) and then rasterises the triangles one after the other (:554-579), a later
// triangle overwriting an earlier one where their sampled pixels overlap. With vertices 5 px apart a fifth of all pixels lie
// on such overlaps (measured: tests/tools/reference_program.py), so the prior a pixel gets depends on the order of the
// triangle list, and that order is a property of the algorithm: the list walks the quad-edge records in the order they
// were allocated by the incremental insertion. A Delaunay triangulation from any other algorithm has the same triangles
// (up to hull slivers) in another order and gives 20 % of the pixels their neighbour's plane.
//
// OpenCV is a dependency of the reference, not part of it and not in this image as C++ (only cv2). This header restates
// the published algorithm of its planar subdivision (Guibas-Stolfi quad-edge structure, incremental insertion with edge
// swaps: walk from the most recent edge to the containing facet, connect the new point to the facet's vertices, swap edges
// until the in-circle test holds; three far-away virtual vertices bound the plane) with the same arithmetic -- float
// coordinates, double predicates, the same tolerances -- and the same allocation order of edge records, which is what fixes
// the order of the triangle list. tests/test_subdiv.py checks it against cv2.Subdiv2D itself: identical triangle lists
// (coordinates and order) on random, gridded, collinear and real vertex sets.
#ifndef MPMVS_PM_SUBDIV_H
#define MPMVS_PM_SUBDIV_H
#include <cfloat>
#include <cmath>
#include <utility>
#include <vector>

namespace pmsd {

class Subdivision {
  public:
    // the rectangle [x, x + w) x [y, y + h) that every inserted point must lie in
    Subdivision(int x, int y, int w, int h) {
        // the three virtual vertices: 6 x the longer side away, as cv2.Subdiv2D(rect).getVertex(1..3) of OpenCV 4.13 reports them
        // (older OpenCV releases used 3 x; the reference pins no version, README.md:5 -- a few hull slivers depend on it)
        const float big = 6.f * (float)(w > h ? w : h);
        const float rx = (float)x, ry = (float)y;
        x0_ = rx; y0_ = ry; x1_ = rx + w; y1_ = ry + h;
        vtx_.push_back(Vertex{0.f, 0.f, 0});               // record 0 is never used: 0 means "none" in both free lists
        quads_.push_back(Quad());
        free_quad_ = 0; free_vertex_ = 0;
        const int a = new_vertex(rx + big, ry), b = new_vertex(rx, ry + big), c = new_vertex(rx - big, ry - big);
        const int ab = new_edge(), bc = new_edge(), ca = new_edge();
        set_ends(ab, a, b); set_ends(bc, b, c); set_ends(ca, c, a);
        splice(ab, sym(ca)); splice(bc, sym(ab)); splice(ca, sym(bc));
        recent_ = ab;
    }

    // Inserts a point; returns its vertex number (4 for the first one; an existing vertex's number if the point coincides
    // with it), or -1 if the point is outside the rectangle or could not be located.
    int insert(float px, float py) {
        int edge = 0, vertex = 0;
        const int where = locate(px, py, edge, vertex);
        if (where == AT_VERTEX) return vertex;
        if (where != INSIDE && where != ON_EDGE) return -1;
        if (where == ON_EDGE) {                              // the edge the point lies on goes; its record is the next one reused
            const int doomed = edge;
            recent_ = edge = step(edge, PREV_AROUND_ORG);
            delete_edge(doomed);
        }
        const int point = new_vertex(px, py);
        int base = new_edge();
        const int first = org(edge);
        set_ends(base, first, point);
        splice(base, edge);
        do {                                                 // spokes from the new point to every vertex of the facet
            base = connect(edge, sym(base));
            edge = step(base, PREV_AROUND_ORG);
        } while (dst(edge) != first);
        edge = step(base, PREV_AROUND_ORG);
        const int limit = (int)quads_.size() * 4;
        for (int i = 0; i < limit; ++i) {                    // restore the empty-circle property around the new point
            const int t = step(edge, PREV_AROUND_ORG);
            const int t_dst = dst(t), e_org = org(edge), e_dst = dst(edge);
            if (right_of(vtx_[t_dst].x, vtx_[t_dst].y, edge) > 0 &&
                in_circle(vtx_[e_org], vtx_[t_dst], vtx_[e_dst], vtx_[point]) < 0) {
                swap_edge(edge);
                edge = step(edge, PREV_AROUND_ORG);
            } else if (e_org == first) {
                break;
            } else {
                edge = step(next(edge), PREV_AROUND_LEFT);
            }
        }
        return point;
    }

    // The facets whose three corners lie inside the rectangle, as vertex numbers (a, b, c per triangle), in the order of
    // the edge records: every directed edge with an even index is tried as the first side of the facet on its left.
    void triangles(std::vector<int>& out) const {
        out.clear();
        const int total = (int)quads_.size() * 4;
        std::vector<char> seen(total, 0);
        for (int e = 4; e < total; e += 2) {
            if (seen[e]) continue;
            const int a = org(e);
            if (!inside(a)) continue;
            const int eb = step(e, NEXT_AROUND_LEFT), b = org(eb);
            if (!inside(b)) continue;
            const int ec = step(eb, NEXT_AROUND_LEFT), c = org(ec);
            if (!inside(c)) continue;
            seen[e] = seen[eb] = seen[ec] = 1;
            out.push_back(a); out.push_back(b); out.push_back(c);
        }
    }

    float vx(int v) const { return vtx_[v].x; }
    float vy(int v) const { return vtx_[v].y; }

  private:
    struct Vertex { float x, y; int first; };
    // One undirected edge = four directed ones: e, its dual rotated by 90 degrees, its reverse, the reverse of the dual.
    // Directed edge number = 4 * record + rotation; next[r] = the next edge counter-clockwise around the origin of rotation r.
    struct Quad {
        int next[4], end[4];
        Quad() { for (int i = 0; i < 4; ++i) next[i] = end[i] = 0; }
        explicit Quad(int e) { next[0] = e; next[1] = e + 3; next[2] = e + 2; next[3] = e + 1; end[0] = end[1] = end[2] = end[3] = 0; }
    };
    // how to step from an edge to a neighbour: low nibble = rotation before following `next`, high nibble = rotation after
    enum { NEXT_AROUND_ORG = 0x00, NEXT_AROUND_DST = 0x22, PREV_AROUND_ORG = 0x11, PREV_AROUND_DST = 0x33,
           NEXT_AROUND_LEFT = 0x13, NEXT_AROUND_RIGHT = 0x31, PREV_AROUND_LEFT = 0x20, PREV_AROUND_RIGHT = 0x02 };
    enum { FAILED = -2, INSIDE = 0, AT_VERTEX = 1, ON_EDGE = 2 };

    std::vector<Vertex> vtx_;
    std::vector<Quad> quads_;
    int free_quad_, free_vertex_, recent_;
    float x0_, y0_, x1_, y1_;

    static int sym(int e) { return e ^ 2; }
    static int rot(int e, int r) { return (e & ~3) + ((e + r) & 3); }
    int next(int e) const { return quads_[e >> 2].next[e & 3]; }
    int step(int e, int how) const {
        e = quads_[e >> 2].next[(e + how) & 3];
        return (e & ~3) + ((e + (how >> 4)) & 3);
    }
    int org(int e) const { return quads_[e >> 2].end[e & 3]; }
    int dst(int e) const { return quads_[e >> 2].end[(e + 2) & 3]; }
    bool inside(int v) const { return x0_ <= vtx_[v].x && vtx_[v].x < x1_ && y0_ <= vtx_[v].y && vtx_[v].y < y1_; }

    int new_vertex(float x, float y) {
        if (free_vertex_ == 0) { vtx_.push_back(Vertex{0.f, 0.f, 0}); free_vertex_ = (int)vtx_.size() - 1; }
        const int v = free_vertex_;
        free_vertex_ = vtx_[v].first;
        vtx_[v] = Vertex{x, y, 0};
        return v;
    }
    int new_edge() {
        if (free_quad_ <= 0) { quads_.push_back(Quad()); free_quad_ = (int)quads_.size() - 1; }
        const int e = free_quad_ * 4;
        free_quad_ = quads_[e >> 2].next[1];
        quads_[e >> 2] = Quad(e);
        return e;
    }
    void delete_edge(int e) {
        splice(e, step(e, PREV_AROUND_ORG));
        const int s = sym(e);
        splice(s, step(s, PREV_AROUND_ORG));
        const int q = e >> 2;
        quads_[q].next[0] = 0;
        quads_[q].next[1] = free_quad_;
        free_quad_ = q;
    }
    void set_ends(int e, int o, int d) {
        quads_[e >> 2].end[e & 3] = o;
        quads_[e >> 2].end[(e + 2) & 3] = d;
        vtx_[o].first = e;
        vtx_[d].first = e ^ 2;
    }
    void splice(int a, int b) {                              // exchanges the successors of a and b, and of their duals
        int& an = quads_[a >> 2].next[a & 3];
        int& bn = quads_[b >> 2].next[b & 3];
        const int ar = rot(an, 1), br = rot(bn, 1);
        int& arn = quads_[ar >> 2].next[ar & 3];
        int& brn = quads_[br >> 2].next[br & 3];
        std::swap(an, bn);
        std::swap(arn, brn);
    }
    int connect(int a, int b) {                              // a new edge from the end of a to the start of b
        const int e = new_edge();
        splice(e, step(a, NEXT_AROUND_LEFT));
        splice(sym(e), b);
        set_ends(e, dst(a), org(b));
        return e;
    }
    void swap_edge(int e) {                                  // the other diagonal of the quadrilateral around e
        const int s = sym(e), a = step(e, PREV_AROUND_ORG), b = step(s, PREV_AROUND_ORG);
        splice(e, a);
        splice(s, b);
        set_ends(e, dst(a), dst(b));
        splice(e, step(a, NEXT_AROUND_LEFT));
        splice(s, step(b, NEXT_AROUND_LEFT));
    }
    static double area2(float ax, float ay, float bx, float by, float cx, float cy) {
        return ((double)bx - ax) * ((double)cy - ay) - ((double)by - ay) * ((double)cx - ax);
    }
    int right_of(float px, float py, int e) const {
        const Vertex &o = vtx_[org(e)], &d = vtx_[dst(e)];
        const double a = area2(px, py, d.x, d.y, o.x, o.y);
        return (a > 0) - (a < 0);
    }
    static int in_circle(const Vertex& p, const Vertex& a, const Vertex& b, const Vertex& c) {
        const double eps = FLT_EPSILON * 0.125;
        double v = ((double)a.x * a.x + (double)a.y * a.y) * area2(b.x, b.y, c.x, c.y, p.x, p.y);
        v -= ((double)b.x * b.x + (double)b.y * b.y) * area2(a.x, a.y, c.x, c.y, p.x, p.y);
        v += ((double)c.x * c.x + (double)c.y * c.y) * area2(a.x, a.y, b.x, b.y, p.x, p.y);
        v -= ((double)p.x * p.x + (double)p.y * p.y) * area2(a.x, a.y, b.x, b.y, c.x, c.y);
        return v > eps ? 1 : v < -eps ? -1 : 0;
    }
    // Walks from the most recently used edge towards the point; ends on an edge of the facet that contains it.
    int locate(float px, float py, int& out_edge, int& out_vertex) {
        out_edge = out_vertex = 0;
        if (px < x0_ || py < y0_ || px >= x1_ || py >= y1_) return FAILED;
        const int limit = (int)quads_.size() * 4;
        int e = recent_;
        int r_cur = right_of(px, py, e);
        if (r_cur > 0) { e = sym(e); r_cur = -r_cur; }
        bool found = false;
        for (int i = 0; i < limit; ++i) {
            const int on = next(e), dp = step(e, PREV_AROUND_DST);
            const int r_on = right_of(px, py, on), r_dp = right_of(px, py, dp);
            if (r_dp > 0) {
                if (r_on > 0 || (r_on == 0 && r_cur == 0)) { found = true; break; }
                r_cur = r_on; e = on;
            } else if (r_on > 0) {
                if (r_dp == 0 && r_cur == 0) { found = true; break; }
                r_cur = r_dp; e = dp;
            } else if (r_cur == 0 && right_of(vtx_[dst(on)].x, vtx_[dst(on)].y, e) >= 0) {
                e = sym(e);
            } else {
                r_cur = r_on; e = on;
            }
        }
        recent_ = e;
        if (!found) return FAILED;
        const Vertex &o = vtx_[org(e)], &d = vtx_[dst(e)];
        double t1 = fabs(px - o.x); t1 += fabs(py - o.y);
        double t2 = fabs(px - d.x); t2 += fabs(py - d.y);
        double t3 = fabs(o.x - d.x); t3 += fabs(o.y - d.y);
        if (t1 < FLT_EPSILON) { out_vertex = org(e); return AT_VERTEX; }
        if (t2 < FLT_EPSILON) { out_vertex = dst(e); return AT_VERTEX; }
        out_edge = e;
        if ((t1 < t3 || t2 < t3) && fabs(area2(px, py, o.x, o.y, d.x, d.y)) < FLT_EPSILON) return ON_EDGE;
        return INSIDE;
    }
};

// Triangulates integer pixel positions (inserted in the given order) inside the image rectangle. `tris` receives vertex
// INDICES into xy (three per triangle, in the list order described above). Returns false if a point could not be inserted.
inline bool triangulate(const int* xy, int n, int W, int H, std::vector<int>& tris) {
    Subdivision sd(0, 0, W, H);
    std::vector<int> index_of;                               // subdivision vertex number -> first input index that created it
    index_of.assign(4, -1);
    bool ok = true;
    for (int i = 0; i < n; ++i) {
        const int v = sd.insert((float)xy[2 * i], (float)xy[2 * i + 1]);
        if (v < 0) { ok = false; continue; }
        if (v >= (int)index_of.size()) index_of.resize(v + 1, -1);
        if (index_of[v] < 0) index_of[v] = i;
    }
    std::vector<int> t;
    sd.triangles(t);
    tris.clear();
    tris.reserve(t.size());
    for (int v : t) tris.push_back(index_of[v]);
    return ok;
}

}  // namespace pmsd
#endif
