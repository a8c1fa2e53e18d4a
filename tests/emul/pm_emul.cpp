// tests/emul/pm_emul.cpp -- TEST-ONLY host instantiation of mp-mvs_b200/csrc/pm_core.cuh.
//
// There is no GPU in the build container, so the per-pixel logic of the CUDA kernels (the templates
// in pm_core.cuh) is instantiated here with a host context and compared against the CPU oracle
// (oracle/pm_oracle.c) by tests/test_emul_vs_oracle.py. This catches logic errors before GPU time
// is spent. It is NOT a fallback: nothing in the product library (libmpmvs_b200.so) links, loads or
// calls this file; it is built into tests/emul/libpm_emul.so by the tests only.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mp-mvs_b200/csrc/pm_views.h"

namespace {

struct Emu {
    int n = 0, W = 0, H = 0;
    mpmvs_camera cams[MPMVS_MAX_VIEWS];
    std::vector<std::vector<float>> images, depths;
    PmView views[PM_MAX_SRC];
    int max_iterations = 3, top_k = 4, max_scale = 2;
    bool geom = false, geomPlanarPrior = false, planar = false;
    float depth_min = 0, depth_max = 1;
    std::vector<pm_f4> planes, prior;
    std::vector<float> costs, geomc;
    std::vector<uint32_t> vmask, rng, mask;
};

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// software model of the texture unit: un-normalised coordinates, clamp, 8 fractional weight bits
float tex2d(const float* img, int w, int h, float x, float y) {
    const float xb = x - 0.5f, yb = y - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float a = floorf((xb - fx) * 256.f + 0.5f) / 256.f, b = floorf((yb - fy) * 256.f + 0.5f) / 256.f;
    int i0 = xb > -1e9f ? (int)fx : 0, j0 = yb > -1e9f ? (int)fy : 0;
    const int i1 = clampi(i0 + 1, 0, w - 1), j1 = clampi(j0 + 1, 0, h - 1);
    i0 = clampi(i0, 0, w - 1); j0 = clampi(j0, 0, h - 1);
    const float t00 = img[(size_t)j0 * w + i0], t10 = img[(size_t)j0 * w + i1];
    const float t01 = img[(size_t)j1 * w + i0], t11 = img[(size_t)j1 * w + i1];
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}

struct HostCtx {
    const Emu* E;
    int x, y;
    mutable float wtab[36];   // the thread's bilateral-weight table (shared memory in the kernels)
    float& wt(int k) const { return wtab[k]; }
    float ref(int dx, int dy) const {
        return E->images[0][(size_t)clampi(y + dy, 0, E->H - 1) * E->W + clampi(x + dx, 0, E->W - 1)];
    }
    float src(int v, float xs, float ys) const {
        return tex2d(E->images[v + 1].data(), E->cams[v + 1].width, E->cams[v + 1].height, xs, ys);
    }
    float src_depth(int v, int xi, int yi) const {
        const PmView& V = E->views[v];
        return V.depth[(size_t)clampi(yi, 0, V.dh - 1) * V.dpitch + clampi(xi, 0, V.dw - 1)];
    }
    const PmView& view(int v) const { return E->views[v]; }
};

PmFrame frame(const Emu* E) {
    return pm_make_frame(E->cams[0], E->n, E->depth_min, E->depth_max, 5.0f, 3.0f, E->top_k, E->geom, E->planar);
}
PmState state(Emu* E) {
    PmState S;
    S.planes = E->planes.data(); S.costs = E->costs.data(); S.views = E->vmask.data(); S.rng = E->rng.data();
    S.geom = E->geomc.data(); S.prior = E->prior.data(); S.mask = E->mask.data(); S.counters = nullptr;
    return S;
}

template <int SCALE>
void sweep(Emu* E, int red, int iter) {
    const PmFrame F = frame(E);
    const PmState S = state(E);
    const int ty_end = 16 * (((E->H / 2) + 15) / 16);
    std::vector<float> ca(8 * PM_MAX_SRC);
    for (int ty = 0; ty < ty_end; ++ty)
        for (int x = 0; x < E->W; ++x) {
            const int y = 2 * ty + ((x & 1) ^ (red ? 1 : 0));
            if (y >= E->H) continue;
            HostCtx c{E, x, y};
            pm_sweep_pixel<SCALE>(c, F, S, x, y, iter, PmTableLocal{ca.data(), F.nsrc});
        }
}
void sweep_any(Emu* E, int red, int iter, int scale) {
    if (scale == 0) sweep<0>(E, red, iter);
    else if (scale == 1) sweep<1>(E, red, iter);
    else sweep<2>(E, red, iter);
}
void init(Emu* E, uint64_t seed) {
    const PmFrame F = frame(E);
    const PmState S = state(E);
    for (int y = 0; y < E->H; ++y)
        for (int x = 0; x < E->W; ++x) {
            HostCtx c{E, x, y};
            pm_init_pixel<2>(c, F, S, x, y, seed);
        }
}
void finalize(Emu* E) {
    const PmFrame F = frame(E);
    for (int y = 0; y < E->H; ++y)
        for (int x = 0; x < E->W; ++x) E->planes[y * E->W + x] = pm_depth_normal(F, E->planes[y * E->W + x], x, y);
    const int ty_end = 16 * (((E->H / 2) + 15) / 16);
    for (int red = 0; red < 2; ++red)
        for (int ty = 0; ty < ty_end; ++ty)
            for (int x = 0; x < E->W; ++x) {
                const int y = 2 * ty + ((x & 1) ^ red);
                if (y >= E->H) continue;
                E->planes[y * E->W + x].w = pm_median_depth(E->planes.data(), E->costs.data(), E->W, E->H, x, y);
            }
}
template <int SCALE>
void ncc_map(Emu* E, const pm_f4* pl, float* out) {
    const PmFrame F = frame(E);
    for (int y = 0; y < E->H; ++y)
        for (int x = 0; x < E->W; ++x) {
            HostCtx c{E, x, y};
            const PmRefStats st = pm_ref_stats<SCALE>(c, F);
            const PmHyp hyp = pm_hyp(F, pl[y * E->W + x], x, y);
            uint32_t nexec = 0;
            for (int v = 0; v < F.nsrc; ++v)
                out[(size_t)v * E->W * E->H + (size_t)y * E->W + x] = pm_ncc<SCALE>(c, F, st, v, hyp, x, y, nexec);
        }
}
}  // namespace

extern "C" {
void* emu_create() { return new Emu(); }
int emu_sizeof_camera() { return (int)sizeof(mpmvs_camera); }
int emu_set_problem(void* h, int n, const float* const* images, const void* cams) {
    Emu* E = (Emu*)h;
    E->n = n;
    memcpy(E->cams, cams, sizeof(mpmvs_camera) * n);
    E->W = E->cams[0].width; E->H = E->cams[0].height;
    E->depth_min = E->cams[0].depth_min * 0.6f;
    E->depth_max = E->cams[0].depth_max * 1.2f;
    E->images.resize(n);
    for (int i = 0; i < n; ++i) E->images[i].assign(images[i], images[i] + (size_t)E->cams[i].width * E->cams[i].height);
    for (int v = 1; v < n; ++v) { pm_build_view_consts(E->cams[0], E->cams[v], E->views[v - 1]); E->views[v - 1].layer = v; }
    const size_t wh = (size_t)E->W * E->H;
    E->planes.assign(wh, pm_f4{0, 0, 0, 0}); E->prior.assign(wh, pm_f4{0, 0, 0, 0});
    E->costs.assign(wh, 0.f); E->geomc.assign(wh, 0.f); E->vmask.assign(wh, 0u); E->mask.assign(wh, 0u); E->rng.assign(wh * 6, 0u);
    return 0;
}
void emu_set_geom_consistency_params(void* h, int geom, int planar) {
    Emu* E = (Emu*)h;
    E->geom = geom != 0;
    if (geom) { E->max_iterations = 2; E->geomPlanarPrior = planar != 0; } else E->max_iterations = 3;
}
void emu_set_planar_prior_params(void* h) { ((Emu*)h)->planar = true; }
int emu_set_src_depths(void* h, const float* const* d) {
    Emu* E = (Emu*)h;
    E->depths.resize(E->n - 1);
    for (int v = 0; v < E->n - 1; ++v) {
        E->depths[v].assign(d[v], d[v] + (size_t)E->cams[v + 1].width * E->cams[v + 1].height);
        E->views[v].depth = E->depths[v].data();
    }
    return 0;
}
int emu_set_state(void* h, const float* planes4, const float* costs) {
    Emu* E = (Emu*)h;
    memcpy(E->planes.data(), planes4, E->planes.size() * sizeof(pm_f4));
    memcpy(E->costs.data(), costs, E->costs.size() * sizeof(float));
    return 0;
}
int emu_set_prior(void* h, const float* prior4, const uint32_t* mask) {
    Emu* E = (Emu*)h;
    memcpy(E->prior.data(), prior4, E->prior.size() * sizeof(pm_f4));
    memcpy(E->mask.data(), mask, E->mask.size() * sizeof(uint32_t));
    return 0;
}
int emu_init_only(void* h, uint64_t seed) { init((Emu*)h, seed); return 0; }
int emu_half_sweep(void* h, int red, int iter, int scale) { sweep_any((Emu*)h, red, iter, scale); return 0; }
int emu_finalize(void* h) { finalize((Emu*)h); return 0; }
float emu_run(void* h, uint64_t seed) {
    Emu* E = (Emu*)h;
    init(E, seed);
    if (E->geom || E->planar) {
        for (int i = 0; i < E->max_iterations; ++i) { sweep_any(E, 0, i, 0); sweep_any(E, 1, i, 0); }
    } else {
        for (int s = E->max_scale; s >= 0; --s)
            for (int i = 0; i < E->max_iterations; ++i) { sweep_any(E, 0, i, s); sweep_any(E, 1, i, s); }
    }
    finalize(E);
    return 0.f;
}
int emu_get_result(void* h, float* planes4, float* costs, float* geom) {
    Emu* E = (Emu*)h;
    if (planes4) memcpy(planes4, E->planes.data(), E->planes.size() * sizeof(pm_f4));
    if (costs) memcpy(costs, E->costs.data(), E->costs.size() * sizeof(float));
    if (geom) memcpy(geom, E->geomc.data(), E->geomc.size() * sizeof(float));
    return 0;
}
float emu_depth_min(void* h) { return ((Emu*)h)->depth_min; }
float emu_depth_max(void* h) { return ((Emu*)h)->depth_max; }
int emu_get_device_state(void* h, float* planes4, float* costs, uint32_t* views, uint32_t* rng6, float* geom) {
    Emu* E = (Emu*)h;
    emu_get_result(h, planes4, costs, geom);
    if (views) memcpy(views, E->vmask.data(), E->vmask.size() * 4);
    if (rng6) memcpy(rng6, E->rng.data(), E->rng.size() * 4);
    return 0;
}
int emu_set_device_state(void* h, const float* planes4, const float* costs, const uint32_t* views, const uint32_t* rng6,
                         const float* geom) {
    Emu* E = (Emu*)h;
    if (planes4) memcpy(E->planes.data(), planes4, E->planes.size() * sizeof(pm_f4));
    if (costs) memcpy(E->costs.data(), costs, E->costs.size() * 4);
    if (views) memcpy(E->vmask.data(), views, E->vmask.size() * 4);
    if (rng6) memcpy(E->rng.data(), rng6, E->rng.size() * 4);
    if (geom) memcpy(E->geomc.data(), geom, E->geomc.size() * 4);
    return 0;
}
int emu_ncc_map(void* h, const float* planes4, int scale, float* out) {
    Emu* E = (Emu*)h;
    const pm_f4* pl = (const pm_f4*)planes4;
    if (scale == 0) ncc_map<0>(E, pl, out);
    else if (scale == 1) ncc_map<1>(E, pl, out);
    else ncc_map<2>(E, pl, out);
    return 0;
}
int emu_geom_map(void* h, const float* planes4, float* out) {
    Emu* E = (Emu*)h;
    if (E->depths.empty()) return -1;
    const PmFrame F = frame(E);
    const pm_f4* pl = (const pm_f4*)planes4;
    for (int y = 0; y < E->H; ++y)
        for (int x = 0; x < E->W; ++x) {
            HostCtx c{E, x, y};
            for (int v = 0; v < F.nsrc; ++v)
                out[(size_t)v * E->W * E->H + (size_t)y * E->W + x] = pm_geom_cost(c, F, v, pl[y * E->W + x], x, y);
        }
    return 0;
}
int emu_uniform_stream(uint64_t seed, int x, int y, int n, float* out) {
    PmRng rs;
    pm_rng_init(rs, pm_mix_seed(seed, (uint32_t)x, (uint32_t)y));
    for (int i = 0; i < n; ++i) out[i] = pm_uniform(rs);
    return 0;
}
void emu_destroy(void* h) { delete (Emu*)h; }
}
