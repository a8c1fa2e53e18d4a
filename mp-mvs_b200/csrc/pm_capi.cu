// pm_capi.cu -- implementation of the C ABI declared in include/mpmvs_b200.h.
//
// Host-side counterpart of the members the reference defines in src/PatchMatch.cpp:640-1139
// (PatchMatchInit / AllocatePatchMatch / CudaMemInit / CudaPlanarPriorInitialization / Release /
// getters) and of the launcher PatchMatchCUDA::Run (src/PatchMatch.cu:1188-1254), redesigned:
//   * all images of a GPU live in ONE layered float texture (mpmvs_image_cache); a problem only
//     records layer indices -- no per-problem cudaMallocArray / texture creation;
//   * per-pixel state buffers and pinned host mirrors are pooled inside the handle and reused when
//     the next reference image has the same size;
//   * a run is 22 launches on one stream with no host synchronisation in between (the reference
//     calls cudaDeviceSynchronize after every launch); results come back with async copies;
//   * source depth maps are read from linear device memory (so the buffer an NCCL all-gather fills
//     can be used in place) instead of one cudaArray + texture per source.
// There is no CPU fallback: without a CUDA device every entry point fails with MPMVS_E_NO_DEVICE.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/mpmvs_b200.h"
#include "pm_core.cuh"
#include "pm_subdiv.h"
#include "pm_kernels.h"
#include "pm_views.h"

static_assert(sizeof(mpmvs_camera) == 112, "mpmvs_camera must stay binary compatible with struct Camera");
static_assert(sizeof(PmView) % 4 == 0, "PmView is staged word-wise");

#define CK(call)                                  \
    do {                                          \
        cudaError_t e_ = (call);                  \
        if (e_ != cudaSuccess) return (int)e_;    \
    } while (0)

struct mpmvs_image_cache {
    int device = 0;
    int W = 0, H = 0, capacity = 0, used = 0;
    int fmt = MPMVS_TEX_F32;
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    std::unordered_map<int, int> id2layer;
    std::vector<int> lw, lh;  // per-layer image size
    void *stage_in = nullptr, *stage_out = nullptr;   // device staging for format conversion (W*H*4 bytes each)
};

struct mpmvs_problem {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n = 0, W = 0, H = 0;
    size_t wh = 0, wh_alloc = 0;
    mpmvs_camera cams[MPMVS_MAX_VIEWS];
    mpmvs_image_cache* cache = nullptr;
    bool own_cache = false;
    int tex_fmt = MPMVS_TEX_F32;       // storage format of the private cache (mpmvs_set_tex_format)
    int layers[MPMVS_MAX_VIEWS];
    // PatchMatchParams mirror (include/PatchMatch.h:48-67)
    int max_iterations = 3, top_k = 4, max_scale = 2;
    float sigma_spatial = 5.0f, sigma_color = 3.0f;
    bool geom = false, geomPlanarPrior = false, planar = false;
    float depth_min = 0.f, depth_max = 1.f;
    // device state
    PmState S{};
    pm_f4* d_prior = nullptr;
    uint32_t* d_mask = nullptr;
    PmView hviews[PM_MAX_SRC];
    PmView* dviews = nullptr;
    float* d_depths[PM_MAX_SRC];       // owned copies of source depth maps (host-upload path)
    size_t d_depths_cap[PM_MAX_SRC];
    bool has_depths = false, has_prior = false;
    // pinned host mirrors
    float *h_planes = nullptr, *h_costs = nullptr, *h_geom = nullptr;
    size_t h_alloc = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int last_launches = 0;
    bool ran = false;
    // profiling (mpmvs_set_profiling): bit 0 = an event between all launches, bit 1 = count executed NCC evaluations
    int profiling = 0;
    std::vector<cudaEvent_t> pev;          // pev[0] before init, pev[1] after init, pev[2+k] after sweep k, last after finalize
    int pev_used = 0;
    unsigned long long* d_counters = nullptr;
    // planar-prior stage scratch (mpmvs_build_prior)
    short2* d_cell_xy = nullptr; unsigned char* d_cell_n = nullptr; size_t cells_cap = 0;
    int2* d_vxy = nullptr; int3* d_tris = nullptr; pm_f4* d_tri_planes = nullptr; size_t vtx_cap = 0, tri_cap = 0;
    unsigned int* d_prior_count = nullptr;
    std::vector<short> h_cell_xy; std::vector<unsigned char> h_cell_n;
    int arith = MPMVS_ARITH_EXACT;     // which of the two compiled arithmetics the launches use (mpmvs_set_arithmetic)
    float lit_table[20];               // pm_literal_table evaluated on this device (mpmvs_create)
};

namespace {

int ensure_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return MPMVS_E_NO_DEVICE;
    if (device < 0 || device >= count) return MPMVS_E_ARG;
    CK(cudaSetDevice(device));
    return MPMVS_OK;
}

PmFrame make_frame(const mpmvs_problem* p) {
    PmFrame F = pm_make_frame(p->cams[0], p->n, p->depth_min, p->depth_max, p->sigma_spatial, p->sigma_color, p->top_k,
                              p->geom, p->planar);
    F.ref_layer = p->layers[0];
    F.tex_scale = p->cache->fmt == MPMVS_TEX_U8 ? 255.0f : 1.0f;
    int soft = 0;
    for (int i = 0; i < p->n; ++i)
        if (p->cams[i].width != p->cache->W || p->cams[i].height != p->cache->H) soft = 1;
    F.soft_clamp = soft;
    F.tex = (unsigned long long)p->cache->tex;
    for (int k = 0; k < 18; ++k) F.lit_sd[k / 6][k % 6] = p->lit_table[k];
    F.lit_rcp_spatial = p->lit_table[18];
    F.lit_rcp_color = p->lit_table[19];
    return F;
}

int cache_alloc(mpmvs_image_cache* c, int W, int H, int capacity, int fmt) {
    cudaChannelFormatDesc desc;
    if (fmt == MPMVS_TEX_F32) desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    else if (fmt == MPMVS_TEX_F16) desc = cudaCreateChannelDescHalf();
    else if (fmt == MPMVS_TEX_U8) desc = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    else return MPMVS_E_ARG;
    c->fmt = fmt;
    CK(cudaMalloc3DArray(&c->arr, &desc, make_cudaExtent((size_t)W, (size_t)H, (size_t)capacity), cudaArrayLayered));
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = c->arr;
    // CudaMemInit's descriptor (PatchMatch.cpp:1012-1018): linear filter, element type, un-normalised
    // coordinates; its Wrap mode is not defined for un-normalised coordinates and acts as clamp.
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeClamp;
    td.addressMode[1] = cudaAddressModeClamp;
    td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    // 8-bit texels are fetched as UNORM floats (v/255, the only filterable integer read mode); the kernels scale by 255
    td.readMode = fmt == MPMVS_TEX_U8 ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&c->tex, &res, &td, nullptr));
    c->W = W; c->H = H; c->capacity = capacity; c->used = 0;
    c->id2layer.clear();
    c->lw.assign(capacity, 0);
    c->lh.assign(capacity, 0);
    return MPMVS_OK;
}

void cache_free(mpmvs_image_cache* c) {
    cudaFree(c->stage_in); cudaFree(c->stage_out);
    c->stage_in = c->stage_out = nullptr;
    if (c->tex) cudaDestroyTextureObject(c->tex);
    if (c->arr) cudaFreeArray(c->arr);
    c->tex = 0; c->arr = nullptr; c->capacity = c->used = 0;
    c->id2layer.clear();
}

size_t texel_bytes(int fmt) { return fmt == MPMVS_TEX_F32 ? 4 : (fmt == MPMVS_TEX_F16 ? 2 : 1); }

// `src` holds texels already in the cache's storage format
int cache_upload(mpmvs_image_cache* c, int layer, const void* src, size_t pitch, int w, int h, cudaMemcpyKind kind,
                 cudaStream_t st) {
    cudaMemcpy3DParms prm;
    memset(&prm, 0, sizeof(prm));
    prm.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), pitch, (size_t)w, (size_t)h);
    prm.dstArray = c->arr;
    prm.dstPos = make_cudaPos(0, 0, (size_t)layer);
    prm.extent = make_cudaExtent((size_t)w, (size_t)h, 1);
    prm.kind = kind;
    CK(cudaMemcpy3DAsync(&prm, st));
    c->lw[layer] = w; c->lh[layer] = h;
    return MPMVS_OK;
}

// Any grey image (float32 or uint8, host or device, pitched) into layer `layer`, converting to the cache's storage
// format on the device when the two differ. All work is ordered on `st`.
int cache_put_any(mpmvs_image_cache* c, int layer, const void* src, bool src_u8, size_t pitch, int w, int h, bool on_device,
                  cudaStream_t st) {
    const bool same = (src_u8 && c->fmt == MPMVS_TEX_U8) || (!src_u8 && c->fmt == MPMVS_TEX_F32);
    if (same) return cache_upload(c, layer, src, pitch, w, h, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
    if (!c->stage_in) {
        CK(cudaMalloc(&c->stage_in, (size_t)c->W * c->H * 4));
        CK(cudaMalloc(&c->stage_out, (size_t)c->W * c->H * 4));
    }
    const size_t elt = src_u8 ? 1 : 4;
    const void* in = src;
    size_t in_pitch = pitch;
    if (!on_device) {
        CK(cudaMemcpy2DAsync(c->stage_in, (size_t)w * elt, src, pitch, (size_t)w * elt, (size_t)h, cudaMemcpyHostToDevice, st));
        in = c->stage_in;
        in_pitch = (size_t)w * elt;
    }
    CK(pm_launch_convert(in, in_pitch, src_u8 ? 1 : 0, c->stage_out, c->fmt, w, h, st));
    return cache_upload(c, layer, c->stage_out, (size_t)w * texel_bytes(c->fmt), w, h, cudaMemcpyDeviceToDevice, st);
}

int alloc_state(mpmvs_problem* p) {
    const size_t wh = p->wh;
    if (wh > p->wh_alloc) {
        cudaFree(p->S.planes); cudaFree(p->S.costs); cudaFree(p->S.views); cudaFree(p->S.rng); cudaFree(p->S.geom);
        cudaFree(p->d_prior); cudaFree(p->d_mask);
        p->S = PmState{}; p->d_prior = nullptr; p->d_mask = nullptr; p->wh_alloc = 0;
        CK(cudaMalloc((void**)&p->S.planes, wh * sizeof(pm_f4)));
        CK(cudaMalloc((void**)&p->S.costs, wh * sizeof(float)));
        CK(cudaMalloc((void**)&p->S.views, wh * sizeof(uint32_t)));
        CK(cudaMalloc((void**)&p->S.rng, wh * 6 * sizeof(uint32_t)));
        CK(cudaMalloc((void**)&p->S.geom, wh * sizeof(float)));
        CK(cudaMalloc((void**)&p->d_prior, wh * sizeof(pm_f4)));
        CK(cudaMalloc((void**)&p->d_mask, wh * sizeof(uint32_t)));
        p->wh_alloc = wh;
    }
    p->S.prior = p->d_prior;
    p->S.mask = p->d_mask;
    p->S.counters = nullptr;
    CK(cudaMemsetAsync(p->S.views, 0, wh * sizeof(uint32_t), p->stream));
    CK(cudaMemsetAsync(p->S.geom, 0, wh * sizeof(float), p->stream));
    CK(cudaMemsetAsync(p->d_mask, 0, wh * sizeof(uint32_t), p->stream));
    if (!p->dviews) CK(cudaMalloc((void**)&p->dviews, PM_MAX_SRC * sizeof(PmView)));
    return MPMVS_OK;
}

int ensure_mirrors(mpmvs_problem* p) {
    if (p->wh > p->h_alloc) {
        cudaFreeHost(p->h_planes); cudaFreeHost(p->h_costs); cudaFreeHost(p->h_geom);
        p->h_planes = p->h_costs = p->h_geom = nullptr; p->h_alloc = 0;
        CK(cudaMallocHost((void**)&p->h_planes, p->wh * sizeof(pm_f4)));
        CK(cudaMallocHost((void**)&p->h_costs, p->wh * sizeof(float)));
        CK(cudaMallocHost((void**)&p->h_geom, p->wh * sizeof(float)));
        p->h_alloc = p->wh;
    }
    return MPMVS_OK;
}

// common tail of the three set_views flavours: cameras, depth range, per-view constants, state buffers
int finish_views(mpmvs_problem* p, int n, const mpmvs_camera* cams) {
    p->n = 0;   // published at the end: after a failure every entry point that needs views answers MPMVS_E_ARG
    memcpy(p->cams, cams, sizeof(mpmvs_camera) * n);
    p->W = cams[0].width; p->H = cams[0].height;
    p->wh = (size_t)p->W * p->H;
    p->depth_min = cams[0].depth_min * 0.6f;   // PatchMatch.cpp:929-930
    p->depth_max = cams[0].depth_max * 1.2f;
    for (int v = 1; v < n; ++v) {
        pm_build_view_consts(cams[0], cams[v], p->hviews[v - 1]);
        p->hviews[v - 1].layer = p->layers[v];
    }
    p->has_depths = false;
    p->has_prior = false;
    p->ran = false;
    int rc = alloc_state(p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(p->dviews, p->hviews, sizeof(PmView) * (n - 1), cudaMemcpyHostToDevice, p->stream));
    p->n = n;
    return MPMVS_OK;
}

int check_views_args(mpmvs_problem* p, int n, const void* a, const mpmvs_camera* cams) {
    if (!p || !a || !cams || n < 2 || n > MPMVS_MAX_VIEWS) return MPMVS_E_ARG;
    for (int i = 0; i < n; ++i)
        if (cams[i].width <= 0 || cams[i].height <= 0) return MPMVS_E_ARG;
    return MPMVS_OK;
}

int private_cache(mpmvs_problem* p, int n, const mpmvs_camera* cams) {
    int mw = 0, mh = 0;
    for (int i = 0; i < n; ++i) { mw = cams[i].width > mw ? cams[i].width : mw; mh = cams[i].height > mh ? cams[i].height : mh; }
    if (p->cache && !p->own_cache) p->cache = nullptr;
    if (p->cache && (p->cache->W != mw || p->cache->H != mh || p->cache->capacity < n || p->cache->fmt != p->tex_fmt)) {
        cache_free(p->cache);
        delete p->cache;
        p->cache = nullptr;
    }
    if (!p->cache) {
        p->cache = new (std::nothrow) mpmvs_image_cache();
        if (!p->cache) return MPMVS_E_ARG;
        p->cache->device = p->device;
        p->own_cache = true;
        int rc = cache_alloc(p->cache, mw, mh, n, p->tex_fmt);
        if (rc) return rc;
    }
    for (int i = 0; i < n; ++i) p->layers[i] = i;
    return MPMVS_OK;
}

int sync_frame_views(mpmvs_problem* p) {
    // depth pointers may have changed since finish_views
    CK(cudaMemcpyAsync(p->dviews, p->hviews, sizeof(PmView) * (p->n - 1), cudaMemcpyHostToDevice, p->stream));
    return MPMVS_OK;
}

// PatchMatchCUDA::Run's schedule, PatchMatch.cu:1198-1244
int enqueue_run(mpmvs_problem* p, uint64_t seed) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    if (p->geom && !p->has_depths) return MPMVS_E_STATE;
    if (p->planar && !p->has_prior) return MPMVS_E_STATE;
    CK(cudaSetDevice(p->device));
    const PmFrame F = make_frame(p);
    const PmLaunchers& L = pm_launchers(p->arith);
    PmState S = p->S;
    S.counters = nullptr;
    if (p->profiling & 2) {
        if (!p->d_counters) CK(cudaMalloc((void**)&p->d_counters, 2 * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(p->d_counters, 0, 2 * sizeof(unsigned long long), p->stream));
        S.counters = p->d_counters;
    }
    p->pev_used = 0;
    auto mark = [&]() -> cudaError_t {
        if (!(p->profiling & 1)) return cudaSuccess;
        if ((size_t)p->pev_used >= p->pev.size()) {
            cudaEvent_t e;
            cudaError_t rc = cudaEventCreate(&e);
            if (rc != cudaSuccess) return rc;
            p->pev.push_back(e);
        }
        return cudaEventRecord(p->pev[p->pev_used++], p->stream);
    };
    CK(cudaEventRecord(p->ev0, p->stream));
    CK(mark());
    int launches = 0;
    CK(L.init(F, S, p->dviews, seed, p->stream));
    CK(mark());
    ++launches;
    if (p->geom || p->planar) {
        for (int i = 0; i < p->max_iterations; ++i)
            for (int red = 0; red < 2; ++red) {
                CK(L.sweep(F, S, p->dviews, red, i, 0, p->stream));
                CK(mark());
                ++launches;
            }
    } else {
        for (int s = p->max_scale; s >= 0; --s)
            for (int i = 0; i < p->max_iterations; ++i)
                for (int red = 0; red < 2; ++red) {
                    CK(L.sweep(F, S, p->dviews, red, i, s, p->stream));
                    CK(mark());
                    ++launches;
                }
    }
    CK(L.finalize(F, S, p->stream));
    CK(mark());
    launches += 3;
    CK(cudaEventRecord(p->ev1, p->stream));
    p->last_launches = launches;
    p->ran = true;
    return MPMVS_OK;
}

}  // namespace

extern "C" {

int mpmvs_version(void) { return 100; }

const char* mpmvs_build_flavor(void) { return "exact+fast"; }

const char* mpmvs_arithmetic_name(int arithmetic) {
    return arithmetic == MPMVS_ARITH_EXACT ? "exact" : (arithmetic == MPMVS_ARITH_FAST ? "fast" : "?");
}

int mpmvs_default_arithmetic(void) {
    const char* e = getenv("MPMVS_ARITHMETIC");
    if (e && (!strcmp(e, "fast") || !strcmp(e, "1"))) return MPMVS_ARITH_FAST;
    return MPMVS_ARITH_EXACT;
}

int mpmvs_set_arithmetic(mpmvs_problem* p, int arithmetic) {
    if (!p || (arithmetic != MPMVS_ARITH_EXACT && arithmetic != MPMVS_ARITH_FAST)) return MPMVS_E_ARG;
    p->arith = arithmetic;
    return MPMVS_OK;
}

int mpmvs_get_arithmetic(mpmvs_problem* p, int* arithmetic) {
    if (!p || !arithmetic) return MPMVS_E_ARG;
    *arithmetic = p->arith;
    return MPMVS_OK;
}

const char* mpmvs_error_string(int code) {
    switch (code) {
        case MPMVS_OK: return "ok";
        case MPMVS_E_ARG: return "invalid argument or call order";
        case MPMVS_E_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case MPMVS_E_STATE: return "required input missing for the selected mode";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int mpmvs_create(int device, void* stream, mpmvs_problem** out) {
    if (!out) return MPMVS_E_ARG;
    int rc = ensure_device(device);
    if (rc) return rc;
    mpmvs_problem* p = new (std::nothrow) mpmvs_problem();
    if (!p) return MPMVS_E_ARG;
    p->device = device;
    memset(p->d_depths, 0, sizeof(p->d_depths));
    memset(p->d_depths_cap, 0, sizeof(p->d_depths_cap));
    if (stream) {
        p->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete p; return (int)e; }
        p->own_stream = true;
    }
    cudaEventCreate(&p->ev0);
    cudaEventCreate(&p->ev1);
    p->arith = mpmvs_default_arithmetic();
    {   // the constants of the exact arithmetic's bilateral weight as THIS device's MUFU unit evaluates them
        float* d = nullptr;
        cudaError_t e = cudaMalloc((void**)&d, sizeof(p->lit_table));
        if (e == cudaSuccess) e = pm_launch_literal_table(p->sigma_spatial, p->sigma_color, d, p->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->lit_table, d, sizeof(p->lit_table), cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        cudaFree(d);
        if (e != cudaSuccess) { mpmvs_destroy(p); return (int)e; }
    }
    *out = p;
    return MPMVS_OK;
}

int mpmvs_destroy(mpmvs_problem* p) {
    if (!p) return MPMVS_E_ARG;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    cudaFree(p->S.planes); cudaFree(p->S.costs); cudaFree(p->S.views); cudaFree(p->S.rng); cudaFree(p->S.geom);
    cudaFree(p->d_prior); cudaFree(p->d_mask); cudaFree(p->dviews);
    for (int i = 0; i < PM_MAX_SRC; ++i) cudaFree(p->d_depths[i]);
    cudaFreeHost(p->h_planes); cudaFreeHost(p->h_costs); cudaFreeHost(p->h_geom);
    if (p->own_cache && p->cache) { cache_free(p->cache); delete p->cache; }
    cudaEventDestroy(p->ev0);
    cudaEventDestroy(p->ev1);
    for (cudaEvent_t e : p->pev) cudaEventDestroy(e);
    cudaFree(p->d_counters);
    cudaFree(p->d_cell_xy); cudaFree(p->d_cell_n); cudaFree(p->d_vxy); cudaFree(p->d_tris); cudaFree(p->d_tri_planes);
    cudaFree(p->d_prior_count);
    if (p->own_stream) cudaStreamDestroy(p->stream);
    delete p;
    return MPMVS_OK;
}

// ---------------------------------------------------------------------------------------------- image cache
int mpmvs_cache_create(int device, int max_width, int max_height, int capacity, mpmvs_image_cache** out) {
    return mpmvs_cache_create_fmt(device, max_width, max_height, capacity, MPMVS_TEX_F32, out);
}

int mpmvs_cache_create_fmt(int device, int max_width, int max_height, int capacity, int tex_format, mpmvs_image_cache** out) {
    if (!out || max_width <= 0 || max_height <= 0 || capacity <= 0 || capacity > 2048) return MPMVS_E_ARG;
    int rc = ensure_device(device);
    if (rc) return rc;
    mpmvs_image_cache* c = new (std::nothrow) mpmvs_image_cache();
    if (!c) return MPMVS_E_ARG;
    c->device = device;
    rc = cache_alloc(c, max_width, max_height, capacity, tex_format);
    if (rc) { delete c; return rc; }
    *out = c;
    return MPMVS_OK;
}

int mpmvs_cache_destroy(mpmvs_image_cache* c) {
    if (!c) return MPMVS_E_ARG;
    cudaSetDevice(c->device);
    cache_free(c);
    delete c;
    return MPMVS_OK;
}

static int cache_layer_for(mpmvs_image_cache* c, int image_id, int w, int h) {
    if (w > c->W || h > c->H) return -1;
    auto it = c->id2layer.find(image_id);
    if (it != c->id2layer.end()) return it->second;
    if (c->used >= c->capacity) return -1;
    const int layer = c->used++;
    c->id2layer[image_id] = layer;
    return layer;
}

static int cache_put_host(mpmvs_image_cache* c, int image_id, const void* gray_host, bool u8, int width, int height) {
    if (!c || !gray_host) return MPMVS_E_ARG;
    CK(cudaSetDevice(c->device));
    const int layer = cache_layer_for(c, image_id, width, height);
    if (layer < 0) return MPMVS_E_ARG;
    int rc = cache_put_any(c, layer, gray_host, u8, (size_t)width * (u8 ? 1 : 4), width, height, false, 0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(0));
    return MPMVS_OK;
}

int mpmvs_cache_put(mpmvs_image_cache* c, int image_id, const float* gray_host, int width, int height) {
    return cache_put_host(c, image_id, gray_host, false, width, height);
}

// uint8 grey levels; in a float cache they become exactly what PatchMatchInit's convertTo(CV_32FC1) gives (PatchMatch.cpp:882)
int mpmvs_cache_put_u8(mpmvs_image_cache* c, int image_id, const uint8_t* gray_host, int width, int height) {
    return cache_put_host(c, image_id, gray_host, true, width, height);
}

int mpmvs_cache_has(mpmvs_image_cache* c, int image_id) { return c && c->id2layer.count(image_id) ? 1 : 0; }

// ---------------------------------------------------------------------------------------------- inputs
static int set_views_any(mpmvs_problem* p, int n, const void* const* imgs, bool u8, const size_t* pitch_bytes, bool on_device,
                         const mpmvs_camera* cams) {
    int rc = check_views_args(p, n, imgs, cams);
    if (rc) return rc;
    CK(cudaSetDevice(p->device));
    p->n = 0;
    rc = private_cache(p, n, cams);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        const size_t pitch = pitch_bytes ? pitch_bytes[i] : (size_t)cams[i].width * (u8 ? 1 : 4);
        rc = cache_put_any(p->cache, i, imgs[i], u8, pitch, cams[i].width, cams[i].height, on_device, p->stream);
        if (rc) return rc;
    }
    return finish_views(p, n, cams);
}

int mpmvs_set_views(mpmvs_problem* p, int n, const float* const* gray_host, const mpmvs_camera* cams) {
    return set_views_any(p, n, (const void* const*)gray_host, false, nullptr, false, cams);
}

int mpmvs_set_views_u8(mpmvs_problem* p, int n, const uint8_t* const* gray_host, const mpmvs_camera* cams) {
    return set_views_any(p, n, (const void* const*)gray_host, true, nullptr, false, cams);
}

int mpmvs_set_views_device(mpmvs_problem* p, int n, const float* const* gray_dev, const size_t* pitch_bytes,
                           const mpmvs_camera* cams) {
    return set_views_any(p, n, (const void* const*)gray_dev, false, pitch_bytes, true, cams);
}

int mpmvs_set_tex_format(mpmvs_problem* p, int tex_format) {
    if (!p || tex_format < MPMVS_TEX_F32 || tex_format > MPMVS_TEX_U8) return MPMVS_E_ARG;
    p->tex_fmt = tex_format;
    return MPMVS_OK;
}

int mpmvs_set_views_cached(mpmvs_problem* p, mpmvs_image_cache* c, int n, const int* image_ids, const mpmvs_camera* cams) {
    int rc = check_views_args(p, n, image_ids, cams);
    if (rc) return rc;
    if (!c || c->device != p->device) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    p->n = 0;
    if (p->own_cache && p->cache) { cache_free(p->cache); delete p->cache; }
    p->cache = c;
    p->own_cache = false;
    for (int i = 0; i < n; ++i) {
        auto it = c->id2layer.find(image_ids[i]);
        if (it == c->id2layer.end()) return MPMVS_E_STATE;
        if (c->lw[it->second] != cams[i].width || c->lh[it->second] != cams[i].height) return MPMVS_E_ARG;
        p->layers[i] = it->second;
    }
    return finish_views(p, n, cams);
}

int mpmvs_set_geom_consistency_params(mpmvs_problem* p, int geom_consistency, int planar_prior) {
    if (!p) return MPMVS_E_ARG;
    p->geom = geom_consistency != 0;   // PatchMatch.cpp:655-665
    if (geom_consistency) {
        p->max_iterations = 2;
        p->geomPlanarPrior = planar_prior != 0;
    } else {
        p->max_iterations = 3;
    }
    return MPMVS_OK;
}

int mpmvs_reset_params(mpmvs_problem* p) {
    if (!p) return MPMVS_E_ARG;
    p->max_iterations = 3; p->geom = false; p->geomPlanarPrior = false; p->planar = false;  // PatchMatch.h:48-67 defaults
    p->has_prior = false;
    return MPMVS_OK;
}

int mpmvs_set_planar_prior_params(mpmvs_problem* p) {
    if (!p) return MPMVS_E_ARG;
    p->planar = true;                  // PatchMatch.cpp:667-670
    return MPMVS_OK;
}

int mpmvs_set_src_depths(mpmvs_problem* p, const float* const* depth_host) {
    if (!p || !depth_host || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    for (int v = 0; v < p->n - 1; ++v) {
        const mpmvs_camera& sc = p->cams[v + 1];
        const size_t bytes = (size_t)sc.width * sc.height * sizeof(float);
        if (bytes > p->d_depths_cap[v]) {
            cudaFree(p->d_depths[v]);
            p->d_depths[v] = nullptr; p->d_depths_cap[v] = 0;
            CK(cudaMalloc((void**)&p->d_depths[v], bytes));
            p->d_depths_cap[v] = bytes;
        }
        CK(cudaMemcpyAsync(p->d_depths[v], depth_host[v], bytes, cudaMemcpyHostToDevice, p->stream));
        p->hviews[v].depth = p->d_depths[v];
        p->hviews[v].dw = sc.width; p->hviews[v].dh = sc.height; p->hviews[v].dpitch = sc.width;
    }
    p->has_depths = true;
    int rc = sync_frame_views(p);
    if (rc) return rc;
    CK(cudaStreamSynchronize(p->stream));  // the caller may free or reuse its buffers on return (as in mpmvs_set_state)
    return MPMVS_OK;
}

int mpmvs_set_src_depths_device(mpmvs_problem* p, const float* const* depth_dev, const size_t* pitch_bytes) {
    if (!p || !depth_dev || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    for (int v = 0; v < p->n - 1; ++v) {
        const mpmvs_camera& sc = p->cams[v + 1];
        p->hviews[v].depth = depth_dev[v];   // used in place: e.g. a slot of the all-gathered depth buffer
        p->hviews[v].dw = sc.width; p->hviews[v].dh = sc.height;
        p->hviews[v].dpitch = pitch_bytes ? (int)(pitch_bytes[v] / sizeof(float)) : sc.width;
    }
    p->has_depths = true;
    return sync_frame_views(p);
}

int mpmvs_set_state(mpmvs_problem* p, const float* planes4_host, const float* costs_host) {
    if (!p || !planes4_host || !costs_host || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(p->S.planes, planes4_host, p->wh * sizeof(pm_f4), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->S.costs, costs_host, p->wh * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));  // the caller may free its buffers on return
    return MPMVS_OK;
}

int mpmvs_set_prior(mpmvs_problem* p, const float* prior_planes4_host, const uint32_t* mask_host) {
    if (!p || !prior_planes4_host || !mask_host || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(p->d_prior, prior_planes4_host, p->wh * sizeof(pm_f4), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->d_mask, mask_host, p->wh * sizeof(uint32_t), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->has_prior = true;
    return MPMVS_OK;
}

// ---------------------------------------------------------------------------------------------- run
int mpmvs_run_async(mpmvs_problem* p, uint64_t seed) { return enqueue_run(p, seed); }

int mpmvs_synchronize(mpmvs_problem* p) {
    if (!p) return MPMVS_E_ARG;
    CK(cudaStreamSynchronize(p->stream));
    return MPMVS_OK;
}

int mpmvs_run_into(mpmvs_problem* p, uint64_t seed, float* planes4_host, float* costs_host, float* geom_costs_host) {
    int rc = enqueue_run(p, seed);
    if (rc) return rc;
    // PatchMatch.cu:1246-1251
    if (planes4_host) CK(cudaMemcpyAsync(planes4_host, p->S.planes, p->wh * sizeof(pm_f4), cudaMemcpyDeviceToHost, p->stream));
    if (costs_host) CK(cudaMemcpyAsync(costs_host, p->S.costs, p->wh * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (geom_costs_host) CK(cudaMemcpyAsync(geom_costs_host, p->S.geom, p->wh * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return MPMVS_OK;
}

int mpmvs_get_results_async(mpmvs_problem* p, float* planes4_host, float* costs_host, float* geom_costs_host) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    if (planes4_host) CK(cudaMemcpyAsync(planes4_host, p->S.planes, p->wh * sizeof(pm_f4), cudaMemcpyDeviceToHost, p->stream));
    if (costs_host) CK(cudaMemcpyAsync(costs_host, p->S.costs, p->wh * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (geom_costs_host) CK(cudaMemcpyAsync(geom_costs_host, p->S.geom, p->wh * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    return MPMVS_OK;
}

int mpmvs_run(mpmvs_problem* p, uint64_t seed) {
    if (!p) return MPMVS_E_ARG;
    int rc = ensure_mirrors(p);
    if (rc) return rc;
    return mpmvs_run_into(p, seed, p->h_planes, p->h_costs, p->geomPlanarPrior ? p->h_geom : nullptr);
}

int mpmvs_last_run_ms(mpmvs_problem* p, float* ms) {
    if (!p || !ms || !p->ran) return MPMVS_E_ARG;
    CK(cudaEventSynchronize(p->ev1));
    CK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return MPMVS_OK;
}

int mpmvs_set_profiling(mpmvs_problem* p, int flags) {
    if (!p) return MPMVS_E_ARG;
    p->profiling = flags;
    return MPMVS_OK;
}

int mpmvs_last_run_profile(mpmvs_problem* p, float* init_ms, float* sweep_ms, int* n_sweeps, float* finalize_ms,
                           uint64_t* ncc_evaluations) {
    if (!p || !p->ran) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaEventSynchronize(p->ev1));
    if (p->profiling & 1) {
        const int ne = p->pev_used;
        if (ne < 3) return MPMVS_E_STATE;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, p->pev[0], p->pev[1]));
        if (init_ms) *init_ms = ms;
        CK(cudaEventElapsedTime(&ms, p->pev[1], p->pev[ne - 2]));
        if (sweep_ms) *sweep_ms = ms;
        if (n_sweeps) *n_sweeps = ne - 3;
        CK(cudaEventElapsedTime(&ms, p->pev[ne - 2], p->pev[ne - 1]));
        if (finalize_ms) *finalize_ms = ms;
    } else if (init_ms || sweep_ms || finalize_ms) {
        return MPMVS_E_STATE;
    }
    if (ncc_evaluations) {
        if (!(p->profiling & 2) || !p->d_counters) return MPMVS_E_STATE;
        unsigned long long c[2];
        CK(cudaMemcpy(c, p->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        *ncc_evaluations = c[0];
    }
    return MPMVS_OK;
}

int mpmvs_last_run_launches(mpmvs_problem* p, int* launches) {
    if (!p || !launches) return MPMVS_E_ARG;
    *launches = p->last_launches;
    return MPMVS_OK;
}

// ---------------------------------------------------------------------------------------------- results
int mpmvs_get_size(mpmvs_problem* p, int* width, int* height) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    if (width) *width = p->W;
    if (height) *height = p->H;
    return MPMVS_OK;
}

int mpmvs_get_depth_range(mpmvs_problem* p, float* depth_min, float* depth_max) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    if (depth_min) *depth_min = p->depth_min;
    if (depth_max) *depth_max = p->depth_max;
    return MPMVS_OK;
}

int mpmvs_get_planes(mpmvs_problem* p, float* planes4_host) {
    if (!p || !planes4_host || !p->h_planes || p->h_alloc < p->wh) return MPMVS_E_ARG;
    memcpy(planes4_host, p->h_planes, p->wh * sizeof(pm_f4));
    return MPMVS_OK;
}
int mpmvs_get_costs(mpmvs_problem* p, float* costs_host) {
    if (!p || !costs_host || !p->h_costs || p->h_alloc < p->wh) return MPMVS_E_ARG;
    memcpy(costs_host, p->h_costs, p->wh * sizeof(float));
    return MPMVS_OK;
}
int mpmvs_get_geom_costs(mpmvs_problem* p, float* geom_costs_host) {
    if (!p || !geom_costs_host || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(geom_costs_host, p->S.geom, p->wh * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return MPMVS_OK;
}

int mpmvs_device_planes(mpmvs_problem* p, const float** planes4_dev) {
    if (!p || !planes4_dev || p->n < 2) return MPMVS_E_ARG;
    *planes4_dev = (const float*)p->S.planes;
    return MPMVS_OK;
}
int mpmvs_device_costs(mpmvs_problem* p, const float** costs_dev) {
    if (!p || !costs_dev || p->n < 2) return MPMVS_E_ARG;
    *costs_dev = p->S.costs;
    return MPMVS_OK;
}
int mpmvs_export_depth_device(mpmvs_problem* p, float* depth_dev, size_t pitch_bytes) {
    if (!p || !depth_dev || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    const int pitch_f = pitch_bytes ? (int)(pitch_bytes / sizeof(float)) : p->W;
    CK(pm_launch_export_depth(p->S.planes, depth_dev, p->W, p->H, pitch_f, p->stream));
    return MPMVS_OK;
}

// ---------------------------------------------------------------------------------------------- stage hooks
int mpmvs_init_only(mpmvs_problem* p, uint64_t seed) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(pm_launchers(p->arith).init(make_frame(p), p->S, p->dviews, seed, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return MPMVS_OK;
}

int mpmvs_half_sweep(mpmvs_problem* p, int red, int iter, int scale) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    if (p->geom && !p->has_depths) return MPMVS_E_STATE;
    CK(cudaSetDevice(p->device));
    CK(pm_launchers(p->arith).sweep(make_frame(p), p->S, p->dviews, red ? 1 : 0, iter, scale, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return MPMVS_OK;
}

int mpmvs_finalize(mpmvs_problem* p) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(pm_launchers(p->arith).finalize(make_frame(p), p->S, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return MPMVS_OK;
}

int mpmvs_get_device_state(mpmvs_problem* p, float* planes4, float* costs, uint32_t* views, uint32_t* rng6, float* geom) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    if (planes4) CK(cudaMemcpy(planes4, p->S.planes, p->wh * sizeof(pm_f4), cudaMemcpyDeviceToHost));
    if (costs) CK(cudaMemcpy(costs, p->S.costs, p->wh * sizeof(float), cudaMemcpyDeviceToHost));
    if (views) CK(cudaMemcpy(views, p->S.views, p->wh * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (rng6) CK(cudaMemcpy(rng6, p->S.rng, p->wh * 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (geom) CK(cudaMemcpy(geom, p->S.geom, p->wh * sizeof(float), cudaMemcpyDeviceToHost));
    return MPMVS_OK;
}

int mpmvs_set_device_state(mpmvs_problem* p, const float* planes4, const float* costs, const uint32_t* views,
                           const uint32_t* rng6, const float* geom) {
    if (!p || p->n < 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    if (planes4) CK(cudaMemcpy(p->S.planes, planes4, p->wh * sizeof(pm_f4), cudaMemcpyHostToDevice));
    if (costs) CK(cudaMemcpy(p->S.costs, costs, p->wh * sizeof(float), cudaMemcpyHostToDevice));
    if (views) CK(cudaMemcpy(p->S.views, views, p->wh * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (rng6) CK(cudaMemcpy(p->S.rng, rng6, p->wh * 6 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (geom) CK(cudaMemcpy(p->S.geom, geom, p->wh * sizeof(float), cudaMemcpyHostToDevice));
    return MPMVS_OK;
}

int mpmvs_ncc_map(mpmvs_problem* p, const float* planes4_host, int scale, float* out_host) {
    if (!p || !planes4_host || !out_host || p->n < 2 || scale < 0 || scale > 2) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    pm_f4* dp = nullptr;
    float* dout = nullptr;
    const size_t out_bytes = p->wh * (p->n - 1) * sizeof(float);
    CK(cudaMalloc((void**)&dp, p->wh * sizeof(pm_f4)));
    cudaError_t e = cudaMalloc((void**)&dout, out_bytes);
    if (e != cudaSuccess) { cudaFree(dp); return (int)e; }
    cudaMemcpyAsync(dp, planes4_host, p->wh * sizeof(pm_f4), cudaMemcpyHostToDevice, p->stream);
    e = pm_launchers(p->arith).ncc_map(make_frame(p), p->dviews, dp, scale, dout, p->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, dout, out_bytes, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(dp);
    cudaFree(dout);
    return (int)e;
}

int mpmvs_ncc_bench(mpmvs_problem* p, const float* planes4_host, int scale, int taps_per_side, int n_views, int reps, float* ms,
                    uint64_t* ncc_evaluations) {
    if (!p || !planes4_host || p->n < 2 || n_views < 1 || n_views > p->n - 1 || reps < 1 || !ms) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    pm_f4* dp = nullptr;
    float* dout = nullptr;
    unsigned long long* dc = nullptr;
    CK(cudaMalloc((void**)&dp, p->wh * sizeof(pm_f4)));
    cudaError_t e = cudaMalloc((void**)&dout, p->wh * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dc, sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(dp); cudaFree(dout); return (int)e; }
    cudaMemcpyAsync(dp, planes4_host, p->wh * sizeof(pm_f4), cudaMemcpyHostToDevice, p->stream);
    cudaMemsetAsync(dc, 0, sizeof(unsigned long long), p->stream);
    const PmFrame F = make_frame(p);
    // one counting pass (also warm-up), then the timed pass without the counter
    e = pm_launchers(p->arith).ncc_bench(F, p->dviews, dp, scale, taps_per_side, n_views, reps, dout, dc, p->stream);
    if (e == cudaSuccess) e = cudaEventRecord(p->ev0, p->stream);
    if (e == cudaSuccess) e = pm_launchers(p->arith).ncc_bench(F, p->dviews, dp, scale, taps_per_side, n_views, reps, dout, nullptr, p->stream);
    if (e == cudaSuccess) e = cudaEventRecord(p->ev1, p->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(p->ev1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, p->ev0, p->ev1);
    unsigned long long cnt = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&cnt, dc, sizeof(cnt), cudaMemcpyDeviceToHost);
    if (ncc_evaluations) *ncc_evaluations = cnt;
    cudaFree(dp); cudaFree(dout); cudaFree(dc);
    return (int)e;
}

int mpmvs_geom_map(mpmvs_problem* p, const float* planes4_host, float* out_host) {
    if (!p || !planes4_host || !out_host || p->n < 2) return MPMVS_E_ARG;
    if (!p->has_depths) return MPMVS_E_STATE;
    CK(cudaSetDevice(p->device));
    pm_f4* dp = nullptr;
    float* dout = nullptr;
    const size_t out_bytes = p->wh * (p->n - 1) * sizeof(float);
    CK(cudaMalloc((void**)&dp, p->wh * sizeof(pm_f4)));
    cudaError_t e = cudaMalloc((void**)&dout, out_bytes);
    if (e != cudaSuccess) { cudaFree(dp); return (int)e; }
    cudaMemcpyAsync(dp, planes4_host, p->wh * sizeof(pm_f4), cudaMemcpyHostToDevice, p->stream);
    e = pm_launchers(p->arith).geom_map(make_frame(p), p->dviews, dp, dout, p->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, dout, out_bytes, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(dp);
    cudaFree(dout);
    return (int)e;
}

// ---------------------------------------------------------------------------------------------- planar-prior host stage
int mpmvs_delaunay(const int* xy, int n, int width, int height, int* tris_out, int max_tris, int* n_tris) {
    if (!xy || n < 0 || width <= 0 || height <= 0 || !n_tris) return MPMVS_E_ARG;
    for (int i = 0; i < n; ++i)
        if (xy[2 * i] < 0 || xy[2 * i] >= width || xy[2 * i + 1] < 0 || xy[2 * i + 1] >= height) return MPMVS_E_ARG;
    std::vector<int> t;
    if (!pmsd::triangulate(xy, n, width, height, t)) return MPMVS_E_ARG;
    *n_tris = (int)(t.size() / 3);
    if (tris_out) {
        if ((int)(t.size() / 3) > max_tris) return MPMVS_E_ARG;
        memcpy(tris_out, t.data(), t.size() * sizeof(int));
    }
    return MPMVS_OK;
}

int mpmvs_pick_vertices(mpmvs_problem* p, int geom_variant, int* xy_out, int max_vertices, int* n_out) {
    if (!p || p->n < 2 || !n_out) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    const int cx = (p->W + 4) / 5, cy = (p->H + 4) / 5;
    const size_t cells = (size_t)cx * cy;
    if (cells > p->cells_cap) {
        cudaFree(p->d_cell_xy); cudaFree(p->d_cell_n);
        p->d_cell_xy = nullptr; p->d_cell_n = nullptr; p->cells_cap = 0;
        CK(cudaMalloc((void**)&p->d_cell_xy, cells * 3 * sizeof(short2)));
        CK(cudaMalloc((void**)&p->d_cell_n, cells));
        p->cells_cap = cells;
    }
    p->h_cell_xy.resize(cells * 6);
    p->h_cell_n.resize(cells);
    CK(pm_launch_pick_vertices(p->S.costs, p->S.geom, p->W, p->H, geom_variant ? 1 : 0, p->d_cell_xy, p->d_cell_n, p->stream));
    CK(cudaMemcpyAsync(p->h_cell_xy.data(), p->d_cell_xy, cells * 3 * sizeof(short2), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(p->h_cell_n.data(), p->d_cell_n, cells, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    int n = 0;
    for (size_t c = 0; c < cells; ++c)      // row-major cells, the order GetTriangulateVertices pushes them (PatchMatch.cpp:788-851)
        for (int k = 0; k < p->h_cell_n[c]; ++k) {
            if (xy_out) {
                if (n >= max_vertices) return MPMVS_E_ARG;
                xy_out[2 * n] = p->h_cell_xy[(c * 3 + k) * 2];
                xy_out[2 * n + 1] = p->h_cell_xy[(c * 3 + k) * 2 + 1];
            }
            ++n;
        }
    *n_out = n;
    return MPMVS_OK;
}

// upload + rasterisation + plane fit; indices are trusted (mpmvs_build_prior passes its own triangulation)
static int prior_from_trusted_triangles(mpmvs_problem* p, const int* xy, int n_vertices, const int* tris, int n_tris,
                                        int* n_prior_pixels) {
    CK(cudaSetDevice(p->device));
    // Grow with headroom: the counts differ a little from image to image and from pass to pass, and every cudaFree here
    // would wait for the whole device -- i.e. for the runs of the other images in flight -- before this image can go on.
    if ((size_t)n_vertices > p->vtx_cap) {
        const size_t cap = (size_t)n_vertices + (size_t)n_vertices / 2 + 1024;
        cudaFree(p->d_vxy); p->d_vxy = nullptr; p->vtx_cap = 0;
        CK(cudaMalloc((void**)&p->d_vxy, sizeof(int2) * cap));
        p->vtx_cap = cap;
    }
    if ((size_t)n_tris > p->tri_cap) {
        const size_t cap = (size_t)n_tris + (size_t)n_tris / 2 + 2048;
        cudaFree(p->d_tris); cudaFree(p->d_tri_planes); p->d_tris = nullptr; p->d_tri_planes = nullptr; p->tri_cap = 0;
        CK(cudaMalloc((void**)&p->d_tris, sizeof(int3) * cap));
        CK(cudaMalloc((void**)&p->d_tri_planes, sizeof(pm_f4) * cap));
        p->tri_cap = cap;
    }
    if (!p->d_prior_count) CK(cudaMalloc((void**)&p->d_prior_count, sizeof(unsigned int)));
    if (n_vertices) CK(cudaMemcpyAsync(p->d_vxy, xy, sizeof(int2) * (size_t)n_vertices, cudaMemcpyHostToDevice, p->stream));
    if (n_tris) CK(cudaMemcpyAsync(p->d_tris, tris, sizeof(int3) * (size_t)n_tris, cudaMemcpyHostToDevice, p->stream));
    CK(pm_launch_prior(make_frame(p), p->S.planes, p->d_vxy, p->d_tris, n_tris, p->d_tri_planes, p->d_mask, p->d_prior, p->d_prior_count,
                       p->stream));
    p->has_prior = true;
    // the count is only fetched (and the stream only drained) when the caller asks for it: without this the host would
    // sit here while it could already be enqueuing the next run or triangulating the next image
    if (n_prior_pixels) return mpmvs_get_prior_pixels(p, n_prior_pixels);
    return MPMVS_OK;
}

int mpmvs_prior_from_triangles(mpmvs_problem* p, const int* xy, int n_vertices, const int* tris, int n_tris, int* n_prior_pixels) {
    if (!p || p->n < 2 || n_vertices < 0 || n_tris < 0 || (n_tris > 0 && (!xy || !tris))) return MPMVS_E_ARG;
    // a caller's triangulation is read by the device as is: reject indices and positions that would leave the buffers
    for (size_t i = 0; i < (size_t)3 * n_tris; ++i)
        if (tris[i] < 0 || tris[i] >= n_vertices) return MPMVS_E_ARG;
    for (int i = 0; i < n_vertices; ++i)
        if (xy[2 * i] < 0 || xy[2 * i] >= p->W || xy[2 * i + 1] < 0 || xy[2 * i + 1] >= p->H) return MPMVS_E_ARG;
    return prior_from_trusted_triangles(p, xy, n_vertices, tris, n_tris, n_prior_pixels);
}

int mpmvs_get_prior_pixels(mpmvs_problem* p, int* n_prior_pixels) {
    if (!p || !n_prior_pixels || !p->has_prior || !p->d_prior_count) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    unsigned int count = 0;
    CK(cudaMemcpyAsync(&count, p->d_prior_count, sizeof(count), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    *n_prior_pixels = (int)count;
    return MPMVS_OK;
}

int mpmvs_build_prior(mpmvs_problem* p, mpmvs_prior_stats* stats) {
    if (!p || p->n < 2 || !p->ran) return MPMVS_E_ARG;
    using clk = std::chrono::steady_clock;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<float, std::milli>(b - a).count(); };
    const auto t0 = clk::now();
    const int max_v = 3 * ((p->W + 4) / 5) * ((p->H + 4) / 5);
    std::vector<int> xy((size_t)2 * max_v);
    int nv = 0;
    int rc = mpmvs_pick_vertices(p, p->geomPlanarPrior ? 1 : 0, xy.data(), max_v, &nv);
    if (rc) return rc;
    const auto t1 = clk::now();
    std::vector<int> tris;
    if (nv >= 3 && !pmsd::triangulate(xy.data(), nv, p->W, p->H, tris)) return MPMVS_E_ARG;   // triangle ORDER matters: pm_subdiv.h
    const auto t2 = clk::now();
    // vertices and triangles are copied with cudaMemcpyAsync from pageable vectors: the runtime stages them before returning
    rc = prior_from_trusted_triangles(p, xy.data(), nv, tris.data(), (int)(tris.size() / 3), nullptr);
    if (rc) return rc;
    const auto t3 = clk::now();
    if (stats) {
        stats->n_vertices = nv; stats->n_triangles = (int)(tris.size() / 3); stats->n_prior_pixels = -1;  /* mpmvs_get_prior_pixels */
        stats->pick_ms = ms(t0, t1); stats->delaunay_ms = ms(t1, t2); stats->raster_ms = ms(t2, t3); stats->total_ms = ms(t0, t3);
    }
    return MPMVS_OK;
}

int mpmvs_get_prior(mpmvs_problem* p, float* prior_planes4_host, uint32_t* mask_host) {
    if (!p || p->n < 2 || !p->has_prior) return MPMVS_E_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    if (prior_planes4_host) CK(cudaMemcpy(prior_planes4_host, p->d_prior, p->wh * sizeof(pm_f4), cudaMemcpyDeviceToHost));
    if (mask_host) CK(cudaMemcpy(mask_host, p->d_mask, p->wh * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return MPMVS_OK;
}

int mpmvs_selftest_ex2_monotone(int device, uint64_t* violations) {
    if (!violations) return MPMVS_E_ARG;
    int rc = ensure_device(device);
    if (rc) return rc;
    unsigned long long* d = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(*d)));
    cudaError_t e = cudaMemset(d, 0, sizeof(*d));
    if (e == cudaSuccess) e = pm_launch_ex2_monotone(-160.0f, d, 0);
    unsigned long long v = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&v, d, sizeof(v), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return (int)e;
    *violations = v;
    return MPMVS_OK;
}

int mpmvs_uniform_stream(uint64_t seed, int x, int y, int n, float* out_host) {
    if (!out_host || n <= 0) return MPMVS_E_ARG;
    int rc = ensure_device(0);
    if (rc) return rc;
    float* d = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(float) * n));
    cudaError_t e = pm_launch_uniform_stream(seed, x, y, n, d, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d, sizeof(float) * n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return (int)e;
}

}  // extern "C"
