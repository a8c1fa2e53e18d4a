// pm_kernels.cu -- sm_100a kernels of the PatchMatch path and their launchers.
//
// Kernels (reference counterparts in /root/reference/src/PatchMatch.cu):
//   pm_init_kernel<2>        InitializeScore                 cu:536-573
//   pm_sweep_kernel<S>       BlackPixelUpdate/RedPixelUpdate cu:724-1019   (S = window scale 0/1/2)
//   pm_depth_normal_kernel   GetDepthandNormal               cu:1021-1034
//   pm_filter_kernel         Black/RedPixelFilter            cu:1036-1174
//   pm_ncc_map_kernel<S>, pm_geom_map_kernel                 test hooks over pm_ncc / pm_geom_cost
//   pm_export_depth_kernel   depth channel -> dense map (feeds the inter-GPU depth all-gather)
//
// Block = 32 x PM_BH (4) threads, one thread per pixel. A sweep block owns the pixels of one checkerboard
// colour in a 32 x 8 pixel tile; the reference-image window of the tile (tile + 5*2^S halo) is staged once in
// shared memory through the texture unit (exact texel fetch, hardware clamp-to-edge at the borders),
// the per-view constants sit next to it. Source samples go through the texture unit: they are
// homography-warped scattered bilinear reads, which is what the unit is built for, and its filter is not
// reproducible in software bit for bit (tools/tex_model_check.cu, profiles/r02_tex_model_check.json).
//
// Dynamic shared memory of a block: [nsrc x PmView][reference tile][36 bilateral weights x threads][8 x nsrc candidate
// costs x threads (sweep only)]; the two per-thread tables are laid out [entry][thread] (conflict-free).
//
// This file is compiled TWICE into the library: as is (namespace pm_fast) and with -DPM_EXACT=1 (namespace pm_exact, the
// reference's arithmetic bit for bit; pm_core.cuh). The helper kernels that do no arithmetic of the path (format
// conversion, depth export, the RNG stream hook, the device-evaluated constant table) exist once, in the exact build.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <type_traits>

#include "pm_core.cuh"
#include "pm_kernels.h"

namespace PM_ARITH_NS {   // pm_exact or pm_fast: the kernels of the two arithmetics carry the name in their symbols

// Tuning knobs (compile-time; the defaults are what the B200 measurements in profiles/ selected).
#ifndef PM_BH
#define PM_BH 4            // thread rows per block: 32 x PM_BH threads (a 32 x 8 pixel tile of one checkerboard colour)
#endif
#ifndef PM_MIN_BLOCKS
#define PM_MIN_BLOCKS 3    // resident blocks per SM the register allocation must allow
#endif
#ifndef PM_CA_SMEM
#define PM_CA_SMEM 1       // 1: the 8 x nsrc candidate cost table lives in shared memory instead of local memory
#endif
constexpr int BW = 32, BH = PM_BH;
constexpr int MIN_BLOCKS = PM_MIN_BLOCKS;

template <int SCALE, int ROWS>
struct Tile {
    static constexpr int R = 5 << SCALE;      // halo = largest tap offset
    static constexpr int TW = BW + 2 * R;     // 42 / 52 / 72
    static constexpr int TH = ROWS + 2 * R;
    static constexpr int PITCH = TW;          // even pitch: a checkerboard warp alternates rows r, r+1 with x -> even lanes hit
                                              // even banks, odd lanes odd banks (conflict-free); an odd pitch would 2-way conflict
    static constexpr int FLOATS = PITCH * TH;
};

constexpr int WT_FLOATS = PM_WTAB ? 36 * BW * BH : 0;   // bilateral weights of the 6 x 6 window, per thread

// candidate cost table of one thread in shared memory (stride = threads per block; PmTableLocal is the local-memory one)
struct TableShared {
    float* p; int nsrc;   // p already offset by the thread index; element k at p[k * threads]
    __device__ __forceinline__ float& operator()(int r, int v) const { return p[(r * nsrc + v) * (BW * BH)]; }
};

template <int PITCH, bool SOFT_CLAMP, bool SCALED = false>
struct DevCtx {
    const float* centre;   // shared-memory address of this thread's pixel inside the tile
    float* wtab;           // this thread's column of the weight table (shared memory, stride = threads per block)
    const PmView* views;   // shared memory
    cudaTextureObject_t tex;
    float scale;           // texel value scale (255 for 8-bit UNORM storage); only read when SCALED
    __device__ __forceinline__ float ref(int dx, int dy) const { return centre[dy * PITCH + dx]; }
    __device__ __forceinline__ float& wt(int k) const { return wtab[k * (BW * BH)]; }
    __device__ __forceinline__ float src(int v, float xs, float ys) const {
        const PmView& V = views[v];
        if (SOFT_CLAMP) {  // views smaller than the layered array: hardware clamp would hit the padding
            xs = fminf(fmaxf(xs, 0.5f), V.w - 0.5f);
            ys = fminf(fmaxf(ys, 0.5f), V.h - 0.5f);
        }
        const float t = tex2DLayered<float>(tex, xs, ys, V.layer);
        return SCALED ? t * scale : t;
    }
    __device__ __forceinline__ float src_depth(int v, int xi, int yi) const {
        const PmView& V = views[v];
        xi = min(max(xi, 0), V.dw - 1);
        yi = min(max(yi, 0), V.dh - 1);
        return __ldg(V.depth + (size_t)yi * V.dpitch + xi);
    }
    __device__ __forceinline__ const PmView& view(int v) const { return views[v]; }
};

__host__ __device__ __forceinline__ size_t views_bytes(int nsrc) { return ((size_t)nsrc * sizeof(PmView) + 15) & ~(size_t)15; }

// cooperative staging: per-view constants, then the reference window of the tile whose top-left pixel is (x0, y0)
template <int SCALE, int ROWS>
__device__ __forceinline__ void stage_tile(float* tile, PmView* sviews, const PmView* gviews, const PmFrame& F,
                                           int x0, int y0) {
    using T = Tile<SCALE, ROWS>;
    const int tid = threadIdx.y * BW + threadIdx.x;
    {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(gviews);
        uint32_t* s = reinterpret_cast<uint32_t*>(sviews);
        const int words = F.nsrc * (int)(sizeof(PmView) / 4);
        for (int i = tid; i < words; i += BW * BH) s[i] = __ldg(g + i);
    }
    for (int i = tid; i < T::TW * T::TH; i += BW * BH) {
        const int ty = i / T::TW, tx = i - ty * T::TW;
        // exact texel fetch (weights 1/0 at texel centres); clamp-to-edge like the reference's texture (PatchMatch.cpp:1012-1018)
        const int gx = min(max(x0 - T::R + tx, 0), F.W - 1), gy = min(max(y0 - T::R + ty, 0), F.H - 1);
        const float t = tex2DLayered<float>((cudaTextureObject_t)F.tex, (float)gx + 0.5f, (float)gy + 0.5f, F.ref_layer);
        // 8-bit UNORM storage returns v/255: v is recovered exactly by rounding (reference-side values stay bit-identical)
        tile[ty * T::PITCH + tx] = F.tex_scale != 1.0f ? rintf(t * F.tex_scale) : t;
    }
    __syncthreads();
}

// Variant = (window scale, soft clamp for mixed image sizes, scaled texel values)
template <int SCALE, bool CL, bool SC>
__global__ void __launch_bounds__(BW* BH, MIN_BLOCKS) pm_sweep_kernel(const PmFrame F, const PmState S, const PmView* gviews,
                                                           int red, int iter) {
    using T = Tile<SCALE, 2 * BH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PmView* sviews = reinterpret_cast<PmView*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw + views_bytes(F.nsrc));
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * 2 * BH;
    stage_tile<SCALE, 2 * BH>(tile, sviews, gviews, F, x0, y0);
    const int tx = threadIdx.x, ly = 2 * threadIdx.y + ((tx & 1) ^ red);
    const int x = x0 + tx, y = y0 + ly;
    if (x >= F.W || y >= F.H) return;
    DevCtx<T::PITCH, CL, SC> c;
    c.centre = tile + (ly + T::R) * T::PITCH + (tx + T::R);
    c.wtab = tile + T::FLOATS + threadIdx.y * BW + tx;
    c.views = sviews;
    c.tex = (cudaTextureObject_t)F.tex;
    c.scale = F.tex_scale;
#if PM_CA_SMEM
    float* cab = c.wtab + WT_FLOATS;
    pm_sweep_pixel<SCALE>(c, F, S, x, y, iter, TableShared{cab, F.nsrc});
#else
    float ca[8 * PM_MAX_SRC];
    pm_sweep_pixel<SCALE>(c, F, S, x, y, iter, PmTableLocal{ca, F.nsrc});
#endif
}

template <int SCALE, bool CL, bool SC>
__global__ void __launch_bounds__(BW* BH, MIN_BLOCKS) pm_init_kernel(const PmFrame F, const PmState S, const PmView* gviews,
                                                          unsigned long long seed) {
    using T = Tile<SCALE, BH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PmView* sviews = reinterpret_cast<PmView*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw + views_bytes(F.nsrc));
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    stage_tile<SCALE, BH>(tile, sviews, gviews, F, x0, y0);
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= F.W || y >= F.H) return;
    DevCtx<T::PITCH, CL, SC> c;
    c.centre = tile + (threadIdx.y + T::R) * T::PITCH + (threadIdx.x + T::R);
    c.wtab = tile + T::FLOATS + threadIdx.y * BW + threadIdx.x;
    c.views = sviews;
    c.tex = (cudaTextureObject_t)F.tex;
    c.scale = F.tex_scale;
    pm_init_pixel<SCALE>(c, F, S, x, y, seed);
}

template <int SCALE, bool CL, bool SC>
__global__ void __launch_bounds__(BW* BH, MIN_BLOCKS) pm_ncc_map_kernel(const PmFrame F, const PmView* gviews,
                                                             const pm_f4* planes, float* out) {
    using T = Tile<SCALE, BH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PmView* sviews = reinterpret_cast<PmView*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw + views_bytes(F.nsrc));
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    stage_tile<SCALE, BH>(tile, sviews, gviews, F, x0, y0);
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= F.W || y >= F.H) return;
    DevCtx<T::PITCH, CL, SC> c;
    c.centre = tile + (threadIdx.y + T::R) * T::PITCH + (threadIdx.x + T::R);
    c.wtab = tile + T::FLOATS + threadIdx.y * BW + threadIdx.x;
    c.views = sviews;
    c.tex = (cudaTextureObject_t)F.tex;
    c.scale = F.tex_scale;
    const int idx = y * F.W + x;
    const PmRefStats st = pm_ref_stats<SCALE>(c, F);
    const PmHyp hyp = pm_hyp(F, planes[idx], x, y);
    uint32_t nexec = 0;
    for (int v = 0; v < F.nsrc; ++v) out[(size_t)v * F.W * F.H + idx] = pm_ncc<SCALE>(c, F, st, v, hyp, x, y, nexec);
}

// NCC microbenchmark (BASELINE.json config 5): TAPS x TAPS window, the first `nviews` source views, `reps` evaluations per
// (pixel, view) with the plane nudged per repetition so the fetches are not loop-invariant; out[idx] = sum of the costs.
template <int SCALE, int TAPS, bool SC>
__global__ void __launch_bounds__(BW* BH, MIN_BLOCKS) pm_ncc_bench_kernel(const PmFrame F, const PmView* gviews, const pm_f4* planes,
                                                               int nviews, int reps, float* out, unsigned long long* counter) {
    using T = Tile<SCALE, BH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PmView* sviews = reinterpret_cast<PmView*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw + views_bytes(F.nsrc));
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    stage_tile<SCALE, BH>(tile, sviews, gviews, F, x0, y0);
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= F.W || y >= F.H) return;
    DevCtx<T::PITCH, false, SC> c;
    c.centre = tile + (threadIdx.y + T::R) * T::PITCH + (threadIdx.x + T::R);
    c.wtab = tile + T::FLOATS + threadIdx.y * BW + threadIdx.x;
    c.views = sviews;
    c.tex = (cudaTextureObject_t)F.tex;
    c.scale = F.tex_scale;
    const int idx = y * F.W + x;
    const PmRefStats st = pm_ref_stats<SCALE, decltype(c), TAPS>(c, F);
    pm_f4 pl = planes[idx];
    float acc = 0.f;
    uint32_t nexec = 0;
    for (int r = 0; r < reps; ++r) {
        const PmHyp hyp = pm_hyp(F, pl, x, y);
        for (int v = 0; v < nviews; ++v) acc += pm_ncc<SCALE, decltype(c), TAPS>(c, F, st, v, hyp, x, y, nexec);
        pl.w *= 1.0005f;
    }
    out[idx] = acc;
    if (counter) atomicAdd(counter, (unsigned long long)nexec);
}

__global__ void __launch_bounds__(BW* BH) pm_geom_map_kernel(const PmFrame F, const PmView* gviews, const pm_f4* planes,
                                                              float* out) {
    __shared__ PmView sviews[PM_MAX_SRC];
    {
        const int tid = threadIdx.y * BW + threadIdx.x;
        const uint32_t* g = reinterpret_cast<const uint32_t*>(gviews);
        uint32_t* s = reinterpret_cast<uint32_t*>(sviews);
        for (int i = tid; i < F.nsrc * (int)(sizeof(PmView) / 4); i += BW * BH) s[i] = g[i];
        __syncthreads();
    }
    const int x = blockIdx.x * BW + threadIdx.x, y = blockIdx.y * BH + threadIdx.y;
    if (x >= F.W || y >= F.H) return;
    DevCtx<1, false, false> c;
    c.centre = nullptr;
    c.wtab = nullptr;
    c.views = sviews;
    c.tex = (cudaTextureObject_t)F.tex;
    const int idx = y * F.W + x;
    for (int v = 0; v < F.nsrc; ++v) out[(size_t)v * F.W * F.H + idx] = pm_geom_cost(c, F, v, planes[idx], x, y);
}

__global__ void __launch_bounds__(256) pm_depth_normal_kernel(const PmFrame F, pm_f4* planes) {
    const int x = blockIdx.x * BW + threadIdx.x, y = blockIdx.y * BH + threadIdx.y;
    if (x >= F.W || y >= F.H) return;
    const int idx = y * F.W + x;
    planes[idx] = pm_depth_normal(F, planes[idx], x, y);
}

__global__ void __launch_bounds__(256) pm_filter_kernel(const PmFrame F, pm_f4* planes, const float* costs, int red) {
    const int tx = threadIdx.x;
    const int x = blockIdx.x * BW + tx;
    const int y = 2 * (blockIdx.y * BH + threadIdx.y) + ((tx & 1) ^ red);
    if (x >= F.W || y >= F.H) return;
    planes[y * F.W + x].w = pm_median_depth(planes, costs, F.W, F.H, x, y);
}

#if PM_EXACT
__global__ void __launch_bounds__(256) pm_export_depth_kernel(const pm_f4* planes, float* out, int W, int H, int pitch_f) {
    const int x = blockIdx.x * BW + threadIdx.x, y = blockIdx.y * BH + threadIdx.y;
    if (x >= W || y >= H) return;
    out[(size_t)y * pitch_f + x] = planes[y * W + x].w;
}

// grey image (float or 8-bit, pitched) -> texels in the storage format of the image cache (dense rows)
template <class In, class Out>
__global__ void __launch_bounds__(256) pm_convert_kernel(const unsigned char* in, size_t in_pitch, Out* out, int W, int H) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const float v = (float)reinterpret_cast<const In*>(in + (size_t)y * in_pitch)[x];
    Out o;
    if constexpr (std::is_same<Out, unsigned char>::value) o = (unsigned char)__float2uint_rn(fminf(fmaxf(v, 0.f), 255.f));
    else if constexpr (std::is_same<Out, __half>::value) o = __float2half_rn(v);
    else o = v;
    out[(size_t)y * W + x] = o;
}

__global__ void pm_uniform_stream_kernel(unsigned long long seed, int x, int y, int n, float* out) {
    PmRng rs;
    pm_rng_init(rs, pm_mix_seed(seed, (uint32_t)x, (uint32_t)y));
    for (int i = 0; i < n; ++i) out[i] = pm_uniform(rs);
}
#endif  // PM_EXACT

template <int SCALE, int ROWS>
inline size_t smem_bytes(int nsrc) {
    return views_bytes(nsrc) + (Tile<SCALE, ROWS>::FLOATS + WT_FLOATS) * sizeof(float);
}
inline size_t sweep_smem(size_t base, int nsrc) { return base + (PM_CA_SMEM ? (size_t)8 * nsrc * BW * BH * sizeof(float) : 0); }

inline dim3 grid_full(int W, int H) { return dim3((W + BW - 1) / BW, (H + BH - 1) / BH, 1); }
// Same row coverage as the reference's 32x16 blocks over H/2 rows (cu:1196): 16*ceil((H/2)/16) thread rows.
inline dim3 grid_checker(int W, int H) { return dim3((W + BW - 1) / BW, (16 / BH) * (((H / 2) + 15) / 16), 1); }

// compile-time variant from the run-time (scale, soft_clamp, scaled) triple
template <class Fn>
inline cudaError_t dispatch(int scale, const PmFrame& F, Fn&& fn) {
    const bool cl = F.soft_clamp != 0, sc = F.tex_scale != 1.0f;
#define PM_CASE(S)                                                                  \
    case S:                                                                         \
        if (cl) { if (sc) fn(std::integral_constant<int, S>{}, std::true_type{}, std::true_type{});   \
                  else fn(std::integral_constant<int, S>{}, std::true_type{}, std::false_type{}); }   \
        else    { if (sc) fn(std::integral_constant<int, S>{}, std::false_type{}, std::true_type{});  \
                  else fn(std::integral_constant<int, S>{}, std::false_type{}, std::false_type{}); }  \
        break;
    switch (scale) {
        PM_CASE(0) PM_CASE(1) PM_CASE(2)
        default: return cudaErrorInvalidValue;
    }
#undef PM_CASE
    return cudaGetLastError();
}

template <class K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace PM_ARITH_NS

// ---------------------------------------------------------------------------------------------- launchers
namespace PM_ARITH_NS {
cudaError_t pm_launch_init(const PmFrame& F, const PmState& S, const PmView* gviews, unsigned long long seed, cudaStream_t st) {
    return dispatch(2, F, [&](auto s, auto cl, auto sc) {
        pm_init_kernel<decltype(s)::value, decltype(cl)::value, decltype(sc)::value>
            <<<grid_full(F.W, F.H), dim3(BW, BH), smem_bytes<decltype(s)::value, BH>(F.nsrc), st>>>(F, S, gviews, seed);
    });
}

cudaError_t pm_launch_sweep(const PmFrame& F, const PmState& S, const PmView* gviews, int red, int iter, int scale,
                            cudaStream_t st) {
    cudaError_t attr = cudaSuccess;
    cudaError_t rc = dispatch(scale, F, [&](auto s, auto cl, auto sc) {
        auto k = pm_sweep_kernel<decltype(s)::value, decltype(cl)::value, decltype(sc)::value>;
        const size_t bytes = sweep_smem(smem_bytes<decltype(s)::value, 2 * BH>(F.nsrc), F.nsrc);
        attr = allow_smem(k, bytes);
        if (attr == cudaSuccess) k<<<grid_checker(F.W, F.H), dim3(BW, BH), bytes, st>>>(F, S, gviews, red, iter);
    });
    return attr != cudaSuccess ? attr : rc;
}

cudaError_t pm_launch_finalize(const PmFrame& F, const PmState& S, cudaStream_t st) {
    pm_depth_normal_kernel<<<grid_full(F.W, F.H), dim3(BW, BH), 0, st>>>(F, S.planes);
    pm_filter_kernel<<<grid_checker(F.W, F.H), dim3(BW, BH), 0, st>>>(F, S.planes, S.costs, 0);
    pm_filter_kernel<<<grid_checker(F.W, F.H), dim3(BW, BH), 0, st>>>(F, S.planes, S.costs, 1);
    return cudaGetLastError();
}

cudaError_t pm_launch_ncc_map(const PmFrame& F, const PmView* gviews, const pm_f4* planes, int scale, float* out,
                              cudaStream_t st) {
    return dispatch(scale, F, [&](auto s, auto cl, auto sc) {
        pm_ncc_map_kernel<decltype(s)::value, decltype(cl)::value, decltype(sc)::value>
            <<<grid_full(F.W, F.H), dim3(BW, BH), smem_bytes<decltype(s)::value, BH>(F.nsrc), st>>>(F, gviews, planes, out);
    });
}

cudaError_t pm_launch_ncc_bench(const PmFrame& F, const PmView* gviews, const pm_f4* planes, int scale, int taps, int nviews, int reps,
                                float* out, unsigned long long* counter, cudaStream_t st) {
    if (F.soft_clamp || taps < 3 || taps > 6 || scale < 0 || scale > 2) return cudaErrorInvalidValue;
    const bool sc = F.tex_scale != 1.0f;
    const dim3 g = grid_full(F.W, F.H), b(BW, BH);
#define PM_NB(S, T)                                                                                                    \
    if (scale == S && taps == T) {                                                                                     \
        if (sc) pm_ncc_bench_kernel<S, T, true><<<g, b, smem_bytes<S, BH>(F.nsrc), st>>>(F, gviews, planes, nviews, reps, out, counter);  \
        else pm_ncc_bench_kernel<S, T, false><<<g, b, smem_bytes<S, BH>(F.nsrc), st>>>(F, gviews, planes, nviews, reps, out, counter);    \
    }
    PM_NB(0, 3) PM_NB(0, 4) PM_NB(0, 5) PM_NB(0, 6) PM_NB(1, 6) PM_NB(2, 6)
#undef PM_NB
    if (scale > 0 && taps != 6) return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t pm_launch_geom_map(const PmFrame& F, const PmView* gviews, const pm_f4* planes, float* out, cudaStream_t st) {
    pm_geom_map_kernel<<<grid_full(F.W, F.H), dim3(BW, BH), 0, st>>>(F, gviews, planes, out);
    return cudaGetLastError();
}

}  // namespace PM_ARITH_NS

#if PM_EXACT   // helper kernels without arithmetic of the path: once per library
cudaError_t pm_launch_convert(const void* in, size_t in_pitch, int in_is_u8, void* out, int out_fmt, int W, int H, cudaStream_t st) {
    const dim3 g((W + 31) / 32, (H + 7) / 8), b(32, 8);
    const unsigned char* i = (const unsigned char*)in;
    if (in_is_u8) {
        if (out_fmt == 0) pm_convert_kernel<unsigned char, float><<<g, b, 0, st>>>(i, in_pitch, (float*)out, W, H);
        else if (out_fmt == 1) pm_convert_kernel<unsigned char, __half><<<g, b, 0, st>>>(i, in_pitch, (__half*)out, W, H);
        else pm_convert_kernel<unsigned char, unsigned char><<<g, b, 0, st>>>(i, in_pitch, (unsigned char*)out, W, H);
    } else {
        if (out_fmt == 0) pm_convert_kernel<float, float><<<g, b, 0, st>>>(i, in_pitch, (float*)out, W, H);
        else if (out_fmt == 1) pm_convert_kernel<float, __half><<<g, b, 0, st>>>(i, in_pitch, (__half*)out, W, H);
        else pm_convert_kernel<float, unsigned char><<<g, b, 0, st>>>(i, in_pitch, (unsigned char*)out, W, H);
    }
    return cudaGetLastError();
}

cudaError_t pm_launch_export_depth(const pm_f4* planes, float* out, int W, int H, int pitch_floats, cudaStream_t st) {
    pm_export_depth_kernel<<<grid_full(W, H), dim3(BW, BH), 0, st>>>(planes, out, W, H, pitch_floats);
    return cudaGetLastError();
}

namespace {
__global__ void pm_literal_table_kernel(float one, float sigma_spatial, float sigma_color, float* out20) {
    pm_literal_table(one, sigma_spatial, sigma_color, out20);
}
}  // namespace
cudaError_t pm_launch_literal_table(float sigma_spatial, float sigma_color, float* out20_dev, cudaStream_t st) {
    pm_literal_table_kernel<<<1, 1, 0, st>>>(1.0f, sigma_spatial, sigma_color, out20_dev);
    return cudaGetLastError();
}

namespace {
// every float in [lo, hi]: ex2.approx.ftz(x) <= ex2.approx.ftz(next float above x)? (__expf = this instruction on x * log2 e)
__global__ void pm_ex2_monotone_kernel(unsigned int bits_lo, unsigned int count, unsigned long long* violations) {
    unsigned long long bad = 0;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        // negative floats: increasing bit pattern = decreasing value
        const float x = __uint_as_float(bits_lo + i + 1u), x_up = __uint_as_float(bits_lo + i);
        float a, b;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(x_up));
        bad += !(a <= b);
    }
    if (bad) atomicAdd(violations, bad);
}
}  // namespace
cudaError_t pm_launch_ex2_monotone(float lo_negative, unsigned long long* violations_dev, cudaStream_t st) {
    // all floats from -0.0f (0x80000000) down to lo_negative
    unsigned int last;
    memcpy(&last, &lo_negative, sizeof(last));
    const unsigned int first = 0x80000000u;
    pm_ex2_monotone_kernel<<<148 * 8, 256, 0, st>>>(first, last - first, violations_dev);
    return cudaGetLastError();
}

cudaError_t pm_launch_uniform_stream(unsigned long long seed, int x, int y, int n, float* out, cudaStream_t st) {
    pm_uniform_stream_kernel<<<1, 1, 0, st>>>(seed, x, y, n, out);
    return cudaGetLastError();
}
#endif  // PM_EXACT
