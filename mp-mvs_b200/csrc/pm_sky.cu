// pm_sky.cu -- joint-bilateral upsampling of the sky mask (the only joint-bilateral code of the
// reference): GenerateSkyRegionMask -> bilateral_filter -> Pixel_bilateral_filter,
// /root/reference/src/PatchMatch.cpp:4-57 and /root/reference/SkySegment/src/SkyRegionDetect.cu:3-66 ("sky.cu:NNN").
//
// A low-resolution sky probability map (the segmentation network's output; the network itself needs ncnn and is out of
// scope) is brought to image size with cv::resize(INTER_LINEAR) (sky.cu:41-42) and then filtered over a 37 x 37 window
// with weights exp(-|dp| / 72 - |dBGR| / 8) guided by the full-resolution colour image; a pixel is sky when the weighted
// mean exceeds 0.6 (sky.cu:9-34). The result gates depth fusion (PatchMatch.cpp:385-388, mpmvs_fusion_set_sky_mask).
//
// Kernels:
//   pm_sky_upsample_kernel   cv::resize INTER_LINEAR on CV_32FC1 (pixel-centre mapping, source index clamped)
//   pm_sky_filter_kernel     two vertically adjacent pixels per thread; the (32 x 32 tile + 18 px halo) of {b, g, r, mask} is
//                            staged once in shared memory as float4, so two taps share one 16-byte shared load, and a tap
//                            costs 2 MUFU (sqrt, ex2) and ~12 FP32 instead of the reference's four uncoalesced global loads
//                            and three int->float conversions. Taps are accumulated in the reference's order (x offset
//                            outer, y offset inner): the sums are the same floats the reference kernel adds in sequence.
// Bound: 1369 taps/pixel x 2 MUFU on the 16-lane XU pipe (4 clk per warp-tap); with one pixel per thread the shared-memory
// pipe was the limiter (95 % of peak, profiles/r01_ncu_sky_filter.txt). HBM traffic is 11 B/pixel and irrelevant.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mpmvs_b200.h"

#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace {

constexpr int SKY_R = 18;                       // half_windows, sky.cu:13
constexpr int SKY_BW = 32, SKY_BH = 16;         // threads per block
constexpr int SKY_PH = 2 * SKY_BH;              // pixel rows per block: two per thread
constexpr int SKY_TW = SKY_BW + 2 * SKY_R;      // 68
constexpr int SKY_TH = SKY_PH + 2 * SKY_R;      // 68
constexpr size_t SKY_SMEM = (size_t)SKY_TW * SKY_TH * sizeof(float4) + (2 * SKY_R + 1) * (2 * SKY_R + 2) * sizeof(float);

// cv::resize(src, dst, Size(W, H), ..., INTER_LINEAR) for one float channel
__global__ void __launch_bounds__(256) pm_sky_upsample_kernel(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst,
                                                              int W, int H, double scale_x, double scale_y) {
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= W || dy >= H) return;
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int ix = (int)floorf(fx);
    fx -= (float)ix;
    if (ix < 0) { ix = 0; fx = 0.f; }
    if (ix >= sw - 1) { ix = sw - 1; fx = 0.f; }
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int iy = (int)floorf(fy);
    fy -= (float)iy;
    if (iy < 0) { iy = 0; fy = 0.f; }
    if (iy >= sh - 1) { iy = sh - 1; fy = 0.f; }
    const int ix1 = min(ix + 1, sw - 1), iy1 = min(iy + 1, sh - 1);
    const float* r0 = src + (size_t)iy * sw;
    const float* r1 = src + (size_t)iy1 * sw;
    // horizontal pass then vertical pass, products rounded separately (no contraction), as the two-pass library code does
    const float top = __fadd_rn(__fmul_rn(__ldg(r0 + ix), 1.f - fx), __fmul_rn(__ldg(r0 + ix1), fx));
    const float bot = __fadd_rn(__fmul_rn(__ldg(r1 + ix), 1.f - fx), __fmul_rn(__ldg(r1 + ix1), fx));
    dst[(size_t)dy * W + dx] = __fadd_rn(__fmul_rn(top, 1.f - fy), __fmul_rn(bot, fy));
}

// Two vertically adjacent pixels per thread: tap (i, j) of pixel (x, y) and tap (i, j - 1) of pixel (x, y + 1) are the same
// tile element, so one 16-byte shared load and one spatial-weight load serve two taps. Each pixel still adds its 1369 taps
// in the reference's order.
__global__ void __launch_bounds__(SKY_BW* SKY_BH) pm_sky_filter_kernel(const unsigned char* __restrict__ bgr, const float* __restrict__ mask,
                                                                       float* __restrict__ result, float* __restrict__ prob_out, int W,
                                                                       int H) {
    extern __shared__ float4 sky_smem[];
    float4 (*tile)[SKY_TW] = reinterpret_cast<float4 (*)[SKY_TW]>(sky_smem);                        // {b, g, r, mask}; 74 KB
    float (*spatial)[2 * SKY_R + 2] = reinterpret_cast<float (*)[2 * SKY_R + 2]>(sky_smem + SKY_TW * SKY_TH);   // 5.6 KB
    const int x0 = blockIdx.x * SKY_BW, y0 = blockIdx.y * SKY_PH;
    const int tid = threadIdx.y * SKY_BW + threadIdx.x;
    for (int i = tid; i < SKY_TW * SKY_TH; i += SKY_BW * SKY_BH) {
        const int ty = i / SKY_TW, tx = i - ty * SKY_TW;
        const int gx = x0 - SKY_R + tx, gy = y0 - SKY_R + ty;
        float4 t;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t g = (size_t)gy * W + gx;
            t = make_float4((float)bgr[3 * g], (float)bgr[3 * g + 1], (float)bgr[3 * g + 2], __ldg(mask + g));
        } else {
            // taps outside the image are skipped by the reference (sky.cu:20-21): a colour this far away gives the tap an
            // exp(-inf) = 0 weight, which adds exactly nothing to either sum
            t = make_float4(3e18f, 3e18f, 3e18f, 0.f);
        }
        tile[ty][tx] = t;
    }
    const float sigma_spatial = 2.0 * 6.0 * 6.0, sigma_color = 2.0 * 2.0 * 2.0;     // sky.cu:9-10
    for (int i = tid; i < (2 * SKY_R + 1) * (2 * SKY_R + 2); i += SKY_BW * SKY_BH) {
        const int a = i / (2 * SKY_R + 2) - SKY_R, b = i % (2 * SKY_R + 2) - SKY_R;    // column b = SKY_R + 1 is padding
        spatial[a + SKY_R][b + SKY_R] = sqrt((float)(a * a + b * b)) / sigma_spatial;   // sky.cu:26-27, hypothesis-invariant
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, ly = 2 * threadIdx.y, y = y0 + ly;
    if (x >= W || y >= H) return;
    const float4 c0 = tile[ly + SKY_R][threadIdx.x + SKY_R], c1 = tile[ly + 1 + SKY_R][threadIdx.x + SKY_R];
    float wsum0 = 0.0f, prob0 = 0.0f, wsum1 = 0.0f, prob1 = 0.0f;
#pragma unroll 1
    for (int i = -SKY_R; i <= SKY_R; ++i) {             // x offset (outer, as sky.cu:16)
        float sp_prev = 0.f;
#pragma unroll
        for (int r = 0; r <= 2 * SKY_R + 1; ++r) {      // tile rows ly + r: y offset r - R for the upper pixel, r - 1 - R for the lower
            const float4 t = tile[ly + r][threadIdx.x + SKY_R + i];
            const float sp = spatial[i + SKY_R][r];
            if (r <= 2 * SKY_R) {
                const float b = t.x - c0.x, g = t.y - c0.y, q = t.z - c0.z;
                const float dis_color = sqrt(b * b + g * g + q * q);
                const float w = exp(-sp - dis_color / sigma_color);
                wsum0 += w;
                prob0 += w * t.w;
            }
            if (r >= 1) {
                const float b = t.x - c1.x, g = t.y - c1.y, q = t.z - c1.z;
                const float dis_color = sqrt(b * b + g * g + q * q);
                const float w = exp(-sp_prev - dis_color / sigma_color);
                wsum1 += w;
                prob1 += w * t.w;
            }
            sp_prev = sp;
        }
    }
    prob0 = prob0 / wsum0;
    const size_t idx = (size_t)y * W + x;
    result[idx] = prob0 > 0.6 ? 255 : 0;                // sky.cu:34 (compared in double, as written there)
    if (prob_out) prob_out[idx] = prob0;
    if (y + 1 < H) {
        prob1 = prob1 / wsum1;
        result[idx + W] = prob1 > 0.6 ? 255 : 0;
        if (prob_out) prob_out[idx + W] = prob1;
    }
}

}  // namespace

extern "C" {

// Host buffers in, host buffers out; blocks until the result is in `result`. bgr [h][w][3] uint8 (cv::imread order),
// mask [mh][mw] float probabilities (any size: brought to w x h first), result [h][w] 0 / 255, prob (optional) the weighted
// mean before the threshold, ms (optional) device time of the two kernels.
int mpmvs_sky_mask_refine(int device, void* stream, const uint8_t* bgr, int width, int height, const float* mask, int mask_width,
                          int mask_height, float* result, float* prob, float* ms) {
    if (!bgr || !mask || !result || width <= 0 || height <= 0 || mask_width <= 0 || mask_height <= 0) return MPMVS_E_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return MPMVS_E_NO_DEVICE;
    if (device < 0 || device >= count) return MPMVS_E_ARG;
    SCK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t wh = (size_t)width * height, mwh = (size_t)mask_width * mask_height;
    struct Temps {
        unsigned char* bgr = nullptr;
        float *mask_lo = nullptr, *mask = nullptr, *res = nullptr, *prob = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Temps() {
            cudaFree(bgr); cudaFree(mask_lo); cudaFree(mask); cudaFree(res); cudaFree(prob);
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
        }
    } T;
    SCK(cudaMalloc((void**)&T.bgr, wh * 3));
    SCK(cudaMalloc((void**)&T.mask, wh * 4));
    SCK(cudaMalloc((void**)&T.res, wh * 4));
    if (prob) SCK(cudaMalloc((void**)&T.prob, wh * 4));
    SCK(cudaEventCreate(&T.e0));
    SCK(cudaEventCreate(&T.e1));
    SCK(cudaMemcpyAsync(T.bgr, bgr, wh * 3, cudaMemcpyHostToDevice, st));
    const bool same = mask_width == width && mask_height == height;      // cv::resize to the same size is a copy
    if (same) {
        SCK(cudaMemcpyAsync(T.mask, mask, wh * 4, cudaMemcpyHostToDevice, st));
        SCK(cudaEventRecord(T.e0, st));
    } else {
        SCK(cudaMalloc((void**)&T.mask_lo, mwh * 4));
        SCK(cudaMemcpyAsync(T.mask_lo, mask, mwh * 4, cudaMemcpyHostToDevice, st));
        SCK(cudaEventRecord(T.e0, st));
        pm_sky_upsample_kernel<<<dim3((width + 31) / 32, (height + 7) / 8), dim3(32, 8), 0, st>>>(
            T.mask_lo, mask_width, mask_height, T.mask, width, height, (double)mask_width / width, (double)mask_height / height);
    }
    SCK(cudaFuncSetAttribute(pm_sky_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKY_SMEM));
    pm_sky_filter_kernel<<<dim3((width + SKY_BW - 1) / SKY_BW, (height + SKY_PH - 1) / SKY_PH), dim3(SKY_BW, SKY_BH), SKY_SMEM, st>>>(
        T.bgr, T.mask, T.res, T.prob, width, height);
    SCK(cudaGetLastError());
    SCK(cudaEventRecord(T.e1, st));
    SCK(cudaMemcpyAsync(result, T.res, wh * 4, cudaMemcpyDeviceToHost, st));
    if (prob) SCK(cudaMemcpyAsync(prob, T.prob, wh * 4, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    if (ms) SCK(cudaEventElapsedTime(ms, T.e0, T.e1));
    return MPMVS_OK;
}

}  // extern "C"
