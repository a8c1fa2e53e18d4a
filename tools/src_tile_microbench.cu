// tools/src_tile_microbench.cu -- the north star's "per-block source tiles in shared memory" against the texture unit, measured.
//
// Question: would the coherent part of the NCC work -- the candidate phase, where the 128 pixels of a
// block warp almost the same window into a source view -- run faster if the block staged that window of the source view in
// shared memory (cp.async / TMA) and filtered it in software, instead of sending every tap through the texture unit
// (/root/reference/src/PatchMatch.cu:373-377)? This program measures exactly that inner loop, isolated: one block = the
// sweep kernel's 32 x 8 pixel checkerboard tile, every thread evaluates `reps` 36-tap NCC accumulations at scale S with the
// exact arithmetic's per-tap work (three numerators, one reciprocal, the bilateral weight, three sums), and the source sample
// of a tap comes from
//   mode 0  the texture unit                        (float32 layered texture, hardware bilinear: what the product does)
//   mode 1  a shared-memory float32 tile            (staged per NCC with cp.async.cg 16-byte copies, software bilinear)
//   mode 2  a shared-memory 8-bit tile              (staged with cp.async, 4x smaller, software bilinear + conversions)
//   mode 3  global memory through L1                (LSU loads of the linear float32 image, software bilinear, nothing staged)
//   mode 4  hybrid                                  (taps alternate between the texture unit and the shared float32 tile)
// Coherent planes (all lanes share one near-identity homography, +- `jitter` px per lane), so a tile of (tile + halo + margin)
// source pixels serves every tap -- the best case for staging; scattered hypotheses would add fallback fetches on top.
// The software filter uses 8 fractional weight bits; it is NOT the hardware's filter bit for bit (tools/tex_model_check.cu:
// the float32 texture path rounds its four weights, not the two fractions), so a shared-memory path could serve the fast
// arithmetic only -- the exact arithmetic needs the unit's own bits.
//
// Build: nvcc -O3 --use_fast_math -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/src_tile_microbench tools/src_tile_microbench.cu
// Run:   tools/src_tile_microbench [reps]      prints one JSON line (taps/clk/SM per mode and scale)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int BW = 32, BH = 4, MARGIN = 4;

template <int SCALE>
struct Geo {
    static constexpr int R = 5 << SCALE;
    static constexpr int TW = BW + 2 * R + 2 * MARGIN;          // source tile = pixel tile + halo + margin for the warp and jitter
    static constexpr int TH = 2 * BH + 2 * R + 2 * MARGIN;
    static constexpr int PITCH = (TW + 7) & ~3;                  // rows start 16-byte aligned; the origin is rounded down to 4 pixels
    static constexpr int P8 = (TW + 30) & ~15;                   // 8-bit tile: origin rounded down to 16 pixels
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// software bilinear, 8 fractional weight bits, on a tile whose texel (0,0) is source pixel (ox, oy)
template <class T>
__device__ __forceinline__ float soft_bilinear(const T* tile, int pitch, float xs, float ys, int ox, int oy) {
    const float xb = xs - 0.5f, yb = ys - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float a = rintf((xb - fx) * 256.0f) * (1.0f / 256.0f), b = rintf((yb - fy) * 256.0f) * (1.0f / 256.0f);
    const T* p = tile + ((int)fy - oy) * pitch + ((int)fx - ox);
    const float t00 = (float)p[0], t10 = (float)p[1], t01 = (float)p[pitch], t11 = (float)p[pitch + 1];
    const float l0 = fmaf(a, t10 - t00, t00), l1 = fmaf(a, t11 - t01, t01);
    return fmaf(b, l1 - l0, l0);
}

template <int SCALE, int MODE>
__global__ void __launch_bounds__(BW* BH, 3) ncc_kernel(cudaTextureObject_t tex, const float* img_f, const unsigned char* img_u8, int W, int H,
                                                        int layers, int reps, float jitter, float* out) {
    using G = Geo<SCALE>;
    constexpr int HS = 1 << SCALE;
    extern __shared__ __align__(16) unsigned char smem[];
    float* ref_tile = reinterpret_cast<float*>(smem);                         // the reference window, as in the product
    constexpr int RTW = BW + 2 * G::R, RTH = 2 * BH + 2 * G::R;
    unsigned char* src_raw = smem + RTW * RTH * sizeof(float);
    const int tid = threadIdx.y * BW + threadIdx.x;
    const int x0 = 64 + (blockIdx.x * BW) % (W - 256), y0 = 64 + (blockIdx.y * 2 * BH) % (H - 256);
    for (int i = tid; i < RTW * RTH; i += BW * BH) {
        const int ty = i / RTW, tx = i - ty * RTW;
        ref_tile[i] = img_f[(size_t)(y0 - G::R + ty) * W + (x0 - G::R + tx)];
    }
    const int tx = threadIdx.x, ly = 2 * threadIdx.y + (tx & 1);
    const int x = x0 + tx, y = y0 + ly;
    const float* centre = ref_tile + (ly + G::R) * RTW + (tx + G::R);
    // per-lane hash -> sub-pixel jitter: neighbouring pixels carry slightly different planes
    unsigned hsh = (blockIdx.x * 131 + blockIdx.y * 7919 + tid) * 2654435761u;
    const float jx = (((hsh >> 8) & 1023) * (1.0f / 1023.0f) - 0.5f) * 2.0f * jitter, jy = (((hsh >> 18) & 1023) * (1.0f / 1023.0f) - 0.5f) * 2.0f * jitter;
    // homography of a near-fronto-parallel plane: H = I + small terms (the reference's per-tap arithmetic runs on its 9 entries)
    float Hm[9] = {1.002f, 0.001f, jx + 0.37f, -0.0015f, 0.999f, jy + 0.61f, 1e-7f, -2e-7f, 1.0f};
    const int ox = x0 - G::R - MARGIN, oy = y0 - G::R - MARGIN;
    __syncthreads();
    const float r0 = centre[0];
    float acc = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        const int layer = rep % layers;
        const size_t lay_off = (size_t)layer * W * H;
        if (MODE == 1 || MODE == 4) {               // stage the float32 source tile of this view: 16-byte cp.async copies
            float* st = reinterpret_cast<float*>(src_raw);
            __syncthreads();
            constexpr int VEC = G::PITCH / 4;
            for (int i = tid; i < VEC * G::TH; i += BW * BH) {
                const int ty = i / VEC, v = i - ty * VEC;
                cp_async16(st + ty * G::PITCH + 4 * v, img_f + lay_off + (size_t)(oy + ty) * W + ((ox & ~3) + 4 * v));
            }
            cp_async_wait();
            __syncthreads();
        }
        if (MODE == 2) {                            // ... or the 8-bit one
            unsigned char* st = src_raw;
            __syncthreads();
            constexpr int P8 = G::P8;
            constexpr int VEC = P8 / 16;
            for (int i = tid; i < VEC * G::TH; i += BW * BH) {
                const int ty = i / VEC, v = i - ty * VEC;
                cp_async16(st + ty * P8 + 16 * v, img_u8 + lay_off + (size_t)(oy + ty) * W + ((ox & ~15) + 16 * v));
            }
            cp_async_wait();
            __syncthreads();
        }
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int i = (2 * a - 5) * HS;
            const float px = (float)(x + i);
            const float hx = __fmul_rn(Hm[0], px), hy = __fmul_rn(Hm[3], px), hz = __fmul_rn(Hm[6], px);
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const int j = (2 * b - 5) * HS;
                const float py = (float)(y + j);
                const float X = __fadd_rn(__fmaf_rn(Hm[1], py, hx), Hm[2]);
                const float Y = __fadd_rn(__fmaf_rn(Hm[4], py, hy), Hm[5]);
                const float Z = __fadd_rn(__fmaf_rn(Hm[7], py, hz), Hm[8]);
                const float rz = 1.0f / Z;                    // MUFU.RCP under --use_fast_math
                const float xs = __fmaf_rn(X, rz, 0.5f), ys = __fmaf_rn(Y, rz, 0.5f);
                float s;
                const bool use_tex = MODE == 0 || (MODE == 4 && ((a + b) & 1));
                if (use_tex) s = tex2DLayered<float>(tex, xs, ys, layer);
                else if (MODE == 1 || MODE == 4) s = soft_bilinear(reinterpret_cast<const float*>(src_raw), G::PITCH, xs, ys, ox & ~3, oy);
                else if (MODE == 2) s = soft_bilinear(src_raw, G::P8, xs, ys, ox & ~15, oy);
                else s = soft_bilinear(img_f + lay_off, W, xs, ys, 0, 0);
                const float r = centre[j * RTW + i];
                const float w = exp2f(__fmul_rn(__fmaf_rn(-1.4142135f * HS * (float)((a + 1) * (b + 1)) * 0.1f, 0.02f, -__fmul_rn(fabsf(r - r0), 0.0555555f)), 1.442695f));
                const float sw = __fmul_rn(w, s), rw = __fmul_rn(w, r);
                s1 = __fadd_rn(s1, sw);
                s2 = __fmaf_rn(sw, s, s2);
                s3 = __fmaf_rn(rw, s, s3);
            }
        }
        acc += s1 + s2 * 1e-3f + s3 * 1e-6f;
        Hm[2] += 0.013f; Hm[5] += 0.007f;          // the next hypothesis: a slightly different plane
    }
    out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * BW * BH + tid] = acc;
}

template <int SCALE, int MODE>
double run(cudaTextureObject_t tex, const float* img_f, const unsigned char* img_u8, int W, int H, int layers, int reps, float jitter, float* out, int sm_count, int clock_khz) {
    using G = Geo<SCALE>;
    constexpr int RTW = BW + 2 * G::R, RTH = 2 * BH + 2 * G::R;
    size_t smem = RTW * RTH * sizeof(float);
    if (MODE == 1 || MODE == 4) smem += (size_t)G::PITCH * G::TH * sizeof(float);
    if (MODE == 2) smem += (size_t)G::P8 * G::TH;
    smem += 40960;       // the product's candidate cost table (8 x 10 sources x 128 threads x 4 B): the same L1 carve-out pressure
    auto k = ncc_kernel<SCALE, MODE>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(96, 148 * 3 / 96 * 8 + 8), block(BW, BH);
    k<<<grid, block, smem>>>(tex, img_f, img_u8, W, H, layers, 2, jitter, out);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<grid, block, smem>>>(tex, img_f, img_u8, W, H, layers, reps, jitter, out);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double taps = (double)grid.x * grid.y * BW * BH * reps * 36.0;
    return taps / (ms * 1e-3) / ((double)clock_khz * 1e3) / sm_count;       // taps per clock per SM
}

int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 60;
    const int W = 3200, H = 2130, L = 11;
    std::vector<float> hf((size_t)W * H * L);
    std::vector<unsigned char> hu(hf.size());
    unsigned s = 12345;
    for (size_t i = 0; i < hf.size(); ++i) { s = s * 1664525u + 1013904223u; hu[i] = (unsigned char)(100 + ((s >> 24) & 63)); hf[i] = (float)hu[i]; }
    float* d_f; unsigned char* d_u; float* d_out;
    CK(cudaMalloc(&d_f, hf.size() * 4)); CK(cudaMalloc(&d_u, hu.size())); CK(cudaMalloc(&d_out, (size_t)96 * 64 * 128 * 4));
    CK(cudaMemcpy(d_f, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_u, hu.data(), hu.size(), cudaMemcpyHostToDevice));
    cudaArray_t arr;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    CK(cudaMalloc3DArray(&arr, &desc, make_cudaExtent(W, H, L), cudaArrayLayered));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(hf.data(), W * 4, W, H);
    p.dstArray = arr;
    p.extent = make_cudaExtent(W, H, L);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &res, &td, nullptr));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const char* names[5] = {"texture_unit", "smem_tile_f32", "smem_tile_u8", "global_lsu", "hybrid_tex_smem"};
    printf("{\"reps\": %d, \"sm\": %d, \"clock_khz\": %d, \"unit\": \"taps/clk/SM (36-tap NCC accumulations with the exact arithmetic's per-tap work)\"", reps, prop.multiProcessorCount, clock_khz);
    for (int jit = 0; jit < 2; ++jit) {
        const float jitter = jit ? 1.5f : 0.25f;
        printf(", \"jitter_%s_px\": {", jit ? "1.5" : "0.25");
        double v[3][5];
#define RUN(S, M) v[S][M] = run<S, M>(tex, d_f, d_u, W, H, L, reps, jitter, d_out, prop.multiProcessorCount, clock_khz);
        RUN(0, 0) RUN(0, 1) RUN(0, 2) RUN(0, 3) RUN(0, 4) RUN(1, 0) RUN(1, 1) RUN(1, 2) RUN(1, 3) RUN(1, 4) RUN(2, 0) RUN(2, 1) RUN(2, 2) RUN(2, 3) RUN(2, 4)
#undef RUN
        for (int sc = 0; sc < 3; ++sc) {
            printf("%s\"scale%d\": {", sc ? ", " : "", sc);
            for (int m = 0; m < 5; ++m) printf("%s\"%s\": %.3f", m ? ", " : "", names[m], v[sc][m]);
            printf("}");
        }
        printf("}");
    }
    printf("}\n");
    return 0;
}
