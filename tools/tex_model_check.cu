// tools/tex_model_check.cu -- is the texture unit's bilinear filter reproducible in software, bit for bit?
//
// The PatchMatch views are 8-bit grey levels (cv::imread(IMREAD_GRAYSCALE) -> float, PatchMatch.cpp:877-882). With texels
// that are integers 0..255 and interpolation weights held in 1.8 fixed point, the bilinear value is a multiple of 2^-16
// below 256, i.e. exactly representable in float32 -- so IF the unit forms its weights the way the model below does, a
// software filter over plain loads (global or shared memory, 8-bit storage) returns the very bits the float32 texture
// returns. This program measures that on the GPU: random images, coordinates of every kind the NCC produces (interior,
// borders, outside, exact ties of the weight rounding), four weight models, and the UNORM8 texture path as well.
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tex_model_check tools/tex_model_check.cu
// Run on the GPU box: tools/tex_model_check [W H] ; prints one JSON line.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum { M_HALF_UP = 0, M_TRUNC = 1, M_NEAREST_EVEN = 2, M_FIXED_COORD = 3, N_MODELS = 4 };

__device__ __forceinline__ unsigned xs32(unsigned& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// weights a, b in units of 1/256 (0..256) and the top-left texel (i0, j0) for coordinate (x, y) under model M
template <int M>
__device__ __forceinline__ void weights(float x, float y, int& i0, int& j0, int& a, int& b) {
    if (M == M_FIXED_COORD) {           // the coordinate itself goes to 8 fractional bits first
        const int xi = __float2int_rd(x * 256.0f + 0.5f) - 128, yi = __float2int_rd(y * 256.0f + 0.5f) - 128;
        i0 = xi >> 8; a = xi & 255; j0 = yi >> 8; b = yi & 255;
        return;
    }
    const float xb = x - 0.5f, yb = y - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float ua = (xb - fx) * 256.0f, ub = (yb - fy) * 256.0f;
    i0 = (int)fx; j0 = (int)fy;
    if (M == M_HALF_UP) { a = (int)floorf(ua + 0.5f); b = (int)floorf(ub + 0.5f); }
    else if (M == M_TRUNC) { a = (int)ua; b = (int)ub; }
    else { a = __float2int_rn(ua); b = __float2int_rn(ub); }
}

template <int M>
__device__ __forceinline__ float model(const unsigned char* img, int W, int H, float x, float y) {
    int i0, j0, a, b;
    weights<M>(x, y, i0, j0, a, b);
    const int i1 = min(max(i0 + 1, 0), W - 1), j1 = min(max(j0 + 1, 0), H - 1);
    i0 = min(max(i0, 0), W - 1); j0 = min(max(j0, 0), H - 1);
    const int t00 = img[(size_t)j0 * W + i0], t10 = img[(size_t)j0 * W + i1], t01 = img[(size_t)j1 * W + i0], t11 = img[(size_t)j1 * W + i1];
    const int v = (256 - a) * (256 - b) * t00 + a * (256 - b) * t10 + (256 - a) * b * t01 + a * b * t11;   // < 2^24
    return (float)v * (1.0f / 65536.0f);
}

struct Counts {
    unsigned long long n, mism[N_MODELS];
    unsigned long long u8_times255_equal, u8_recovered_equal, u8_is_div255, u8_is_mul_inv255;
    float first_bad[8][5];    // x, y, hw, model0, model3
    unsigned n_bad;
    float u8_max_abs;
};

__global__ void check_kernel(cudaTextureObject_t texf, cudaTextureObject_t texu, const unsigned char* img, int W, int H, int layer,
                             int per_thread, Counts* out) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned long long n = 0, mm[N_MODELS] = {0, 0, 0, 0}, e255 = 0, erec = 0, ediv = 0, emul = 0;
    float umax = 0.f;
    for (int k = 0; k < per_thread; ++k) {
        const unsigned kind = xs32(s) % 8;
        float x, y;
        const float rx = (xs32(s) >> 8) * (1.0f / 16777216.0f), ry = (xs32(s) >> 8) * (1.0f / 16777216.0f);
        if (kind < 4) { x = rx * (W + 4) - 2.0f; y = ry * (H + 4) - 2.0f; }                       // anywhere, full float precision
        else if (kind == 4) { x = floorf(rx * W) + 0.5f + (ry - 0.5f) * 1e-3f; y = ry * H; }     // next to a texel centre
        else if (kind == 5) { x = floorf(rx * W * 512.0f) / 512.0f; y = floorf(ry * H * 512.0f) / 512.0f; }   // exact ties of the weight rounding
        else if (kind == 6) { x = rx * 3.0f - 1.0f; y = ry * H; }                                 // left border
        else { x = W - 2.0f + rx * 3.0f; y = H - 2.0f + ry * 3.0f; }                              // bottom-right corner
        const float hw = tex2DLayered<float>(texf, x, y, layer);
        const float m0 = model<M_HALF_UP>(img, W, H, x, y), m1 = model<M_TRUNC>(img, W, H, x, y);
        const float m2 = model<M_NEAREST_EVEN>(img, W, H, x, y), m3 = model<M_FIXED_COORD>(img, W, H, x, y);
        ++n;
        mm[0] += hw != m0; mm[1] += hw != m1; mm[2] += hw != m2; mm[3] += hw != m3;
        if (hw != m0 && hw != m3) {
            const unsigned slot = atomicAdd(&out->n_bad, 1u);
            if (slot < 8) { out->first_bad[slot][0] = x; out->first_bad[slot][1] = y; out->first_bad[slot][2] = hw; out->first_bad[slot][3] = m0; out->first_bad[slot][4] = m3; }
        }
        const float hu = tex2DLayered<float>(texu, x, y, layer);
        e255 += (hu * 255.0f == hw);
        erec += (rintf(hu * 255.0f * 65536.0f) * (1.0f / 65536.0f) == hw);
        ediv += (hu == __fdiv_rn(hw, 255.0f));
        emul += (hu == hw * (1.0f / 255.0f));
        umax = fmaxf(umax, fabsf(hu * 255.0f - hw));
    }
    atomicAdd(&out->n, n);
    for (int m = 0; m < N_MODELS; ++m) atomicAdd(&out->mism[m], mm[m]);
    atomicAdd(&out->u8_times255_equal, e255);
    atomicAdd(&out->u8_recovered_equal, erec);
    atomicAdd(&out->u8_is_div255, ediv);
    atomicAdd(&out->u8_is_mul_inv255, emul);
    atomicMax((int*)&out->u8_max_abs, __float_as_int(umax));
}

template <class T>
cudaTextureObject_t make_tex(const std::vector<unsigned char>& img, int W, int H, int L, cudaChannelFormatDesc desc, bool unorm) {
    cudaArray_t arr;
    CK(cudaMalloc3DArray(&arr, &desc, make_cudaExtent(W, H, L), cudaArrayLayered));
    std::vector<T> host((size_t)W * H);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (T)img[i];
    for (int l = 0; l < L; ++l) {
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(host.data(), W * sizeof(T), W, H);
        p.dstArray = arr;
        p.dstPos = make_cudaPos(0, 0, l);
        p.extent = make_cudaExtent(W, H, 1);
        p.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3D(&p));
    }
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td = {};      // the product's descriptor (pm_capi.cu: cache_alloc), which is the reference's (PatchMatch.cpp:1012-1018)
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = unorm ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &res, &td, nullptr));
    return tex;
}

int main(int argc, char** argv) {
    const int W = argc > 2 ? atoi(argv[1]) : 3200, H = argc > 2 ? atoi(argv[2]) : 2130;
    std::vector<unsigned char> img((size_t)W * H);
    unsigned s = 777;
    for (auto& v : img) { s = s * 1664525u + 1013904223u; v = (unsigned char)(s >> 24); }
    // a band of extreme contrast and a flat band, like the synthetic scenes have
    for (int y = 0; y < H / 8; ++y) for (int x = 0; x < W; ++x) img[(size_t)y * W + x] = ((x ^ y) & 1) ? 255 : 0;
    for (int y = H / 8; y < H / 4; ++y) for (int x = 0; x < W; ++x) img[(size_t)y * W + x] = 200 + ((x + y) % 3);
    unsigned char* d_img;
    CK(cudaMalloc(&d_img, img.size()));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    cudaTextureObject_t texf = make_tex<float>(img, W, H, 2, cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat), false);
    cudaTextureObject_t texu = make_tex<unsigned char>(img, W, H, 2, cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned), true);
    Counts* d_c;
    CK(cudaMalloc(&d_c, sizeof(Counts)));
    CK(cudaMemset(d_c, 0, sizeof(Counts)));
    check_kernel<<<148 * 8, 256>>>(texf, texu, d_img, W, H, 1, 2048, d_c);
    CK(cudaDeviceSynchronize());
    Counts c;
    CK(cudaMemcpy(&c, d_c, sizeof(c), cudaMemcpyDeviceToHost));
    printf("{\"image\": [%d, %d], \"samples\": %llu, \"f32_texture_vs_model_mismatches\": {\"round_half_up\": %llu, \"truncate\": %llu, \"nearest_even\": %llu, "
           "\"fixed_point_coordinate\": %llu}, \"unorm8_texture\": {\"times255_equals_f32\": %llu, \"rounded_to_2^-16_equals_f32\": %llu, "
           "\"is_f32_div_255\": %llu, \"is_f32_times_inv255\": %llu, \"max_abs_diff_times255\": %g}, \"neither_model_0_nor_3\": %u, \"examples\": [",
           W, H, c.n, c.mism[0], c.mism[1], c.mism[2], c.mism[3], c.u8_times255_equal, c.u8_recovered_equal, c.u8_is_div255, c.u8_is_mul_inv255,
           c.u8_max_abs, c.n_bad);
    for (unsigned i = 0; i < (c.n_bad < 8 ? c.n_bad : 8); ++i)
        printf("%s{\"x\": %.9g, \"y\": %.9g, \"hw\": %.9g, \"half_up\": %.9g, \"fixed\": %.9g}", i ? ", " : "", c.first_bad[i][0], c.first_bad[i][1],
               c.first_bad[i][2], c.first_bad[i][3], c.first_bad[i][4]);
    printf("]}\n");
    return 0;
}
