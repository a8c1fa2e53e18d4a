// tools/tex_microbench.cu -- texture-unit roofline microbenchmark for B200 (BASELINE.json config 5 groundwork).
// Measures sustained bilinear (and point) fetch rate in samples/clk/SM for the access patterns of the PatchMatch NCC:
// a warp = 32 neighbouring pixels (row strip or checkerboard), 36 independent taps per "NCC", texel formats f32/f16/u8,
// with optional per-lane position jitter (different plane hypotheses per pixel).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tex_microbench tools/tex_microbench.cu ; run on the GPU box.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// pattern: 0 = row strip (lane i -> x+i, same row), 1 = checkerboard (lane i -> x+i, row + (i&1)), 2 = 8x4 block
// lane_mask: which lanes of every warp take part (0xffffffff = all). Idle lanes skip the fetches (divergent branch), as
// lanes of the PatchMatch kernel do when their pixel does not need the view being scored.
template <int PATTERN, int STEP>
__global__ void __launch_bounds__(256) fetch_kernel(cudaTextureObject_t tex, int W, int H, int layers, float jitter, int reps, float scale,
                                                    float* out, unsigned lane_mask = 0xffffffffu) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int px, py;
    if (PATTERN == 0) { px = lane; py = 0; }
    else if (PATTERN == 1) { px = lane; py = lane & 1; }
    else { px = lane & 7; py = lane >> 3; }
    const int bx = (blockIdx.x * 8 + warp) * 32 % (W - 128) + 48, by = (blockIdx.x * 37 + warp * 5) % (H - 128) + 48;
    // per-lane hash for jitter
    unsigned h = (blockIdx.x * 256 + threadIdx.x) * 2654435761u;
    const float jx = ((h >> 8) & 1023) * (1.0f / 1023.0f) * jitter, jy = ((h >> 18) & 1023) * (1.0f / 1023.0f) * jitter;
    float acc = 0.f;
    float fx = bx + px * scale + jx + 0.37f, fy = by + py * scale + jy + 0.61f;
    if (!((lane_mask >> lane) & 1u)) return;
    for (int r = 0; r < reps; ++r) {
        const int layer = r % layers;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b)
                acc += tex2DLayered<float>(tex, fx + (a - 2.5f) * STEP * scale, fy + (b - 2.5f) * STEP * scale, layer);
        fx += 0.013f; fy += 0.007f;
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <class T>
cudaTextureObject_t make_tex(int W, int H, int L, cudaChannelFormatDesc desc, bool linear, bool norm, cudaArray_t* arr) {
    CK(cudaMalloc3DArray(arr, &desc, make_cudaExtent(W, H, L), cudaArrayLayered));
    std::vector<T> host((size_t)W * H);
    for (int l = 0; l < L; ++l) {
        for (size_t i = 0; i < host.size(); ++i) host[i] = (T)((i * 7 + l * 13) % 251);
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(host.data(), W * sizeof(T), W, H);
        p.dstArray = *arr;
        p.dstPos = make_cudaPos(0, 0, l);
        p.extent = make_cudaExtent(W, H, 1);
        p.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3D(&p));
    }
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = *arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = norm ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &res, &td, nullptr));
    return tex;
}

template <int PATTERN, int STEP>
void run(const char* name, cudaTextureObject_t tex, int W, int H, int L, float jitter, float scale, float* dout, int clock_khz, int sms) {
    const int blocks = sms * 24, reps = 64;
    fetch_kernel<PATTERN, STEP><<<blocks, 256>>>(tex, W, H, L, jitter, 4, scale, dout);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fetch_kernel<PATTERN, STEP><<<blocks, 256>>>(tex, W, H, L, jitter, reps, scale, dout);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double taps = (double)blocks * 256 * reps * 36;
    const double per_clk_sm = taps / (ms * 1e-3) / (clock_khz * 1e3) / sms;
    printf("%-44s pattern %d step %d jitter %5.1f scale %.2f : %8.3f ms  %7.1f Gtaps/s  %.3f taps/clk/SM (at %d MHz)\n", name, PATTERN, STEP, jitter,
           scale, ms, taps / ms / 1e6, per_clk_sm, clock_khz / 1000);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s: %d SMs, clock %d kHz\n", prop.name, sms, clock_khz);
    const int W = 1600, H = 1200, L = 11;
    float* dout;
    CK(cudaMalloc(&dout, 4));
    cudaArray_t a;
    struct Fmt { const char* name; cudaTextureObject_t tex; };
    std::vector<Fmt> fmts;
    fmts.push_back({"f32 linear", make_tex<float>(W, H, L, cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat), true, false, &a)});
    fmts.push_back({"f32 point", make_tex<float>(W, H, L, cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat), false, false, &a)});
    fmts.push_back({"f16 linear", make_tex<__half>(W, H, L, cudaCreateChannelDescHalf(), true, false, &a)});
    fmts.push_back({"u8 unorm linear", make_tex<unsigned char>(W, H, L, cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned), true, true, &a)});
    fmts.push_back({"u8 unorm point", make_tex<unsigned char>(W, H, L, cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned), false, true, &a)});
    for (auto& f : fmts) {
        for (float jitter : {0.0f, 2.0f, 16.0f}) {
            run<0, 2>(f.name, f.tex, W, H, L, jitter, 1.0f, dout, clock_khz, sms);
            run<1, 2>(f.name, f.tex, W, H, L, jitter, 1.0f, dout, clock_khz, sms);
            run<2, 2>(f.name, f.tex, W, H, L, jitter, 1.0f, dout, clock_khz, sms);
            run<1, 8>(f.name, f.tex, W, H, L, jitter, 1.0f, dout, clock_khz, sms);
        }
        run<1, 2>(f.name, f.tex, W, H, L, 0.0f, 0.5f, dout, clock_khz, sms);   // minified source (texels shared by neighbours)
        run<1, 2>(f.name, f.tex, W, H, L, 0.0f, 2.0f, dout, clock_khz, sms);   // magnified 2x (sparser footprint)
    }
    // partially active warps: is the unit's cost per warp instruction or per active quad / lane?
    struct M { const char* name; unsigned mask; int lanes; };
    const M masks[] = {{"all 32 lanes", 0xffffffffu, 32}, {"lanes 0-15 (4 full quads)", 0x0000ffffu, 16}, {"every other quad (4 full quads)", 0x0f0f0f0fu, 16},
                       {"2 lanes of every quad", 0x33333333u, 16}, {"26 lanes (6 full quads + 2 lanes)", 0x03ffffffu, 26}, {"3 lanes of every quad", 0x77777777u, 24}};
    for (const M& m : masks) {
        const int blocks = sms * 24, reps = 64;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        fetch_kernel<1, 2><<<blocks, 256>>>(fmts[3].tex, W, H, L, 0.f, 4, 1.0f, dout, m.mask);
        cudaEventRecord(e0);
        fetch_kernel<1, 2><<<blocks, 256>>>(fmts[3].tex, W, H, L, 0.f, reps, 1.0f, dout, m.mask);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double taps = (double)blocks * 8 * m.lanes * reps * 36;
        printf("u8 unorm linear, %-36s : %8.3f ms  %7.1f Gtaps/s  %.3f taps/clk/SM  (%.3f warp-instr/clk/SM x 8)\n", m.name, ms, taps / ms / 1e6,
               taps / (ms * 1e-3) / (clock_khz * 1e3) / sms, (double)blocks * 8 * reps * 36 / (ms * 1e-3) / (clock_khz * 1e3) / sms * 8);
    }
    return 0;
}
