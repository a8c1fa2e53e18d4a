// ref_sky_harness.cu -- TEST INFRASTRUCTURE: the reference's own sky-mask joint-bilateral kernel, compiled where it lies
// (/root/reference/SkySegment/src/SkyRegionDetect.cu, flags of /root/reference/CMakeLists.txt:18 retargeted to sm_100)
// and launched with the reference's grid (SkyRegionDetect.cu:55-60) on raw arrays. Only tests, smoke() and the golden
// generator may load the resulting oracle/_ref/libmpmvs_ref_sky.so; the product never does.
#include "SkySegment/src/SkyRegionDetect.cu"

extern "C" int ref_sky_filter(const unsigned char* bgr, const float* mask, float* result, int height, int width) {
    unsigned char* src = nullptr;
    float *dmask = nullptr, *dres = nullptr;
    const size_t wh = (size_t)height * width;
    if (cudaMalloc((void**)&src, wh * 3) != cudaSuccess) return 1;
    if (cudaMalloc((void**)&dmask, wh * 4) != cudaSuccess) return 1;
    if (cudaMalloc((void**)&dres, wh * 4) != cudaSuccess) return 1;
    cudaMemcpy(src, bgr, wh * 3, cudaMemcpyHostToDevice);
    cudaMemcpy(dmask, mask, wh * 4, cudaMemcpyHostToDevice);
    const int BLOCK_W = 32, BLOCK_H = BLOCK_W / 2;
    const dim3 blockSize(BLOCK_W, BLOCK_H, 1);
    const dim3 gridSize((width + BLOCK_W - 1) / BLOCK_W, (height + BLOCK_H - 1) / BLOCK_H, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    Pixel_bilateral_filter<<<gridSize, blockSize>>>(src, dmask, dres, height, width);
    cudaEventRecord(e1);
    const cudaError_t rc = cudaDeviceSynchronize();
    cudaMemcpy(result, dres, wh * 4, cudaMemcpyDeviceToHost);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(src); cudaFree(dmask); cudaFree(dres);
    return rc == cudaSuccess ? 0 : (int)rc;
}

static float g_last_ms = 0.f;
extern "C" float ref_sky_filter_timed(const unsigned char* bgr, const float* mask, float* result, int height, int width, int reps) {
    unsigned char* src = nullptr;
    float *dmask = nullptr, *dres = nullptr;
    const size_t wh = (size_t)height * width;
    cudaMalloc((void**)&src, wh * 3); cudaMalloc((void**)&dmask, wh * 4); cudaMalloc((void**)&dres, wh * 4);
    cudaMemcpy(src, bgr, wh * 3, cudaMemcpyHostToDevice);
    cudaMemcpy(dmask, mask, wh * 4, cudaMemcpyHostToDevice);
    const dim3 blockSize(32, 16, 1), gridSize((width + 31) / 32, (height + 15) / 16, 1);
    Pixel_bilateral_filter<<<gridSize, blockSize>>>(src, dmask, dres, height, width);   // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) Pixel_bilateral_filter<<<gridSize, blockSize>>>(src, dmask, dres, height, width);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaMemcpy(result, dres, wh * 4, cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&g_last_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(src); cudaFree(dmask); cudaFree(dres);
    return g_last_ms / (reps > 0 ? reps : 1);
}
