// pm_fusion.cu -- depth-map fusion on the GPU: the per-pixel consistency test of RunFusion
// (/root/reference/src/PatchMatch.cpp:287-504, "cpp:NNN") with all depth/normal maps of a scene resident in HBM.
//
// The reference visits images in order and pixels in raster order on one host thread; a fused point masks the source
// pixels it used, so they are neither fused again nor used as evidence later (cpp:380,404-405,467-471). Kept here:
// images are processed in the reference's order, one launch per image, and masks written while processing image i are
// visible to every later image. Relaxed, deliberately: the pixels of ONE image are processed in parallel against the
// masks as they were when that image started (double-buffered, so the result is deterministic), whereas the reference
// lets a pixel see the masks set by earlier pixels of the same image; and the reference's `used_list`, which is not
// reset per pixel and therefore also masks stale positions of earlier pixels (cpp:371), is per pixel here.
// Points come out in raster order per image (compaction by prefix sum), like the reference's push_back order.
#include <cuda_runtime.h>

#include <cub/device/device_select.cuh>

#include <vector>

#include "../../include/mpmvs_b200.h"

#define FCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace {

struct FusionView {
    mpmvs_camera cam;
    const float* depth;          // [H][W]
    const float* normal;         // [H][W][3] world frame
    const unsigned char* gray;   // [H][W]
    const unsigned char* bgr;    // optional [H][W][3], cv::imread(IMREAD_COLOR) order (cpp:322): the colour the reference averages
    unsigned char* mask_prev;    // masks as of the start of the current image's launch
    unsigned char* mask_next;    // masks set during the current launch
    const unsigned char* sky;    // optional [H][W], > 0 = sky (skymask_refine, cpp:358-373)
    int W, H;
};

struct alignas(4) FusedPoint { float coord[3], normal[3], color[3]; };   // struct PointList, PatchMatch.h:29-33

__device__ __forceinline__ void world_point(float x, float y, float depth, const mpmvs_camera& c, float* X) {   // cpp:211-230
    const float px = depth * (x - c.K[2]) / c.K[0], py = depth * (y - c.K[5]) / c.K[4], pz = depth;
    X[0] = c.R[0] * px + c.R[3] * py + c.R[6] * pz + c.C[0];
    X[1] = c.R[1] * px + c.R[4] * py + c.R[7] * pz + c.C[1];
    X[2] = c.R[2] * px + c.R[5] * py + c.R[8] * pz + c.C[2];
}
__device__ __forceinline__ void project(const float* X, const mpmvs_camera& c, float& u, float& v, float& depth) {   // cpp:252-262
    const float tx = c.R[0] * X[0] + c.R[1] * X[1] + c.R[2] * X[2] + c.t[0];
    const float ty = c.R[3] * X[0] + c.R[4] * X[1] + c.R[5] * X[2] + c.t[1];
    const float tz = c.R[6] * X[0] + c.R[7] * X[1] + c.R[8] * X[2] + c.t[2];
    depth = c.K[6] * tx + c.K[7] * ty + c.K[8] * tz;
    u = (c.K[0] * tx + c.K[1] * ty + c.K[2] * tz) / depth;
    v = (c.K[3] * tx + c.K[4] * ty + c.K[5] * tz) / depth;
}

// one thread per pixel of reference view `ref`; srcs = its source views (indices into `views`, -1 = not estimated)
__global__ void __launch_bounds__(256) pm_fuse_kernel(const FusionView* views, int ref, const int* srcs, int num_ngb, int dynamic,
                                                      FusedPoint* out, unsigned char* keep) {
    const FusionView& R = views[ref];
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (c >= R.W || r >= R.H) return;
    const size_t idx = (size_t)r * R.W + c;
    keep[idx] = 0;
    if (R.mask_prev[idx] == 1) return;                                   // cpp:380
    const float ref_depth = R.depth[idx];
    if (ref_depth <= 0.0f) return;                                       // cpp:389-391
    const float rn[3] = {R.normal[3 * idx], R.normal[3 * idx + 1], R.normal[3 * idx + 2]};
    float PX[3];
    world_point((float)c, (float)r, ref_depth, R.cam, PX);
    float sumP[3] = {PX[0], PX[1], PX[2]}, sumN[3] = {rn[0], rn[1], rn[2]};
    // consistent_Color (cpp:399,443-445): B, G, R of the colour image; a view without one contributes its grey level thrice
    float sumC[3];
    for (int k = 0; k < 3; ++k) sumC[k] = R.bgr ? (float)R.bgr[3 * idx + k] : (float)R.gray[idx];
    int num_consistent = 0;
    float dyn = 0.f;
    int used[MPMVS_MAX_VIEWS];
    for (int j = 1; j < num_ngb; ++j) {
        used[j] = -1;
        if (j == num_ngb - 1 && num_consistent == 0) break;             // cpp:401-402
        const int s = srcs[j];
        if (s < 0) continue;
        const FusionView& S = views[s];
        float u, v, pd;
        project(PX, S.cam, u, v, pd);
        const int sr = (int)(v + 0.5f), sc = (int)(u + 0.5f);
        if (!(sc >= 0 && sc < S.W && sr >= 0 && sr < S.H)) continue;
        const size_t sidx = (size_t)sr * S.W + sc;
        if (S.mask_prev[sidx] == 1) continue;
        const float src_depth = S.depth[sidx];
        if (src_depth <= 0.0f) continue;
        float TX[3], bu, bv;
        world_point((float)sc, (float)sr, src_depth, S.cam, TX);
        project(TX, R.cam, bu, bv, pd);
        const float reproj = sqrtf((c - bu) * (c - bu) + (r - bv) * (r - bv));
        if (reproj >= 2.0f) continue;                                    // cpp:423-424
        const float rel = fabsf(pd - ref_depth) / ref_depth;
        if (rel >= 0.01f) continue;                                      // cpp:427-428
        const float sn[3] = {S.normal[3 * sidx], S.normal[3 * sidx + 1], S.normal[3 * sidx + 2]};
        float angle = acosf(rn[0] * sn[0] + rn[1] * sn[1] + rn[2] * sn[2]);   // GetAngle, cpp:241-250
        if (angle != angle) angle = 0.f;
        if (angle < 0.174533f) {
            used[j] = (int)sidx;
            for (int k = 0; k < 3; ++k) { sumP[k] += TX[k]; sumN[k] += sn[k]; }
            for (int k = 0; k < 3; ++k) sumC[k] += S.bgr ? (float)S.bgr[3 * sidx + k] : (float)S.gray[sidx];
            dyn += expf(-(reproj + 200 * rel + angle * 10));             // cpp:444-446
            ++num_consistent;
        }
    }
    const bool ok = dynamic ? (num_consistent >= 1 && dyn > 0.3f * num_consistent) : (num_consistent >= 2);   // cpp:451,474
    if (!ok) return;
    FusedPoint p;
    // cpp:454-460: the point and the colour are DIVIDED by (n + 1); the normal is a cv::Vec3f, whose `operator/=(float)` (OpenCV's
    // core/matx.hpp) multiplies by the float reciprocal 1.f / alpha instead. IEEE operations whatever --use_fast_math says.
    const float n1 = num_consistent + 1.0f, ialpha = __frcp_rn(n1);
    for (int k = 0; k < 3; ++k) { p.coord[k] = __fdiv_rn(sumP[k], n1); p.normal[k] = __fmul_rn(sumN[k], ialpha); p.color[k] = __fdiv_rn(sumC[k], n1); }
    out[idx] = p;
    keep[idx] = 1;
    for (int j = 1; j < num_ngb; ++j)
        if (used[j] >= 0 && srcs[j] >= 0) views[srcs[j]].mask_next[used[j]] = 1;    // cpp:467-471 (idempotent store)
}

__global__ void __launch_bounds__(256) pm_mask_merge_kernel(unsigned char* prev, const unsigned char* next, size_t n) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n && next[i]) prev[i] = 1;
}

// cpp:385-388: when an image's turn comes, its sky pixels are masked and skipped (earlier images could still use them as evidence)
__global__ void __launch_bounds__(256) pm_mask_sky_kernel(unsigned char* prev, unsigned char* next, const unsigned char* sky, size_t n) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n && sky[i] > 0) { prev[i] = 1; next[i] = 1; }
}

}  // namespace

struct mpmvs_fusion {
    int device = 0, n = 0;
    std::vector<FusionView> hviews;
    std::vector<void*> owned;
    FusionView* dviews = nullptr;
    FusedPoint* d_points = nullptr;      // all fused points, image after image
    size_t n_points = 0, cap_points = 0;
};

extern "C" {

int mpmvs_fusion_create(int device, int n_images, mpmvs_fusion** out) {
    if (!out || n_images <= 0) return MPMVS_E_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return MPMVS_E_NO_DEVICE;
    if (device < 0 || device >= count) return MPMVS_E_ARG;
    FCK(cudaSetDevice(device));
    mpmvs_fusion* f = new mpmvs_fusion();
    f->device = device;
    f->n = n_images;
    f->hviews.resize(n_images);
    for (FusionView& v : f->hviews) { v = FusionView{}; v.W = v.H = 0; }
    *out = f;
    return MPMVS_OK;
}

int mpmvs_fusion_destroy(mpmvs_fusion* f) {
    if (!f) return MPMVS_E_ARG;
    cudaSetDevice(f->device);
    for (void* p : f->owned) cudaFree(p);
    cudaFree(f->dviews);
    cudaFree(f->d_points);
    delete f;
    return MPMVS_OK;
}

// depth [h][w], normal [h][w][3] (world frame, as normals.dmb), gray [h][w] uint8; host pointers, copied to the device
int mpmvs_fusion_set_view(mpmvs_fusion* f, int index, const mpmvs_camera* cam, const float* depth, const float* normal,
                          const uint8_t* gray) {
    if (!f || index < 0 || index >= f->n || !cam || !depth || !normal || !gray || cam->width <= 0 || cam->height <= 0) return MPMVS_E_ARG;
    FCK(cudaSetDevice(f->device));
    FusionView& v = f->hviews[index];
    if (v.W) return MPMVS_E_ARG;     // already set
    const size_t wh = (size_t)cam->width * cam->height;
    float *dd, *dn;
    unsigned char *dg, *m0, *m1;
    FCK(cudaMalloc((void**)&dd, wh * 4)); f->owned.push_back(dd);
    FCK(cudaMalloc((void**)&dn, wh * 12)); f->owned.push_back(dn);
    FCK(cudaMalloc((void**)&dg, wh)); f->owned.push_back(dg);
    FCK(cudaMalloc((void**)&m0, wh)); f->owned.push_back(m0);
    FCK(cudaMalloc((void**)&m1, wh)); f->owned.push_back(m1);
    FCK(cudaMemcpy(dd, depth, wh * 4, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(dn, normal, wh * 12, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(dg, gray, wh, cudaMemcpyHostToDevice));
    FCK(cudaMemset(m0, 0, wh));
    FCK(cudaMemset(m1, 0, wh));
    v.cam = *cam; v.depth = dd; v.normal = dn; v.gray = dg; v.mask_prev = m0; v.mask_next = m1; v.W = cam->width; v.H = cam->height;
    return MPMVS_OK;
}

// bgr [h][w][3] uint8 (cv::imread(IMREAD_COLOR) channel order, at the size of the depth map): the colour RunFusion averages
// per point (cpp:322,399,443-445). Optional: a view without it contributes its grey level to all three channels.
int mpmvs_fusion_set_color(mpmvs_fusion* f, int index, const uint8_t* bgr) {
    if (!f || index < 0 || index >= f->n || !bgr) return MPMVS_E_ARG;
    FusionView& v = f->hviews[index];
    if (!v.W) return MPMVS_E_ARG;    // mpmvs_fusion_set_view first
    FCK(cudaSetDevice(f->device));
    const size_t bytes = (size_t)v.W * v.H * 3;
    unsigned char* d = const_cast<unsigned char*>(v.bgr);
    if (!d) { FCK(cudaMalloc((void**)&d, bytes)); f->owned.push_back(d); v.bgr = d; }
    FCK(cudaMemcpy(d, bgr, bytes, cudaMemcpyHostToDevice));
    return MPMVS_OK;
}

int mpmvs_fusion_set_sky_mask(mpmvs_fusion* f, int index, const uint8_t* sky) {
    if (!f || index < 0 || index >= f->n || !sky) return MPMVS_E_ARG;
    FusionView& v = f->hviews[index];
    if (!v.W) return MPMVS_E_ARG;    // mpmvs_fusion_set_view first
    FCK(cudaSetDevice(f->device));
    const size_t wh = (size_t)v.W * v.H;
    unsigned char* d = const_cast<unsigned char*>(v.sky);
    if (!d) { FCK(cudaMalloc((void**)&d, wh)); f->owned.push_back(d); v.sky = d; }
    FCK(cudaMemcpy(d, sky, wh, cudaMemcpyHostToDevice));
    return MPMVS_OK;
}

// src_lists: n_images rows of `max_list` view indices; row i = [i, its sources...] padded with -2 (end); -1 = a listed
// source that was not estimated. use_dynamic_consistency: "Use dynamic_consistency to fuse" (cpp:451 vs 474).
int mpmvs_fusion_run(mpmvs_fusion* f, const int* src_lists, int max_list, int use_dynamic_consistency, uint64_t* n_points, float* ms) {
    if (!f || !src_lists || max_list < 2 || max_list > MPMVS_MAX_VIEWS) return MPMVS_E_ARG;
    FCK(cudaSetDevice(f->device));
    if (!f->dviews) FCK(cudaMalloc((void**)&f->dviews, sizeof(FusionView) * f->n));
    FCK(cudaMemcpy(f->dviews, f->hviews.data(), sizeof(FusionView) * f->n, cudaMemcpyHostToDevice));
    size_t total = 0, max_wh = 0;
    for (const FusionView& v : f->hviews) { const size_t wh = (size_t)v.W * v.H; total += wh; max_wh = wh > max_wh ? wh : max_wh; }
    // The point buffer grows with the points actually fused (typically a small fraction of the pixels), not with the pixel
    // count of the scene: 36 B per pixel of every view would be 23 GB for 96 views of 3200 x 2130.
    auto reserve = [&](size_t need) -> int {
        if (need <= f->cap_points) return MPMVS_OK;
        const size_t cap = need > 2 * f->cap_points ? need : 2 * f->cap_points;
        FusedPoint* grown = nullptr;
        FCK(cudaMalloc((void**)&grown, cap * sizeof(FusedPoint)));
        if (f->n_points) FCK(cudaMemcpy(grown, f->d_points, f->n_points * sizeof(FusedPoint), cudaMemcpyDeviceToDevice));
        cudaFree(f->d_points);
        f->d_points = grown;
        f->cap_points = cap;
        return MPMVS_OK;
    };
    (void)total;
    f->n_points = 0;
    { int rc0 = reserve(2 * max_wh); if (rc0) return rc0; }
    struct Temps {                       // freed on every return path
        FusedPoint* tmp = nullptr;
        unsigned char* keep = nullptr;
        int* d_src = nullptr;
        size_t* d_num = nullptr;
        void* d_cub = nullptr;
        ~Temps() { cudaFree(tmp); cudaFree(keep); cudaFree(d_src); cudaFree(d_num); cudaFree(d_cub); }
    } T;
    FusedPoint*& tmp = T.tmp;
    unsigned char*& keep = T.keep;
    int*& d_src = T.d_src;
    size_t*& d_num = T.d_num;
    void*& d_cub = T.d_cub;
    size_t cub_bytes = 0;
    FCK(cudaMalloc((void**)&tmp, max_wh * sizeof(FusedPoint)));
    FCK(cudaMalloc((void**)&keep, max_wh));
    FCK(cudaMalloc((void**)&d_src, sizeof(int) * (size_t)f->n * max_list));
    FCK(cudaMalloc((void**)&d_num, sizeof(size_t)));
    FCK(cudaMemcpy(d_src, src_lists, sizeof(int) * (size_t)f->n * max_list, cudaMemcpyHostToDevice));
    cub::DeviceSelect::Flagged(nullptr, cub_bytes, tmp, keep, f->d_points, d_num, (int)max_wh);
    FCK(cudaMalloc(&d_cub, cub_bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int rc = MPMVS_OK;
    for (int i = 0; i < f->n && rc == MPMVS_OK; ++i) {
        const FusionView& v = f->hviews[i];
        const int* row = src_lists + (size_t)i * max_list;
        if (!v.W || row[0] < 0) continue;                 // not estimated
        int num_ngb = 0;
        while (num_ngb < max_list && row[num_ngb] != -2) ++num_ngb;
        const size_t wh = (size_t)v.W * v.H;
        rc = reserve(f->n_points + wh);               // worst case: every pixel of this image fuses
        if (rc) break;
        if (v.sky) pm_mask_sky_kernel<<<(unsigned)((wh + 255) / 256), 256>>>(v.mask_prev, v.mask_next, v.sky, wh);
        pm_fuse_kernel<<<dim3((v.W + 31) / 32, (v.H + 7) / 8), dim3(32, 8)>>>(f->dviews, i, d_src + (size_t)i * max_list, num_ngb,
                                                                             use_dynamic_consistency ? 1 : 0, tmp, keep);
        cub::DeviceSelect::Flagged(d_cub, cub_bytes, tmp, keep, f->d_points + f->n_points, d_num, (int)wh);
        size_t got = 0;
        cudaError_t e = cudaMemcpy(&got, d_num, sizeof(got), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = (int)e; break; }
        f->n_points += got;
        for (int j = 1; j < num_ngb; ++j) {              // publish the masks this image set (cpp:467-471) before the next image
            const int s = row[j];
            if (s < 0 || !f->hviews[s].W) continue;
            const size_t swh = (size_t)f->hviews[s].W * f->hviews[s].H;
            pm_mask_merge_kernel<<<(unsigned)((swh + 255) / 256), 256>>>(f->hviews[s].mask_prev, f->hviews[s].mask_next, swh);
        }
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc) return rc;
    FCK(cudaGetLastError());
    if (n_points) *n_points = f->n_points;
    if (ms) *ms = t;
    return MPMVS_OK;
}

// points: n x 9 floats (x y z nx ny nz c0 c1 c2), the PointList layout of PatchMatch.h:29-33
int mpmvs_fusion_get_points(mpmvs_fusion* f, float* points9_host, uint64_t capacity) {
    if (!f || !points9_host || capacity < f->n_points) return MPMVS_E_ARG;
    FCK(cudaSetDevice(f->device));
    FCK(cudaMemcpy(points9_host, f->d_points, f->n_points * sizeof(FusedPoint), cudaMemcpyDeviceToHost));
    return MPMVS_OK;
}

}  // extern "C"
