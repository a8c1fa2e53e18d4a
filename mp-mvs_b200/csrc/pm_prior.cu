// pm_prior.cu -- device side of the planar-prior stage that ProcessProblem runs on the host between two Run()s
// (/root/reference/src/PatchMatch.cpp:532-609, "cpp:NNN" below): vertex picking, triangle rasterisation to an id mask,
// 3-point plane fit, depth-range check. Only the triangulation itself stays on the host (pm_subdiv.h).
// Working on the state that the first Run() left in HBM avoids the reference's round trip: D2H of planes/costs, a
// single-threaded host loop over every 5x5 cell, ~0.5 M cv::SVD::solveZ calls and a per-pixel H2D of float4 planes.
#include <cuda_runtime.h>

#include "pm_core.cuh"
#include "pm_kernels.h"

namespace {

// GetTriangulateVertices, cpp:782-853. One thread per 5x5 cell; out: up to 3 (x, y) per cell + count, in the order the
// reference pushes them.
__global__ void __launch_bounds__(128) pm_pick_vertices_kernel(const float* costs, const float* geom, int W, int H, int cells_x, int cells_y,
                                                               int geom_variant, short2* out_xy, unsigned char* out_n) {
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
    if (cx >= cells_x || cy >= cells_y) return;
    const int col = cx * 5, row = cy * 5;
    const int c_bound = min(W, col + 5), r_bound = min(H, row + 5);
    const int cell = cy * cells_x + cx;
    if (!geom_variant) {
        float min_cost = 2.0f;
        int bx = 0, by = 0;
        for (int r = row; r < r_bound; ++r)
            for (int c = col; c < c_bound; ++c) {
                const float cost = costs[r * W + c];
                if (cost < 2.0f && min_cost > cost) { bx = c; by = r; min_cost = cost; }
            }
        const bool ok = min_cost < 0.1f;
        out_n[cell] = ok ? 1 : 0;
        out_xy[cell * 3] = make_short2((short)bx, (short)by);
    } else {
        float mc[3] = {2.0f, 2.0f, 2.0f};
        short2 pt[3] = {make_short2(0, 0), make_short2(0, 0), make_short2(0, 0)};
        float cost_sum = 0.0f;
        for (int r = row; r < r_bound; ++r)
            for (int c = col; c < c_bound; ++c) {
                const float cost = costs[r * W + c];
                cost_sum += cost;
                if (cost < 1.0f && geom[r * W + c] < 0.4f && cost < mc[2]) {
                    mc[2] = cost;
                    pt[2] = make_short2((short)c, (short)r);
                    for (int i = 1; i >= 0; --i) {
                        if (mc[i] <= mc[i + 1]) break;
                        const float t = mc[i + 1]; mc[i + 1] = mc[i]; mc[i] = t;
                        const short2 tp = pt[i + 1]; pt[i + 1] = pt[i]; pt[i] = tp;
                    }
                }
            }
        // QUIRK cpp:841: the cell sum is divided by r_bound * c_bound (absolute coordinates), not by the cell area
        cost_sum = (float)((double)(cost_sum / (float)(r_bound * c_bound)) * 0.85);
        const float thresh = fmaxf(cost_sum, 0.2f);
        int n = 0;
        for (int i = 0; i < 3; ++i) {
            if (mc[i] < thresh) { out_xy[cell * 3 + n] = pt[i]; ++n; }
            else break;
        }
        out_n[cell] = (unsigned char)n;
    }
}

// GetPriorPlaneParams (cpp:723-755) + the rasterisation loop of ProcessProblem (cpp:554-579), one thread per triangle.
// The plane through the three back-projected vertices replaces cv::SVD::solveZ of the 3x4 system (same null space when
// the points are not collinear; collinear points give a NaN plane, which the range check below discards).
__global__ void __launch_bounds__(128) pm_triangle_kernel(const PmFrame F, const pm_f4* planes, const int2* vxy, const int3* tris, int n_tris,
                                                          pm_f4* tri_planes, unsigned int* mask) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const int3 tr = tris[t];
    const int2 p1 = vxy[tr.x], p2 = vxy[tr.y], p3 = vxy[tr.z];
    double X[3][3];
    const int2 pp[3] = {p1, p2, p3};
    for (int k = 0; k < 3; ++k) {   // Get3DPointonRefCam, cpp:200-209 (float arithmetic)
        const float depth = planes[pp[k].y * F.W + pp[k].x].w;
        X[k][0] = depth * (pp[k].x - F.cx) / F.fx;
        X[k][1] = depth * (pp[k].y - F.cy) / F.fy;
        X[k][2] = depth;
    }
    const double ux = X[1][0] - X[0][0], uy = X[1][1] - X[0][1], uz = X[1][2] - X[0][2];
    const double vx = X[2][0] - X[0][0], vy = X[2][1] - X[0][1], vz = X[2][2] - X[0][2];
    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    double nn = sqrt(nx * nx + ny * ny + nz * nz);
    double d = -(nx * X[0][0] + ny * X[0][1] + nz * X[0][2]);
    if (d < 0) nn = -nn;           // cpp:746-749: normalise so that the plane distance is positive
    pm_f4 pl;
    pl.x = (float)(nx / nn); pl.y = (float)(ny / nn); pl.z = (float)(nz / nn); pl.w = (float)(d / nn);
    tri_planes[t] = pl;
    // cpp:556-569, arithmetic types as written there: p, q, step float; (1.0 - p - q) double
    const float L01 = (float)sqrt((double)((p1.x - p2.x) * (p1.x - p2.x) + (p1.y - p2.y) * (p1.y - p2.y)));
    const float L02 = (float)sqrt((double)((p1.x - p3.x) * (p1.x - p3.x) + (p1.y - p3.y) * (p1.y - p3.y)));
    const float L12 = (float)sqrt((double)((p2.x - p3.x) * (p2.x - p3.x) + (p2.y - p3.y) * (p2.y - p3.y)));
    const float max_edge = fmaxf(L01, fmaxf(L02, L12));
    const float step = (float)(1.0 / (double)max_edge);
    for (float p = 0; p < 1.0; p = __fadd_rn(p, step)) {
        for (float q = 0; (double)q < __dsub_rn(1.0, (double)p); q = __fadd_rn(q, step)) {
            // no FMA contraction: the host code this restates rounds every product and sum separately (x86-64, no -mfma),
            // and the truncation to int below is sensitive to the last bit at integer boundaries
            const double r = __dsub_rn(__dsub_rn(1.0, (double)p), (double)q);
            const float fx = __fadd_rn(__fmul_rn(p, (float)p1.x), __fmul_rn(q, (float)p2.x));
            const float fy = __fadd_rn(__fmul_rn(p, (float)p1.y), __fmul_rn(q, (float)p2.y));
            const int x = (int)__dadd_rn((double)fx, __dmul_rn(r, (double)p3.x));
            const int y = (int)__dadd_rn((double)fy, __dmul_rn(r, (double)p3.y));
            if (x >= 0 && x < F.W && y >= 0 && y < F.H) atomicMax(&mask[y * F.W + x], (unsigned int)(t + 1));  // later triangle wins
        }
    }
}

// depth-range check (cpp:583-595) + per-pixel expansion (CudaPlanarPriorInitialization, cpp:984-993)
__global__ void __launch_bounds__(256) pm_prior_expand_kernel(const PmFrame F, const pm_f4* tri_planes, unsigned int* mask, pm_f4* prior,
                                                              unsigned int* count) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= F.W || y >= F.H) return;
    const int idx = y * F.W + x;
    const unsigned int m = mask[idx];
    if (m == 0) return;
    const pm_f4 pl = tri_planes[m - 1];
    const float d = pm_depth_from_plane(F, pl, x, y);   // GetDepthFromPlaneParam, cpp:650-653
    if (d <= F.depth_max && d >= F.depth_min) {
        prior[idx] = pl;
        atomicAdd(count, 1u);
    } else {
        mask[idx] = 0;
    }
}

}  // namespace

cudaError_t pm_launch_pick_vertices(const float* costs, const float* geom, int W, int H, int geom_variant, short2* out_xy,
                                    unsigned char* out_n, cudaStream_t st) {
    const int cx = (W + 4) / 5, cy = (H + 4) / 5;
    pm_pick_vertices_kernel<<<dim3((cx + 127) / 128, cy), 128, 0, st>>>(costs, geom, W, H, cx, cy, geom_variant, out_xy, out_n);
    return cudaGetLastError();
}

cudaError_t pm_launch_prior(const PmFrame& F, const pm_f4* planes, const int2* vxy, const int3* tris, int n_tris, pm_f4* tri_planes,
                            unsigned int* mask, pm_f4* prior, unsigned int* count, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(mask, 0, sizeof(unsigned int) * (size_t)F.W * F.H, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(count, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    if (n_tris > 0) pm_triangle_kernel<<<(n_tris + 127) / 128, 128, 0, st>>>(F, planes, vxy, tris, n_tris, tri_planes, mask);
    pm_prior_expand_kernel<<<dim3((F.W + 31) / 32, (F.H + 7) / 8), dim3(32, 8), 0, st>>>(F, tri_planes, mask, prior, count);
    return cudaGetLastError();
}
