// pm_views.h -- host-side preparation of the per-view constants (PmView) and the frame constants (PmFrame).
// Plain C++ (no CUDA), double precision; used by pm_capi.cu and by the test-only host emulation.
#ifndef MPMVS_PM_VIEWS_H
#define MPMVS_PM_VIEWS_H
#include "../../include/mpmvs_b200.h"
#include <float.h>
#include <math.h>

#include "pm_core.cuh"

namespace pmv {
// ---- small double-precision 3x3 helpers for the per-view constants
struct M3 { double m[9]; };
inline M3 mul(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}
inline M3 transpose(const M3& a) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * j + i];
    return r;
}
inline void mulv(const M3& a, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = a.m[3 * i] * v[0] + a.m[3 * i + 1] * v[1] + a.m[3 * i + 2] * v[2];
}
inline M3 from(const float* f) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = f[i];
    return r;
}
// the zero-skew inverse intrinsics the reference uses when it back-projects (cu:163-168, 260-268, 587-588)
inline M3 kinv_zero_skew(const float* K) {
    M3 r = {{1.0 / K[0], 0, -(double)K[2] / K[0], 0, 1.0 / K[4], -(double)K[5] / K[4], 0, 0, 1}};
    return r;
}
// the part of K_s the homography uses (cu:270-278): fx, cx, fy, cy and K[8]
inline M3 k_homography(const float* K) {
    M3 r = {{K[0], 0, K[2], 0, K[4], K[5], 0, 0, K[8]}};
    return r;
}

}  // namespace pmv

// The hypothesis-invariant head of ComputeHomography, cu:233-247 (exact arithmetic): R_relative = R_s R_r^T and
// t_relative = R_s (C_r - C_s) with the roundings of the reference's compiled kernels. nvcc turns its `a*b + c*d + e*f`
// into FFMA(e, f, FFMA(a, b, FMUL(c, d))) -- the MIDDLE product is the rounded one (SASS of oracle/_ref, RefNccMap and
// BlackPixelUpdate) -- all with .FTZ. In the library (this header compiled by nvcc's host pass) the same operations are
// done with fmaf and flush-to-zero, so the values are the bits the reference forms on the device; in the test-only host
// emulation (plain g++) it is the oracle's unfused left-to-right arithmetic, like everything else there.
namespace pmv {
inline float ftz(float x) { return (x > -FLT_MIN && x < FLT_MIN) ? (x < 0 ? -0.0f : 0.0f) : x; }
inline float dot3_ref(float a, float b, float c, float d, float e, float f) {
#if defined(__CUDACC__)
    a = ftz(a); b = ftz(b); c = ftz(c); d = ftz(d); e = ftz(e); f = ftz(f);
    const float m = ftz(c * d);
    return ftz(fmaf(e, f, ftz(fmaf(a, b, m))));
#else
    return a * b + c * d + e * f;
#endif
}
}  // namespace pmv
inline void pm_view_prep(const mpmvs_camera& rc, PmView& V) {
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            V.Rrel[3 * a + b] = pmv::dot3_ref(V.sR[3 * a], rc.R[3 * b], V.sR[3 * a + 1], rc.R[3 * b + 1], V.sR[3 * a + 2], rc.R[3 * b + 2]);
#if defined(__CUDACC__)
    const float c0 = pmv::ftz(pmv::ftz(rc.C[0]) - pmv::ftz(V.sC[0])), c1 = pmv::ftz(pmv::ftz(rc.C[1]) - pmv::ftz(V.sC[1])),
                c2 = pmv::ftz(pmv::ftz(rc.C[2]) - pmv::ftz(V.sC[2]));
#else
    const float c0 = rc.C[0] - V.sC[0], c1 = rc.C[1] - V.sC[1], c2 = rc.C[2] - V.sC[2];
#endif
    for (int a = 0; a < 3; ++a) V.trel[a] = pmv::dot3_ref(V.sR[3 * a], c0, V.sR[3 * a + 1], c1, V.sR[3 * a + 2], c2);
}

inline void pm_build_view_consts(const mpmvs_camera& rc, const mpmvs_camera& sc, PmView& V) {
    using namespace pmv;
    const M3 Rr = from(rc.R), Rs = from(sc.R);
    const M3 Rrel = mul(Rs, transpose(Rr));                  // R_s R_r^T          (cu:233-241)
    const double Crel[3] = {(double)rc.C[0] - sc.C[0], (double)rc.C[1] - sc.C[1], (double)rc.C[2] - sc.C[2]};
    double trel[3];
    mulv(Rs, Crel, trel);                                    // R_s (C_r - C_s)    (cu:242-247)
    const M3 Kri = kinv_zero_skew(rc.K), Ksh = k_homography(sc.K);
    const M3 A = mul(Ksh, mul(Rrel, Kri));
    double b[3];
    mulv(Ksh, trel, b);
    for (int i = 0; i < 9; ++i) V.A[i] = (float)A.m[i];
    for (int i = 0; i < 3; ++i) V.b[i] = (float)(-b[i]);
    V.w = (float)sc.width;
    V.h = (float)sc.height;
    // geometric consistency: forward  x_s ~ K_s (R_s (R_r^T (z K_r^-1 p) + C_r) + t_s)   (cu:582-615, full K in ProjectPoint)
    const M3 Ksf = from(sc.K), Krf = from(rc.K);
    const M3 Mf = mul(Ksf, mul(Rrel, Kri));
    double tmp[3], Cr[3] = {rc.C[0], rc.C[1], rc.C[2]}, Cs[3] = {sc.C[0], sc.C[1], sc.C[2]}, vf[3], vb[3];
    mulv(Rs, Cr, tmp);
    for (int i = 0; i < 3; ++i) tmp[i] += sc.t[i];
    mulv(Ksf, tmp, vf);
    // backward  x_r ~ K_r (R_r (R_s^T (z_s K_s^-1 q) + C_s) + t_r)
    const M3 Mb = mul(Krf, mul(transpose(Rrel), kinv_zero_skew(sc.K)));
    mulv(Rr, Cs, tmp);
    for (int i = 0; i < 3; ++i) tmp[i] += rc.t[i];
    mulv(Krf, tmp, vb);
    for (int i = 0; i < 9; ++i) { V.Mf[i] = (float)Mf.m[i]; V.Mb[i] = (float)Mb.m[i]; }
    for (int i = 0; i < 3; ++i) { V.vf[i] = (float)vf[i]; V.vb[i] = (float)vb[i]; }
    V.dw = sc.width;
    V.dh = sc.height;
    V.dpitch = sc.width;
    V.depth = nullptr;
    for (int i = 0; i < 9; ++i) { V.sK[i] = sc.K[i]; V.sR[i] = sc.R[i]; }
    for (int i = 0; i < 3; ++i) { V.st[i] = sc.t[i]; V.sC[i] = sc.C[i]; }
    pm_view_prep(rc, V);
}


inline PmFrame pm_make_frame(const mpmvs_camera& c, int n, float depth_min, float depth_max, float sigma_spatial,
                             float sigma_color, int top_k, bool geom, bool planar) {
    PmFrame F{};
    F.fx = c.K[0]; F.fy = c.K[4]; F.cx = c.K[2]; F.cy = c.K[5];
    F.ifx = 1.0f / c.K[0]; F.ify = 1.0f / c.K[4];
    for (int i = 0; i < 9; ++i) F.R[i] = c.R[i];
    F.depth_min = depth_min; F.depth_max = depth_max;
    const double log2e = 1.4426950408889634;
    F.spat_k = (float)(-log2e / (2.0 * sigma_spatial * sigma_spatial));
    F.col_k = (float)(log2e / (2.0 * sigma_color * sigma_color));
    F.W = c.width; F.H = c.height; F.nsrc = n - 1; F.top_k = top_k;
    F.geom = geom ? 1 : 0; F.planar = planar ? 1 : 0;
    for (int i = 0; i < 9; ++i) F.K[i] = c.K[i];
    for (int i = 0; i < 3; ++i) { F.t[i] = c.t[i]; F.C[i] = c.C[i]; }
    F.one = 1;
    F.sigma_spatial = sigma_spatial; F.sigma_color = sigma_color;
    {   // host values (libm); the CUDA library overwrites them with the device's own (pm_capi.cu: mpmvs_create)
        float t[20];
        pm_literal_table(1.0f, sigma_spatial, sigma_color, t);
        for (int k = 0; k < 18; ++k) F.lit_sd[k / 6][k % 6] = t[k];
        F.lit_rcp_spatial = t[18]; F.lit_rcp_color = t[19];
    }
    return F;
}
#endif
