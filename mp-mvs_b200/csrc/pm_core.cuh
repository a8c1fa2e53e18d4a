// pm_core.cuh -- per-pixel arithmetic of the PatchMatch path, written once for the sm_100a kernels.
//
// Behavioural spec: /root/reference/src/PatchMatch.cu ("cu:NNN" below). This is not a transcription:
//   * the reference side of the NCC -- the weighted reference mean/variance of the window -- depends only on (pixel,
//     scale) and is computed once per pixel per sweep (pm_ref_stats), not once per (hypothesis, view) as the reference
//     does (cu:355-414);
//   * the reference window comes from a shared-memory tile (Ctx::ref), source samples from the texture
//     unit (Ctx::src), source depths from plain global loads (Ctx::src_depth);
//   * views with zero sampling weight are skipped where the reference multiplies their cost by 0
//     (cu:684-694, 904-913): result-identical, costs are always finite;
//   * per-pixel scratch that the reference keeps in 1.5 KB of local arrays is reduced to the 8 x nsrc
//     candidate cost table; view weights are 4-bit counters packed in two 64-bit registers.
// All reference quirks that change results are kept and marked QUIRK.
//
// TWO ARITHMETICS, one library. The file is compiled twice into libmpmvs_b200.so (pm_kernels.cu with and without
// -DPM_EXACT=1; everything below the types lives in a namespace named after the arithmetic, so the two sets of
// kernels and launchers link side by side) and a handle picks one at run time (mpmvs_set_arithmetic):
//   * exact (PM_EXACT=1, the default of the library): every floating-point operation of the reference's compiled
//     kernels in the reference's order with the reference's roundings -- which products are fused into FMAs is written
//     down explicitly (pm_rmul / pm_radd / pm_ffma), MUFU.RCP / MUFU.SQRT / MUFU.EX2 where its SASS has them. Results
//     are BIT-IDENTICAL to the reference's kernels (tests/test_zz_fidelity_build_gpu.py). What is still hoisted is
//     bit-safe: the reference-side sums, the weights (pure functions of the reference window), the relative pose
//     R_s R_r^T, R_s (C_r - C_s) (hypothesis-invariant, formed once per problem by the same instructions);
//   * fast (PM_EXACT=0): hoisted homography H = A_v + b_v m^T (per-view constants prepared on the host in double
//     precision; a warped tap costs 5 FMA + 1 RCP), folded weight constants, rsqrt. Statistically equal results.
//
// The functions are templates over a context `Ctx` so that the very same code can be instantiated by
// the test-only host emulation (tests/emul), which is how the logic is debugged without a GPU. The
// product library only ever instantiates the device context in pm_kernels.cu.
#ifndef MPMVS_PM_CORE_CUH
#define MPMVS_PM_CORE_CUH

#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#define PM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PM_HD inline
#define PM_HD_NOINLINE
#endif

#define PM_MAX_SRC 32
#ifndef PM_UNIFORM_VIEWS
#define PM_UNIFORM_VIEWS 1   // profiles/r01_variant_sweep_*: 5-8 % faster once the cost table is in shared memory
#endif
#ifndef PM_EARLY_OUT
#define PM_EARLY_OUT 1   // stop scoring a refinement proposal once it can no longer be accepted (result-identical)
#endif
#ifndef PM_WTAB
// 1: the 36 bilateral weights of a pixel's window are computed once per launch into a per-thread shared-memory table (Ctx::wt)
// instead of once per tap of every NCC (result-identical: they depend on the reference window only). Measured on B200 and NOT
// adopted: the table's 18 KB per block come out of the L1/TEX cache the source fetches live on, and the photometric Run() of
// the exact arithmetic takes 404 ms with it against 354 ms without (profiles/r02_wtab_experiment.md).
#define PM_WTAB 0
#endif
#ifndef PM_PRIOR_EARLY_OUT
// The same for proposals scored under the planar prior (cu:696-711): their acceptance test `exp(-tc^2 / beta) * prior >
// restricted_cost` is monotone in the running cost too -- every step of the chain (FMUL by a positive constant, FMUL(t, -t)
// for t >= 0, MUFU.EX2, FMUL by prior >= 0.5) is non-increasing in tc, and rounding keeps weak monotonicity -- so a proposal
// whose PARTIAL cost already fails it can never pass. MUFU.EX2's monotonicity over the whole argument range is a GPU test
// (mpmvs_selftest_ex2_monotone, tests/test_zz_fidelity_build_gpu.py). A NaN prior fails at the first view, as it does at the end.
#define PM_PRIOR_EARLY_OUT 1
#endif
#ifndef PM_PACKED
// 1 (exact arithmetic, device): the tap loop of the NCC issues Blackwell's packed FP32 instructions (FFMA2 / FADD2 / FMUL2,
// PTX fma/add/mul.rn.ftz.f32x2, new with sm_100): the numerators and coordinates of two neighbouring taps share one
// instruction per operation, and so do the two products (w s, w r) and the two accumulations (sum w s s, sum w r s) of a tap.
// Each lane of a packed instruction is the same IEEE operation with the same rounding and flush-to-zero as the scalar
// instruction it replaces, so the results stay the reference's bits; what changes is the issue-slot count per tap (the exact
// kernel is limited by instruction issue and the texture pipe about equally, profiles/r02_ncu_sweep_exact_v2.txt).
#define PM_PACKED 1
#endif
#ifndef PM_VIEW_MAJOR
// 1: the 8 x (N-1) candidate costs of a pixel are evaluated view by view (all candidates against source 0, then source 1,
// ...) instead of candidate by candidate: eight consecutive NCCs of a block then sample the SAME source image, whose warped
// window stays in the L1/TEX cache. Result-identical (the costs are independent of one another).
#define PM_VIEW_MAJOR 0
#endif
#ifndef PM_NCC_NOINLINE
#define PM_NCC_NOINLINE 0   // 1: the 36-tap NCC is a real function (one copy in the instruction cache however many call sites)
#endif
#ifndef PM_EXACT
#define PM_EXACT 0
#endif
#if PM_EXACT
#define PM_ARITH_NS pm_exact
#else
#define PM_ARITH_NS pm_fast
#endif
#define PM_PI_F 3.14159265358979323846f
// The reference multiplies by M_PI, a double literal: `perturbation * M_PI` (cu:671), `3 * perturbation * M_PI` (cu:559)
// and `M_PI * (5.0f / 180.0f)` (cu:651,929) are double products narrowed to float -- 0.06283185 and 0.08726646, one
// ulp below the float products 0.06283186 and 0.08726647. Compile-time constants either way.
#define PM_PI_D 3.14159265358979323846

#ifndef MPMVS_PM_TYPES
#define MPMVS_PM_TYPES
struct alignas(16) pm_f4 { float x, y, z, w; };

// Per-source-view constants (pm_views.h: pm_build_view_consts, host, double precision; Rrel / trel: pm_view_prep).
struct PmView {
    float A[9];   // K_s R_rel K_r^-1        (fast arithmetic: homography, cu:228-279 with n/d factored out)
    float b[3];   // -K_s t_rel
    float w, h;   // source image size as float (bounds test of the warped centre, cu:351)
    float Mf[9];  // K_s R_s R_r^T K_r^-1     ref pixel (x,y,1)*z -> source homogeneous pixel (cu:582-615)
    float vf[3];  // K_s (R_s C_r + t_s)
    float Mb[9];  // K_r R_r R_s^T K_s^-1     source pixel*z_s -> reference homogeneous pixel
    float vb[3];  // K_r (R_r C_s + t_r)
    int dw, dh;   // source depth-map size
    int dpitch;   // in floats
    int layer;    // layer of this view in the resident layered image texture
    const float* depth;        // source depth map (geom pass), linear device memory
    // exact arithmetic (the layout is the same for both arithmetics of the library):
    float sK[9], sR[9], st[3], sC[3];   // the source camera as the reference holds it (struct Camera, PatchMatch.h:35-46)
    float Rrel[9], trel[3];             // R_s R_r^T and R_s (C_r - C_s) with the reference's roundings (cu:233-247), formed
                                        // once per problem with the reference's instructions (pm_view_prep): hypothesis-invariant
};

// Reference-camera constants and run flags.
struct PmFrame {
    float fx, fy, cx, cy;      // K[0], K[4], K[2], K[5] of the reference view
    float ifx, ify;            // 1/fx, 1/fy
    float R[9];                // reference rotation (cu:89-97, 308-316)
    float depth_min, depth_max;
    float spat_k;              // -log2(e) / (2 sigma_spatial^2)
    float col_k;               //  log2(e) / (2 sigma_color^2)
    int W, H, nsrc, top_k;
    int geom, planar;
    int ref_layer;             // layer of the reference image
    float tex_scale;           // 255 when the views are stored as 8-bit UNORM texels (fetch returns v/255), else 1
    int soft_clamp;            // 1 when the views differ in size: clamp-to-edge is then done on the coordinates
    unsigned long long tex;    // cudaTextureObject_t of the layered image array (one handle, warp-uniform:
                               // a per-view handle makes the compiler serialise every fetch per unique handle)
    // exact arithmetic:
    float K[9], t[3], C[3];    // the rest of the reference camera (R is above): ComputeHomography, BackProjectPoint2W, ProjectPoint
    int one;                   // 1, as a run-time value (windows other than 6 x 6: keeps the tap offsets out of constant folding)
    float sigma_spatial, sigma_color;
    // sqrt(i^2 + j^2) of the six tap classes at the three scales and the two reciprocals of the weight, as the DEVICE
    // evaluates them under --use_fast_math (pm_literal_table; the host emulation fills in libm's values)
    float lit_sd[3][6];
    float lit_rcp_spatial, lit_rcp_color;
};

// Device-resident per-pixel state (SoA).
struct PmState {
    pm_f4* planes;             // (n, d) camera frame while sweeping; (world n, depth) after finalize
    float* costs;
    uint32_t* views;           // selected-view bit mask
    uint32_t* rng;             // 6 words per pixel: d, v0..v4 (XORWOW)
    float* geom;
    const pm_f4* prior;
    const uint32_t* mask;
    unsigned long long* counters;  // optional (profiling): [0] += NCC evaluations that ran their 36 taps
};
#endif  // MPMVS_PM_TYPES

namespace PM_ARITH_NS {
// ------------------------------------------------------------------------------------------------ math
PM_HD float pm_ex2(float x) { return exp2f(x); }
#if defined(__CUDA_ARCH__)
PM_HD float pm_exp(float x) { return __expf(x); }
PM_HD float pm_rsqrt(float x) { return rsqrtf(x); }
PM_HD int pm_ffs(uint32_t m) { return __ffs((int)m); }
PM_HD int pm_f2i(float f) { return __float2int_rz(f); }  // saturating, NaN -> 0 (what `(int)` compiles to)
#else
PM_HD float pm_exp(float x) { return expf(x); }
PM_HD float pm_rsqrt(float x) { return 1.0f / sqrtf(x); }
PM_HD int pm_ffs(uint32_t m) { return __builtin_ffs((int)m); }
PM_HD int pm_f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
#endif

PM_HD void pm_count(unsigned long long* ctr, uint32_t n) {
#if defined(__CUDA_ARCH__)
    atomicAdd(ctr, (unsigned long long)n);
#else
    *ctr += n;
#endif
}

#if PM_EXACT
// Pinned roundings (exact arithmetic): on the device exactly one instruction each, never contracted with a neighbour.
#if defined(__CUDA_ARCH__)
PM_HD float pm_rmul(float a, float b) { return __fmul_rn(a, b); }
PM_HD float pm_radd(float a, float b) { return __fadd_rn(a, b); }
PM_HD float pm_ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// MUFU.RCP / MUFU.SQRT as such: `1.0f / sqrtf(x)` would be turned into one MUFU.RSQ, which the reference does not execute
__device__ __forceinline__ float pm_rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float pm_sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// two floats in an aligned register pair and the packed FP32 instructions of sm_100 on them (FFMA2 / FADD2 / FMUL2). NOTE:
// ptxas contracts a packed mul feeding a packed add into one FFMA2 even with .rn on both, so no pm_mul2 result may be the
// addend-side input of a pm_add2 (none is: tests/test_sass_equivalence.py counts the instructions).
struct pm_f2 { unsigned long long v; };
__device__ __forceinline__ pm_f2 pm_pk(float lo, float hi) { pm_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void pm_upk(pm_f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ pm_f2 pm_fma2(pm_f2 a, pm_f2 b, pm_f2 c) { pm_f2 r; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ pm_f2 pm_mul2(pm_f2 a, pm_f2 b) { pm_f2 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pm_f2 pm_add2(pm_f2 a, pm_f2 b) { pm_f2 r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
#else
PM_HD float pm_rmul(float a, float b) { return a * b; }
PM_HD float pm_radd(float a, float b) { return a + b; }
PM_HD float pm_ffma(float a, float b, float c) { return a * b + c; }   // the oracle is plain C: every product rounded
#endif
#endif
// i^2 + j^2 of a tap in units of (step/2)^2 -> its class 0..5: 2, 10, 18, 26, 34, 50
PM_HD constexpr int pm_tap_class(int u, int v) {
    return (u * u + v * v == 2) ? 0 : (u * u + v * v == 10) ? 1 : (u * u + v * v == 18) ? 2 : (u * u + v * v == 26) ? 3 : (u * u + v * v == 34) ? 4 : 5;
}
// The table behind PmFrame::lit_sd / lit_rcp_*: 18 square roots and 2 reciprocals, evaluated wherever this is compiled
// (device: MUFU.SQRT / MUFU.RCP under --use_fast_math, as in the reference's kernels; `one` = 1.0f at run time keeps the
// compiler from folding them).
PM_HD void pm_literal_table(float one, float sigma_spatial, float sigma_color, float* out20) {
    const int cls[6] = {2, 10, 18, 26, 34, 50};
    for (int s = 0; s < 3; ++s)
        for (int k = 0; k < 6; ++k) out20[s * 6 + k] = sqrtf((float)(cls[k] << (2 * s)) * one);
    out20[18] = one / (2.0f * sigma_spatial * sigma_spatial);
    out20[19] = one / (2.0f * sigma_color * sigma_color);
}

// ------------------------------------------------------------------------------------------------ RNG
// XORWOW with cuRAND's start state for curand_init(seed, 0, 0) and curand_uniform's output map, so a
// pixel draws the same stream as the fixed-seed reference build (oracle/ref_harness.cu).
struct PmRng { uint32_t d, v0, v1, v2, v3, v4; };

PM_HD uint64_t pm_mix_seed(uint64_t seed, uint32_t x, uint32_t y) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)y) << 32) | (uint64_t)x) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
PM_HD void pm_rng_init(PmRng& s, uint64_t seed) {
    const uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s.d = 6615241u + t1 + t0;
    s.v0 = 123456789u + t0; s.v1 = 362436069u ^ t0; s.v2 = 521288629u + t1; s.v3 = 88675123u ^ t1; s.v4 = 5783321u + t0;
}
PM_HD float pm_uniform(PmRng& s) {  // (0, 1]
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return fmaf((float)(s.v4 + s.d), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}
PM_HD PmRng pm_rng_load(const uint32_t* g, int idx) {
    const uint32_t* p = g + 6 * (size_t)idx;
    PmRng s; s.d = p[0]; s.v0 = p[1]; s.v1 = p[2]; s.v2 = p[3]; s.v3 = p[4]; s.v4 = p[5];
    return s;
}
PM_HD void pm_rng_store(uint32_t* g, int idx, const PmRng& s) {
    uint32_t* p = g + 6 * (size_t)idx;
    p[0] = s.d; p[1] = s.v0; p[2] = s.v1; p[3] = s.v2; p[4] = s.v3; p[5] = s.v4;
}

// ------------------------------------------------------------------------------------------------ geometry
// Vec3DotVec3, cu:9-12. Exact arithmetic: nvcc contracts the reference's `a.x*b.x + a.y*b.y + a.z*b.z` to
// FFMA(az, bz, FFMA(ax, bx, FMUL(ay, by))) in every inlined copy (the MIDDLE product is the rounded one: SASS of oracle/_ref,
// all ten acos sites of Black/RedPixelUpdate). Left to the compiler OUR copies do not all come out that way (the one in the
// refinement rounded the first product), and one ulp in |n|^2 decides whether acos(n . n) is 0 or NaN where a hypothesis
// carries the prior's own normal -- 0.015 % of the pixels of a full-size planar-prior sweep. Pinned.
PM_HD float pm_dot3(const pm_f4& a, const pm_f4& b) {
#if PM_EXACT && defined(__CUDA_ARCH__)
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
#else
    return a.x * b.x + a.y * b.y + a.z * b.z;
#endif
}

// ComputeDepthfromPlaneHypothesis, cu:84-87.
// K[0] / K[4] is formed on the device by the reference: under --use_fast_math an approximate division (x * rcp(x) is
// not always 1 when the two focal lengths are equal), so it is formed here and not on the host.
//
// Exact arithmetic: nvcc contracts the reference's three-term denominator in TWO ways depending on the inlined copy
// (SASS of oracle/_ref, BlackPixelUpdate / RedPixelUpdate; tests/tools/sass_expr.py compares them):
//   form 0, every copy but one:            FFMA(fx, nz, FFMA(nx, x - cx,  FMUL(ny, yt)))     yt = (fx * RCP(fy)) * (y - cy)
//   form 1, ComputeGeomConsistencyCost's copy inside PlaneHypothesisRefinement (cu:687):
//                                          FFMA(fx, nz, FFMA(ny, yt,      FMUL(x - cx, nx)))
// and depth = FMUL(FMUL(fx, -w), RCP(den)); where the reference subtracts the prior's depth from a depth that has no
// other use (cu:939-940; not cu:949-950, whose depth_now is the outer one) the last product is fused into the subtraction: FFMA(fx * -w, RCP(den), -depth_prior).
// Left to the compiler, which copy gets which form follows OUR inlining, not the reference's -- so the device code pins
// every operation. Host (emulation, tests only): the oracle's plain expression, where all forms coincide.
#if PM_EXACT && defined(__CUDA_ARCH__)
PM_HD float pm_depth_den(const PmFrame& F, const pm_f4& pl, int x, int y, int form) {
    const float xt = __fadd_rn((float)x, -F.cx);
    const float yt = __fmul_rn(__fmul_rn(F.fx, pm_rcp_approx(F.fy)), __fadd_rn((float)y, -F.cy));
    const float inner = form ? __fmaf_rn(pl.y, yt, __fmul_rn(xt, pl.x)) : __fmaf_rn(pl.x, xt, __fmul_rn(pl.y, yt));
    return __fmaf_rn(F.fx, pl.z, inner);
}
PM_HD float pm_depth_from_plane(const PmFrame& F, const pm_f4& pl, int x, int y, int form = 0) {
    return __fmul_rn(__fmul_rn(F.fx, -pl.w), pm_rcp_approx(pm_depth_den(F, pl, x, y, form)));
}
PM_HD float pm_depth_minus(const PmFrame& F, const pm_f4& pl, int x, int y, float sub) {
    return __fmaf_rn(__fmul_rn(F.fx, -pl.w), pm_rcp_approx(pm_depth_den(F, pl, x, y, 0)), -sub);
}
#else
PM_HD float pm_depth_from_plane(const PmFrame& F, const pm_f4& pl, int x, int y, int form = 0) {
    (void)form;
    return -pl.w * F.fx / ((x - F.cx) * pl.x + (F.fx / F.fy) * (y - F.cy) * pl.y + F.fx * pl.z);
}
PM_HD float pm_depth_minus(const PmFrame& F, const pm_f4& pl, int x, int y, float sub) {
    return pm_depth_from_plane(F, pl, x, y) - sub;
}
#endif
// GetPlane2Origin, cu:163-176
PM_HD float pm_plane_distance(const PmFrame& F, int x, int y, float depth, const pm_f4& n) {
    const float X0 = depth * (x - F.cx) / F.fx, X1 = depth * (y - F.cy) / F.fy;
    return -(n.x * X0 + n.y * X1 + n.z * depth);
}
PM_HD void pm_normalize(pm_f4& v) {  // cu:188-195
    const float inv = pm_rsqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    v.x *= inv; v.y *= inv; v.z *= inv;
}
// GenerateRandomNormal, cu:197-219 (Marsaglia; flipped to face the camera)
PM_HD pm_f4 pm_random_normal(const PmFrame& F, int x, int y, PmRng& rs) {
    float q1, q2, s;
    do {
        q1 = 2.f * pm_uniform(rs) - 1.f;
        q2 = 2.f * pm_uniform(rs) - 1.f;
        s = q1 * q1 + q2 * q2;
    } while (s >= 1.f);
    const float sq = sqrtf(1.f - s);
    pm_f4 n = {2.0f * q1 * sq, 2.0f * q2 * sq, 1.0f - 2.0f * s, 0.f};
    const float vx = (x - F.cx) / F.fx, vy = (y - F.cy) / F.fy;
    if (n.x * vx + n.y * vy + n.z > 0.0f) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
    pm_normalize(n);
    return n;
}
// GenerateRandomPlaneHypothesis, cu:221-226
PM_HD pm_f4 pm_random_plane(const PmFrame& F, int x, int y, PmRng& rs) {
    pm_f4 pl = pm_random_normal(F, x, y, rs);
    const float depth = pm_uniform(rs) * (F.depth_max - F.depth_min) + F.depth_min;
    pl.w = pm_plane_distance(F, x, y, depth, pl);
    return pl;
}
// GeneratePerturbedNormal, cu:460-495 (three Euler angles; keeps the old normal if the new one faces away)
PM_HD pm_f4 pm_perturbed_normal(const PmFrame& F, int x, int y, const pm_f4& normal, PmRng& rs, float perturbation) {
    const float vx = (x - F.cx) / F.fx, vy = (y - F.cy) / F.fy;
    const float a1 = (pm_uniform(rs) - 0.5f) * perturbation;
    const float a2 = (pm_uniform(rs) - 0.5f) * perturbation;
    const float a3 = (pm_uniform(rs) - 0.5f) * perturbation;
    const float s1 = sinf(a1), s2 = sinf(a2), s3 = sinf(a3), c1 = cosf(a1), c2 = cosf(a2), c3 = cosf(a3);
    pm_f4 o;
    o.x = (c2 * c3) * normal.x + (c3 * s1 * s2 - c1 * s3) * normal.y + (s1 * s3 + c1 * c3 * s2) * normal.z;
    o.y = (c2 * s3) * normal.x + (c1 * c3 + s1 * s2 * s3) * normal.y + (c1 * s2 * s3 - c3 * s1) * normal.z;
    o.z = (-s2) * normal.x + (c2 * s1) * normal.y + (c1 * c2) * normal.z;
    o.w = 0.f;
    if (o.x * vx + o.y * vy + o.z >= 0.0f) return normal;
    pm_normalize(o);
    return o;
}

// ------------------------------------------------------------------------------------------------ NCC
// Spatial tap distance |(i,j)| in units of step/2, i,j in {+-1,+-3,+-5}: one of six constants.
PM_HD constexpr float pm_tap_dist(int u, int v) {
    return (u * u + v * v == 2)    ? 1.41421356237f
           : (u * u + v * v == 10) ? 3.16227766017f
           : (u * u + v * v == 26) ? 5.09901951359f
           : (u * u + v * v == 18) ? 4.24264068712f
           : (u * u + v * v == 34) ? 5.83095189485f
                                   : 7.07106781187f;
}

// General window (NCC microbenchmark, BASELINE.json config 5): TAPS x TAPS taps at offsets (2a - (TAPS-1)) * step/2, i.e. a
// (2*(TAPS-1)+1)-pixel window at step 2. The reference only has TAPS = 6 (cu:342-346,365,373), which keeps its exact constants.
PM_HD constexpr float pm_csqrt(float x) {
    float r = x > 1.f ? x : 1.f;
    for (int i = 0; i < 40; ++i) r = 0.5f * (r + x / r);
    return r;
}
template <int TAPS>
PM_HD constexpr float pm_tap_dist_n(int u, int v) {
    return TAPS == 6 ? pm_tap_dist(u, v) : pm_csqrt((float)(u * u + v * v));
}

// Reference-side statistics of one pixel at one scale: hypothesis-invariant (hoisted out of cu:341-414).
struct PmRefStats {
    float r0;        // centre intensity
    float inv_sw;    // 1 / sum w
    float mean_r;    // sum w r / sum w
    float var_r;     // weighted variance of the reference patch
};

#if PM_EXACT
// ComputeBilateralWeight, cu:318-323, for windows other than the reference's 6 x 6 (NCC microbenchmark only): the offsets
// arrive as run-time values, so the square root is the device's approximate one under --use_fast_math.
PM_HD float pm_weight_literal(float x_dist, float y_dist, float pix, float center_pix, float sigma_spatial, float sigma_color) {
    const float spatial_dist = sqrtf(x_dist * x_dist + y_dist * y_dist);
    const float color_dist = fabsf(pix - center_pix);
    return expf(-spatial_dist / (2.0f * sigma_spatial * sigma_spatial) - color_dist / (2.0f * sigma_color * sigma_color));
}
// The same weight as the reference's SASS evaluates it (RefNccMap / BlackPixelUpdate in oracle/_ref):
//   EX2(1.4426950216 * FFMA(-SQRT(i^2 + j^2), RCP(2 s_s s_s), -(|r - r0| * RCP(2 s_c s_c))))
// with the square root and the reciprocals taken from the device-computed table. Host: the oracle's expression.
template <int SCALE>
PM_HD float pm_weight_pinned(const PmFrame& F, int cls, float r, float r0) {
#if defined(__CUDA_ARCH__)
    const float t = __fmul_rn(fabsf(r - r0), F.lit_rcp_color);
    const float u = __fmaf_rn(-F.lit_sd[SCALE][cls], F.lit_rcp_spatial, -t);
    return exp2f(__fmul_rn(u, 1.4426950216293334961f));
#else
    return expf(-F.lit_sd[SCALE][cls] / (2.0f * F.sigma_spatial * F.sigma_spatial) - fabsf(r - r0) / (2.0f * F.sigma_color * F.sigma_color));
#endif
}
#endif

// The bilateral weight of tap (a, b) of the window (cu:318-323): a pure function of the reference window.
template <int SCALE, int TAPS>
PM_HD float pm_tap_weight(const PmFrame& F, int a, int b, float r, float r0) {
    constexpr int HS = 1 << SCALE;
    const int u = 2 * a - (TAPS - 1), v = 2 * b - (TAPS - 1);
#if PM_EXACT
    // the table holds the reference's 6 x 6 window; other windows (NCC microbenchmark only) take the square root at run time
    return TAPS == 6 ? pm_weight_pinned<SCALE>(F, pm_tap_class(u, v), r, r0)
                     : pm_weight_literal((float)(u * HS * F.one), (float)(v * HS * F.one), r, r0, F.sigma_spatial, F.sigma_color);
#else
    return pm_ex2(fmaf(fabsf(r - r0), -F.col_k, (HS * pm_tap_dist_n<TAPS>(u, v)) * F.spat_k));
#endif
}

// Reference side of ComputeBilateralNCC for one pixel: the TAPS x TAPS weights go to the thread's table c.wt(a * TAPS + b)
// (every NCC of this pixel in this launch reads them back: 14 x (N-1) times in a half-sweep), the weighted sums to PmRefStats.
template <int SCALE, class Ctx, int TAPS = 6>
PM_HD PmRefStats pm_ref_stats(const Ctx& c, const PmFrame& F) {
    constexpr int HS = 1 << SCALE;  // step/2: taps at (2a-5)*HS
    PmRefStats st;
    st.r0 = c.ref(0, 0);
#if PM_EXACT
    // reference-side half of cu:355-405 with pinned roundings: FMUL r*w, FADD sum w, FFMA(r, r*w, .), FADD sum r*w per
    // tap, rows added with FADD; 1/sum through MUFU.RCP, mean by FMUL, variance FFMA(sum rr, inv, -mean^2)
    float sum_ref = 0.0f, sum_ref_ref = 0.0f, weight_sum = 0.0f;
#pragma unroll
    for (int a = 0; a < TAPS; ++a) {
        const int i = (2 * a - (TAPS - 1)) * HS;
        float row_ref = 0.0f, row_ref_ref = 0.0f, row_weight = 0.0f;
#pragma unroll
        for (int b = 0; b < TAPS; ++b) {
            const int j = (2 * b - (TAPS - 1)) * HS;
            const float r = c.ref(i, j);
            const float w = pm_tap_weight<SCALE, TAPS>(F, a, b, r, st.r0);
            if (PM_WTAB) c.wt(a * TAPS + b) = w;
            const float rw = pm_rmul(w, r);
            row_ref = pm_radd(row_ref, rw);
            row_ref_ref = pm_ffma(rw, r, row_ref_ref);
            row_weight = pm_radd(row_weight, w);
        }
        sum_ref = pm_radd(sum_ref, row_ref);
        sum_ref_ref = pm_radd(sum_ref_ref, row_ref_ref);
        weight_sum = pm_radd(weight_sum, row_weight);
    }
    st.inv_sw = 1.0f / weight_sum;
    st.mean_r = pm_rmul(sum_ref, st.inv_sw);
#if defined(__CUDA_ARCH__)
    st.var_r = __fmaf_rn(sum_ref_ref, st.inv_sw, -__fmul_rn(st.mean_r, st.mean_r));
#else
    sum_ref_ref *= st.inv_sw;
    st.var_r = sum_ref_ref - st.mean_r * st.mean_r;
#endif
#else
    float sw = 0.f, swr = 0.f, swrr = 0.f;
#pragma unroll
    for (int a = 0; a < TAPS; ++a) {
#pragma unroll
        for (int b = 0; b < TAPS; ++b) {
            const float r = c.ref((2 * a - (TAPS - 1)) * HS, (2 * b - (TAPS - 1)) * HS);
            const float w = pm_tap_weight<SCALE, TAPS>(F, a, b, r, st.r0);
            if (PM_WTAB) c.wt(a * TAPS + b) = w;
            sw += w;
            swr = fmaf(w, r, swr);
            swrr = fmaf(w * r, r, swrr);
        }
    }
    st.inv_sw = 1.0f / sw;
    st.mean_r = swr * st.inv_sw;
    st.var_r = swrr * st.inv_sw - st.mean_r * st.mean_r;
#endif
    return st;
}

// Per-hypothesis part of the homography. fast: m = K_r^-T n / d, q0 = m . (x, y, 1)  (= -1/z at the pixel);
// exact: the plane itself (the reference rebuilds H from it per view, cu:228-279).
struct PmHyp {
#if PM_EXACT
    pm_f4 pl;
#else
    float mx, my, q0;
#endif
};
PM_HD PmHyp pm_hyp(const PmFrame& F, const pm_f4& pl, int x, int y) {
    PmHyp h;
#if PM_EXACT
    (void)F; (void)x; (void)y;
    h.pl = pl;
#else
    const float id = 1.0f / pl.w;
    h.mx = pl.x * F.ifx * id;
    h.my = pl.y * F.ify * id;
    const float mz = (pl.z - pl.x * F.cx * F.ifx - pl.y * F.cy * F.ify) * id;
    h.q0 = fmaf(h.mx, (float)x, fmaf(h.my, (float)y, mz));
#endif
    return h;
}

#if PM_EXACT
// ComputeHomography, cu:228-279. Its hypothesis-invariant head -- R_relative = R_s R_r^T and t_relative = R_s (C_r - C_s),
// cu:233-247 -- is formed once per problem with the reference's roundings (pm_views.h: pm_view_prep) and arrives as
// PmView::Rrel / trel; the rest, cu:248-279, follows operation by operation (plane term, K_r^-1 with zero skew, the part
// of K_s it uses): plain expressions, so nvcc contracts them the way it contracts the reference's (checked on the SASS by
// tests/test_sass_equivalence.py).
PM_HD void pm_homography_literal(const PmFrame& F, const PmView& V, const pm_f4& pl, float* H) {
    float tmp[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        H[3 * a] = V.Rrel[3 * a] - V.trel[a] * pl.x / pl.w;
        H[3 * a + 1] = V.Rrel[3 * a + 1] - V.trel[a] * pl.y / pl.w;
        H[3 * a + 2] = V.Rrel[3 * a + 2] - V.trel[a] * pl.z / pl.w;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        tmp[3 * a] = H[3 * a] / F.fx;
        tmp[3 * a + 1] = H[3 * a + 1] / F.fy;
        tmp[3 * a + 2] = -H[3 * a] * F.cx / F.fx - H[3 * a + 1] * F.cy / F.fy + H[3 * a + 2];
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        H[b] = V.sK[0] * tmp[b] + V.sK[2] * tmp[6 + b];
        H[3 + b] = V.sK[4] * tmp[3 + b] + V.sK[5] * tmp[6 + b];
        H[6 + b] = V.sK[8] * tmp[6 + b];
    }
}
// ComputeCorrespondingPoint, cu:281-288, on an integer pixel position
PM_HD void pm_warp_literal(const float* H, int px, int py, float& ox, float& oy) {
    const float X = H[0] * px + H[1] * py + H[2];
    const float Y = H[3] * px + H[4] * py + H[5];
    const float Z = H[6] * px + H[7] * py + H[8];
    ox = X / Z;
    oy = Y / Z;
}
#endif

// ComputeBilateralNCC, cu:325-414: cost of hypothesis `hyp` at pixel (x,y) against source view v.
template <int SCALE, class Ctx, int TAPS = 6>
PM_HD float pm_ncc(const Ctx& c, const PmFrame& F, const PmRefStats& st, int v, const PmHyp& hyp, int x, int y,
                   uint32_t& nexec) {
    constexpr int HS = 1 << SCALE;
    const PmView& V = c.view(v);
#if PM_EXACT
    float Hm[9];
    pm_homography_literal(F, V, hyp.pl, Hm);
    float pcx, pcy;
    pm_warp_literal(Hm, x, y, pcx, pcy);
    if (pcx >= V.w || pcx < 0.0f || pcy >= V.h || pcy < 0.0f) return 2.0f;  // cu:351-353
    if (st.var_r < 1e-5f) return 2.0f;                                     // cu:407 (hypothesis-invariant)
    ++nexec;
    // source-side half of cu:355-413 as the reference's SASS evaluates it, taps unrolled. Per row: FMUL H0 px, H3 px, H6 px.
    // Per tap: numerators FADD(H2, FFMA(H1, py, H0 px)), ...; MUFU.RCP(Z); coordinates FFMA(X, rcp, 0.5); FMUL r*w, FMUL s*w,
    // FADD sum(w s), FFMA(s, s*w, .), FFMA(s, r*w, .). Host: the oracle's divisions and unfused products, same order.
    float sum_src = 0.0f, sum_src_src = 0.0f, sum_ref_src = 0.0f;
#if PM_PACKED && defined(__CUDA_ARCH__)
    if (TAPS % 2 == 0) {
        // The same operations, two per instruction (PM_PACKED above): taps (b, b + 1) of a row share the numerator and coordinate
        // instructions; (w s, w r) of a tap are one FMUL2, (sum w s s, sum w r s) one FFMA2. The per-row and per-window sums are
        // added in the reference's order (sequential FADDs per quantity), so nothing is reassociated.
        const pm_f2 H1 = pm_pk(Hm[1], Hm[1]), H4 = pm_pk(Hm[4], Hm[4]), H7 = pm_pk(Hm[7], Hm[7]);
        const pm_f2 H2 = pm_pk(Hm[2], Hm[2]), H5 = pm_pk(Hm[5], Hm[5]), H8 = pm_pk(Hm[8], Hm[8]);
        const pm_f2 half2 = pm_pk(0.5f, 0.5f);
#pragma unroll
        for (int a = 0; a < TAPS; ++a) {
            const int i = (2 * a - (TAPS - 1)) * HS;
            const float px = (float)(x + i);
            const float hx = pm_rmul(Hm[0], px), hy = pm_rmul(Hm[3], px), hz = pm_rmul(Hm[6], px);
            const pm_f2 hx2 = pm_pk(hx, hx), hy2 = pm_pk(hy, hy), hz2 = pm_pk(hz, hz);
            float row_src = 0.0f;
            pm_f2 row2 = pm_pk(0.0f, 0.0f);          // (row_src_src, row_ref_src)
#pragma unroll
            for (int b = 0; b < TAPS; b += 2) {
                const int j0 = (2 * b - (TAPS - 1)) * HS, j1 = (2 * b + 2 - (TAPS - 1)) * HS;
                const pm_f2 py = pm_pk((float)(y + j0), (float)(y + j1));
                const pm_f2 X = pm_add2(pm_fma2(H1, py, hx2), H2), Y = pm_add2(pm_fma2(H4, py, hy2), H5), Z = pm_add2(pm_fma2(H7, py, hz2), H8);
                float z0, z1;
                pm_upk(Z, z0, z1);
                const pm_f2 rz = pm_pk(pm_rcp_approx(z0), pm_rcp_approx(z1));
                float xs0, xs1, ys0, ys1;
                pm_upk(pm_fma2(X, rz, half2), xs0, xs1);
                pm_upk(pm_fma2(Y, rz, half2), ys0, ys1);
                const float s0 = c.src(v, xs0, ys0), s1 = c.src(v, xs1, ys1);
                const float r0 = c.ref(i, j0), r1 = c.ref(i, j1);
                const float w0 = PM_WTAB ? c.wt(a * TAPS + b) : pm_tap_weight<SCALE, TAPS>(F, a, b, r0, st.r0);
                const float w1 = PM_WTAB ? c.wt(a * TAPS + b + 1) : pm_tap_weight<SCALE, TAPS>(F, a, b + 1, r1, st.r0);
                const pm_f2 p0 = pm_mul2(pm_pk(w0, w0), pm_pk(s0, r0)), p1 = pm_mul2(pm_pk(w1, w1), pm_pk(s1, r1));   // (w s, w r)
                float sw0, rw0, sw1, rw1;
                pm_upk(p0, sw0, rw0);
                pm_upk(p1, sw1, rw1);
                row_src = pm_radd(row_src, sw0);
                row2 = pm_fma2(p0, pm_pk(s0, s0), row2);
                row_src = pm_radd(row_src, sw1);
                row2 = pm_fma2(p1, pm_pk(s1, s1), row2);
            }
            float row_src_src, row_ref_src;
            pm_upk(row2, row_src_src, row_ref_src);
            sum_src = pm_radd(sum_src, row_src);
            sum_src_src = pm_radd(sum_src_src, row_src_src);
            sum_ref_src = pm_radd(sum_ref_src, row_ref_src);
        }
    } else
#endif
#pragma unroll
    for (int a = 0; a < TAPS; ++a) {
        const int i = (2 * a - (TAPS - 1)) * HS;
        const float px = (float)(x + i);
        const float hx = pm_rmul(Hm[0], px), hy = pm_rmul(Hm[3], px), hz = pm_rmul(Hm[6], px);
        float row_src = 0.0f, row_src_src = 0.0f, row_ref_src = 0.0f;
#pragma unroll
        for (int b = 0; b < TAPS; ++b) {
            const int j = (2 * b - (TAPS - 1)) * HS;
            const float py = (float)(y + j);
            const float X = pm_radd(pm_ffma(Hm[1], py, hx), Hm[2]);
            const float Y = pm_radd(pm_ffma(Hm[4], py, hy), Hm[5]);
            const float Z = pm_radd(pm_ffma(Hm[7], py, hz), Hm[8]);
#if defined(__CUDA_ARCH__)
            const float rz = pm_rcp_approx(Z);
            const float s = c.src(v, __fmaf_rn(X, rz, 0.5f), __fmaf_rn(Y, rz, 0.5f));
#else
            const float s = c.src(v, X / Z + 0.5f, Y / Z + 0.5f);
#endif
            const float r = c.ref(i, j);
            const float w = PM_WTAB ? c.wt(a * TAPS + b) : pm_tap_weight<SCALE, TAPS>(F, a, b, r, st.r0);
            const float sw = pm_rmul(w, s), rw = pm_rmul(w, r);
            row_src = pm_radd(row_src, sw);
            row_src_src = pm_ffma(sw, s, row_src_src);
            row_ref_src = pm_ffma(rw, s, row_ref_src);
        }
        sum_src = pm_radd(sum_src, row_src);
        sum_src_src = pm_radd(sum_src_src, row_src_src);
        sum_ref_src = pm_radd(sum_ref_src, row_ref_src);
    }
    sum_src = pm_rmul(sum_src, st.inv_sw);
#if defined(__CUDA_ARCH__)
    const float var_src = __fmaf_rn(sum_src_src, st.inv_sw, -__fmul_rn(sum_src, sum_src));
    sum_ref_src = __fmul_rn(sum_ref_src, st.inv_sw);
    if (var_src < 1e-5f) return 2.0f;
    const float covar_src_ref = __fmaf_rn(-st.mean_r, sum_src, sum_ref_src);
    const float inv_dev = pm_rcp_approx(pm_sqrt_approx(__fmul_rn(st.var_r, var_src)));   // MUFU.SQRT, then MUFU.RCP
    return fmaxf(0.0f, fminf(2.0f, __fmaf_rn(-covar_src_ref, inv_dev, 1.0f)));
#else
    sum_src_src *= st.inv_sw;
    sum_ref_src *= st.inv_sw;
    const float var_src = sum_src_src - sum_src * sum_src;
    if (var_src < 1e-5f) return 2.0f;
    const float covar_src_ref = sum_ref_src - st.mean_r * sum_src;
    return fmaxf(0.0f, fminf(2.0f, 1.0f - covar_src_ref / sqrtf(st.var_r * var_src)));
#endif
#else
    const float fxp = (float)x, fyp = (float)y;
    // warped centre  P0 = A (x,y,1) + b q0 ; tap steps  U = H e_x, Vv = H e_y
    const float X0 = fmaf(V.b[0], hyp.q0, fmaf(V.A[0], fxp, fmaf(V.A[1], fyp, V.A[2])));
    const float Y0 = fmaf(V.b[1], hyp.q0, fmaf(V.A[3], fxp, fmaf(V.A[4], fyp, V.A[5])));
    const float Z0 = fmaf(V.b[2], hyp.q0, fmaf(V.A[6], fxp, fmaf(V.A[7], fyp, V.A[8])));
    const float rz0 = 1.0f / Z0;
    const float pcx = X0 * rz0, pcy = Y0 * rz0;
    if (pcx >= V.w || pcx < 0.0f || pcy >= V.h || pcy < 0.0f) return 2.0f;  // cu:351-353
    if (st.var_r < 1e-5f) return 2.0f;                                     // cu:407 (hypothesis-invariant)
    ++nexec;
    const float Ux = fmaf(V.b[0], hyp.mx, V.A[0]), Uy = fmaf(V.b[1], hyp.mx, V.A[3]), Uz = fmaf(V.b[2], hyp.mx, V.A[6]);
    const float Vx = fmaf(V.b[0], hyp.my, V.A[1]), Vy = fmaf(V.b[1], hyp.my, V.A[4]), Vz = fmaf(V.b[2], hyp.my, V.A[7]);
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int a = 0; a < TAPS; ++a) {
        const int i = (2 * a - (TAPS - 1)) * HS;
        const float Xa = fmaf((float)i, Ux, X0), Ya = fmaf((float)i, Uy, Y0), Za = fmaf((float)i, Uz, Z0);
#pragma unroll
        for (int b = 0; b < TAPS; ++b) {
            const int j = (2 * b - (TAPS - 1)) * HS;
            const float Z = fmaf((float)j, Vz, Za);
            const float rz = 1.0f / Z;
            const float xs = fmaf(fmaf((float)j, Vx, Xa), rz, 0.5f);
            const float ys = fmaf(fmaf((float)j, Vy, Ya), rz, 0.5f);
            const float s = c.src(v, xs, ys);
            const float r = c.ref(i, j);
            const float ws = (PM_WTAB ? c.wt(a * TAPS + b) : pm_tap_weight<SCALE, TAPS>(F, a, b, r, st.r0)) * s;
            s1 += ws;
            s2 = fmaf(ws, s, s2);
            s3 = fmaf(ws, r, s3);
        }
    }
    const float mean_s = s1 * st.inv_sw;
    const float var_s = s2 * st.inv_sw - mean_s * mean_s;
    if (var_s < 1e-5f) return 2.0f;
    const float cov = s3 * st.inv_sw - st.mean_r * mean_s;
    return fmaxf(0.0f, fminf(2.0f, 1.0f - cov * pm_rsqrt(st.var_r * var_s)));
#endif
}

#if PM_NCC_NOINLINE && defined(__CUDACC__)
template <int SCALE, class Ctx>
__device__ __noinline__ float pm_ncc_call(const Ctx& c, const PmFrame& F, const PmRefStats& st, int v, const PmHyp& hyp, int x, int y, uint32_t& nexec) {
    return pm_ncc<SCALE>(c, F, st, v, hyp, x, y, nexec);
}
#else
template <int SCALE, class Ctx>
PM_HD float pm_ncc_call(const Ctx& c, const PmFrame& F, const PmRefStats& st, int v, const PmHyp& hyp, int x, int y, uint32_t& nexec) {
    return pm_ncc<SCALE>(c, F, st, v, hyp, x, y, nexec);
}
#endif

// ComputeGeomConsistencyCost, cu:617-640.
#if PM_EXACT
// BackProjectPoint2W, cu:582-603, and ProjectPoint, cu:605-615, on a camera given as (K, R, t, C)
PM_HD void pm_backproject_literal(float x, float y, float depth, const float* K, const float* R, const float* C, float* X) {
    const float px = depth * (x - K[2]) / K[0];
    const float py = depth * (y - K[5]) / K[4];
    const float pz = depth;
    const float tx = R[0] * px + R[3] * py + R[6] * pz;
    const float ty = R[1] * px + R[4] * py + R[7] * pz;
    const float tz = R[2] * px + R[5] * py + R[8] * pz;
    X[0] = tx + C[0];
    X[1] = ty + C[1];
    X[2] = tz + C[2];
}
PM_HD void pm_project_literal(const float* X, const float* K, const float* R, const float* t, float& u, float& v) {
    const float tx = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const float ty = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const float tz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const float depth = K[6] * tx + K[7] * ty + K[8] * tz;
    u = (K[0] * tx + K[1] * ty + K[2] * tz) / depth;
    v = (K[3] * tx + K[4] * ty + K[5] * tz) / depth;
}
#endif

// depth_form: which contraction of the depth denominator the reference's inlined copy has (pm_depth_from_plane)
template <class Ctx>
PM_HD float pm_geom_cost(const Ctx& c, const PmFrame& F, int v, const pm_f4& pl, int x, int y, int depth_form = 0) {
    const PmView& V = c.view(v);
#if PM_EXACT
    // cu:617-640 through the two cameras, not through the pre-multiplied per-view transforms
    const float depth = pm_depth_from_plane(F, pl, x, y, depth_form);
    float P3[3], sx, sy;
    pm_backproject_literal((float)x, (float)y, depth, F.K, F.R, F.C, P3);
    pm_project_literal(P3, V.sK, V.sR, V.st, sx, sy);
    const float sd = c.src_depth(v, pm_f2i(sx), pm_f2i(sy));
    if (sd == 0.0f) return 3.0f;
    float Q3[3], bx, by;
    pm_backproject_literal(sx, sy, sd, V.sK, V.sR, V.sC, Q3);
    pm_project_literal(Q3, F.K, F.R, F.t, bx, by);
    const float diff_col = x - bx, diff_row = y - by;
    return fminf(3.0f, sqrtf(diff_col * diff_col + diff_row * diff_row));
#else
    // the four camera transforms pre-multiplied per view
    (void)depth_form;
    const float z = pm_depth_from_plane(F, pl, x, y);
    const float fxp = (float)x, fyp = (float)y;
    const float hx = fmaf(z, fmaf(V.Mf[0], fxp, fmaf(V.Mf[1], fyp, V.Mf[2])), V.vf[0]);
    const float hy = fmaf(z, fmaf(V.Mf[3], fxp, fmaf(V.Mf[4], fyp, V.Mf[5])), V.vf[1]);
    const float hz = fmaf(z, fmaf(V.Mf[6], fxp, fmaf(V.Mf[7], fyp, V.Mf[8])), V.vf[2]);
    const float sx = hx / hz, sy = hy / hz;
    const float sd = c.src_depth(v, pm_f2i(sx), pm_f2i(sy));
    if (sd == 0.0f) return 3.0f;
    const float bx = fmaf(sd, fmaf(V.Mb[0], sx, fmaf(V.Mb[1], sy, V.Mb[2])), V.vb[0]);
    const float by = fmaf(sd, fmaf(V.Mb[3], sx, fmaf(V.Mb[4], sy, V.Mb[5])), V.vb[1]);
    const float bz = fmaf(sd, fmaf(V.Mb[6], sx, fmaf(V.Mb[7], sy, V.Mb[8])), V.vb[2]);
    const float dc = fxp - bx / bz, dr = fyp - by / bz;
    return fminf(3.0f, sqrtf(dc * dc + dr * dr));
#endif
}

// ------------------------------------------------------------------------------------------------ view weights
// Monte-Carlo view weights are integer counts <= 15 (15 samples, cu:856): 4 bits per view, 32 views.
struct PmWeights {
    uint64_t lo, hi;
    PM_HD int get(int v) const { return (int)(((v < 16 ? lo : hi) >> ((v & 15) * 4)) & 15ull); }
    PM_HD void inc(int v) { if (v < 16) lo += 1ull << (v * 4); else hi += 1ull << ((v - 16) * 4); }
};

// ComputeMultiViewInitialCostandSelectedViews, cu:497-534: mean of the top_k smallest costs; mask of views <= k-th.
template <int SCALE, class Ctx>
PM_HD float pm_initial_cost(const Ctx& c, const PmFrame& F, const PmRefStats& st, const pm_f4& pl, int x, int y,
                            uint32_t& selected, uint32_t& nexec) {
    const PmHyp hyp = pm_hyp(F, pl, x, y);
    float cv[PM_MAX_SRC];
    int valid = 0;
    for (int v = 0; v < F.nsrc; ++v) {
        const float cst = pm_ncc<SCALE>(c, F, st, v, hyp, x, y, nexec);
        cv[v] = cst;
        if (cst < 2.0f) ++valid;
    }
    selected = 0;
    const int k = valid < F.top_k ? valid : F.top_k;
    if (k <= 0) return 2.0f;
    // k-th smallest by repeated selection (k <= 4): same values as the reference's full insertion sort
    float sum = 0.f, thr = -1.0f;
    uint32_t used = 0;
    for (int t = 0; t < k; ++t) {
        float best = FLT_MAX; int bi = 0;
        for (int v = 0; v < F.nsrc; ++v)
            if (!((used >> v) & 1u) && cv[v] < best) { best = cv[v]; bi = v; }
        used |= 1u << bi;
        sum += best;
        thr = best;
    }
    for (int v = 0; v < F.nsrc; ++v) if (cv[v] <= thr) selected |= 1u << v;
    return sum / k;
}

// ------------------------------------------------------------------------------------------------ propagation
// Candidate offsets of the 8 sampling regions, cu:769-779, generated instead of tabulated:
// regions 0-3 are V-shaped (12 samples), regions 4-7 axial rays at odd distances 5..23 (10 samples).
PM_HD void pm_region_offset(int r, int k, int& dx, int& dy) {
    if (r < 4) {
        const int t = k >> 1, sgn = (k & 1) ? 1 : -1;   // pairs (-a, ..), (+a, ..)
        const int a = 5 + t, b = 6 + t;
        if (r == 0) { dx = sgn * a; dy = -b; }
        else if (r == 1) { dx = sgn * a; dy = b; }
        else if (r == 2) { dx = -b; dy = sgn * a; }
        else { dx = b; dy = sgn * a; }
    } else {
        const int d = 5 + 2 * k;
        if (r == 4) { dx = 0; dy = -d; }
        else if (r == 5) { dx = 0; dy = d; }
        else if (r == 6) { dx = -d; dy = 0; }
        else { dx = d; dy = 0; }
    }
}

// One pixel of one checkerboard half-sweep: CheckerboardPropagation (cu:724-998) with
// PlaneHypothesisRefinement (cu:642-722) folded in.
//
// The 14 hypotheses a pixel evaluates per half-sweep -- 8 propagation candidates, its current plane,
// 5 refinement proposals -- run through ONE loop with ONE inlined copy of the 36-tap NCC, so the hot
// code stays small enough for the instruction cache; the decisions the reference takes between those
// evaluations (view selection, candidate choice, proposal generation) sit at the top of iterations 8 and 9.
// `ca` is the thread's private candidate cost table, 8 x nsrc floats (cost_array, cu:795).
// candidate cost table of one thread in thread-private memory (stride 1)
struct PmTableLocal {
    float* p; int nsrc;
    PM_HD float& operator()(int r, int v) const { return p[r * nsrc + v]; }
};

// `ca` is any object with float& operator()(int region, int view): the thread's candidate cost table.
template <int SCALE, class Ctx, class Table>
PM_HD void pm_sweep_pixel(const Ctx& c, const PmFrame& F, const PmState& S, int x, int y, int iter, Table ca) {
    const int W = F.W, H = F.H, nsrc = F.nsrc;
    const int idx = y * W + x;
    PmRng rs = pm_rng_load(S.rng, idx);
    const PmRefStats st = pm_ref_stats<SCALE>(c, F);
    uint32_t nexec = 0;
    const uint32_t all_views = nsrc >= 32 ? 0xffffffffu : ((1u << nsrc) - 1u);

    // ---- best-cost sample of each region (cu:798-816)
    int pos[8];
    uint32_t flags = 0;
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const int nk = r < 4 ? 12 : 10;
        float best = FLT_MAX; int bpos = -1;
        for (int k = 0; k < nk; ++k) {
            int dx, dy;
            pm_region_offset(r, k, dx, dy);
            const int nx = x + dx, ny = y + dy;
            if (!(nx >= 0 && ny >= 0 && nx < W && ny < H)) continue;
            const float nc = S.costs[ny * W + nx];
            if (best > nc) { best = nc; bpos = ny * W + nx; }
        }
        pos[r] = bpos;
        if (best < FLT_MAX) flags |= 1u << r;
    }
    // QUIRK cu:795: `float cost_array[8][32] = {2.0f}` -> element [0][0] is 2, everything else 0, and regions
    // without an in-bounds sample keep those values.
    for (int r = 0; r < 8; ++r)
        for (int v = 0; v < nsrc; ++v) ca(r, v) = 0.0f;
    ca(0, 0) = 2.0f;

    const float depth_sigma = (F.depth_max - F.depth_min) / 64.0f;
    const float two_ds2 = 2 * depth_sigma * depth_sigma;
    const float angle_sigma = (float)(PM_PI_D * (5.0f / 180.0f));
    const float two_as2 = 2 * angle_sigma * angle_sigma;

    const pm_f4 cur = S.planes[idx];
    PmWeights vw = {0ull, 0ull};
    uint32_t sel = 0;
    float wnorm = 0.f;
    float fc[8];
    int kmin = 0;
    float cost_now = 0.f, g_now = 0.f, depth_now = 0.f, restricted_cost = 0.f;
    pm_f4 plane_now = cur;
    // refinement proposal ingredients (set at h == 9)
    float depth_rand = 0.f, depth_pert = 0.f, depth_base = 0.f, depth_prior = 0.f;
    pm_f4 rand_n = cur, pert_n = cur, base_n = cur, prior_pl = cur;
    bool has_prior = false;

#if PM_VIEW_MAJOR
    // ---- candidate cost table, view-major (cu:798-819 evaluates it candidate-major; the entries are independent)
#pragma unroll 1
    for (int v = 0; v < nsrc; ++v) {
#pragma unroll 1
        for (int h = 0; h < 8; ++h) {
            if (!((flags >> h) & 1u)) continue;
            const PmHyp hyp = pm_hyp(F, S.planes[pos[h]], x, y);
            ca(h, v) = pm_ncc_call<SCALE>(c, F, st, v, hyp, x, y, nexec);
        }
    }
#endif
#pragma unroll 1
    for (int h = PM_VIEW_MAJOR ? 8 : 0; h < 14; ++h) {
        if (h == 8) {
            // ---- view selection (cu:821-878)
            uint32_t nb[4];  // neighbour masks: up / down / left / right, gated by the flags of regions 0..3 (cu:824-830)
            nb[0] = (flags & 1u) ? S.views[idx - W] : 0u;
            nb[1] = (flags & 2u) ? S.views[idx + W] : 0u;
            nb[2] = (flags & 4u) ? S.views[idx - 1] : 0u;
            nb[3] = (flags & 8u) ? S.views[idx + 1] : 0u;
            const float thr = (float)(0.8 * (double)pm_exp((float)(iter * iter) / (-90.0f)));
            const float fallback = pm_exp(thr * thr / (-0.32f));
            float probs[PM_MAX_SRC];
            float psum = 0.f;
            for (int v = 0; v < nsrc; ++v) {
                float prior = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((flags >> i) & 1u) prior += ((nb[i] >> v) & 1u) ? 0.9f : 0.1f;
                float count = 0.f, tmpw = 0.f; int count_false = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float cc = ca(r, v);
                    if (cc < thr) { tmpw += pm_exp(cc * cc / (-0.18f)); count += 1.f; }
                    if (cc > 1.2f) ++count_false;
                }
                float pr;
                if (count > 2 && count_false < 3) pr = prior * tmpw / count;
                else if (count_false < 3) pr = prior * fallback;
                else pr = 0.f;
                probs[v] = pr;
                psum += pr;
            }
            // TransformPDFToCDF, cu:42-56 (0-sum -> NaNs, last bin forced to 1)
            const float inv = 1.0f / psum;
            float cum = 0.f;
            for (int v = 0; v < nsrc; ++v) { cum += probs[v] * inv; probs[v] = cum; }
            probs[nsrc - 1] = 1.f;
            for (int s = 0; s < 15; ++s) {
                const float u = pm_uniform(rs) - FLT_EPSILON;
                for (int v = 0; v < nsrc; ++v)
                    if (probs[v] > u) { vw.inc(v); break; }
            }
            for (int v = 0; v < nsrc; ++v) {
                const int wv = vw.get(v);
                if (wv > 0) { sel |= 1u << v; wnorm += (float)wv; }
            }
            // ---- candidate scores (cu:880-899)
#pragma unroll 1
            for (int r = 0; r < 8; ++r) {
                float acc = 0.f;
                const bool fl = (flags >> r) & 1u;
                pm_f4 pl = cur;
                if (F.geom && fl) pl = S.planes[pos[r]];
                for (uint32_t m = sel; m; m &= m - 1) {
                    const int v = pm_ffs(m) - 1;
                    float term = ca(r, v);
                    if (F.geom) term += fl ? 0.2f * pm_geom_cost(c, F, v, pl, x, y) : 0.1f * 3.0f;
                    acc += (float)vw.get(v) * term;
                }
                fc[r] = acc / wnorm;
            }
            float mn = fc[0];  // FindMinCostIndex, cu:58-69: last minimum wins
#pragma unroll
            for (int r = 1; r < 8; ++r) if (fc[r] <= mn) { mn = fc[r]; kmin = r; }
        }
        if (h == 9) {
            // ---- planar-prior selection (cu:924-978)
            if (F.planar && !F.geom) {
                if (S.mask[idx] > 0) {
                    const pm_f4 pp = S.prior[idx];
                    const float dprior = pm_depth_from_plane(F, pp, x, y);
                    float rbest = 0.f; int kmax = 0;   // FindMaxCostIndex (0 for unflagged), last max wins
                    pm_f4 cand_best = cur;
#pragma unroll 1
                    for (int r = 0; r < 8; ++r) {
                        float rf = 0.f;
                        pm_f4 pl = cur;
                        if ((flags >> r) & 1u) {
                            pl = S.planes[pos[r]];
                            const float dd = pm_depth_minus(F, pl, x, y, dprior);
                            const float ad = acosf(pm_dot3(pp, pl));
                            const float prior = 0.5f + pm_exp(-dd * dd / two_ds2) * pm_exp(-ad * ad / two_as2);
                            rf = pm_exp(-fc[r] * fc[r] / 0.18f) * prior;
                        }
                        if (r == 0 || rf >= rbest) { rbest = rf; kmax = r; cand_best = pl; }
                    }
                    const float dd = depth_now - dprior;   // cu:949-950: the local depth_now is the outer one (CSE): plain FADD
                    const float ad = acosf(pm_dot3(pp, cur));
                    const float prior = 0.5f + pm_exp(-dd * dd / two_ds2) * pm_exp(-ad * ad / two_as2);
                    const float rc_now = pm_exp(-cost_now * cost_now / 0.18f) * prior;
                    if ((flags >> kmax) & 1u) {
                        const float db = pm_depth_from_plane(F, cand_best, x, y);
                        if (db >= F.depth_min && db <= F.depth_max && rbest > rc_now) {
                            // QUIRK cu:950-966: the adopted depth goes to a shadowed local and cost_now is not
                            // updated: refinement continues with the OLD depth and OLD cost but the NEW plane.
                            plane_now = cand_best;
                            restricted_cost = rbest;
                            S.views[idx] = sel;
                        }
                    }
                } else if ((flags >> kmin) & 1u) {
                    const pm_f4 pl = S.planes[pos[kmin]];
                    const float db = pm_depth_from_plane(F, pl, x, y);
                    if (db >= F.depth_min && db <= F.depth_max && fc[kmin] < cost_now) {
                        // QUIRK cu:969-977: plane and depth adopted, cost_now and the view mask are not
                        depth_now = db;
                        plane_now = pl;
                    }
                }
            }
            // ---- plain selection (cu:981-991)
            if (!F.planar && ((flags >> kmin) & 1u)) {
                const pm_f4 pl = S.planes[pos[kmin]];
                const float db = pm_depth_from_plane(F, pl, x, y);
                if (db >= F.depth_min && db <= F.depth_max && fc[kmin] < cost_now) {
                    depth_now = db; plane_now = pl; cost_now = fc[kmin];
                    S.views[idx] = sel;
                }
            }
            // ---- refinement proposals (cu:644-675)
            const float perturbation = 0.02f;
            has_prior = F.planar && S.mask[idx] > 0;
            if (has_prior) {
                // QUIRK cu:656-663: the prior-guided proposal is drawn (4 uniforms) and then overwritten, because
                // the block that follows has no `else`. Only the RNG stream advances.
                prior_pl = S.prior[idx];
                depth_prior = pm_depth_from_plane(F, prior_pl, x, y);
                (void)pm_uniform(rs);
                (void)pm_perturbed_normal(F, x, y, prior_pl, rs, angle_sigma);
            }
            depth_rand = pm_uniform(rs) * (F.depth_max - F.depth_min) + F.depth_min;
            rand_n = pm_random_normal(F, x, y, rs);
            // QUIRK cu:668-670: `while (d < min && d > max)` can never hold, so exactly one draw
            const float dlo = (1 - perturbation) * depth_now, dhi = (1 + perturbation) * depth_now;
            depth_pert = pm_uniform(rs) * (dhi - dlo) + dlo;
            pert_n = pm_perturbed_normal(F, x, y, plane_now, rs, (float)(perturbation * PM_PI_D));
            base_n = plane_now;
            depth_base = depth_now;
        }

        // ---- the hypothesis of this iteration and the views it is scored on
        pm_f4 tp;
        uint32_t mask;
        float hd = 0.f;
        if (h < 8) {
            if (!((flags >> h) & 1u)) continue;
            tp = S.planes[pos[h]];
            mask = all_views;                       // cu:817: every source view
        } else if (h == 8) {
            tp = cur;
            mask = sel;                             // cu:903-913; zero-weight views contribute exactly 0
        } else {
            // (d_rand, n) (d, n_rand) (d_rand, n_rand) (d, n_pert) (d_pert, n)   cu:674-675
            const int i = h - 9;
            hd = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : depth_base);
            tp = (i == 1 || i == 2) ? rand_n : (i == 3 ? pert_n : base_n);
            tp.w = pm_plane_distance(F, x, y, hd, tp);
            mask = sel;                             // cu:684-694
#if PM_EARLY_OUT
            // a proposal whose depth leaves the range is scored by the reference and then always rejected (cu:707,713)
            const float dchk = pm_depth_from_plane(F, tp, x, y);
            if (!(dchk >= F.depth_min && dchk <= F.depth_max)) continue;
#endif
        }
        const PmHyp hyp = pm_hyp(F, tp, x, y);
        float tc = 0.f, tg = 0.f;
        float prior_f = 0.f;       // the prior factor of a refinement proposal (cu:696-700): known before its views are scored
        if (h > 8 && has_prior) {
            const float dd = hd - depth_prior;
            const float ad = acosf(pm_dot3(prior_pl, tp));
            prior_f = 0.5f + pm_exp(-dd * dd / two_ds2) * pm_exp(-ad * ad / two_as2);
        }
#pragma unroll 1
#if PM_UNIFORM_VIEWS
        // every lane of a warp walks the views in the same order (lanes that do not need view v idle through it), so
        // a warp-wide texture fetch stays inside one layer of the image array
        for (int v = 0; v < nsrc; ++v) {
            if (!((mask >> v) & 1u)) continue;
#else
        for (uint32_t m = mask; m; m &= m - 1) {
            const int v = pm_ffs(m) - 1;
#endif
            const float cst = pm_ncc_call<SCALE>(c, F, st, v, hyp, x, y, nexec);
            if (h < 8) {
                ca(h, v) = cst;
            } else {
                const float wv = (float)vw.get(v);
                if (F.geom) {
                    const float g = 0.2f * pm_geom_cost(c, F, v, tp, x, y, h > 8);
                    tc += wv * (cst + g);
                    // QUIRK cu:689: in refinement the geometric part is weighted by view_weights[hypothesis index]
                    tg += (h == 8 ? wv : (float)vw.get(h - 9)) * g;
                } else {
                    tc += wv * cst;
                }
#if PM_EARLY_OUT
                // Costs are >= 0 and summed in the reference's order, so the running sum only grows: once it fails the
                // acceptance test `tc / wnorm < cost_now` (cu:713) the remaining views cannot rescue the proposal.
                if (h > 8 && !has_prior && !(tc / wnorm < cost_now)) break;
#if PM_PRIOR_EARLY_OUT
                if (h > 8 && has_prior) {
                    const float t = tc / wnorm;
                    if (!(pm_exp(-t * t / 0.18f) * prior_f > restricted_cost)) break;
                }
#endif
#endif
            }
        }
        if (h == 8) {
            cost_now = tc / wnorm;
            if (F.geom) g_now = tg / wnorm;
            depth_now = pm_depth_from_plane(F, cur, x, y);
        } else if (h > 8) {
            tc /= wnorm;
            if (F.geom) tg /= wnorm;
            const float dbefore = pm_depth_from_plane(F, tp, x, y);
            const bool in_range = dbefore >= F.depth_min && dbefore <= F.depth_max;
            if (has_prior) {
                const float rtc = pm_exp(-tc * tc / 0.18f) * prior_f;
                // QUIRK cu:707-710: restricted_cost is never refreshed after an accept
                if (in_range && rtc > restricted_cost) { plane_now = tp; cost_now = tc; }
            } else if (in_range && tc < cost_now) {
                plane_now = tp; cost_now = tc; g_now = tg;
            }
        }
    }
    // ---- write-back (cu:993-997)
    S.costs[idx] = cost_now;
    S.planes[idx] = plane_now;
    if (F.geom) S.geom[idx] = g_now;
    pm_rng_store(S.rng, idx, rs);
    if (S.counters) pm_count(S.counters, nexec);
}

// ------------------------------------------------------------------------------------------------ init / finalize
// InitializeScore, cu:536-573 (always evaluated at the widest window, SCALE = max_scale = 2).
template <int SCALE, class Ctx>
PM_HD void pm_init_pixel(const Ctx& c, const PmFrame& F, const PmState& S, int x, int y, uint64_t seed) {
    const int idx = y * F.W + x;
    PmRng rs;
    pm_rng_init(rs, pm_mix_seed(seed, (uint32_t)x, (uint32_t)y));
    pm_f4 pl;
    if (!F.geom && !F.planar) {
        pl = pm_random_plane(F, x, y, rs);
    } else if (F.planar && S.mask[idx] > 0 && S.costs[idx] >= 0.1f) {
        // prior-guided re-initialisation: plane distance uniform in [0.94, 1.06] * w_prior, normal perturbed by 0.06 pi
        const float perturbation = 0.02f;
        const pm_f4 pp = S.prior[idx];
        const float lo = (1 - 3 * perturbation) * pp.w, hi = (1 + 3 * perturbation) * pp.w;
        const float dp = pm_uniform(rs) * (hi - lo) + lo;
        pl = pm_perturbed_normal(F, x, y, pp, rs, (float)(3 * perturbation * PM_PI_D));
        pl.w = dp;
    } else {
        // stored (world normal, depth) -> (camera normal, plane distance)
        const pm_f4 q = S.planes[idx];
        pl.x = F.R[0] * q.x + F.R[1] * q.y + F.R[2] * q.z;
        pl.y = F.R[3] * q.x + F.R[4] * q.y + F.R[5] * q.z;
        pl.z = F.R[6] * q.x + F.R[7] * q.y + F.R[8] * q.z;
        pl.w = pm_plane_distance(F, x, y, q.w, pl);
    }
    const PmRefStats st = pm_ref_stats<SCALE>(c, F);
    uint32_t sel, nexec = 0;
    const float cost = pm_initial_cost<SCALE>(c, F, st, pl, x, y, sel, nexec);
    S.planes[idx] = pl;
    S.costs[idx] = cost;
    S.views[idx] = sel;
    pm_rng_store(S.rng, idx, rs);
    if (S.counters) pm_count(S.counters, nexec);
}

// GetDepthandNormal, cu:1021-1034
PM_HD pm_f4 pm_depth_normal(const PmFrame& F, const pm_f4& pl, int x, int y) {
    pm_f4 o;
    o.w = pm_depth_from_plane(F, pl, x, y);
    o.x = F.R[0] * pl.x + F.R[3] * pl.y + F.R[6] * pl.z;
    o.y = F.R[1] * pl.x + F.R[4] * pl.y + F.R[7] * pl.z;
    o.z = F.R[2] * pl.x + F.R[5] * pl.y + F.R[8] * pl.z;
    return o;
}

// CheckerboardFilter, cu:1036-1150: median of the depth over the pixel and up to 20 opposite-colour neighbours.
// Returns the filtered depth (or the unchanged one when cost < 0.001).
PM_HD float pm_median_depth(const pm_f4* planes, const float* costs, int W, int H, int x, int y) {
    const int ctr = y * W + x;
    const float self = planes[ctr].w;
    if (costs[ctr] < 0.001f) return self;
    float f[21];
    int n = 0;
    f[n++] = self;
#define PM_TAP(cond, off) if (cond) f[n++] = planes[ctr + (off)].w;
    PM_TAP(y > 0, -W) PM_TAP(y > 2, -3 * W) PM_TAP(y > 4, -5 * W)
    PM_TAP(y < H - 1, W) PM_TAP(y < H - 3, 3 * W) PM_TAP(y < H - 5, 5 * W)
    PM_TAP(x > 0, -1) PM_TAP(x > 2, -3) PM_TAP(x > 4, -5)
    PM_TAP(x < W - 1, 1) PM_TAP(x < W - 3, 3) PM_TAP(x < W - 5, 5)
    // knight-like taps; bounds tests are the reference's, asymmetries included (cu:1107-1141)
    PM_TAP(y > 0 && x < W - 2, -W + 2) PM_TAP(y < H - 1 && x < W - 2, W + 2)
    PM_TAP(y > 0 && x > 1, -W - 2) PM_TAP(y < H - 1 && x > 1, W - 2)
    PM_TAP(x > 0 && y > 2, -1 - 2 * W) PM_TAP(x < W - 1 && y > 2, 1 - 2 * W)
    PM_TAP(x > 0 && y < H - 2, -1 + 2 * W) PM_TAP(x < W - 1 && y < H - 2, 1 + 2 * W)
#undef PM_TAP
    for (int i = 1; i < n; ++i) {  // insertion sort, cu:14-23
        const float t = f[i];
        int j = i;
        for (; j >= 1 && t < f[j - 1]; --j) f[j] = f[j - 1];
        f[j] = t;
    }
    const int m = n / 2;
    return (n % 2 == 0) ? (f[m - 1] + f[m]) / 2 : f[m];
}

}  // namespace PM_ARITH_NS
using namespace PM_ARITH_NS;

#endif  // MPMVS_PM_CORE_CUH
