/* oracle/pm_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of MP-MVS's PatchMatch hot path.
 *
 * Plain C, one function per reference device function / kernel, each citing the lines of
 * /root/reference/src/PatchMatch.cu it follows (all "cu:NNN" below refer to that file).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library;
 * the product (mp-mvs_b200/csrc) never links or calls it.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself: oracle/_ref/libmpmvs_ref.so (the
 * reference .cu compiled in place) run on a B200 produced the fixtures under tests/golden/
 * (script: tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against them
 * on CPU, and tests/test_parity_gpu.py re-checks it live against oracle/_ref on the GPU box.
 *
 * Hardware behaviour that has to be modelled on a CPU:
 *   - cudaFilterModeLinear on a float texture with un-normalised coordinates and Wrap addressing
 *     (PatchMatch.cpp:1012-1018): behaves as clamp-to-edge; coordinates are shifted by -0.5 and the
 *     fractional weights are held in 9-bit fixed point with 8 fractional bits.
 *   - cuRAND XORWOW (curand_init(seed,0,0) + curand_uniform), restated from the published generator.
 *   - --use_fast_math approximations (ex2/rcp/rsqrt/sin/cos.approx) are replaced by libm calls; the
 *     difference is far below the 1e-3 cost tolerance the parity tests use.
 * The only deliberate change is the seed: mix(seed, x, y) instead of clock64(), the
 * same redirection oracle/ref_harness.cu applies to the reference build.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef PMO_TEX_FRAC_BITS
#define PMO_TEX_FRAC_BITS 8
#endif

typedef struct { float x, y, z, w; } f4;

/* struct Camera, include/PatchMatch.h:35-46 (112 bytes) */
typedef struct {
    float K[9], R[9], t[3], C[3];
    int height, width;
    float depth_min, depth_max;
} PmoCamera;

/* the fields of PatchMatchParams (include/PatchMatch.h:48-67) that the kernels read */
typedef struct {
    int max_iterations;   /* 3 */
    int num_images;
    float sigma_spatial;  /* 5 */
    float sigma_color;    /* 3 */
    int top_k;            /* 4 */
    float depth_min, depth_max;
    int max_scale;        /* 2 */
    int geom_consistency, geomPlanarPrior, planar_prior;
} PmoParams;

typedef struct { uint32_t d, v[5]; } PmoRng; /* curandStateXORWOW without the Box-Muller fields */

typedef struct {
    int n, width, height;
    PmoCamera cams[33];
    float *images[33];   /* owned copies */
    float *depths[32];   /* source depth maps (geom), owned */
    PmoParams params;
    f4 *planes;  float *costs;  PmoRng *rng;  uint32_t *views;  float *geom;
    f4 *prior;   uint32_t *mask;
} Pmo;

/* ------------------------------------------------------------------ RNG */
static uint64_t mix_seed(uint64_t seed, uint32_t x, uint32_t y) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)y) << 32) | (uint64_t)x) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* curand_init(seed, 0, 0): XORWOW seeding (Marsaglia xorwow, cuRAND's salted start state) */
static void rng_init(PmoRng *s, uint64_t seed) {
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;  s->v[1] = 362436069u ^ t0;  s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;   s->v[4] = 5783321u + t0;
}
static uint32_t rng_next(PmoRng *s) {
    uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}
static float rng_uniform(PmoRng *s) { /* curand_uniform: (0,1], fused multiply-add like the device code */
    return fmaf((float)rng_next(s), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

/* ------------------------------------------------------------------ row-parallel helper (pthreads; no OpenMP runtime in this image) */
#include <pthread.h>
#include <unistd.h>
typedef void (*row_fn)(void *ctx, int row);
typedef struct { row_fn fn; void *ctx; int nrows; volatile int next; } RowJob;
static void *row_worker(void *arg) {
    RowJob *j = (RowJob *)arg;
    for (;;) {
        int r = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (r >= j->nrows) break;
        j->fn(j->ctx, r);
    }
    return NULL;
}
static int g_threads = 0;
void pmo_set_threads(int n) { g_threads = n; }
int pmo_get_threads(void) {
    if (g_threads > 0) return g_threads;
    const char *e = getenv("PMO_THREADS");
    int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : n);
}
static void parallel_rows(row_fn fn, void *ctx, int nrows) {
    RowJob job = { fn, ctx, nrows, 0 };
    int nt = pmo_get_threads();
    if (nt > nrows) nt = nrows;
    if (nt <= 1) { row_worker(&job); return; }
    pthread_t th[256];
    for (int i = 1; i < nt; ++i) pthread_create(&th[i], NULL, row_worker, &job);
    row_worker(&job);
    for (int i = 1; i < nt; ++i) pthread_join(th[i], NULL);
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------ texture unit model */
static float tex2d(const float *img, int w, int h, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    float fx = floorf(xb), fy = floorf(yb);
    const float q = (float)(1 << PMO_TEX_FRAC_BITS);
    float a = floorf((xb - fx) * q + 0.5f) / q, b = floorf((yb - fy) * q + 0.5f) / q;
    int i0 = (int)fx, j0 = (int)fy, i1 = i0 + 1, j1 = j0 + 1;
    if (!(xb > -1e9f)) i0 = i1 = 0; /* NaN / -inf guard */
    if (!(yb > -1e9f)) j0 = j1 = 0;
    i0 = clampi(i0, 0, w - 1); i1 = clampi(i1, 0, w - 1);
    j0 = clampi(j0, 0, h - 1); j1 = clampi(j1, 0, h - 1);
    float t00 = img[(size_t)j0 * w + i0], t10 = img[(size_t)j0 * w + i1];
    float t01 = img[(size_t)j1 * w + i0], t11 = img[(size_t)j1 * w + i1];
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}

/* ------------------------------------------------------------------ small helpers */
static float dot3(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* cu:9-12 */

static void sort_small(float *d, int n) { /* cu:14-23 insertion sort */
    for (int i = 1; i < n; i++) {
        float tmp = d[i]; int j;
        for (j = i; j >= 1 && tmp < d[j - 1]; j--) d[j] = d[j - 1];
        d[j] = tmp;
    }
}

static void pdf_to_cdf(float *p, int n) { /* cu:42-56, including the 0-sum NaN behaviour */
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s += p[i];
    const float inv = 1.0f / s;
    float c = 0.0f;
    for (int i = 0; i < n; ++i) { c += p[i] * inv; p[i] = c; }
    p[n - 1] = 1.f;
}

static int arg_min_last(const float *c, int n) { /* cu:58-69 ("<=": last minimum wins) */
    float m = c[0]; int k = 0;
    for (int i = 1; i < n; ++i) if (c[i] <= m) { m = c[i]; k = i; }
    return k;
}
static int arg_max_last(const float *c, int n) { /* cu:71-82 */
    float m = c[0]; int k = 0;
    for (int i = 1; i < n; ++i) if (c[i] >= m) { m = c[i]; k = i; }
    return k;
}

static float depth_from_plane(const PmoCamera *c, f4 pl, int x, int y) { /* cu:84-87 */
    return -pl.w * c->K[0] / ((x - c->K[2]) * pl.x + (c->K[0] / c->K[4]) * (y - c->K[5]) * pl.y + c->K[0] * pl.z);
}

static f4 normal_cam_to_world(const PmoCamera *c, f4 p) { /* TransformNormal cu:89-97 (R^T n) */
    f4 o;
    o.x = c->R[0] * p.x + c->R[3] * p.y + c->R[6] * p.z;
    o.y = c->R[1] * p.x + c->R[4] * p.y + c->R[7] * p.z;
    o.z = c->R[2] * p.x + c->R[5] * p.y + c->R[8] * p.z;
    o.w = p.w;
    return o;
}
static f4 normal_world_to_cam(const PmoCamera *c, f4 p) { /* TransformNormal2RefCam cu:308-316 (R n) */
    f4 o;
    o.x = c->R[0] * p.x + c->R[1] * p.y + c->R[2] * p.z;
    o.y = c->R[3] * p.x + c->R[4] * p.y + c->R[5] * p.z;
    o.z = c->R[6] * p.x + c->R[7] * p.y + c->R[8] * p.z;
    o.w = p.w;
    return o;
}

static float plane_distance(const PmoCamera *c, int x, int y, float depth, f4 n) { /* GetPlane2Origin cu:163-176 */
    float X0 = depth * (x - c->K[2]) / c->K[0], X1 = depth * (y - c->K[5]) / c->K[4], X2 = depth;
    return -(n.x * X0 + n.y * X1 + n.z * X2);
}

static f4 view_dir(const PmoCamera *c, int x, int y) { /* cu:179-186 */
    f4 v = { (x - c->K[2]) / c->K[0], (y - c->K[5]) / c->K[4], 1.0f, 0.0f };
    return v;
}

static void normalize3(f4 *v) { /* cu:188-195 (rsqrtf) */
    float inv = 1.0f / sqrtf(v->x * v->x + v->y * v->y + v->z * v->z);
    v->x *= inv; v->y *= inv; v->z *= inv;
}

static f4 random_normal(const PmoCamera *c, int x, int y, PmoRng *rs) { /* cu:197-219 */
    float q1, q2, s;
    do {
        q1 = 2.f * rng_uniform(rs) - 1.f;
        q2 = 2.f * rng_uniform(rs) - 1.f;
        s = q1 * q1 + q2 * q2;
    } while (s >= 1.f);
    const float sq = sqrtf(1.f - s);
    f4 n = { 2.0f * q1 * sq, 2.0f * q2 * sq, 1.0f - 2.0f * s, 0.0f };
    f4 vd = view_dir(c, x, y);
    if (n.x * vd.x + n.y * vd.y + n.z * vd.z > 0.0f) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
    normalize3(&n);
    return n;
}

static f4 random_plane(const PmoCamera *c, int x, int y, PmoRng *rs, float dmin, float dmax) { /* cu:221-226 */
    f4 pl = random_normal(c, x, y, rs);
    float depth = rng_uniform(rs) * (dmax - dmin) + dmin;
    pl.w = plane_distance(c, x, y, depth, pl);
    return pl;
}

static f4 perturbed_normal(const PmoCamera *c, int x, int y, f4 normal, PmoRng *rs, float perturbation) { /* cu:460-495 */
    f4 vd = view_dir(c, x, y);
    const float a1 = (rng_uniform(rs) - 0.5f) * perturbation;
    const float a2 = (rng_uniform(rs) - 0.5f) * perturbation;
    const float a3 = (rng_uniform(rs) - 0.5f) * perturbation;
    const float s1 = sinf(a1), s2 = sinf(a2), s3 = sinf(a3), c1 = cosf(a1), c2 = cosf(a2), c3 = cosf(a3);
    float R[9];
    R[0] = c2 * c3;                 R[1] = c3 * s1 * s2 - c1 * s3;  R[2] = s1 * s3 + c1 * c3 * s2;
    R[3] = c2 * s3;                 R[4] = c1 * c3 + s1 * s2 * s3;  R[5] = c1 * s2 * s3 - c3 * s1;
    R[6] = -s2;                     R[7] = c2 * s1;                 R[8] = c1 * c2;
    f4 o;
    o.x = R[0] * normal.x + R[1] * normal.y + R[2] * normal.z;   /* Mat33DotVec3 cu:35-40: .w left unset */
    o.y = R[3] * normal.x + R[4] * normal.y + R[5] * normal.z;
    o.z = R[6] * normal.x + R[7] * normal.y + R[8] * normal.z;
    o.w = 0.0f;
    if (dot3(o, vd) >= 0.0f) return normal;
    normalize3(&o);
    return o;
}

/* ------------------------------------------------------------------ homography + NCC */
static void homography(const PmoCamera *r, const PmoCamera *s, f4 pl, float *H) { /* cu:228-279 */
    float Rr[9], Cr[3], tr[3], tmp[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            Rr[3 * a + b] = s->R[3 * a + 0] * r->R[3 * b + 0] + s->R[3 * a + 1] * r->R[3 * b + 1] + s->R[3 * a + 2] * r->R[3 * b + 2];
    for (int a = 0; a < 3; ++a) Cr[a] = r->C[a] - s->C[a];
    for (int a = 0; a < 3; ++a) tr[a] = s->R[3 * a + 0] * Cr[0] + s->R[3 * a + 1] * Cr[1] + s->R[3 * a + 2] * Cr[2];
    for (int a = 0; a < 3; ++a) {
        H[3 * a + 0] = Rr[3 * a + 0] - tr[a] * pl.x / pl.w;
        H[3 * a + 1] = Rr[3 * a + 1] - tr[a] * pl.y / pl.w;
        H[3 * a + 2] = Rr[3 * a + 2] - tr[a] * pl.z / pl.w;
    }
    for (int a = 0; a < 3; ++a) {
        tmp[3 * a + 0] = H[3 * a + 0] / r->K[0];
        tmp[3 * a + 1] = H[3 * a + 1] / r->K[4];
        tmp[3 * a + 2] = -H[3 * a + 0] * r->K[2] / r->K[0] - H[3 * a + 1] * r->K[5] / r->K[4] + H[3 * a + 2];
    }
    for (int b = 0; b < 3; ++b) {
        H[0 + b] = s->K[0] * tmp[0 + b] + s->K[2] * tmp[6 + b];
        H[3 + b] = s->K[4] * tmp[3 + b] + s->K[5] * tmp[6 + b];
        H[6 + b] = s->K[8] * tmp[6 + b];
    }
}

static void warp(const float *H, int x, int y, float *ox, float *oy) { /* ComputeCorrespondingPoint cu:281-288 */
    float X = H[0] * x + H[1] * y + H[2], Y = H[3] * x + H[4] * y + H[5], Z = H[6] * x + H[7] * y + H[8];
    *ox = X / Z; *oy = Y / Z;
}

static float bilateral_weight(float dx, float dy, float pix, float cpix, float ss, float sc) { /* cu:318-323 */
    return expf(-sqrtf(dx * dx + dy * dy) / (2.0f * ss * ss) - fabsf(pix - cpix) / (2.0f * sc * sc));
}

static float bilateral_ncc(const Pmo *P, int v, int px, int py, f4 pl, int scale) { /* cu:325-414 */
    const PmoCamera *rc = &P->cams[0], *sc = &P->cams[v];
    const float *ri = P->images[0], *si = P->images[v];
    const float cost_max = 2.0f;
    int step = 2;
    for (int i = 0; i < scale; ++i) step *= 2;
    const int radius = 5 * step / 2;
    float H[9], cx, cy;
    homography(rc, sc, pl, H);
    warp(H, px, py, &cx, &cy);
    if (cx >= sc->width || cx < 0.0f || cy >= sc->height || cy < 0.0f) return cost_max;

    float sr = 0, srr = 0, ss = 0, sss = 0, srs = 0, sw = 0;
    const float rc0 = tex2d(ri, rc->width, rc->height, px + 0.5f, py + 0.5f);
    for (int i = -radius; i < radius + 1; i += step) {
        float r_sr = 0, r_ss = 0, r_srr = 0, r_sss = 0, r_srs = 0, r_sw = 0;
        for (int j = -radius; j < radius + 1; j += step) {
            const int rx = px + i, ry = py + j;
            const float rp = tex2d(ri, rc->width, rc->height, rx + 0.5f, ry + 0.5f);
            float sx, sy;
            warp(H, rx, ry, &sx, &sy);
            const float sp = tex2d(si, sc->width, sc->height, sx + 0.5f, sy + 0.5f);
            const float w = bilateral_weight((float)i, (float)j, rp, rc0, P->params.sigma_spatial, P->params.sigma_color);
            r_sr += w * rp; r_srr += w * rp * rp; r_ss += w * sp; r_sss += w * sp * sp; r_srs += w * rp * sp; r_sw += w;
        }
        sr += r_sr; srr += r_srr; ss += r_ss; sss += r_sss; srs += r_srs; sw += r_sw;
    }
    const float inv = 1.0f / sw;
    sr *= inv; srr *= inv; ss *= inv; sss *= inv; srs *= inv;
    const float var_r = srr - sr * sr, var_s = sss - ss * ss;
    if (var_r < 1e-5f || var_s < 1e-5f) return cost_max;
    const float cov = srs - sr * ss;
    return fmaxf(0.0f, fminf(cost_max, 1.0f - cov / sqrtf(var_r * var_s)));
}

static void cost_vector(const Pmo *P, int x, int y, f4 pl, float *out, int scale) { /* cu:575-580 */
    for (int i = 1; i < P->params.num_images; ++i) out[i - 1] = bilateral_ncc(P, i, x, y, pl, scale);
}

static float initial_cost(const Pmo *P, int x, int y, f4 pl, uint32_t *sel, int scale) { /* cu:497-534 */
    const float cost_max = 2.0f;
    float cv[32] = { 2.0f }, cc[32] = { 2.0f };
    int count = 0, valid = 0;
    for (int i = 1; i < P->params.num_images; ++i) {
        float c = bilateral_ncc(P, i, x, y, pl, scale);
        cv[i - 1] = c; cc[i - 1] = c; count++;
        if (c < cost_max) valid++;
    }
    sort_small(cv, count);
    *sel = 0;
    int k = valid < P->params.top_k ? valid : P->params.top_k;
    if (k > 0) {
        float cost = 0.0f;
        for (int i = 0; i < k; ++i) cost += cv[i];
        float thr = cv[k - 1];
        for (int i = 0; i < P->params.num_images - 1; ++i) if (cc[i] <= thr) *sel |= (1u << i);
        return cost / k;
    }
    return cost_max;
}

/* ------------------------------------------------------------------ geometric consistency */
static void backproject(float x, float y, float depth, const PmoCamera *c, float *X) { /* cu:582-603 */
    float px = depth * (x - c->K[2]) / c->K[0], py = depth * (y - c->K[5]) / c->K[4], pz = depth;
    X[0] = c->R[0] * px + c->R[3] * py + c->R[6] * pz + c->C[0];
    X[1] = c->R[1] * px + c->R[4] * py + c->R[7] * pz + c->C[1];
    X[2] = c->R[2] * px + c->R[5] * py + c->R[8] * pz + c->C[2];
}
static void project(const float *X, const PmoCamera *c, float *ox, float *oy) { /* cu:605-615 */
    float tx = c->R[0] * X[0] + c->R[1] * X[1] + c->R[2] * X[2] + c->t[0];
    float ty = c->R[3] * X[0] + c->R[4] * X[1] + c->R[5] * X[2] + c->t[1];
    float tz = c->R[6] * X[0] + c->R[7] * X[1] + c->R[8] * X[2] + c->t[2];
    float d = c->K[6] * tx + c->K[7] * ty + c->K[8] * tz;
    *ox = (c->K[0] * tx + c->K[1] * ty + c->K[2] * tz) / d;
    *oy = (c->K[3] * tx + c->K[4] * ty + c->K[5] * tz) / d;
}
static float geom_cost(const Pmo *P, int v, f4 pl, int x, int y) { /* cu:617-640; v = source index 1..n-1 */
    const float max_cost = 3.0f;
    const PmoCamera *rc = &P->cams[0], *sc = &P->cams[v];
    float depth = depth_from_plane(rc, pl, x, y);
    float X[3], sx, sy;
    backproject((float)x, (float)y, depth, rc, X);
    project(X, sc, &sx, &sy);
    /* (int) of a NaN / out-of-range float is undefined in C; the device saturates (cvt.rzi), NaN -> 0 */
    float fsx = sx != sx ? 0.0f : fminf(fmaxf(sx, -2147483648.0f), 2147483520.0f);
    float fsy = sy != sy ? 0.0f : fminf(fmaxf(sy, -2147483648.0f), 2147483520.0f);
    const float sd = tex2d(P->depths[v - 1], sc->width, sc->height, (int)fsx + 0.5f, (int)fsy + 0.5f);
    if (sd == 0.0f) return max_cost;
    float Y[3], bx, by;
    backproject(sx, sy, sd, sc, Y);
    project(Y, rc, &bx, &by);
    const float dc = x - bx, dr = y - by;
    return fminf(max_cost, sqrtf(dc * dc + dr * dr));
}

/* ------------------------------------------------------------------ refinement, cu:642-722 */
static void refine(const Pmo *P, float tex_prior, f4 *plane, float *depth, float *cost, float *gcost, PmoRng *rs,
                   const float *vw, float wnorm, float *restricted_cost, int x, int y, int scale) {
    const PmoParams *pr = &P->params;
    const PmoCamera *c0 = &P->cams[0];
    const float perturbation = 0.02f;
    float depth_rand; f4 rand_n;
    const int idx = y * c0->width + x;
    const float gamma = 0.5f;
    const float depth_sigma = (pr->depth_max - pr->depth_min) / 64.0f;
    const float two_ds2 = 2 * depth_sigma * depth_sigma;
    const float angle_sigma = (float)(M_PI * (5.0f / 180.0f));
    const float two_as2 = 2 * angle_sigma * angle_sigma;
    const float beta = 0.18f;
    float depth_prior = 0.0f;
    const int has_prior = pr->planar_prior && P->mask[idx] > 0;

    if (has_prior) { /* cu:656-660: drawn, then overwritten below (the block that follows has no `else`) */
        depth_prior = depth_from_plane(c0, P->prior[idx], x, y);
        depth_rand = rng_uniform(rs) * 6 * depth_sigma + (depth_prior - 3 * depth_sigma);
        rand_n = perturbed_normal(c0, x, y, P->prior[idx], rs, angle_sigma);
    }
    depth_rand = rng_uniform(rs) * (pr->depth_max - pr->depth_min) + pr->depth_min; /* cu:661-662 */
    rand_n = random_normal(c0, x, y, rs);

    float depth_pert = *depth;
    const float dlo = (1 - perturbation) * depth_pert, dhi = (1 + perturbation) * depth_pert;
    do { /* cu:668-670: condition is an empty set unless min > max */
        depth_pert = rng_uniform(rs) * (dhi - dlo) + dlo;
    } while (depth_pert < pr->depth_min && depth_pert > pr->depth_max);
    f4 pert_n = perturbed_normal(c0, x, y, *plane, rs, (float)(perturbation * M_PI));

    const float depths[5] = { depth_rand, *depth, depth_rand, *depth, depth_pert };
    const f4 normals[5] = { *plane, rand_n, rand_n, pert_n, *plane };

    for (int i = 0; i < 5; ++i) {
        float cv[32] = { 2.0f };
        f4 tp = normals[i];
        tp.w = plane_distance(c0, x, y, depths[i], tp);
        cost_vector(P, x, y, tp, cv, scale);
        float tc = 0.0f, tg = 0.0f;
        for (int j = 0; j < pr->num_images - 1; ++j) {
            if (vw[j] > 0) {
                if (pr->geom_consistency) {
                    float g = 0.2f * geom_cost(P, j + 1, tp, x, y);
                    tc += vw[j] * (cv[j] + g);
                    tg += vw[i] * g; /* cu:689: indexed by the hypothesis, not the view */
                } else {
                    tc += vw[j] * cv[j];
                }
            }
        }
        tc /= wnorm;
        if (pr->geom_consistency) tg /= wnorm;
        float dbefore = depth_from_plane(c0, tp, x, y);
        if (has_prior) {
            float dd = depths[i] - depth_prior;
            float ad = acosf(dot3(P->prior[idx], tp));
            float prior = gamma + expf(-dd * dd / two_ds2) * expf(-ad * ad / two_as2);
            float rtc = expf(-tc * tc / beta) * prior * tex_prior;
            if (dbefore >= pr->depth_min && dbefore <= pr->depth_max && rtc > *restricted_cost) {
                *plane = tp; *cost = tc; /* restricted_cost is not refreshed, cu:707-710 */
            }
        } else if (dbefore >= pr->depth_min && dbefore <= pr->depth_max && tc < *cost) {
            *plane = tp; *cost = tc; *gcost = tg;
        }
    }
}

/* ------------------------------------------------------------------ one pixel of a half-sweep, cu:724-998 */
static const int8_t kDirs[8][12][2] = { /* cu:769-779 */
    {{-5,-6},{5,-6},{-6,-7},{6,-7},{-7,-8},{7,-8},{-8,-9},{8,-9},{-9,-10},{9,-10},{-10,-11},{10,-11}},
    {{-5,6},{5,6},{-6,7},{6,7},{-7,8},{7,8},{-8,9},{8,9},{-9,10},{9,10},{-10,11},{10,11}},
    {{-6,-5},{-6,5},{-7,-6},{-7,6},{-8,-7},{-8,7},{-9,-8},{-9,8},{-10,-9},{-10,9},{-11,-10},{-11,10}},
    {{6,-5},{6,5},{7,-6},{7,6},{8,-7},{8,7},{9,-8},{9,8},{10,-9},{10,9},{11,-10},{11,10}},
    {{0,-5},{0,-7},{0,-9},{0,-11},{0,-13},{0,-15},{0,-17},{0,-19},{0,-21},{0,-23},{0,0},{0,0}},
    {{0,5},{0,7},{0,9},{0,11},{0,13},{0,15},{0,17},{0,19},{0,21},{0,23},{0,0},{0,0}},
    {{-5,0},{-7,0},{-9,0},{-11,0},{-13,0},{-15,0},{-17,0},{-19,0},{-21,0},{-23,0},{0,0},{0,0}},
    {{5,0},{7,0},{9,0},{11,0},{13,0},{15,0},{17,0},{19,0},{21,0},{23,0},{0,0},{0,0}} };
static const int kNumDirs[8] = { 12, 12, 12, 12, 10, 10, 10, 10 };

static void propagate_pixel(Pmo *P, int x, int y, int iter, int scale) {
    const PmoParams *pr = &P->params;
    const PmoCamera *c0 = &P->cams[0];
    const int width = c0->width, height = c0->height;
    if (x >= width || y >= height) return;
    const int idx = y * width + x, nsrc = pr->num_images - 1;
    PmoRng *rs = &P->rng[idx];
    const int nbr[4] = { idx - width, idx + width, idx - 1, idx + 1 };
    int pos[8];
    float ca[8][32];
    memset(ca, 0, sizeof ca); ca[0][0] = 2.0f;               /* `= {2.0f}`, cu:795 */
    int flag[8] = { 0 };

    for (int r = 0; r < 8; ++r) {                            /* cu:798-819 */
        int bx = 0, by = 0; float best = FLT_MAX;
        for (int k = 0; k < kNumDirs[r]; ++k) {
            int nx = x + kDirs[r][k][0], ny = y + kDirs[r][k][1];
            if (!(nx >= 0 && ny >= 0 && nx < width && ny < height)) continue;
            float nc = P->costs[ny * width + nx];
            if (best > nc) { bx = nx; by = ny; best = nc; }
        }
        if (best < FLT_MAX) {
            flag[r] = 1; pos[r] = by * width + bx;
            cost_vector(P, x, y, P->planes[pos[r]], ca[r], scale);
        }
    }

    float vw[32] = { 0.0f };                                  /* cu:821-867 */
    {
        float prior[32] = { 0.0f };
        for (int i = 0; i < 4; ++i)
            if (flag[i])
                for (int j = 0; j < nsrc; ++j) prior[j] += ((P->views[nbr[i]] >> j) & 1u) ? 0.9f : 0.1f;
        float probs[32] = { 0.0f };
        const float thr = (float)(0.8 * expf((iter) * (iter) / (-90.0f)));
        for (int i = 0; i < nsrc; i++) {
            float count = 0; int count_false = 0; float tmpw = 0;
            for (int j = 0; j < 8; j++) {
                if (ca[j][i] < thr) { tmpw += expf(ca[j][i] * ca[j][i] / (-0.18f)); count++; }
                if (ca[j][i] > 1.2f) count_false++;
            }
            if (count > 2 && count_false < 3) probs[i] = prior[i] * tmpw / count;
            else if (count_false < 3) probs[i] = prior[i] * expf(thr * thr / (-0.32f));
            else probs[i] = 0;
        }
        pdf_to_cdf(probs, nsrc);
        for (int s = 0; s < 15; ++s) {
            const float u = rng_uniform(rs) - FLT_EPSILON;
            for (int v = 0; v < nsrc; ++v) if (probs[v] > u) { vw[v] += 1.0f; break; }
        }
    }

    uint32_t tsel = 0; float wnorm = 0;                       /* cu:869-878 */
    for (int i = 0; i < nsrc; ++i) if (vw[i] > 0) { tsel |= (1u << i); wnorm += vw[i]; }

    float fc[8] = { 0.0f };                                   /* cu:880-898 */
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < nsrc; ++j) {
            if (vw[j] > 0) {
                if (pr->geom_consistency) {
                    if (flag[i]) fc[i] += vw[j] * (ca[i][j] + 0.2f * geom_cost(P, j + 1, P->planes[pos[i]], x, y));
                    else fc[i] += vw[j] * (ca[i][j] + 0.1f * 3.0f);
                } else fc[i] += vw[j] * ca[i][j];
            }
        }
        fc[i] /= wnorm;
    }
    const int kmin = arg_min_last(fc, 8);
    float cost_now = 0.0f, g_now = 0.0f;                      /* cu:900-920 */
    {
        float cn[32] = { 2.0f };
        cost_vector(P, x, y, P->planes[idx], cn, scale);
        for (int i = 0; i < nsrc; ++i) {
            if (pr->geom_consistency) {
                float g = 0.2f * geom_cost(P, i + 1, P->planes[idx], x, y);
                cost_now += vw[i] * (cn[i] + g);
                g_now += vw[i] * g;
            } else cost_now += vw[i] * cn[i];
        }
    }
    cost_now /= wnorm;
    if (pr->geom_consistency) { g_now /= wnorm; P->geom[idx] = g_now; }
    P->costs[idx] = cost_now;
    float depth_now = depth_from_plane(c0, P->planes[idx], x, y);
    float restricted_cost = 0.0f;
    const float tex_prior = 1.0f;
    if (pr->planar_prior && !pr->geom_consistency) {          /* cu:924-978 */
        float rfc[8] = { 0.0f };
        const float gamma = 0.5f;
        const float depth_sigma = (pr->depth_max - pr->depth_min) / 64.0f;
        const float two_ds2 = 2 * depth_sigma * depth_sigma;
        const float angle_sigma = (float)(M_PI * (5.0f / 180.0f));
        const float two_as2 = 2 * angle_sigma * angle_sigma;
        const float depth_prior = depth_from_plane(c0, P->prior[idx], x, y);
        const float beta = 0.18f;
        if (P->mask[idx] > 0) {
            for (int i = 0; i < 8; i++) {
                if (flag[i]) {
                    float dn = depth_from_plane(c0, P->planes[pos[i]], x, y);
                    float dd = dn - depth_prior;
                    float ad = acosf(dot3(P->prior[idx], P->planes[pos[i]]));
                    float prior = gamma + tex_prior * expf(-dd * dd / two_ds2) * expf(-ad * ad / two_as2);
                    rfc[i] = expf(-fc[i] * fc[i] / beta) * prior;
                }
            }
            const int kmax = arg_max_last(rfc, 8);
            float dn_shadow = depth_from_plane(c0, P->planes[idx], x, y); /* cu:950 shadows the outer depth_now */
            float dd = dn_shadow - depth_prior;
            float ad = acosf(dot3(P->prior[idx], P->planes[idx]));
            float prior = gamma + tex_prior * expf(-dd * dd / two_ds2) * expf(-ad * ad / two_as2);
            float rc_now = expf(-cost_now * cost_now / beta) * prior;
            if (flag[kmax]) {
                float db = depth_from_plane(c0, P->planes[pos[kmax]], x, y);
                if (db >= pr->depth_min && db <= pr->depth_max && rfc[kmax] > rc_now) {
                    dn_shadow = db; (void)dn_shadow;          /* cu:961 writes the shadow; outer depth_now keeps the old depth */
                    P->planes[idx] = P->planes[pos[kmax]];
                    P->costs[idx] = fc[kmax];
                    restricted_cost = rfc[kmax];
                    P->views[idx] = tsel;
                }
            }
        } else if (flag[kmin]) {
            float db = depth_from_plane(c0, P->planes[pos[kmin]], x, y);
            if (db >= pr->depth_min && db <= pr->depth_max && fc[kmin] < cost_now) {
                depth_now = db;
                P->planes[idx] = P->planes[pos[kmin]];
                P->costs[idx] = fc[kmin];                     /* cost_now itself is not updated, cu:972-976 */
            }
        }
    }

    f4 plane_now = P->planes[idx];                            /* cu:981-991 */
    if (!pr->planar_prior && flag[kmin]) {
        float db = depth_from_plane(c0, P->planes[pos[kmin]], x, y);
        if (db >= pr->depth_min && db <= pr->depth_max && fc[kmin] < cost_now) {
            depth_now = db; plane_now = P->planes[pos[kmin]]; cost_now = fc[kmin];
            P->views[idx] = tsel;
        }
    }
    refine(P, tex_prior, &plane_now, &depth_now, &cost_now, &g_now, rs, vw, wnorm, &restricted_cost, x, y, scale);
    P->costs[idx] = cost_now;                                 /* cu:993-997 */
    P->planes[idx] = plane_now;
    if (pr->geom_consistency) P->geom[idx] = g_now;
}

/* launch geometry of BlackPixelUpdate / RedPixelUpdate, cu:1000-1019 with the grids of cu:1192-1196:
 * thread (tx, ty) with tx < 32*ceil(W/32), ty < 16*ceil((H/2)/16) handles row 2*ty + (tx&1) (black) or
 * 2*ty + 1 - (tx&1) (red) -- so for odd H the last row is never visited. */
typedef struct { Pmo *P; int red, iter, scale; uint64_t seed; const f4 *pl; float *out; } SweepCtx;
static void half_sweep_row(void *c, int ty) {
    SweepCtx *s = (SweepCtx *)c;
    for (int x = 0; x < s->P->width; ++x) {
        int y = s->red ? ((x & 1) ? 2 * ty : 2 * ty + 1) : ((x & 1) ? 2 * ty + 1 : 2 * ty);
        propagate_pixel(s->P, x, y, s->iter, s->scale);
    }
}
static void half_sweep(Pmo *P, int red, int iter, int scale) {
    const int ty_end = 16 * (((P->height / 2) + 15) / 16);
    /* pixels of one colour only read pixels of the other colour (all offsets have odd |dx|+|dy|), so
     * the order within a colour does not matter; rows are independent -> parallel over rows. */
    SweepCtx c = { P, red, iter, scale, 0, NULL, NULL };
    parallel_rows(half_sweep_row, &c, ty_end);
}

/* InitializeScore, cu:536-573 */
static void initialize_row(void *cv_, int y) {
    SweepCtx *sc = (SweepCtx *)cv_;
    Pmo *P = sc->P;
    const uint64_t seed = sc->seed; const int scale = sc->scale;
    const PmoParams *pr = &P->params;
    const PmoCamera *c0 = &P->cams[0];
    const int W = P->width;
    {
        for (int x = 0; x < W; ++x) {
            const int idx = y * W + x;
            rng_init(&P->rng[idx], mix_seed(seed, (uint32_t)x, (uint32_t)y));
            if (!pr->geom_consistency && !pr->planar_prior) {
                P->planes[idx] = random_plane(c0, x, y, &P->rng[idx], pr->depth_min, pr->depth_max);
            } else if (pr->planar_prior && P->mask[idx] > 0 && P->costs[idx] >= 0.1f) {
                const float perturbation = 0.02f;
                f4 ph = P->prior[idx];
                float dp = ph.w;
                const float lo = (1 - 3 * perturbation) * dp, hi = (1 + 3 * perturbation) * dp;
                dp = rng_uniform(&P->rng[idx]) * (hi - lo) + lo;
                f4 pp = perturbed_normal(c0, x, y, ph, &P->rng[idx], (float)(3 * perturbation * M_PI));
                pp.w = dp;
                P->planes[idx] = pp;
            } else {
                f4 ph = normal_world_to_cam(c0, P->planes[idx]);
                float depth = ph.w;
                ph.w = plane_distance(c0, x, y, depth, ph);
                P->planes[idx] = ph;
            }
            P->costs[idx] = initial_cost(P, x, y, P->planes[idx], &P->views[idx], scale);
        }
    }
}
static void initialize_score(Pmo *P, uint64_t seed, int scale) {
    SweepCtx c = { P, 0, 0, scale, seed, NULL, NULL };
    parallel_rows(initialize_row, &c, P->height);
}

/* GetDepthandNormal, cu:1021-1034 */
static void depth_and_normal(Pmo *P) {
    const int W = P->width, H = P->height;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int idx = y * W + x;
            P->planes[idx].w = depth_from_plane(&P->cams[0], P->planes[idx], x, y);
            P->planes[idx] = normal_cam_to_world(&P->cams[0], P->planes[idx]);
        }
}

/* CheckerboardFilter, cu:1036-1150 (bounds tests kept literally, they are asymmetric) */
static void filter_pixel(Pmo *P, int x, int y) {
    const int width = P->width, height = P->height;
    if (x >= width || y >= height) return;
    const int c = y * width + x;
    float f[21]; int n = 0;
    const f4 *pl = P->planes;
    f[n++] = pl[c].w;
    const int left = c - 1, leftleft = c - 3, up = c - width, upup = c - 3 * width;
    const int down = c + width, downdown = c + 3 * width, right = c + 1, rightright = c + 3;
    if (P->costs[c] < 0.001f) return;
    if (y > 0) f[n++] = pl[up].w;
    if (y > 2) f[n++] = pl[upup].w;
    if (y > 4) f[n++] = pl[upup - width * 2].w;
    if (y < height - 1) f[n++] = pl[down].w;
    if (y < height - 3) f[n++] = pl[downdown].w;
    if (y < height - 5) f[n++] = pl[downdown + width * 2].w;
    if (x > 0) f[n++] = pl[left].w;
    if (x > 2) f[n++] = pl[leftleft].w;
    if (x > 4) f[n++] = pl[leftleft - 2].w;
    if (x < width - 1) f[n++] = pl[right].w;
    if (x < width - 3) f[n++] = pl[rightright].w;
    if (x < width - 5) f[n++] = pl[rightright + 2].w;
    if (y > 0 && x < width - 2) f[n++] = pl[up + 2].w;
    if (y < height - 1 && x < width - 2) f[n++] = pl[down + 2].w;
    if (y > 0 && x > 1) f[n++] = pl[up - 2].w;
    if (y < height - 1 && x > 1) f[n++] = pl[down - 2].w;
    if (x > 0 && y > 2) f[n++] = pl[left - width * 2].w;
    if (x < width - 1 && y > 2) f[n++] = pl[right - width * 2].w;
    if (x > 0 && y < height - 2) f[n++] = pl[left + width * 2].w;
    if (x < width - 1 && y < height - 2) f[n++] = pl[right + width * 2].w;
    sort_small(f, n);
    int m = n / 2;
    P->planes[c].w = (n % 2 == 0) ? (f[m - 1] + f[m]) / 2 : f[m];
}
static void filter_sweep(Pmo *P, int red) { /* cu:1152-1174 */
    const int W = P->width, H = P->height;
    const int ty_end = 16 * (((H / 2) + 15) / 16);
    for (int ty = 0; ty < ty_end; ++ty)
        for (int x = 0; x < W; ++x) {
            int y = red ? ((x & 1) ? 2 * ty : 2 * ty + 1) : ((x & 1) ? 2 * ty + 1 : 2 * ty);
            filter_pixel(P, x, y);
        }
}

/* ------------------------------------------------------------------ API (same shape as oracle/ref_harness.cu) */
void *pmo_create(void) {
    Pmo *P = (Pmo *)calloc(1, sizeof(Pmo));
    P->params.max_iterations = 3; P->params.sigma_spatial = 5.0f; P->params.sigma_color = 3.0f;
    P->params.top_k = 4; P->params.max_scale = 2; P->params.depth_min = 0.0f; P->params.depth_max = 1.0f;
    return P;
}
int pmo_sizeof_camera(void) { return (int)sizeof(PmoCamera); }

int pmo_set_problem(void *h, int n, const float *const *images, const void *cams) {
    Pmo *P = (Pmo *)h;
    if (n < 2 || n > 33) return -1;
    memcpy(P->cams, cams, sizeof(PmoCamera) * n);
    P->n = n; P->width = P->cams[0].width; P->height = P->cams[0].height;
    P->params.depth_min = P->cams[0].depth_min * 0.6f;   /* PatchMatch.cpp:929-930 */
    P->params.depth_max = P->cams[0].depth_max * 1.2f;
    P->params.num_images = n;
    for (int i = 0; i < n; ++i) {
        size_t sz = (size_t)P->cams[i].width * P->cams[i].height;
        P->images[i] = (float *)malloc(sz * sizeof(float));
        memcpy(P->images[i], images[i], sz * sizeof(float));
    }
    size_t wh = (size_t)P->width * P->height;
    P->planes = (f4 *)calloc(wh, sizeof(f4)); P->costs = (float *)calloc(wh, sizeof(float));
    P->rng = (PmoRng *)calloc(wh, sizeof(PmoRng)); P->views = (uint32_t *)calloc(wh, sizeof(uint32_t));
    P->geom = (float *)calloc(wh, sizeof(float));
    P->prior = (f4 *)calloc(wh, sizeof(f4)); P->mask = (uint32_t *)calloc(wh, sizeof(uint32_t));
    return 0;
}
void pmo_set_geom_consistency_params(void *h, int geom, int planar) { /* PatchMatch.cpp:655-665 */
    PmoParams *p = &((Pmo *)h)->params;
    p->geom_consistency = geom != 0;
    if (geom) { p->max_iterations = 2; p->geomPlanarPrior = planar != 0; } else p->max_iterations = 3;
}
void pmo_set_planar_prior_params(void *h) { ((Pmo *)h)->params.planar_prior = 1; } /* PatchMatch.cpp:667-670 */

int pmo_set_src_depths(void *h, const float *const *depths) {
    Pmo *P = (Pmo *)h;
    for (int i = 0; i < P->n - 1; ++i) {
        size_t sz = (size_t)P->cams[i + 1].width * P->cams[i + 1].height;
        free(P->depths[i]);
        P->depths[i] = (float *)malloc(sz * sizeof(float));
        memcpy(P->depths[i], depths[i], sz * sizeof(float));
    }
    return 0;
}
int pmo_set_state(void *h, const float *planes4, const float *costs) {
    Pmo *P = (Pmo *)h; size_t wh = (size_t)P->width * P->height;
    memcpy(P->planes, planes4, wh * sizeof(f4)); memcpy(P->costs, costs, wh * sizeof(float));
    return 0;
}
int pmo_set_prior(void *h, const float *prior4, const uint32_t *mask) {
    Pmo *P = (Pmo *)h; size_t wh = (size_t)P->width * P->height;
    memcpy(P->prior, prior4, wh * sizeof(f4)); memcpy(P->mask, mask, wh * sizeof(uint32_t));
    return 0;
}
int pmo_init_only(void *h, uint64_t seed) { Pmo *P = (Pmo *)h; initialize_score(P, seed, P->params.max_scale); return 0; }
int pmo_half_sweep(void *h, int red, int iter, int scale) { half_sweep((Pmo *)h, red, iter, scale); return 0; }
int pmo_finalize(void *h) { Pmo *P = (Pmo *)h; depth_and_normal(P); filter_sweep(P, 0); filter_sweep(P, 1); return 0; }

/* PatchMatchCUDA::Run, cu:1188-1254 */
float pmo_run(void *h, uint64_t seed) {
    Pmo *P = (Pmo *)h;
    const PmoParams *pr = &P->params;
    initialize_score(P, seed, pr->max_scale);
    if (pr->geom_consistency || pr->planar_prior) {
        for (int i = 0; i < pr->max_iterations; ++i) { half_sweep(P, 0, i, 0); half_sweep(P, 1, i, 0); }
    } else {
        for (int s = pr->max_scale; s >= 0; --s)
            for (int i = 0; i < pr->max_iterations; ++i) { half_sweep(P, 0, i, s); half_sweep(P, 1, i, s); }
    }
    pmo_finalize(P);
    return 0.0f;
}
int pmo_get_result(void *h, float *planes4, float *costs, float *geom) {
    Pmo *P = (Pmo *)h; size_t wh = (size_t)P->width * P->height;
    if (planes4) memcpy(planes4, P->planes, wh * sizeof(f4));
    if (costs) memcpy(costs, P->costs, wh * sizeof(float));
    if (geom) memcpy(geom, P->geom, wh * sizeof(float));
    return 0;
}
float pmo_depth_min(void *h) { return ((Pmo *)h)->params.depth_min; }
float pmo_depth_max(void *h) { return ((Pmo *)h)->params.depth_max; }

int pmo_get_device_state(void *h, float *planes4, float *costs, uint32_t *views, uint32_t *rng6, float *geom) {
    Pmo *P = (Pmo *)h; size_t wh = (size_t)P->width * P->height;
    pmo_get_result(h, planes4, costs, geom);
    if (views) memcpy(views, P->views, wh * sizeof(uint32_t));
    if (rng6) memcpy(rng6, P->rng, wh * sizeof(PmoRng));
    return 0;
}
int pmo_set_device_state(void *h, const float *planes4, const float *costs, const uint32_t *views, const uint32_t *rng6,
                         const float *geom) {
    Pmo *P = (Pmo *)h; size_t wh = (size_t)P->width * P->height;
    if (planes4) memcpy(P->planes, planes4, wh * sizeof(f4));
    if (costs) memcpy(P->costs, costs, wh * sizeof(float));
    if (views) memcpy(P->views, views, wh * sizeof(uint32_t));
    if (rng6) memcpy(P->rng, rng6, wh * sizeof(PmoRng));
    if (geom) memcpy(P->geom, geom, wh * sizeof(float));
    return 0;
}
static void ncc_map_row(void *cv_, int y) {
    SweepCtx *s = (SweepCtx *)cv_;
    Pmo *P = s->P; const int W = P->width, H = P->height;
    for (int x = 0; x < W; ++x)
        for (int v = 1; v < P->n; ++v)
            s->out[(size_t)(v - 1) * W * H + (size_t)y * W + x] = bilateral_ncc(P, v, x, y, s->pl[y * W + x], s->scale);
}
int pmo_ncc_map(void *h, const float *planes4, int scale, float *out) {
    Pmo *P = (Pmo *)h;
    SweepCtx c = { P, 0, 0, scale, 0, (const f4 *)planes4, out };
    parallel_rows(ncc_map_row, &c, P->height);
    return 0;
}
int pmo_geom_map(void *h, const float *planes4, float *out) {
    Pmo *P = (Pmo *)h; const int W = P->width, H = P->height; const f4 *pl = (const f4 *)planes4;
    if (!P->depths[0]) return -1;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int v = 1; v < P->n; ++v)
                out[(size_t)(v - 1) * W * H + (size_t)y * W + x] = geom_cost(P, v, pl[y * W + x], x, y);
    return 0;
}
int pmo_uniform_stream(uint64_t seed, int x, int y, int n, float *out) {
    PmoRng s; rng_init(&s, mix_seed(seed, (uint32_t)x, (uint32_t)y));
    for (int i = 0; i < n; ++i) out[i] = rng_uniform(&s);
    return 0;
}
float pmo_tex2d(const float *img, int w, int h, float x, float y) { return tex2d(img, w, h, x, y); }
void pmo_destroy(void *h) {
    Pmo *P = (Pmo *)h;
    for (int i = 0; i < 33; ++i) free(P->images[i]);
    for (int i = 0; i < 32; ++i) free(P->depths[i]);
    free(P->planes); free(P->costs); free(P->rng); free(P->views); free(P->geom); free(P->prior); free(P->mask);
    free(P);
}

/* ------------------------------------------------------------------ host planar-prior stage (single-threaded, like the reference)
 * Restatement of the host code ProcessProblem runs between the two Run()s (/root/reference/src/PatchMatch.cpp:532-609):
 * GetTriangulateVertices (:782-853), the rasterisation loop (:554-579), GetPriorPlaneParams (:723-755), the depth-range
 * check (:583-595) and the per-pixel expansion of CudaPlanarPriorInitialization (:984-993). The triangulation itself is
 * NOT here: callers use cv2.Subdiv2D, the reference's own triangulator (:766-771). Used by the tests (small cases, against
 * oracle/prior_oracle.py) and by bench.py's reference arm as the host stage of the reference pipeline, timed at -O2 on one
 * core. cv::SVD::solveZ of the 3x4 system [X 1] is restated as the plane through the three points (its null vector). */
int pmo_pick_vertices(const float *costs, const float *geom, int w, int h, int geom_variant, int *xy_out) {
    int n = 0;
    for (int row = 0; row < h; row += 5)
        for (int col = 0; col < w; col += 5) {
            const int c_bound = w < col + 5 ? w : col + 5, r_bound = h < row + 5 ? h : row + 5;
            if (!geom_variant) {
                float min_cost = 2.0f; int bx = 0, by = 0;
                for (int r = row; r < r_bound; ++r)
                    for (int c = col; c < c_bound; ++c) {
                        const float cost = costs[r * w + c];
                        if (cost < 2.0f && min_cost > cost) { bx = c; by = r; min_cost = cost; }
                    }
                if (min_cost < 0.1f) { xy_out[2 * n] = bx; xy_out[2 * n + 1] = by; ++n; }
            } else {
                float mc[3] = { 2.0f, 2.0f, 2.0f }; int px[3] = { 0, 0, 0 }, py[3] = { 0, 0, 0 };
                float cost_sum = 0.0f;
                for (int r = row; r < r_bound; ++r)
                    for (int c = col; c < c_bound; ++c) {
                        const float cost = costs[r * w + c];
                        cost_sum += cost;
                        if (cost < 1.0f && geom[r * w + c] < 0.4f && cost < mc[2]) {
                            mc[2] = cost; px[2] = c; py[2] = r;
                            for (int i = 1; i >= 0; --i) {
                                if (mc[i] <= mc[i + 1]) break;
                                float t = mc[i + 1]; mc[i + 1] = mc[i]; mc[i] = t;
                                int tx = px[i + 1]; px[i + 1] = px[i]; px[i] = tx;
                                int ty = py[i + 1]; py[i + 1] = py[i]; py[i] = ty;
                            }
                        }
                    }
                cost_sum = cost_sum / (r_bound * c_bound) * 0.85; /* :841, absolute coordinates in the divisor */
                const float thresh = cost_sum > 0.2f ? cost_sum : 0.2f;
                for (int i = 0; i < 3; ++i) {
                    if (mc[i] < thresh) { xy_out[2 * n] = px[i]; xy_out[2 * n + 1] = py[i]; ++n; } else break;
                }
            }
        }
    return n;
}

/* tris: n_tris x 3 x (x, y) pixel coordinates. planes4: (h, w, 4) result of the previous Run (w = depth).
 * tri_planes4: the plane of every triangle (n_tris x 4) as the caller fitted it -- the reference's own fit is cv::SVD::solveZ
 * (:743), which lives in OpenCV, so the faithful caller passes OpenCV's results -- or NULL for the closed form below (the
 * plane through the three points in double: the same null space, without the float32 SVD's rounding noise; what bench.py's
 * reference arm times, in the reference's favour).
 * Writes prior4 (h, w, 4) and mask (h, w); returns the number of prior pixels. */
int pmo_prior_from_triangles_fit(const int *tris, int n_tris, const float *tri_planes4, const float *planes4, const float *K, int w, int h,
                                 float depth_min, float depth_max, float *prior4, uint32_t *mask) {
    float *fmask = (float *)calloc((size_t)w * h, sizeof(float));
    f4 *tp = (f4 *)malloc(sizeof(f4) * (size_t)(n_tris > 0 ? n_tris : 1));
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int t = 0; t < n_tris; ++t) {
        const int *T = tris + 6 * t;
        const int x1 = T[0], y1 = T[1], x2 = T[2], y2 = T[3], x3 = T[4], y3 = T[5];
        float L01 = sqrt(pow(x1 - x2, 2) + pow(y1 - y2, 2)), L02 = sqrt(pow(x1 - x3, 2) + pow(y1 - y3, 2));
        float L12 = sqrt(pow(x2 - x3, 2) + pow(y2 - y3, 2));
        float max_edge = L01 > (L02 > L12 ? L02 : L12) ? L01 : (L02 > L12 ? L02 : L12);
        float step = 1.0 / max_edge;
        for (float p = 0; p < 1.0; p += step)
            for (float q = 0; q < 1.0 - p; q += step) {
                int x = p * x1 + q * x2 + (1.0 - p - q) * x3;
                int y = p * y1 + q * y2 + (1.0 - p - q) * y3;
                if (x >= 0 && x < w && y >= 0 && y < h) fmask[y * w + x] = t + 1.0;
            }
        double X[3][3];
        for (int k = 0; k < 3; ++k) { /* Get3DPointonRefCam, :200-209 */
            const int x = T[2 * k], y = T[2 * k + 1];
            const float depth = planes4[((size_t)y * w + x) * 4 + 3];
            X[k][0] = depth * (x - cx) / fx; X[k][1] = depth * (y - cy) / fy; X[k][2] = depth;
        }
        const double ux = X[1][0] - X[0][0], uy = X[1][1] - X[0][1], uz = X[1][2] - X[0][2];
        const double vx = X[2][0] - X[0][0], vy = X[2][1] - X[0][1], vz = X[2][2] - X[0][2];
        double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
        double nn = sqrt(nx * nx + ny * ny + nz * nz), d = -(nx * X[0][0] + ny * X[0][1] + nz * X[0][2]);
        if (d < 0) nn = -nn; /* :746-749 */
        tp[t].x = (float)(nx / nn); tp[t].y = (float)(ny / nn); tp[t].z = (float)(nz / nn); tp[t].w = (float)(d / nn);
        if (tri_planes4) memcpy(&tp[t], tri_planes4 + 4 * (size_t)t, sizeof(f4));
    }
    int count = 0;
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
            const size_t idx = (size_t)j * w + i;
            mask[idx] = 0;
            if (fmask[idx] > 0) {
                const f4 n4 = tp[(int)fmask[idx] - 1];
                float d = -n4.w * fx / ((i - cx) * n4.x + (fx / fy) * (j - cy) * n4.y + fx * n4.z); /* :650-653 */
                if (d <= depth_max && d >= depth_min) {
                    mask[idx] = (uint32_t)fmask[idx];
                    memcpy(prior4 + 4 * idx, &n4, sizeof(f4));
                    ++count;
                }
            }
        }
    free(fmask); free(tp);
    return count;
}

int pmo_prior_from_triangles(const int *tris, int n_tris, const float *planes4, const float *K, int w, int h, float depth_min,
                             float depth_max, float *prior4, uint32_t *mask) {
    return pmo_prior_from_triangles_fit(tris, n_tris, NULL, planes4, K, w, h, depth_min, depth_max, prior4, mask);
}
