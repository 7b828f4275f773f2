/* fgoicp_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's Go-ICP hot path.
 *
 * This file is the parity oracle for the CUDA kernels.  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product path (fast_go_icp_b200/csrc) never links or calls anything in oracle/.
 *
 * PARITY STATUS: PINNED against outputs of the reference itself.  The reference ships no tests, golden
 * vectors or expected outputs of its own and has no CPU path, so the pins are:
 *   (1) golden vectors produced by the UNMODIFIED reference sources (compiled by oracle/build_ref.py into
 *       oracle/_ref/, run on a B200; generator scripts committed next to the vectors):
 *       tests/golden/reference_small.npz (synthetic pair) and tests/golden/reference_clouds.npz (subsamples of
 *       the reference repository's own bunny, skull, dragon and partial-overlap pairs) -- preprocessing and
 *       every grid cell bit-exact, bounds / SSE / ICP / inner searches / run() within the tolerances stated in
 *       tests/test_golden.py and tests/test_golden_clouds.py (CPU suite);
 *   (2) the floating-point association of every device expression below was read off the SASS that nvcc 12.9
 *       emits for the unmodified reference kernels on sm_100 (DESIGN.md "Canonical arithmetic") and is written
 *       here with explicit fmaf() and -ffp-contract=off;
 *   (3) on a GPU box, tests also compare the CUDA path AND this oracle against oracle/_ref live (real tex3D,
 *       real kernels);
 *   (4) independent cross-checks in tests (scipy cKDTree, numpy SVD/Kabsch, numpy trilinear);
 *   (5) the nearest-neighbour search and the grid build exist twice here -- the literal scans of the reference's
 *       kernels and an exact k-d tree (default, so that the oracle answers at full size) -- and are tested to agree
 *       bit for bit (tests/test_oracle.py); driven through the level-synchronous driver the oracle reproduces the
 *       CUDA path's whole search at full size bit for bit (tests/test_fullsize_parity.py).
 * The trimming switch (orc_set_trim_k) is an EXTENSION with no reference behaviour: it is pinned by a numpy
 * restatement only (tests/test_trimming.py) and is off by default.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference repo).
 * Matrices are 9 floats, column-major (glm::mat3 memory order): M[c*3+r].
 * Reductions over points are accumulated in double and rounded to float once: the reference's
 * fp32 CUB tree order is unspecified and not reproducible, see DESIGN.md.
 *
 * Build: oracle/build_oracle.py  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "svd3.h"

#define ORC_PI    3.141592653589793f   /* common.hpp:17 */
#define ORC_INF   1E+10f               /* common.hpp:18 */
#define ORC_SQRT3 1.732050807568877f   /* common.hpp:19 */

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* Tunables pinned empirically against tex3D on a B200 (tests/test_gpu_texture_conformance.py) */
/* ------------------------------------------------------------------------------------------ */
static int g_weight_mode = 0;   /* 0: round-half-up 1.8 fixed point (B200 hardware), 1: truncate, 2: half-even */
static int g_interp_mode = 0;   /* 0: nested lerps x,y,z with fmaf, 1: 8-term weighted sum      */
static int g_sin_mode = 0;      /* 0: host sinf(), 1: table lookup of device values (see below) */
static int g_sin_n = 0;
static float g_sin_span[16];
static float g_sin_val[16];
/* Trimming (EXTENSION; the reference parses `trim` and ignores it): when 0 < g_trim_k < ns every
 * sum over the data points -- per-cube upper/lower bounds, the exact SSE, the ICP's centroids and cross-covariance --
 * runs over the g_trim_k points with the SMALLEST residual only (Go-ICP's trimmed registration).  0 = off = the
 * reference's behaviour, bit for bit. */
static size_t g_trim_k = 0;

ORC_API void orc_set_trim_k(size_t k) { g_trim_k = k; }
ORC_API size_t orc_get_trim_k(void) { return g_trim_k; }
/* inliers kept for a trim fraction rho: ns - floor(float(ns) * rho), fp32 product */
ORC_API size_t orc_trim_count(size_t ns, float rho)
{
    size_t drop = (size_t)((float)ns * rho);
    return drop >= ns ? 1 : ns - drop;
}

static int orc_cmp_float(const void* a, const void* b)
{
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* sum of the k smallest of v[0..n) (v is clobbered), fp64 accumulation in ascending order of value */
static double orc_trimmed_sum(float* v, size_t n, size_t k)
{
    double s = 0.0;
    size_t i;
    if (k == 0 || k >= n) { for (i = 0; i < n; ++i) s += (double)v[i]; return s; }
    qsort(v, n, sizeof(float), orc_cmp_float);
    for (i = 0; i < k; ++i) s += (double)v[i];
    return s;
}

ORC_API void orc_set_modes(int weight_mode, int interp_mode)
{
    g_weight_mode = weight_mode;
    g_interp_mode = interp_mode;
}

/* Device sinf() differs from glibc's in the last ulp for some arguments; a GPU test reads the
 * device values (fgoicp_rot_sin) and installs them here so both sides use identical constants. */
ORC_API void orc_set_sin_table(const float* spans, const float* vals, int n)
{
    int i;
    if (n > 16) n = 16;
    for (i = 0; i < n; ++i) { g_sin_span[i] = spans[i]; g_sin_val[i] = vals[i]; }
    g_sin_n = n;
    g_sin_mode = n > 0;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Canonical device arithmetic (association read from the reference's sm_100 SASS)            */
/* ------------------------------------------------------------------------------------------ */

/* dx*dx + dy*dy + dz*dz  (registration.cu:154-160, 248-254; glm::dot in icp3d.cu:20).
 * SASS of every reference kernel: FMUL dy,dy ; FFMA dx,dx,. ; FFMA dz,dz,.  -- nvcc keeps the SECOND
 * product as the plain multiply and fuses the first and third. */
static inline float orc_sq3(float dx, float dy, float dz)
{
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* glm mat3 * vec3 + t in device code (registration.cu:20, 34; icp3d.cu:35).
 * SASS: FMUL m1r,p.y ; FFMA m0r,p.x,. ; FFMA m2r,p.z,. ; FADD t  (same association as orc_sq3). */
static inline void orc_xform(const float* R, const float* t, const float* p, float* q)
{
    int r;
    for (r = 0; r < 3; ++r)
        q[r] = fmaf(R[6 + r], p[2], fmaf(R[r], p[0], R[3 + r] * p[1])) + t[r];
}

/* ------------------------------------------------------------------------------------------ */
/* Rotation cube helpers (host arithmetic in the reference: no contraction)                    */
/* ------------------------------------------------------------------------------------------ */

/* struct Rotation ctor, common.hpp:37-57.  Returns r (|q| inside the ball, |q|^2 outside). */
ORC_API float orc_rotation(float x, float y, float z, float* R)
{
    float r = x * x + y * y + z * z;
    int i;
    for (i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    if (r > 1.0f) return r;
    {
        float ww = 1.0f - r;
        float w = sqrtf(ww);
        float wx = w * x, xx = x * x;
        float wy = w * y, xy = x * y, yy = y * y;
        float wz = w * z, xz = x * z, yz = y * z, zz = z * z;
        /* glm::mat3(a,b,c, d,e,f, g,h,i) fills COLUMNS (common.hpp:50-54) */
        R[0] = ww + xx - yy - zz; R[1] = 2 * (xy - wz);     R[2] = 2 * (xz + wy);
        R[3] = 2 * (xy + wz);     R[4] = ww - xx + yy - zz; R[5] = 2 * (yz - wx);
        R[6] = 2 * (xz - wy);     R[7] = 2 * (yz + wx);     R[8] = ww - xx - yy + zz;
    }
    return sqrtf(r);
}

/* RotNode::overlaps_SO3, common.hpp:99-103 (uses Rotation::r raw). */
ORC_API int orc_overlaps_so3(float x, float y, float z, float span)
{
    float R[9];
    float r = orc_rotation(x, y, z, R);
    return r - 2 * span * (fabsf(x) + fabsf(y) + fabsf(z)) + 3 * span * span <= 1;
}

/* Rotation::in_SO3, common.hpp:69 */
ORC_API int orc_in_so3(float x, float y, float z)
{
    float R[9];
    return orc_rotation(x, y, z, R) <= 1.0f;
}

/* sin(half_angle), half_angle = span * sqrt3 * pi / 2  (registration.cu:41; SASS: FMUL, FMUL.D2) */
ORC_API float orc_rot_sin(float span)
{
    int i;
    if (g_sin_mode)
        for (i = 0; i < g_sin_n; ++i)
            if (g_sin_span[i] == span) return g_sin_val[i];
    {
        float half_angle = (span * ORC_SQRT3) * ORC_PI / 2.0f;
        return sinf(half_angle);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Pre/post-processing, fgoicp.cpp:176-287, fgoicp.hpp:87-90 (serial fp32, index order)        */
/* ------------------------------------------------------------------------------------------ */

/* center_point_cloud: subtracts the centroid in place, returns -centroid (fgoicp.cpp:176-195) */
ORC_API void orc_center(float* pts, size_t n, float* neg_centroid)
{
    float c[3] = { 0.0f, 0.0f, 0.0f };
    size_t i;
    for (i = 0; i < n; ++i) { c[0] += pts[3 * i]; c[1] += pts[3 * i + 1]; c[2] += pts[3 * i + 2]; }
    c[0] /= (float)n; c[1] /= (float)n; c[2] /= (float)n;
    for (i = 0; i < n; ++i) { pts[3 * i] -= c[0]; pts[3 * i + 1] -= c[1]; pts[3 * i + 2] -= c[2]; }
    neg_centroid[0] = -c[0]; neg_centroid[1] = -c[1]; neg_centroid[2] = -c[2];
}

/* get_scaling_factor (fgoicp.cpp:197-220): 1 / max |coord| */
ORC_API float orc_scaling_factor(const float* pts, size_t n)
{
    float m = -FLT_MAX;
    size_t i;
    for (i = 0; i < 3 * n; ++i) { float a = fabsf(pts[i]); if (a > m) m = a; }
    return 1.0f / m;
}

ORC_API void orc_scale(float* pts, size_t n, float s)
{
    size_t i;
    for (i = 0; i < 3 * n; ++i) pts[i] *= s;
}

/* get_point_cloud_ranges (fgoicp.cpp:222-268) */
ORC_API void orc_ranges(const float* pts, size_t n, float* mn, float* mx)
{
    size_t i; int a;
    for (a = 0; a < 3; ++a) { mn[a] = FLT_MAX; mx[a] = -FLT_MAX; }
    for (i = 0; i < n; ++i)
        for (a = 0; a < 3; ++a)
        {
            float v = pts[3 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
}

/* restore_translation (fgoicp.hpp:87-90): t / s + R * offset_pcs - offset_pct  (host, unfused) */
ORC_API void orc_restore_translation(const float* R, const float* t, float s,
                                     const float* offset_pcs, const float* offset_pct, float* out)
{
    int r;
    for (r = 0; r < 3; ++r)
    {
        float rp = R[r] * offset_pcs[0] + R[3 + r] * offset_pcs[1] + R[6 + r] * offset_pcs[2];
        out[r] = t[r] / s + rp - offset_pct[r];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Nearest-distance grid ("LUT"), registration.cu:180-207, 258-318                             */
/* ------------------------------------------------------------------------------------------ */

ORC_API void orc_lut_dims(const float* bbox_min, const float* bbox_max, float res, int* dims)
{
    int a;
    for (a = 0; a < 3; ++a) dims[a] = (int)ceilf((bbox_max[a] - bbox_min[a]) / res);
}

/* buildLUTKernel (registration.cu:258-278) with the host-side point shift (:289-296).
 * SASS: dx = FFMA(float(x), res, -P.x) -- the node coordinate is never rounded on its own. */
static void orc_lut_build_brute(const float* model, size_t nt, const float* bbox_min, float res,
                                const int* dims, float* out)
{
    float* P = (float*)malloc(sizeof(float) * 3 * (nt ? nt : 1));
    size_t j;
    long long cells = (long long)dims[0] * dims[1] * dims[2];
    long long c;
    for (j = 0; j < nt; ++j)
    {
        P[3 * j + 0] = model[3 * j + 0] + (-bbox_min[0]);
        P[3 * j + 1] = model[3 * j + 1] + (-bbox_min[1]);
        P[3 * j + 2] = model[3 * j + 2] + (-bbox_min[2]);
    }
#pragma omp parallel for schedule(static)
    for (c = 0; c < cells; ++c)
    {
        int x = (int)(c % dims[0]);
        int y = (int)((c / dims[0]) % dims[1]);
        int z = (int)(c / ((long long)dims[0] * dims[1]));
        float fx = (float)x, fy = (float)y, fz = (float)z;
        float best = FLT_MAX;
        size_t k;
        for (k = 0; k < nt; ++k)
        {
            float dx = fmaf(fx, res, -P[3 * k + 0]);
            float dy = fmaf(fy, res, -P[3 * k + 1]);
            float dz = fmaf(fz, res, -P[3 * k + 2]);
            float d = orc_sq3(dx, dy, dz);
            best = best < d ? best : d;
        }
        out[c] = best;
    }
    free(P);
}

typedef struct orc_lut
{
    const float* data;
    int dims[3];
    float scale;       /* 1 / res  */
    float offset[3];   /* -bbox_min */
} orc_lut;

static inline int orc_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* One texture axis: unnormalised coordinate u -> (i0, i1, alpha) under linear filtering with
 * clamp addressing: uB = u - 0.5, i = floor(uB), alpha = frac(uB) kept to 8 fractional bits
 * (CUDA Programming Guide, "Texture Fetching / Linear Filtering"; registration.cu:226-231). */
static inline void orc_tex_axis(float u, int dim, int* i0, int* i1, float* alpha)
{
    int xf, i;
    /* keep the conversion in range; outside [-1, dim] both taps clamp to the same edge texel */
    if (!(u > -2.0f)) u = -2.0f;
    if (u > (float)dim + 2.0f) u = (float)dim + 2.0f;
    if (g_weight_mode == 0)
        xf = (int)floorf(u * 256.0f + 0.5f) - 128;  /* round HALF UP to 1/256: pinned on B200, see DESIGN.md */
    else if (g_weight_mode == 2)
        xf = (int)rintf(u * 256.0f) - 128;          /* round half to even (probe only) */
    else
        xf = (int)floorf(u * 256.0f) - 128;         /* truncate to 1/256      */
    i = xf >> 8;                                    /* arithmetic shift = floor */
    *alpha = (float)(xf & 255) * (1.0f / 256.0f);
    *i0 = orc_clampi(i, 0, dim - 1);
    *i1 = orc_clampi(i + 1, 0, dim - 1);
}

/* NearestNeighborLUT::search (registration.cu:320-328) + tex3D linear filter semantics. */
static inline float orc_lut_sample_one(const orc_lut* L, const float* q)
{
    int x0, x1, y0, y1, z0, z1;
    float a, b, c;
    float ux = (q[0] + L->offset[0]) * L->scale;
    float uy = (q[1] + L->offset[1]) * L->scale;
    float uz = (q[2] + L->offset[2]) * L->scale;
    const float* T = L->data;
    size_t sx = 1, sy = (size_t)L->dims[0], sz = (size_t)L->dims[0] * L->dims[1];
    float t000, t100, t010, t110, t001, t101, t011, t111;
    orc_tex_axis(ux, L->dims[0], &x0, &x1, &a);
    orc_tex_axis(uy, L->dims[1], &y0, &y1, &b);
    orc_tex_axis(uz, L->dims[2], &z0, &z1, &c);
    t000 = T[x0 * sx + y0 * sy + z0 * sz]; t100 = T[x1 * sx + y0 * sy + z0 * sz];
    t010 = T[x0 * sx + y1 * sy + z0 * sz]; t110 = T[x1 * sx + y1 * sy + z0 * sz];
    t001 = T[x0 * sx + y0 * sy + z1 * sz]; t101 = T[x1 * sx + y0 * sy + z1 * sz];
    t011 = T[x0 * sx + y1 * sy + z1 * sz]; t111 = T[x1 * sx + y1 * sy + z1 * sz];
    if (g_interp_mode == 0)
    {
        float c00 = fmaf(a, t100 - t000, t000);
        float c10 = fmaf(a, t110 - t010, t010);
        float c01 = fmaf(a, t101 - t001, t001);
        float c11 = fmaf(a, t111 - t011, t011);
        float c0 = fmaf(b, c10 - c00, c00);
        float c1 = fmaf(b, c11 - c01, c01);
        return fmaf(c, c1 - c0, c0);
    }
    else
    {
        /* weights are exact in fp32: products of k/256 with k <= 256 need <= 24 bits */
        float a0 = 1.0f - a, b0 = 1.0f - b, c0 = 1.0f - c;
        float acc = (a0 * b0 * c0) * t000;
        acc = fmaf(a * b0 * c0, t100, acc);
        acc = fmaf(a0 * b * c0, t010, acc);
        acc = fmaf(a * b * c0, t110, acc);
        acc = fmaf(a0 * b0 * c, t001, acc);
        acc = fmaf(a * b0 * c, t101, acc);
        acc = fmaf(a0 * b * c, t011, acc);
        acc = fmaf(a * b * c, t111, acc);
        return acc;
    }
}

ORC_API void orc_lut_sample(const float* lut, const int* dims, const float* bbox_min, float res,
                            const float* q, size_t n, float* out)
{
    orc_lut L;
    long long i;
    L.data = lut; L.dims[0] = dims[0]; L.dims[1] = dims[1]; L.dims[2] = dims[2];
    L.scale = 1.0f / res;
    L.offset[0] = -bbox_min[0]; L.offset[1] = -bbox_min[1]; L.offset[2] = -bbox_min[2];
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long long)n; ++i) out[i] = orc_lut_sample_one(&L, q + 3 * i);
}

/* ------------------------------------------------------------------------------------------ */
/* Batched bounds: kernComputeBounds + the two reductions (registration.cu:27-60, 88-152)      */
/* ------------------------------------------------------------------------------------------ */

/* data: ns x 3;  tcubes: T x 4 (tx, ty, tz, span);  outputs lb[T], ub[T] */
ORC_API void orc_bounds(const float* lut, const int* dims, const float* bbox_min, float res,
                        const float* data, size_t ns,
                        const float* R, float rot_span, int fix_rot,
                        const float* tcubes, int T, float* lb, float* ub)
{
    orc_lut L;
    float sin_half = orc_rot_sin(rot_span);
    int c;
    L.data = lut; L.dims[0] = dims[0]; L.dims[1] = dims[1]; L.dims[2] = dims[2];
    L.scale = 1.0f / res;
    L.offset[0] = -bbox_min[0]; L.offset[1] = -bbox_min[1]; L.offset[2] = -bbox_min[2];
#pragma omp parallel for schedule(dynamic, 1)
    for (c = 0; c < T; ++c)
    {
        const float* tc = tcubes + 4 * c;
        double sum_ub = 0.0, sum_lb = 0.0;
        const int trim = g_trim_k > 0 && g_trim_k < ns;
        float* vu = trim ? (float*)malloc(sizeof(float) * ns) : NULL;
        float* vl = trim ? (float*)malloc(sizeof(float) * ns) : NULL;
        size_t i;
        for (i = 0; i < ns; ++i)
        {
            const float* p = data + 3 * i;
            float q[3], d2, d, e, ui, li;
            orc_xform(R, tc, p, q);
            d2 = orc_lut_sample_one(&L, q);
            d = sqrtf(d2);
            if (!fix_rot)
            {
                float radius = orc_sq3(p[0], p[1], p[2]);        /* squared norm (Q5) */
                float rot_r = (2.0f * radius) * sin_half;        /* SASS: FADD r,r ; FMUL */
                d -= rot_r;
            }
            ui = d > 0.0f ? d * d : 0.0f;
            e = fmaf(tc[3], -ORC_SQRT3, d);                      /* SASS: FFMA span,-sqrt3,d */
            li = e > 0.0f ? e * e : 0.0f;
            if (trim) { vu[i] = ui; vl[i] = li; }
            else { sum_ub += (double)ui; sum_lb += (double)li; }
        }
        if (trim)
        {
            sum_ub = orc_trimmed_sum(vu, ns, g_trim_k);
            sum_lb = orc_trimmed_sum(vl, ns, g_trim_k);
            free(vu); free(vl);
        }
        ub[c] = (float)sum_ub;
        lb[c] = (float)sum_lb;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Exact nearest neighbour (brute force): K5 registration.cu:160-172, K7 icp3d.cu:11-28        */
/* ------------------------------------------------------------------------------------------ */

/* rooted = 0: compare squared distances, strict <, ascending j (lowest index wins ties)
 * rooted = 1: compare sqrtf(d2) as glm::distance does, strict >, ascending j.
 * queries are transformed by (R, t) first when R != NULL.  idx / d2 may be NULL. */
static void orc_nn_brute(const float* model, size_t nt, const float* q_in, size_t n,
                         const float* R, const float* t, int rooted, int32_t* idx, float* d2out)
{
    long long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long long)n; ++i)
    {
        float q[3];
        float best = ORC_INF, best_d2 = ORC_INF;
        int32_t bi = -1;
        size_t j;
        if (R) orc_xform(R, t, q_in + 3 * i, q);
        else { q[0] = q_in[3 * i]; q[1] = q_in[3 * i + 1]; q[2] = q_in[3 * i + 2]; }
        for (j = 0; j < nt; ++j)
        {
            float dx = q[0] - model[3 * j], dy = q[1] - model[3 * j + 1], dz = q[2] - model[3 * j + 2];
            float dd = orc_sq3(dx, dy, dz);
            float key = rooted ? sqrtf(dd) : dd;
            if (key < best) { best = key; best_d2 = dd; bi = (int32_t)j; }
        }
        if (idx) idx[i] = bi;
        if (d2out) d2out[i] = best_d2;
    }
}

/* The same search through a k-d tree (the CPU baseline's stand-in for the nanoflann ICP BASELINE.json names; nanoflann
 * itself is not in this image).  EXACT, ties included: the winner is the lexicographic minimum of (key, index), which
 * is what the ascending scans above pick, and a subtree is skipped only when the key of its bounding box -- the same
 * fp32 formula on the per-axis gaps, monotone in every rounding step -- is strictly above the best key so far. */
#define ORC_KD_LEAF 12
typedef struct orc_kdnode
{
    float lo[3], hi[3];
    int left, right;            /* children, -1 for a leaf    */
    int begin, end;             /* leaf: range of perm / pts  */
} orc_kdnode;
typedef struct orc_kdtree
{
    orc_kdnode* nodes;
    int n_nodes;
    int32_t* perm;              /* original index of slot p   */
    float* pts;                 /* xyz in slot order          */
    size_t nt;
    uint64_t hash;              /* of the model bytes (cache) */
} orc_kdtree;

static int g_nn_mode = 1;       /* 0: brute force, 1: k-d tree for nt >= 64 */
static orc_kdtree g_kd = { NULL, 0, NULL, NULL, 0, 0 };

ORC_API void orc_set_nn_mode(int mode) { g_nn_mode = mode; }

static int orc_kd_build_rec(orc_kdtree* T, const float* model, int begin, int end)
{
    int id = T->n_nodes++, a, p, axis = 0;
    orc_kdnode* nd = &T->nodes[id];
    for (a = 0; a < 3; ++a) { nd->lo[a] = FLT_MAX; nd->hi[a] = -FLT_MAX; }
    for (p = begin; p < end; ++p)
        for (a = 0; a < 3; ++a)
        {
            float v = model[3 * (size_t)T->perm[p] + a];
            if (v < nd->lo[a]) nd->lo[a] = v;
            if (v > nd->hi[a]) nd->hi[a] = v;
        }
    nd->begin = begin; nd->end = end; nd->left = nd->right = -1;
    if (end - begin <= ORC_KD_LEAF) return id;
    for (a = 1; a < 3; ++a) if (nd->hi[a] - nd->lo[a] > nd->hi[axis] - nd->lo[axis]) axis = a;
    if (!(nd->hi[axis] > nd->lo[axis])) return id;                    /* all points coincide: one leaf */
    {
        /* quickselect of the median along `axis` */
        int mid = begin + (end - begin) / 2, lo = begin, hi = end - 1;
        while (lo < hi)
        {
            float pivot = model[3 * (size_t)T->perm[(lo + hi) / 2] + axis];
            int i = lo, j = hi;
            while (i <= j)
            {
                while (model[3 * (size_t)T->perm[i] + axis] < pivot) ++i;
                while (model[3 * (size_t)T->perm[j] + axis] > pivot) --j;
                if (i <= j) { int32_t tmp = T->perm[i]; T->perm[i] = T->perm[j]; T->perm[j] = tmp; ++i; --j; }
            }
            if (mid <= j) hi = j; else if (mid >= i) lo = i; else break;
        }
        {
            int l = orc_kd_build_rec(T, model, begin, mid);
            int r = orc_kd_build_rec(T, model, mid, end);
            T->nodes[id].left = l; T->nodes[id].right = r;            /* T->nodes never moves: sized up front */
        }
    }
    return id;
}

static const orc_kdtree* orc_kd_get(const float* model, size_t nt)
{
    uint64_t h = 1469598103934665603ull;
    const uint32_t* w = (const uint32_t*)model;
    size_t i;
    for (i = 0; i < 3 * nt; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    if (g_kd.nodes && g_kd.nt == nt && g_kd.hash == h) return &g_kd;
    free(g_kd.nodes); free(g_kd.perm); free(g_kd.pts);
    g_kd.nodes = (orc_kdnode*)malloc(sizeof(orc_kdnode) * (2 * nt + 1));
    g_kd.perm = (int32_t*)malloc(sizeof(int32_t) * nt);
    g_kd.pts = (float*)malloc(sizeof(float) * 3 * nt);
    g_kd.n_nodes = 0; g_kd.nt = nt; g_kd.hash = h;
    for (i = 0; i < nt; ++i) g_kd.perm[i] = (int32_t)i;
    orc_kd_build_rec(&g_kd, model, 0, (int)nt);
    for (i = 0; i < nt; ++i) memcpy(g_kd.pts + 3 * i, model + 3 * (size_t)g_kd.perm[i], 3 * sizeof(float));
    return &g_kd;
}

static inline float orc_kd_box_d2(const orc_kdnode* nd, const float* q)
{
    float g[3];
    int a;
    for (a = 0; a < 3; ++a)
        g[a] = q[a] < nd->lo[a] ? nd->lo[a] - q[a] : (q[a] > nd->hi[a] ? q[a] - nd->hi[a] : 0.0f);
    return orc_sq3(g[0], g[1], g[2]);
}

static void orc_nn_kd(const orc_kdtree* T, const float* q_in, size_t n,
                      const float* R, const float* t, int rooted, int32_t* idx, float* d2out)
{
    long long i;
#pragma omp parallel for schedule(dynamic, 64)
    for (i = 0; i < (long long)n; ++i)
    {
        float q[3];
        float best = ORC_INF, best_d2 = ORC_INF;
        int32_t bi = -1;
        int stack[128], sp = 0;
        if (R) orc_xform(R, t, q_in + 3 * i, q);
        else { q[0] = q_in[3 * i]; q[1] = q_in[3 * i + 1]; q[2] = q_in[3 * i + 2]; }
        stack[sp++] = 0;
        while (sp > 0)
        {
            const orc_kdnode* nd = &T->nodes[stack[--sp]];
            float lb = orc_kd_box_d2(nd, q);
            if ((rooted ? sqrtf(lb) : lb) > best) continue;             /* equal keys may still hold a lower index */
            if (nd->left < 0)
            {
                int p;
                for (p = nd->begin; p < nd->end; ++p)
                {
                    float dx = q[0] - T->pts[3 * p], dy = q[1] - T->pts[3 * p + 1], dz = q[2] - T->pts[3 * p + 2];
                    float dd = orc_sq3(dx, dy, dz);
                    float key = rooted ? sqrtf(dd) : dd;
                    int32_t j = T->perm[p];
                    /* the brute-force scan starts from ORC_INF with a strict compare: keys >= ORC_INF never win */
                    if (key < best || (key == best && bi >= 0 && j < bi)) { best = key; best_d2 = dd; bi = j; }
                }
            }
            else
            {
                float dl = orc_kd_box_d2(&T->nodes[nd->left], q), dr = orc_kd_box_d2(&T->nodes[nd->right], q);
                if (dl <= dr) { stack[sp++] = nd->right; stack[sp++] = nd->left; }     /* nearer child on top */
                else { stack[sp++] = nd->left; stack[sp++] = nd->right; }
            }
        }
        if (idx) idx[i] = bi;
        if (d2out) d2out[i] = best_d2;
    }
}

/* The grid build through the same tree (over the SHIFTED points P = model - bbox_min, registration.cu:289-296): each
 * node's value is the plain minimum of the reference's per-pair expression, so only values matter (no tie rule).
 * A subtree is skipped when its box cannot beat the best value: per axis, with c = x * res exact, every point p of
 * the box has |fl(c - p)| >= |fl(c - lo)| when c <= lo and >= |fl(c - hi)| when c >= hi (one rounding, monotone),
 * and the sign of fl(c - lo) / fl(c - hi) is the sign of the exact difference. */
static inline float orc_kd_box_d2_node(const orc_kdnode* nd, const float* f, float res)
{
    float g[3];
    int a;
    for (a = 0; a < 3; ++a)
    {
        float glo = fmaf(f[a], res, -nd->lo[a]);
        if (glo <= 0.0f) g[a] = -glo;
        else
        {
            float ghi = fmaf(f[a], res, -nd->hi[a]);
            g[a] = ghi >= 0.0f ? ghi : 0.0f;
        }
    }
    return orc_sq3(g[0], g[1], g[2]);
}

static int g_lut_mode = 1;      /* 0: brute force over all points per node, 1: k-d tree */
ORC_API void orc_set_lut_mode(int mode) { g_lut_mode = mode; }

ORC_API void orc_lut_build(const float* model, size_t nt, const float* bbox_min, float res,
                           const int* dims, float* out)
{
    orc_kdtree T = { NULL, 0, NULL, NULL, 0, 0 };
    float* P;
    size_t j;
    long long rows = (long long)dims[1] * dims[2], row;
    if (g_lut_mode != 1 || nt < 64 || nt >= ((size_t)1 << 30)) { orc_lut_build_brute(model, nt, bbox_min, res, dims, out); return; }
    P = (float*)malloc(sizeof(float) * 3 * nt);
    for (j = 0; j < nt; ++j)
    {
        P[3 * j + 0] = model[3 * j + 0] + (-bbox_min[0]);
        P[3 * j + 1] = model[3 * j + 1] + (-bbox_min[1]);
        P[3 * j + 2] = model[3 * j + 2] + (-bbox_min[2]);
    }
    T.nodes = (orc_kdnode*)malloc(sizeof(orc_kdnode) * (2 * nt + 1));
    T.perm = (int32_t*)malloc(sizeof(int32_t) * nt);
    T.pts = (float*)malloc(sizeof(float) * 3 * nt);
    T.nt = nt;
    for (j = 0; j < nt; ++j) T.perm[j] = (int32_t)j;
    orc_kd_build_rec(&T, P, 0, (int)nt);
    for (j = 0; j < nt; ++j) memcpy(T.pts + 3 * j, P + 3 * (size_t)T.perm[j], 3 * sizeof(float));
#pragma omp parallel for schedule(dynamic, 16)
    for (row = 0; row < rows; ++row)
    {
        int y = (int)(row % dims[1]), z = (int)(row / dims[1]), x;
        int seed = 0;                                   /* slot of the previous node's winner: a real candidate, so a valid start */
        for (x = 0; x < dims[0]; ++x)
        {
            float f[3] = { (float)x, (float)y, (float)z };
            float best;
            int stack[128], sp = 0;
            {
                float dx = fmaf(f[0], res, -T.pts[3 * seed]), dy = fmaf(f[1], res, -T.pts[3 * seed + 1]), dz = fmaf(f[2], res, -T.pts[3 * seed + 2]);
                best = orc_sq3(dx, dy, dz);
            }
            stack[sp++] = 0;
            while (sp > 0)
            {
                const orc_kdnode* nd = &T.nodes[stack[--sp]];
                if (orc_kd_box_d2_node(nd, f, res) > best) continue;
                if (nd->left < 0)
                {
                    int p;
                    for (p = nd->begin; p < nd->end; ++p)
                    {
                        float dx = fmaf(f[0], res, -T.pts[3 * p]), dy = fmaf(f[1], res, -T.pts[3 * p + 1]), dz = fmaf(f[2], res, -T.pts[3 * p + 2]);
                        float d = orc_sq3(dx, dy, dz);
                        if (d < best) { best = d; seed = p; }
                    }
                }
                else
                {
                    float dl = orc_kd_box_d2_node(&T.nodes[nd->left], f, res), dr = orc_kd_box_d2_node(&T.nodes[nd->right], f, res);
                    if (dl <= dr) { stack[sp++] = nd->right; stack[sp++] = nd->left; }
                    else { stack[sp++] = nd->left; stack[sp++] = nd->right; }
                }
            }
            out[((size_t)z * dims[1] + y) * dims[0] + x] = best;
        }
    }
    free(T.nodes); free(T.perm); free(T.pts); free(P);
}

ORC_API void orc_nn(const float* model, size_t nt, const float* q_in, size_t n,
                    const float* R, const float* t, int rooted, int32_t* idx, float* d2out)
{
    if (g_nn_mode == 1 && nt >= 64 && nt < ((size_t)1 << 30))
        orc_nn_kd(orc_kd_get(model, nt), q_in, n, R, t, rooted, idx, d2out);
    else
        orc_nn_brute(model, nt, q_in, n, R, t, rooted, idx, d2out);
}

/* Registration::compute_sse_error(R, t), registration.cu:62-86 */
ORC_API float orc_sse(const float* model, size_t nt, const float* data, size_t ns,
                      const float* R, const float* t)
{
    float* d2 = (float*)malloc(sizeof(float) * (ns ? ns : 1));
    double s = 0.0;
    size_t i;
    orc_nn(model, nt, data, ns, R, t, 0, NULL, d2);
    if (g_trim_k > 0 && g_trim_k < ns) s = orc_trimmed_sum(d2, ns, g_trim_k);
    else for (i = 0; i < ns; ++i) s += (double)d2[i];
    free(d2);
    return (float)s;
}

/* ------------------------------------------------------------------------------------------ */
/* ICP, icp3d.cu:55-172                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* glm mat3 * mat3 and mat3 * vec3 on the HOST (icp3d.cu:101-102, :169): unfused, left to right */
static void orc_mat3_mul_host(const float* A, const float* B, float* C)
{
    int c, r;
    for (c = 0; c < 3; ++c)
        for (r = 0; r < 3; ++r)
            C[c * 3 + r] = A[r] * B[c * 3] + A[3 + r] * B[c * 3 + 1] + A[6 + r] * B[c * 3 + 2];
}

static void orc_mat3_vec_host(const float* A, const float* v, float* o)
{
    int r;
    for (r = 0; r < 3; ++r) o[r] = A[r] * v[0] + A[3 + r] * v[1] + A[6 + r] * v[2];
}

/* closest_orthogonal_approximation (icp3d.cu:110-138): H is glm column-major float;
 * the Eigen matrix is its transpose-of-storage, i.e. the mathematical sum a b^T. */
ORC_API void orc_closest_orthogonal(const float* ABt, float* Rout)
{
    double H[9], Rd[9];
    int r, c;
    for (r = 0; r < 3; ++r)
        for (c = 0; c < 3; ++c)
            H[r * 3 + c] = (double)ABt[c * 3 + r];     /* matrix(r,c) = ABt[c][r] */
    orc_closest_rotation(H, Rd);
    for (c = 0; c < 3; ++c)
        for (r = 0; r < 3; ++r)
            Rout[c * 3 + r] = (float)Rd[r * 3 + c];    /* glm column c = (R(0,c),R(1,c),R(2,c)) */
}

/* IterativeClosestPoint3D::run (icp3d.cu:80-108) with procrustes() (:140-172).
 * Returns sse; writes R, t (column-major) and the number of loop iterations executed. */
ORC_API float orc_icp(const float* model, size_t nt, const float* data, size_t ns,
                      int max_iter, float thr, const float* R0, const float* t0,
                      float* Rout, float* tout, int* iters_out)
{
    float* W = (float*)malloc(sizeof(float) * 3 * (ns ? ns : 1));
    int32_t* corr = (int32_t*)malloc(sizeof(int32_t) * (ns ? ns : 1));
    const int trim = g_trim_k > 0 && g_trim_k < ns;
    const size_t n_in = trim ? g_trim_k : ns;                 /* points entering the Procrustes step */
    float* cd2 = (float*)malloc(sizeof(float) * (ns ? ns : 1));
    float* srt = (float*)malloc(sizeof(float) * (ns ? ns : 1));
    unsigned char* inl = (unsigned char*)malloc(ns ? ns : 1);
    float R[9], t[3], lastR[9], lastT[3];
    float sse = ORC_INF, last_sse = 2.0f * ORC_INF;
    size_t i;
    int iter = 0, k;
    memcpy(R, R0, sizeof(R)); memcpy(t, t0, sizeof(t));
    memcpy(lastR, R0, sizeof(R)); memcpy(lastT, t0, sizeof(t));
    for (i = 0; i < ns; ++i) orc_xform(R, t, data + 3 * i, W + 3 * i);          /* icp3d.cu:85 */

    while (iter++ < max_iter && (last_sse - sse) > thr * last_sse)                /* icp3d.cu:94 */
    {
        double sa[3] = { 0, 0, 0 }, sb[3] = { 0, 0, 0 }, H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        float abar[3], bbar[3], ABt[9], Rd[9], td[3], Rn[9], tn[3], tmp[3];
        last_sse = sse; memcpy(lastR, R, sizeof(R)); memcpy(lastT, t, sizeof(t));

        orc_nn(model, nt, W, ns, NULL, NULL, 1, corr, cd2);                       /* icp3d.cu:146 */
        /* trimming: inliers = the n_in correspondences with the smallest rooted distance, ties in point order */
        for (i = 0; i < ns; ++i) inl[i] = 1;
        if (trim)
        {
            size_t less = 0, take_eq;
            float vk;
            for (i = 0; i < ns; ++i) { cd2[i] = sqrtf(cd2[i]); srt[i] = cd2[i]; }
            qsort(srt, ns, sizeof(float), orc_cmp_float);
            vk = srt[n_in - 1];
            for (i = 0; i < ns; ++i) less += cd2[i] < vk;
            take_eq = n_in - less;
            for (i = 0; i < ns; ++i)
            {
                if (cd2[i] < vk) inl[i] = 1;
                else if (cd2[i] == vk && take_eq > 0) { inl[i] = 1; --take_eq; }
                else inl[i] = 0;
            }
        }
        for (i = 0; i < ns; ++i)
            for (k = 0; k < 3 && inl[i]; ++k)
            {
                sa[k] += (double)W[3 * i + k];
                sb[k] += (double)model[3 * (size_t)corr[i] + k];
            }
        for (k = 0; k < 3; ++k)
        {
            abar[k] = (float)sa[k] / (float)n_in;                                 /* icp3d.cu:155-156 */
            bbar[k] = (float)sb[k] / (float)n_in;
        }
        for (i = 0; i < ns; ++i)
        {
            float a[3], b[3];
            int c, r;
            if (!inl[i]) continue;
            for (k = 0; k < 3; ++k)
            {
                a[k] = W[3 * i + k] - abar[k];                                    /* icp3d.cu:158-159 */
                b[k] = model[3 * (size_t)corr[i] + k] - bbar[k];
            }
            /* glm::outerProduct(a, b)[c][r] = a[r] * b[c]  (icp3d.cu:51) */
            for (c = 0; c < 3; ++c)
                for (r = 0; r < 3; ++r)
                    H[c * 3 + r] += (double)(a[r] * b[c]);
        }
        for (k = 0; k < 9; ++k) ABt[k] = (float)H[k];                            /* glm::mat3 sum */
        orc_closest_orthogonal(ABt, Rd);                                          /* icp3d.cu:168 */
        orc_mat3_vec_host(Rd, abar, tmp);
        for (k = 0; k < 3; ++k) td[k] = bbar[k] - tmp[k];                         /* icp3d.cu:169 */

        for (i = 0; i < ns; ++i)                                                  /* icp3d.cu:100 */
        {
            float q[3];
            orc_xform(Rd, td, W + 3 * i, q);
            W[3 * i] = q[0]; W[3 * i + 1] = q[1]; W[3 * i + 2] = q[2];
        }
        orc_mat3_mul_host(Rd, R, Rn);                                             /* icp3d.cu:101 */
        orc_mat3_vec_host(Rd, t, tmp);
        for (k = 0; k < 3; ++k) tn[k] = tmp[k] + td[k];                           /* icp3d.cu:102 */
        memcpy(R, Rn, sizeof(R)); memcpy(t, tn, sizeof(t));
        sse = orc_sse(model, nt, data, ns, R, t);                                 /* icp3d.cu:103 */
    }
    if (iters_out) *iters_out = iter - 1;
    free(W); free(corr); free(cd2); free(srt); free(inl);
    if (sse < last_sse) { memcpy(Rout, R, sizeof(R)); memcpy(tout, t, sizeof(t)); return sse; }
    memcpy(Rout, lastR, sizeof(R)); memcpy(tout, lastT, sizeof(t));
    return last_sse;
}

/* ------------------------------------------------------------------------------------------ */
/* Best-first heaps with a TOTAL order (common.hpp:85-92, 120-127 + insertion sequence, Q13)   */
/* ------------------------------------------------------------------------------------------ */

typedef struct orc_node
{
    float x, y, z, span, lb, ub;
    uint32_t seq;
} orc_node;

/* "a before b": smaller lb first, ties -> larger span first, ties -> earlier insertion first */
static inline int orc_before(const orc_node* a, const orc_node* b)
{
    if (a->lb != b->lb) return a->lb < b->lb;
    if (a->span != b->span) return a->span > b->span;
    return a->seq < b->seq;
}

typedef struct orc_heap
{
    orc_node* v;
    size_t n, cap;
    uint32_t next_seq;
} orc_heap;

static void orc_heap_init(orc_heap* h) { h->v = NULL; h->n = 0; h->cap = 0; h->next_seq = 0; }
static void orc_heap_free(orc_heap* h) { free(h->v); h->v = NULL; h->n = h->cap = 0; }

static void orc_heap_push(orc_heap* h, orc_node nd)
{
    size_t i;
    if (h->n == h->cap)
    {
        h->cap = h->cap ? 2 * h->cap : 256;
        h->v = (orc_node*)realloc(h->v, h->cap * sizeof(orc_node));
    }
    nd.seq = h->next_seq++;
    i = h->n++;
    h->v[i] = nd;
    while (i > 0)
    {
        size_t p = (i - 1) / 2;
        if (!orc_before(&h->v[i], &h->v[p])) break;
        { orc_node tmp = h->v[i]; h->v[i] = h->v[p]; h->v[p] = tmp; }
        i = p;
    }
}

static orc_node orc_heap_pop(orc_heap* h)
{
    orc_node top = h->v[0];
    size_t i = 0;
    h->v[0] = h->v[--h->n];
    for (;;)
    {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->n && orc_before(&h->v[l], &h->v[m])) m = l;
        if (r < h->n && orc_before(&h->v[r], &h->v[m])) m = r;
        if (m == i) break;
        { orc_node tmp = h->v[i]; h->v[i] = h->v[m]; h->v[m] = tmp; }
        i = m;
    }
    return top;
}

/* ------------------------------------------------------------------------------------------ */
/* Problem handle: everything the BnB loops need                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct orc_problem
{
    const float* model; size_t nt;
    const float* data; size_t ns;
    const float* lut; int dims[3]; float bbox_min[3]; float res;
    float sse_threshold;
    int batch;             /* 32 in the reference (fgoicp.cpp:122) */
    float min_tspan;       /* 0.1 (fgoicp.cpp:155) */
    float min_rspan;       /* 0.05 (fgoicp.cpp:53) */
    float icp_trigger;     /* 1.8 (fgoicp.cpp:74) */
} orc_problem;

/* FastGoICP::branch_and_bound_R3 (fgoicp.cpp:102-174).  rot = (x, y, z, span); R is derived.
 * rnode_ub only seeds the root TransNode's ub field (never used for pruning).
 * Returns best_ub; writes best_t[3], *evals (cube x point evaluations), *batches. */
static float orc_bnb_r3_impl(const orc_problem* P, const float* rot, int fix_rot, float best_sse,
                             float rnode_ub, float* best_t, uint64_t* evals, uint32_t* batches)
{
    float R[9];
    float best_error = best_sse;
    float best_ub = ORC_INF;
    orc_heap heap;
    orc_node root;
    orc_node* batch = (orc_node*)malloc(sizeof(orc_node) * P->batch);
    float* tc = (float*)malloc(sizeof(float) * 4 * P->batch);
    float* lb = (float*)malloc(sizeof(float) * P->batch);
    float* ub = (float*)malloc(sizeof(float) * P->batch);
    uint64_t ev = 0; uint32_t nb = 0;

    orc_rotation(rot[0], rot[1], rot[2], R);
    best_t[0] = best_t[1] = best_t[2] = 0.0f;
    orc_heap_init(&heap);
    root.x = root.y = root.z = 0.0f; root.span = 1.0f; root.lb = 0.0f; root.ub = rnode_ub; root.seq = 0;
    orc_heap_push(&heap, root);

    while (heap.n > 0)
    {
        int nbatch = 0, i, k, idx_min = 0;
        if (best_error - heap.v[0].lb < P->sse_threshold) break;                 /* fgoicp.cpp:120 */
        while (heap.n > 0 && nbatch < P->batch)                                   /* :122-130 */
        {
            orc_node nd = orc_heap_pop(&heap);
            if (nd.lb < best_error) batch[nbatch++] = nd;
        }
        if (nbatch == 0) break;   /* cannot happen with the reference's stop rule; guard only */
        for (i = 0; i < nbatch; ++i)
        {
            tc[4 * i] = batch[i].x; tc[4 * i + 1] = batch[i].y; tc[4 * i + 2] = batch[i].z;
            tc[4 * i + 3] = batch[i].span;
        }
        orc_bounds(P->lut, P->dims, P->bbox_min, P->res, P->data, P->ns, R, rot[3], fix_rot,
                   tc, nbatch, lb, ub);                                           /* :135 */
        ev += (uint64_t)nbatch * P->ns; ++nb;
        for (i = 1; i < nbatch; ++i) if (ub[i] < ub[idx_min]) idx_min = i;        /* first min, :139 */
        best_ub = best_ub < ub[idx_min] ? best_ub : ub[idx_min];                  /* :140 */
        if (ub[idx_min] < best_error)                                             /* :141-145 */
        {
            best_error = ub[idx_min];
            best_t[0] = batch[idx_min].x; best_t[1] = batch[idx_min].y; best_t[2] = batch[idx_min].z;
        }
        for (i = 0; i < nbatch; ++i)                                              /* :148-169 */
        {
            float span;
            if (lb[i] >= best_error) continue;
            if (batch[i].span < P->min_tspan) continue;
            span = batch[i].span / 2.0f;
            for (k = 0; k < 8; ++k)
            {
                orc_node ch;
                ch.x = batch[i].x - span + (float)(k >> 0 & 1) * batch[i].span;
                ch.y = batch[i].y - span + (float)(k >> 1 & 1) * batch[i].span;
                ch.z = batch[i].z - span + (float)(k >> 2 & 1) * batch[i].span;
                ch.span = span; ch.lb = lb[i]; ch.ub = ub[i]; ch.seq = 0;
                orc_heap_push(&heap, ch);
            }
        }
    }
    orc_heap_free(&heap);
    free(batch); free(tc); free(lb); free(ub);
    if (evals) *evals = ev;
    if (batches) *batches = nb;
    return best_ub;
}

ORC_API float orc_bnb_r3(const float* model, size_t nt, const float* data, size_t ns,
                         const float* lut, const int* dims, const float* bbox_min, float res,
                         const float* rot_xyz_span, int fix_rot, float best_sse,
                         float sse_threshold, int batch, float* best_t,
                         uint64_t* evals, uint32_t* batches)
{
    orc_problem P;
    P.model = model; P.nt = nt; P.data = data; P.ns = ns; P.lut = lut;
    P.dims[0] = dims[0]; P.dims[1] = dims[1]; P.dims[2] = dims[2];
    P.bbox_min[0] = bbox_min[0]; P.bbox_min[1] = bbox_min[1]; P.bbox_min[2] = bbox_min[2];
    P.res = res; P.sse_threshold = sse_threshold; P.batch = batch > 0 ? batch : 32;
    P.min_tspan = 0.1f; P.min_rspan = 0.05f; P.icp_trigger = 1.8f;
    return orc_bnb_r3_impl(&P, rot_xyz_span, fix_rot, best_sse, 0.0f, best_t, evals, batches);
}

/* FastGoICP::run (fgoicp.cpp:10-30) + branch_and_bound_SO3 (fgoicp.cpp:32-100), best-first,
 * on already centred+scaled clouds.  stats[0]=rotation cubes evaluated, [1]=ICP runs,
 * [2]=bound evals, [3]=inner batches.  Returns best_sse; R, t in the normalised frame. */
ORC_API float orc_run(const float* model, size_t nt, const float* data, size_t ns,
                      const float* lut, const int* dims, const float* bbox_min, float res,
                      float mse_threshold, float* Rout, float* tout, uint64_t* stats)
{
    orc_problem P;
    orc_heap heap;
    orc_node root;
    float best_sse, best_R[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, best_t[3] = { 0, 0, 0 };
    float I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, zero[3] = { 0, 0, 0 };
    float icpR[9], icpT[3];
    uint64_t n_cubes = 0, n_icp = 0, n_evals = 0, n_batches = 0;
    int k;

    P.model = model; P.nt = nt; P.data = data; P.ns = ns; P.lut = lut;
    P.dims[0] = dims[0]; P.dims[1] = dims[1]; P.dims[2] = dims[2];
    P.bbox_min[0] = bbox_min[0]; P.bbox_min[1] = bbox_min[1]; P.bbox_min[2] = bbox_min[2];
    P.res = res;
    P.sse_threshold = (float)((g_trim_k > 0 && g_trim_k < ns) ? g_trim_k : ns) * mse_threshold;   /* fgoicp.hpp:23 */
    P.batch = 32; P.min_tspan = 0.1f; P.min_rspan = 0.05f; P.icp_trigger = 1.8f;

    best_sse = orc_icp(model, nt, data, ns, 100, (float)0.05, I, zero, icpR, icpT, NULL); /* :12-14 */
    ++n_icp;

    orc_heap_init(&heap);
    root.x = root.y = root.z = 0.0f; root.span = 1.0f; root.lb = 0.0f; root.ub = best_sse; root.seq = 0;
    orc_heap_push(&heap, root);
    while (heap.n > 0)
    {
        orc_node nd = orc_heap_pop(&heap);
        float span;
        if (best_sse - nd.lb <= P.sse_threshold) break;                           /* :44 */
        span = nd.span / 2.0f;
        for (k = 0; k < 8; ++k)
        {
            orc_node ch;
            float rot[4], bt[3], dummy_t[3], Rc[9], ub, lb;
            uint64_t ev = 0; uint32_t nb = 0;
            if (span < P.min_rspan) continue;                                     /* :53 */
            ch.x = nd.x - span + (float)(k >> 0 & 1) * nd.span;
            ch.y = nd.y - span + (float)(k >> 1 & 1) * nd.span;
            ch.z = nd.z - span + (float)(k >> 2 & 1) * nd.span;
            ch.span = span; ch.lb = nd.lb; ch.ub = nd.ub; ch.seq = 0;
            if (!orc_overlaps_so3(ch.x, ch.y, ch.z, ch.span)) continue;           /* :61 */
            if (!orc_in_so3(ch.x, ch.y, ch.z)) { orc_heap_push(&heap, ch); continue; } /* :62-66 */
            rot[0] = ch.x; rot[1] = ch.y; rot[2] = ch.z; rot[3] = ch.span;
            ++n_cubes;
            ub = orc_bnb_r3_impl(&P, rot, 1, best_sse, ch.ub, bt, &ev, &nb);      /* :69 */
            n_evals += ev; n_batches += nb;
            if ((double)ub < (double)best_sse * 1.8)                              /* :74 */
            {
                float e;
                orc_rotation(ch.x, ch.y, ch.z, Rc);
                e = orc_icp(model, nt, data, ns, 100, (float)0.005, Rc, bt, icpR, icpT, NULL); /* :76 */
                ++n_icp;
                if (e < best_sse)
                {
                    best_sse = e; memcpy(best_R, icpR, sizeof(best_R)); memcpy(best_t, icpT, sizeof(best_t));
                }
            }
            lb = orc_bnb_r3_impl(&P, rot, 0, best_sse, ch.ub, dummy_t, &ev, &nb); /* :90 */
            n_evals += ev; n_batches += nb;
            if (lb >= best_sse) continue;                                         /* :92 */
            ch.lb = lb; ch.ub = ub;
            orc_heap_push(&heap, ch);
        }
    }
    orc_heap_free(&heap);

    best_sse = orc_icp(model, nt, data, ns, 100, (float)0.0005, best_R, best_t, icpR, icpT, NULL); /* :22-23 */
    ++n_icp;
    memcpy(Rout, icpR, sizeof(icpR)); memcpy(tout, icpT, sizeof(icpT));
    if (stats) { stats[0] = n_cubes; stats[1] = n_icp; stats[2] = n_evals; stats[3] = n_batches; }
    return best_sse;
}
