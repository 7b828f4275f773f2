// common.cuh -- context, error plumbing and the canonical device arithmetic shared by every
// kernel of the B200 Go-ICP hot path.
//
// "Canonical arithmetic": every floating-point expression that the reference evaluates in
// device code is written here with explicit round-to-nearest intrinsics in exactly the
// association nvcc 12.9 gives the reference's own kernels on sm_100 (read from their SASS, see
// DESIGN.md).  That makes per-point values bit-identical between these kernels, the CPU oracle
// (tests only) and -- up to the texture unit's internal filter -- the reference itself.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fgoicp_c.h"

// NVTX ranges around every stage (header-only NVTX3: a no-op unless a profiler is attached)
#include <nvtx3/nvToolsExt.h>
struct FgRange
{
    explicit FgRange(const char* name) { nvtxRangePushA(name); }
    ~FgRange() { nvtxRangePop(); }
    FgRange(const FgRange&) = delete;
    FgRange& operator=(const FgRange&) = delete;
};
#define FG_RANGE(name) FgRange fg_range__(name)

#define FG_SQRT3 1.732050807568877f     // reference fgoicp/common.hpp:19
#define FG_PI    3.141592653589793f     // reference fgoicp/common.hpp:17
#define FG_INF   1E+10f                 // reference fgoicp/common.hpp:18

namespace fg
{
    void set_error(const std::string& msg);
    int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
}

#define FG_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t fg_err__ = (call);                                                  \
        if (fg_err__ != cudaSuccess)                                                    \
            return fg::cuda_fail(fg_err__, #call, __FILE__, __LINE__);                  \
    } while (0)

#define FG_ARG(cond, msg)                                                               \
    do {                                                                                \
        if (!(cond)) { fg::set_error(std::string("bad argument: ") + (msg)); return FGOICP_ERR_ARG; } \
    } while (0)

// Device-side view of the nearest-squared-distance grid in its three physical layouts.
struct LutDev
{
    const float* grid;        // dense, x fastest: grid[(z*dy + y)*dx + x]             (HBM)
    const float* packed;      // corner-packed cells, 8 floats (32 B) per cell, cell index
                              // (cx,cy,cz) in [0,dx]x[0,dy]x[0,dz] <-> texel i = c-1     (HBM)
    cudaTextureObject_t tex;  // cudaArray + linear filter + clamp (the reference's own setup)
    int dx, dy, dz;
    float scale;              // 1 / resolution
    float ox, oy, oz;         // -bbox_min
    // derived, for the manual samplers (filled by fg_lut_finalise):
    float scale256;           // scale * 256 (exact: power-of-two factor)
    float vmx, vmy, vmz;      // 256 * dim - 1: upper clamp of the 24.8 fixed-point texel coordinate
    int dx1, dy1;             // dx + 1, dy + 1: row / slice pitch of the corner-packed grid in cells
    const float* packed0;     // packed + 8 * ((dy1 + 1) * dx1 + 1): cell (ix, iy, iz) = texel index, no +1 needed
    int nbx, nby, nbz;        // FG_BRICKED layout: 4x4x4-cell bricks per axis
};

// FG_BRICKED = 1: the corner-packed grid is stored as 4x4x4-cell bricks (2 KB each, cells in Morton order inside a
// brick, so a 128-byte line is a 2x2x1 block of cells and 256 bytes a 2x2x2 block) instead of x-fastest rows.
#ifndef FG_BRICKED
#define FG_BRICKED 0
#endif

__host__ __device__ inline size_t fg_brick_cell(int cx, int cy, int cz, int nbx, int nby)
{
    size_t brick = ((size_t)(cz >> 2) * (size_t)nby + (size_t)(cy >> 2)) * (size_t)nbx + (size_t)(cx >> 2);
    unsigned intra = (cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2) | ((cx & 2) << 2) | ((cy & 2) << 3) | ((cz & 2) << 4);
    return brick * 64 + intra;
}

inline void fg_lut_finalise(LutDev& L)
{
    L.scale256 = L.scale * 256.0f;
    L.vmx = 256.0f * (float)L.dx - 1.0f; L.vmy = 256.0f * (float)L.dy - 1.0f; L.vmz = 256.0f * (float)L.dz - 1.0f;
    L.dx1 = L.dx + 1; L.dy1 = L.dy + 1;
    L.nbx = (L.dx + 1 + 3) / 4; L.nby = (L.dy + 1 + 3) / 4; L.nbz = (L.dz + 1 + 3) / 4;
    L.packed0 = L.packed ? L.packed + 8 * ((size_t)(L.dy1 + 1) * (size_t)L.dx1 + 1) : nullptr;
}

// Uniform cell grid over the model cloud in LUT space; cells along x are consecutive in the CSR order,
// so a run of cells of one (y, z) row is one contiguous range of points.
struct CellGrid
{
    const int* start;      // [ncell + 1] CSR offsets into the sorted points
    const float4* pts;     // sorted points (LUT-space copy or original-coordinate copy)
    int nx, ny, nz;
    float h, inv_h;        // cell size
    // Coarse boxes: the non-empty blocks of FG_COARSE^3 cells, each with the tight bounding box of its points (LUT
    // space) -- box k = { (lo.x, lo.y, lo.z, bits(X | Y << 10 | Z << 20)), (hi.x, hi.y, hi.z, 0) }.  Lets the NN search
    // of a far query cull whole blocks of rows instead of looking every row of its ball up (nn_icp.cu).
    const float4* coarse = nullptr;
    int n_coarse = 0;
    int coarse_min_rows = 320;     // rows of a ball beyond which the coarse boxes are consulted (+ n_coarse / 8)
};
#ifndef FG_COARSE
#define FG_COARSE 8      // cells per axis of a coarse block (4 measured: see profiles/nn_scan_r02.md); the tight cell bounds of a block use 3 bits each
#endif

struct fgoicp_ctx
{
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    size_t restore_l2_fetch = 0;   // previous cudaLimitMaxL2FetchGranularity when FGOICP_L2_FETCH changed it (0: untouched)

    size_t nt = 0, ns = 0;
    float4* d_model = nullptr;     // (x, y, z, index-as-bits)   [nt]
    float4* d_data = nullptr;      // (x, y, z, |p|^2)           [ns], stored in Morton order of the coordinates
    int* d_data_orig = nullptr;    // storage slot -> index in the caller's cloud [ns]

    float res = 0.f;
    float bbox_min[3] = { 0, 0, 0 }, bbox_max[3] = { 0, 0, 0 };
    LutDev lut{};
    float* d_grid = nullptr;
    float* d_packed = nullptr;
    cudaArray_t arr = nullptr;
    int sampler = FGOICP_SAMPLER_GRID;
    float build_ms = 0.f;

    // scratch (grown on demand, never shrunk)
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void* h_pinned = nullptr;
    size_t pinned_bytes = 0;

    // uniform cell grid over the model cloud (CSR): built once, used by the grid build and by NN search
    int* d_cell_start = nullptr;              // [ncell + 1]
    float4* d_cell_P = nullptr;               // points in LUT space (model - bbox_min), sorted by cell
    float4* d_cell_M = nullptr;               // same order, original coordinates, w = original index
    int cnx = 0, cny = 0, cnz = 0;
    float cell_h = 0.f, cell_inv_h = 0.f;
    float4* d_coarse = nullptr;               // coarse boxes of the cell grid (CellGrid::coarse)
    int n_coarse = 0;
    int nn_mode = 0;                          // 0: cell-grid search, 1: tiled brute force, 2 / 3: cell grid with the coarse boxes forced on / off (test hooks)

    // z-phase-ordered bound evaluation (bounds_phased.cu): per-launch index lists
    void* d_phase = nullptr;
    size_t phase_bytes = 0;
    int phased = 0;                           // 1: fgoicp_bounds_multi* use the phase-ordered kernel

    // round-synchronous inner search (bnb.cu): per-level state in HBM
    void* d_rounds = nullptr;
    size_t rounds_bytes = 0;
    int bnb_mode = 0;                         // 0: auto, 1: persistent per-cube kernel, 2: round-synchronous

    // trimmed registration (extension): 0 = off (the reference's behaviour), else the number of data points whose
    // residuals enter every sum (the smallest ones)
    size_t trim_k = 0;
    void* d_trim = nullptr;                   // per-block term slices of the trimmed bound kernel when 2 x ns floats exceed shared memory
    size_t trim_bytes = 0;
    unsigned char* d_inl = nullptr;           // ICP inlier flags [icp_capacity][ns]
    void* d_icp_part = nullptr;               // Procrustes partial sums + arrival counters (nn_icp.cu)

    // ICP state
    float4* d_work = nullptr;                 // working copy W  [ns]
    unsigned long long* d_nnkey = nullptr;    // packed (value bits << 32 | index) [ns]
    void* d_icp = nullptr;                    // per-instance ICP state blocks, see nn_icp.cu
    int icp_capacity = 0;                     // concurrent ICP instances the three buffers are sized for
    float4* d_nnmemo = nullptr;               // winner memo of the ICP searches: scan position + proven clearance [icp_capacity][ns]
    void* d_icp_jobs = nullptr;               // job queue of a batch of refinements: counters, seed poses, results (nn_icp.cu)
    size_t icp_jobs_bytes = 0;
    void* d_icp_loop = nullptr;               // persistent ICP loop kernel: control block, reduction partials, miss list (nn_icp.cu)
    int icp_mode = 0;                         // 0: automatic (loop kernel for batches that fit the slot pool, else launch chain), 1: launch chain, 2: loop kernel
    int icp_loop_grid = 0;                    // co-resident blocks of k_icp_loop on this device (0: not yet queried)
};

namespace fg
{
    int ensure_scratch(fgoicp_ctx* c, size_t bytes);
    int ensure_pinned(fgoicp_ctx* c, size_t bytes);
}

// ---------------------------------------------------------------------------------------------
// canonical device arithmetic
// ---------------------------------------------------------------------------------------------

// dx*dx + dy*dy + dz*dz  (reference registration.cu:154-160, 248-254; glm::dot in icp3d.cu:20).
// SASS of every reference kernel: FMUL dy,dy ; FFMA dx,dx,. ; FFMA dz,dz,.  -- nvcc keeps the SECOND
// product as the plain multiply and fuses the first and third.
__device__ __forceinline__ float fg_sq3(float dx, float dy, float dz)
{
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// glm mat3 * vec3 (column-major R), reference registration.cu:20, 34 and icp3d.cu:35.
// SASS: FMUL m1r,p.y ; FFMA m0r,p.x,. ; FFMA m2r,p.z,.  (same association as fg_sq3)
__device__ __forceinline__ float3 fg_rotate(const float* R, float px, float py, float pz)
{
    float3 q;
    q.x = __fmaf_rn(R[6], pz, __fmaf_rn(R[0], px, __fmul_rn(R[3], py)));
    q.y = __fmaf_rn(R[7], pz, __fmaf_rn(R[1], px, __fmul_rn(R[4], py)));
    q.z = __fmaf_rn(R[8], pz, __fmaf_rn(R[2], px, __fmul_rn(R[5], py)));
    return q;
}

// Rotation(x, y, z) constructor, reference common.hpp:37-57.  This is HOST arithmetic in the
// reference (no contraction), so every operation is an explicit unfused intrinsic.  Writes the
// column-major matrix; returns Rotation::r.
__device__ __forceinline__ float fg_rotation_matrix(float x, float y, float z, float* R)
{
    float r = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
    if (r > 1.0f) return r;
    float ww = __fsub_rn(1.0f, r);
    float w = __fsqrt_rn(ww);
    float wx = __fmul_rn(w, x), xx = __fmul_rn(x, x);
    float wy = __fmul_rn(w, y), xy = __fmul_rn(x, y), yy = __fmul_rn(y, y);
    float wz = __fmul_rn(w, z), xz = __fmul_rn(x, z), yz = __fmul_rn(y, z), zz = __fmul_rn(z, z);
    R[0] = __fsub_rn(__fsub_rn(__fadd_rn(ww, xx), yy), zz);
    R[1] = __fmul_rn(2.f, __fsub_rn(xy, wz));
    R[2] = __fmul_rn(2.f, __fadd_rn(xz, wy));
    R[3] = __fmul_rn(2.f, __fadd_rn(xy, wz));
    R[4] = __fsub_rn(__fadd_rn(__fsub_rn(ww, xx), yy), zz);
    R[5] = __fmul_rn(2.f, __fsub_rn(yz, wx));
    R[6] = __fmul_rn(2.f, __fsub_rn(xz, wy));
    R[7] = __fmul_rn(2.f, __fadd_rn(yz, wx));
    R[8] = __fadd_rn(__fsub_rn(__fsub_rn(ww, xx), yy), zz);
    return __fsqrt_rn(r);
}

// sin(span * sqrt3 * pi / 2) exactly as the reference's kernel evaluates it per thread
// (registration.cu:41-42; SASS: FMUL, FMUL.D2, then libdevice sinf).
__device__ __forceinline__ float fg_rot_sin(float span)
{
    float half_angle = __fmul_rn(__fmul_rn(__fmul_rn(span, FG_SQRT3), FG_PI), 0.5f);
    return sinf(half_angle);
}

// ---------------------------------------------------------------------------------------------
// grid sampling: the reference's tex3D semantics (registration.cu:214-234, 320-328)
//   u = (q + offset) * scale ; uB = u - 0.5 ; i = floor(uB) ; alpha = frac(uB) in 1.8 fixed point ;
//   texel indices clamped to [0, dim-1] ; tri-linear blend of SQUARED distances.
// FG_WEIGHT_TRUNC / FG_INTERP_WSUM select the alternatives probed by the texture conformance
// test; the defaults are the ones that match tex3D on B200 best.
// ---------------------------------------------------------------------------------------------
#ifndef FG_WEIGHT_TRUNC
#define FG_WEIGHT_TRUNC 0
#endif
#ifndef FG_INTERP_WSUM
#define FG_INTERP_WSUM 0
#endif

#if FG_WEIGHT_TRUNC
#define FG_WEIGHT_BIAS (-128.0f)
#else
#define FG_WEIGHT_BIAS (-127.5f)
#endif

// One axis of the filter footprint: texel index i in [-1, dim-1] and weight alpha.
// Measured on B200 (scripts/tex_conformance.py, 65,536-step sweep across a texel): the texture unit
// keeps the weight in 1.8 fixed point with ROUND-HALF-UP: alpha*256 = floor(uB*256 + 0.5), uB = u - 0.5,
// u = (q + offset) * scale (reference registration.cu:323-325: FADD then FMUL).
//   * u * 256 is an exact power-of-two scaling, so it is folded into the multiplier (scale256): one FMUL
//     gives fl(u) * 256 bit for bit; the bias -127.5 is then ONE rounded add, as before;
//   * the clamp to the grid happens on that float (2 FMNMX) instead of on three integers: v in
//     [-256, 256*dim-1] <=> i in [-1, dim-1].  Wherever the clamp acts both corners of the axis are the
//     same texel (clamp-to-edge), so the weight there is irrelevant: fma(a, t-t, t) = t.  NaN -> -256.
__device__ __forceinline__ void fg_axis(float q, float o, float scale256, float vmax, int& i, float& alpha)
{
    float v = __fadd_rn(__fmul_rn(__fadd_rn(q, o), scale256), FG_WEIGHT_BIAS);
    v = fminf(fmaxf(v, -256.0f), vmax);
    int xf = __float2int_rd(v);
    i = xf >> 8;
    alpha = __fmul_rn((float)(xf & 255), 1.0f / 256.0f);
}

// address of the corner-packed cell of texel index (ix, iy, iz), each in [-1, dim-1]: 32-bit cell arithmetic
// (the +1 per axis lives in packed0), one widening multiply-add for the byte address
__device__ __forceinline__ const float* fg_packed_cell(const LutDev& L, int ix, int iy, int iz)
{
#if FG_BRICKED
    const int cx = ix + 1, cy = iy + 1, cz = iz + 1;
    int brick = ((cz >> 2) * L.nby + (cy >> 2)) * L.nbx + (cx >> 2);
    int intra = (cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2) | ((cx & 2) << 2) | ((cy & 2) << 3) | ((cz & 2) << 4);
    return L.packed + ((long long)brick * 64 + intra) * 8;
#else
    int cell = (iz * L.dy1 + iy) * L.dx1 + ix;
    return L.packed0 + (long long)cell * 8;
#endif
}

__device__ __forceinline__ float fg_trilerp(float a, float b, float c,
                                            float t000, float t100, float t010, float t110,
                                            float t001, float t101, float t011, float t111)
{
#if FG_INTERP_WSUM
    float a0 = __fsub_rn(1.0f, a), b0 = __fsub_rn(1.0f, b), c0 = __fsub_rn(1.0f, c);
    float acc = __fmul_rn(__fmul_rn(__fmul_rn(a0, b0), c0), t000);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a, b0), c0), t100, acc);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a0, b), c0), t010, acc);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a, b), c0), t110, acc);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a0, b0), c), t001, acc);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a, b0), c), t101, acc);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a0, b), c), t011, acc);
    acc = __fmaf_rn(__fmul_rn(__fmul_rn(a, b), c), t111, acc);
    return acc;
#else
    float c00 = __fmaf_rn(a, __fsub_rn(t100, t000), t000);
    float c10 = __fmaf_rn(a, __fsub_rn(t110, t010), t010);
    float c01 = __fmaf_rn(a, __fsub_rn(t101, t001), t001);
    float c11 = __fmaf_rn(a, __fsub_rn(t111, t011), t011);
    float c0 = __fmaf_rn(b, __fsub_rn(c10, c00), c00);
    float c1 = __fmaf_rn(b, __fsub_rn(c11, c01), c01);
    return __fmaf_rn(c, __fsub_rn(c1, c0), c0);
#endif
}

// one 256-bit read-only load (sm_100+): the whole corner-packed cell in a single request
__device__ __forceinline__ void fg_ld256(const float* p, float (&v)[8])
{
    // not volatile: a read-only load with no side effects -- the scheduler is free to hoist it and to keep
    // several of them in flight (the bound kernels issue 4 per lane before consuming any)
#ifndef FG_LD256_QUAL
// cache hints of the 256-bit cell gather.  L1::no_allocate: the gathers hit L1 4 % of the time, so they should not
// evict the point data (measured: inner searches 88.7 -> 83.6 ms, phase-ordered kernel 8.58 -> 8.48 ms);
// ".L2::64B" / ".L2::128B" prefetch-size hints change nothing.
#define FG_LD256_QUAL ".L1::no_allocate"
#endif
    asm("ld.global.nc" FG_LD256_QUAL ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]),
                   "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

template <int SAMPLER>
__device__ __forceinline__ float fg_sample(const LutDev& L, float qx, float qy, float qz)
{
    if (SAMPLER == FGOICP_SAMPLER_TEX)
    {
        // NearestNeighborLUT::search: (query + offset) * scale   (FADD then FMUL in the SASS)
        float ux = __fmul_rn(__fadd_rn(qx, L.ox), L.scale);
        float uy = __fmul_rn(__fadd_rn(qy, L.oy), L.scale);
        float uz = __fmul_rn(__fadd_rn(qz, L.oz), L.scale);
        return tex3D<float>(L.tex, ux, uy, uz);
    }
    int ix, iy, iz;
    float a, b, c;
    fg_axis(qx, L.ox, L.scale256, L.vmx, ix, a);
    fg_axis(qy, L.oy, L.scale256, L.vmy, iy, b);
    fg_axis(qz, L.oz, L.scale256, L.vmz, iz, c);
    if (SAMPLER == FGOICP_SAMPLER_PACKED)
    {
        // the cell of texel index i holds the 8 clamped corner texels i, i+1 per axis
        float v[8];
        fg_ld256(fg_packed_cell(L, ix, iy, iz), v);
        return fg_trilerp(a, b, c, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    }
    else
    {
        int x0 = max(ix, 0), x1 = min(ix + 1, L.dx - 1);
        int y0 = max(iy, 0), y1 = min(iy + 1, L.dy - 1);
        int z0 = max(iz, 0), z1 = min(iz + 1, L.dz - 1);
        size_t sy = (size_t)L.dx, sz = (size_t)L.dx * (size_t)L.dy;
        const float* g = L.grid;
        float t000 = __ldg(g + x0 + y0 * sy + z0 * sz), t100 = __ldg(g + x1 + y0 * sy + z0 * sz);
        float t010 = __ldg(g + x0 + y1 * sy + z0 * sz), t110 = __ldg(g + x1 + y1 * sy + z0 * sz);
        float t001 = __ldg(g + x0 + y0 * sy + z1 * sz), t101 = __ldg(g + x1 + y0 * sy + z1 * sz);
        float t011 = __ldg(g + x0 + y1 * sy + z1 * sz), t111 = __ldg(g + x1 + y1 * sy + z1 * sz);
        return fg_trilerp(a, b, c, t000, t100, t010, t110, t001, t101, t011, t111);
    }
}

// Split form of fg_sample for kernels that want several gathers in flight per lane: *_issue computes the
// filter weights and ISSUES the loads into the request's registers, *_finish blends them.  Issuing all
// requests of a lane before finishing any turns N serial DRAM round trips into one.
struct SampleReq
{
    float a, b, c;      // filter weights
    float v[8];         // the 8 corner texels (TEX sampler: v[0] holds the filtered value)
};

template <int SAMPLER>
__device__ __forceinline__ void fg_sample_issue(const LutDev& L, float qx, float qy, float qz, SampleReq& r)
{
    if (SAMPLER == FGOICP_SAMPLER_TEX)
    {
        float ux = __fmul_rn(__fadd_rn(qx, L.ox), L.scale);
        float uy = __fmul_rn(__fadd_rn(qy, L.oy), L.scale);
        float uz = __fmul_rn(__fadd_rn(qz, L.oz), L.scale);
        r.v[0] = tex3D<float>(L.tex, ux, uy, uz);
        return;
    }
    int ix, iy, iz;
    fg_axis(qx, L.ox, L.scale256, L.vmx, ix, r.a);
    fg_axis(qy, L.oy, L.scale256, L.vmy, iy, r.b);
    fg_axis(qz, L.oz, L.scale256, L.vmz, iz, r.c);
    if (SAMPLER == FGOICP_SAMPLER_PACKED)
    {
        fg_ld256(fg_packed_cell(L, ix, iy, iz), r.v);
    }
    else
    {
        int x0 = max(ix, 0), x1 = min(ix + 1, L.dx - 1);
        int y0 = max(iy, 0), y1 = min(iy + 1, L.dy - 1);
        int z0 = max(iz, 0), z1 = min(iz + 1, L.dz - 1);
        size_t sy = (size_t)L.dx, sz = (size_t)L.dx * (size_t)L.dy;
        const float* g = L.grid;
        r.v[0] = __ldg(g + x0 + y0 * sy + z0 * sz); r.v[1] = __ldg(g + x1 + y0 * sy + z0 * sz);
        r.v[2] = __ldg(g + x0 + y1 * sy + z0 * sz); r.v[3] = __ldg(g + x1 + y1 * sy + z0 * sz);
        r.v[4] = __ldg(g + x0 + y0 * sy + z1 * sz); r.v[5] = __ldg(g + x1 + y0 * sy + z1 * sz);
        r.v[6] = __ldg(g + x0 + y1 * sy + z1 * sz); r.v[7] = __ldg(g + x1 + y1 * sy + z1 * sz);
    }
}

template <int SAMPLER>
__device__ __forceinline__ float fg_sample_finish(const SampleReq& r)
{
    if (SAMPLER == FGOICP_SAMPLER_TEX) return r.v[0];
    return fg_trilerp(r.a, r.b, r.c, r.v[0], r.v[1], r.v[2], r.v[3], r.v[4], r.v[5], r.v[6], r.v[7]);
}

// Per-point body of kernComputeBounds (reference registration.cu:27-60) after the sample:
// d = sqrt(d2); if (!fix_rot) d -= rot_r; ub = d>0 ? d*d : 0; e = fma(span_t, -sqrt3, d);
// lb = e>0 ? e*e : 0.
__device__ __forceinline__ float fg_bound_residual(float d2, float rot_r, bool fix_rot);
__device__ __forceinline__ void fg_bound_from_residual(float d, float span_t, float& ub, float& lb);
__device__ __forceinline__ void fg_bound_terms(float d2, float rot_r, bool fix_rot, float span_t,
                                               float& ub, float& lb)
{
    // branch-free: d - 0 == d bit for bit (d = sqrt >= +0), and x > 0 ? x*x : 0 == sq(max(x, 0)) (NaN -> 0 both ways)
    fg_bound_from_residual(fg_bound_residual(d2, rot_r, fix_rot), span_t, ub, lb);
}

// The same per-point body in two steps, for callers that select on the signed residual before they sum (trimmed bounds):
// fg_bound_residual gives d of fg_bound_terms, fg_bound_from_residual its two terms -- bit for bit the same operations.
__device__ __forceinline__ float fg_bound_residual(float d2, float rot_r, bool fix_rot)
{
    return __fsub_rn(__fsqrt_rn(d2), fix_rot ? 0.0f : rot_r);
}
__device__ __forceinline__ void fg_bound_from_residual(float d, float span_t, float& ub, float& lb)
{
    float du = fmaxf(d, 0.0f);
    ub = __fmul_rn(du, du);
    float e = fmaxf(__fmaf_rn(span_t, -FG_SQRT3, d), 0.0f);
    lb = __fmul_rn(e, e);
}

__device__ __forceinline__ double fg_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
