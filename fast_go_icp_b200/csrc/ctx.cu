// ctx.cu -- context life-cycle and the nearest-squared-distance grid ("LUT") build.
//
// Replaces Registration's constructor and NearestNeighborLUT::{ctor, build} of the reference
// (fgoicp/registration.hpp:68-87, fgoicp/registration.cu:180-318).  The reference scans every
// model point from every grid node (O(cells * nt)); here the grid is built brick by brick with
// an exact candidate cull, and written straight into the layouts the bound kernels read.
#include "common.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fg
{
    static thread_local std::string g_last_error;

    void set_error(const std::string& msg) { g_last_error = msg; }

    int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
        g_last_error = buf;
        return FGOICP_ERR_CUDA;
    }

    int ensure_scratch(fgoicp_ctx* c, size_t bytes)
    {
        if (bytes <= c->scratch_bytes) return FGOICP_OK;
        if (c->d_scratch) { FG_CUDA(cudaStreamSynchronize(c->stream)); FG_CUDA(cudaFree(c->d_scratch)); c->d_scratch = nullptr; c->scratch_bytes = 0; }
        // generous first allocation and geometric growth: a search must not pay cudaFree/cudaMalloc between levels
        size_t want = std::max(std::max(bytes, (size_t)8 << 20), 2 * c->scratch_bytes);
        FG_CUDA(cudaMalloc(&c->d_scratch, want));
        c->scratch_bytes = want;
        return FGOICP_OK;
    }

    int ensure_pinned(fgoicp_ctx* c, size_t bytes)
    {
        if (bytes <= c->pinned_bytes) return FGOICP_OK;
        if (c->h_pinned) { FG_CUDA(cudaStreamSynchronize(c->stream)); FG_CUDA(cudaFreeHost(c->h_pinned)); c->h_pinned = nullptr; c->pinned_bytes = 0; }
        size_t want = std::max(std::max(bytes, (size_t)2 << 20), 2 * c->pinned_bytes);
        FG_CUDA(cudaMallocHost(&c->h_pinned, want));
        c->pinned_bytes = want;
        return FGOICP_OK;
    }
}

// ---------------------------------------------------------------------------------------------
// Grid build kernels
// ---------------------------------------------------------------------------------------------

// Lattice node (x, y, z) sits at (x, y, z) * res in "LUT space" (model shifted by -bbox_min),
// reference registration.cu:266.  The distance uses dx = fma(float(x), res, -P.x): the reference's
// SASS fuses the node coordinate into the subtraction, so it is never rounded on its own.
#define LUT_TILE 1024      // model points staged in shared memory per pass
#define LUT_XPT  4         // nodes per thread along x

// Tiled brute force: every node against every model point.  Exact by construction; used for
// small grids, as the top level of the hierarchical build and as a test hook.
__global__ void __launch_bounds__(256)
k_lut_brute(float* __restrict__ out, int dx, int dy, int dz, float res,
            const float4* __restrict__ P, int nt)
{
    __shared__ float4 tile[LUT_TILE];
    // block covers 32*LUT_XPT nodes in x, 8 rows in y, one z
    int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * LUT_XPT;
    int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    int z = blockIdx.z;
    float fy = (float)y, fz = (float)z;
    float fx[LUT_XPT], best[LUT_XPT];
#pragma unroll
    for (int k = 0; k < LUT_XPT; ++k) { fx[k] = (float)(x0 + k); best[k] = FLT_MAX; }

    for (int base = 0; base < nt; base += LUT_TILE)
    {
        int cnt = min(LUT_TILE, nt - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) tile[i] = P[base + i];
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < cnt; ++i)
        {
            float4 p = tile[i];
            float dyv = __fmaf_rn(fy, res, -p.y);
            float dzv = __fmaf_rn(fz, res, -p.z);
#pragma unroll
            for (int k = 0; k < LUT_XPT; ++k)
            {
                float dxv = __fmaf_rn(fx[k], res, -p.x);
                float d = fg_sq3(dxv, dyv, dzv);
                best[k] = best[k] < d ? best[k] : d;     // registration.cu:273
            }
        }
    }
    if (y < dy && z < dz)
    {
#pragma unroll
        for (int k = 0; k < LUT_XPT; ++k)
            if (x0 + k < dx) out[((size_t)z * dy + y) * dx + x0 + k] = best[k];
    }
}

// ---- exact hierarchical build -------------------------------------------------------------
// Level l is a lattice with origin o_l and spacing s_l (level 0: o = 0, s = res).  A brick is
// 8x8x8 nodes; the next coarser level has one node per brick, at the brick's centre.  Given the
// exact nearest distance D of the brick centre c, every node x of the brick has its nearest
// model point within D + 2*rho of c (rho = distance from c to the farthest node), so the brick
// only needs the model points inside that ball.  Model points are binned in a uniform cell
// grid; the brick gathers candidates cell by cell and takes the exact minimum with the same
// per-pair arithmetic as the brute-force kernel -- the result is bit-identical to it.


__global__ void k_cell_index(const float4* __restrict__ P, int nt, CellGrid g, int* __restrict__ cell_of, int* __restrict__ counts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    float4 p = P[i];
    int cx = min(max((int)floorf(p.x * g.inv_h), 0), g.nx - 1);
    int cy = min(max((int)floorf(p.y * g.inv_h), 0), g.ny - 1);
    int cz = min(max((int)floorf(p.z * g.inv_h), 0), g.nz - 1);
    int c = (cz * g.ny + cy) * g.nx + cx;
    cell_of[i] = c;
    atomicAdd(&counts[c], 1);
}

// single-block exclusive scan (ncell is at most a few hundred thousand)
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ counts, int n, int* __restrict__ start)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024)
    {
        int i = base + threadIdx.x;
        int v = i < n ? counts[i] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0)
        {
            int s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        int prefix = carry + (w > 0 ? warp_sums[w - 1] : 0) + incl - v;
        if (i < n) start[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[n] = carry;
}

__global__ void k_cell_scatter(const float4* __restrict__ P, const float4* __restrict__ M, int nt,
                               const int* __restrict__ cell_of, const int* __restrict__ start,
                               int* __restrict__ fill, float4* __restrict__ sortedP, float4* __restrict__ sortedM)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    int c = cell_of[i];
    int slot = start[c] + atomicAdd(&fill[c], 1);
    sortedP[slot] = P[i];
    sortedM[slot] = M[i];
}

// One warp per block of FG_COARSE^3 cells: number of points and their tight bounding box (LUT space).
__global__ void __launch_bounds__(128)
k_coarse_boxes(CellGrid g, int Cx, int Cy, int Cz, float4* __restrict__ lo_cnt, float4* __restrict__ hi)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= Cx * Cy * Cz) return;
    const int X = w % Cx, Y = (w / Cx) % Cy, Z = w / (Cx * Cy);
    const int x0 = X * FG_COARSE, x1 = min(x0 + FG_COARSE, g.nx) - 1;
    float lo[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, hi3[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
    int cnt = 0;
    for (int row = lane; row < FG_COARSE * FG_COARSE; row += 32)
    {
        const int cy = Y * FG_COARSE + row % FG_COARSE, cz = Z * FG_COARSE + row / FG_COARSE;
        if (cy >= g.ny || cz >= g.nz) continue;
        const int c0 = (cz * g.ny + cy) * g.nx;
        const int b = g.start[c0 + x0], e = g.start[c0 + x1 + 1];
        cnt += e - b;
        for (int k = b; k < e; ++k)
        {
            const float4 p = g.pts[k];
            lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
            hi3[0] = fmaxf(hi3[0], p.x); hi3[1] = fmaxf(hi3[1], p.y); hi3[2] = fmaxf(hi3[2], p.z);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi3[a] = fmaxf(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
        }
    }
    if (lane == 0)
    {
        lo_cnt[w] = make_float4(lo[0], lo[1], lo[2], __int_as_float(cnt));
        hi[w] = make_float4(hi3[0], hi3[1], hi3[2], 0.0f);
    }
}

#define BRICK 8
#define BRICK_CAND 1024

// One block (256 threads, 2 nodes each) per brick of 8^3 nodes of lattice (origin o, spacing s).
// centre_d2: exact squared nearest distance of each brick centre (from the coarser level).
// exact_level0: use the canonical fused node arithmetic dx = fma(float(x), res, -P.x) (origin 0).
__global__ void __launch_bounds__(256)
k_lut_bricks(float* __restrict__ out, int dx, int dy, int dz, float ox, float oy, float oz, float s,
             const float* __restrict__ centre_d2, int bx, int by, int bz, CellGrid g, int exact_level0)
{
    __shared__ float4 cand[BRICK_CAND];
    __shared__ int n_cand;
    __shared__ int cell_lo[3], cell_hi[3];
    __shared__ int row_b[256], row_e[256];

    int b = blockIdx.x;
    int bxi = b % bx, byi = (b / bx) % by, bzi = b / (bx * by);
    // brick centre and reach
    float cx = ox + ((float)(bxi * BRICK) + 3.5f) * s;
    float cy = oy + ((float)(byi * BRICK) + 3.5f) * s;
    float cz = oz + ((float)(bzi * BRICK) + 3.5f) * s;
    float rho = 3.5f * s * 1.7320508f;
    float D = sqrtf(centre_d2[b]);
    float reach = (D + 2.0f * rho) * 1.0001f + 1e-6f;      // inflated: candidates are a superset
    float reach2 = reach * reach;

    if (threadIdx.x < 3)
    {
        float c = threadIdx.x == 0 ? cx : (threadIdx.x == 1 ? cy : cz);
        int n = threadIdx.x == 0 ? g.nx : (threadIdx.x == 1 ? g.ny : g.nz);
        cell_lo[threadIdx.x] = min(max((int)floorf((c - reach) * g.inv_h), 0), n - 1);
        cell_hi[threadIdx.x] = min(max((int)floorf((c + reach) * g.inv_h), 0), n - 1);
    }
    if (threadIdx.x == 0) n_cand = 0;
    __syncthreads();

    // my two nodes: (lx, ly, lz) and (lx, ly, lz + 4)
    int lx = threadIdx.x & 7, ly = (threadIdx.x >> 3) & 7, lz = threadIdx.x >> 6;
    int nx_ = bxi * BRICK + lx, ny_ = byi * BRICK + ly, nz0 = bzi * BRICK + lz, nz1 = nz0 + 4;
    float fxs, fys, fz0s, fz1s;    // node coordinates (only used when !exact_level0)
    fxs = ox + (float)nx_ * s; fys = oy + (float)ny_ * s; fz0s = oz + (float)nz0 * s; fz1s = oz + (float)nz1 * s;
    float fxi = (float)nx_, fyi = (float)ny_, fz0i = (float)nz0, fz1i = (float)nz1;
    float best0 = FLT_MAX, best1 = FLT_MAX;

    int ex = cell_hi[0] - cell_lo[0] + 1, ey = cell_hi[1] - cell_lo[1] + 1, ez = cell_hi[2] - cell_lo[2] + 1;
    int n_rows = ey * ez;    // a "row" = run of cells along x: contiguous in the CSR order

    // take the exact minimum of my two nodes over the staged candidates, then empty the list
    auto flush = [&]()
    {
        int nc = n_cand;
        if (exact_level0)
        {
            for (int i = 0; i < nc; ++i)
            {
                float4 p = cand[i];
                float dxv = __fmaf_rn(fxi, s, -p.x);
                float dyv = __fmaf_rn(fyi, s, -p.y);
                float d0 = fg_sq3(dxv, dyv, __fmaf_rn(fz0i, s, -p.z));
                float d1 = fg_sq3(dxv, dyv, __fmaf_rn(fz1i, s, -p.z));
                best0 = best0 < d0 ? best0 : d0;
                best1 = best1 < d1 ? best1 : d1;
            }
        }
        else
        {
            for (int i = 0; i < nc; ++i)
            {
                float4 p = cand[i];
                float dxv = fxs - p.x, dyv = fys - p.y;
                float d0 = fg_sq3(dxv, dyv, fz0s - p.z);
                float d1 = fg_sq3(dxv, dyv, fz1s - p.z);
                best0 = best0 < d0 ? best0 : d0;
                best1 = best1 < d1 ? best1 : d1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) n_cand = 0;
        __syncthreads();
    };

    // Rows are looked up 256 at a time (one per thread), then walked by the whole block.
    for (int row_base = 0; row_base < n_rows; row_base += 256)
    {
        int row = row_base + threadIdx.x;
        int p_begin = 0, p_end = 0;
        if (row < n_rows)
        {
            int ry = cell_lo[1] + row % ey, rz = cell_lo[2] + row / ey;
            // distance from the centre to this row's (y, z) slab
            float y_lo = (float)ry * g.h, y_hi = y_lo + g.h, z_lo = (float)rz * g.h, z_hi = z_lo + g.h;
            float ddy = fmaxf(fmaxf(y_lo - cy, cy - y_hi), 0.0f);
            float ddz = fmaxf(fmaxf(z_lo - cz, cz - z_hi), 0.0f);
            // boundary cells also hold clamped points lying outside the lattice: never cull them
            bool edge = (ry == 0 || ry == g.ny - 1 || rz == 0 || rz == g.nz - 1);
            if (edge || ddy * ddy + ddz * ddz <= reach2)
            {
                int c0 = (rz * g.ny + ry) * g.nx + cell_lo[0];
                p_begin = g.start[c0];
                p_end = g.start[c0 + ex];
            }
        }
        __syncthreads();            // previous round's readers are done with row_b / row_e
        row_b[threadIdx.x] = p_begin;
        row_e[threadIdx.x] = p_end;
        __syncthreads();
        int rows_here = min(256, n_rows - row_base);
        for (int r = 0; r < rows_here; ++r)
        {
            int b0 = row_b[r], e0 = row_e[r];
            for (int p0 = b0; p0 < e0; p0 += 256)           // uniform across the block
            {
                int pi = p0 + threadIdx.x;
                if (pi < e0)
                {
                    float4 p = g.pts[pi];
                    float ddx = p.x - cx, ddyp = p.y - cy, ddzp = p.z - cz;
                    if (ddx * ddx + ddyp * ddyp + ddzp * ddzp <= reach2)
                    {
                        int slot = atomicAdd(&n_cand, 1);
                        cand[slot] = p;
                    }
                }
                __syncthreads();                            // all appends of this stride are visible
                // second barrier: nobody appends again before everyone has read the count
                if (__syncthreads_or(n_cand > BRICK_CAND - 256)) flush();
            }
        }
    }
    __syncthreads();
    flush();
    if (nx_ < dx && ny_ < dy)
    {
        if (nz0 < dz) out[((size_t)nz0 * dy + ny_) * dx + nx_] = best0;
        if (nz1 < dz) out[((size_t)nz1 * dy + ny_) * dx + nx_] = best1;
    }
}

// Coarse-level brute force over arbitrary lattice (origin o, spacing s): top of the hierarchy.
__global__ void __launch_bounds__(256)
k_lattice_brute(float* __restrict__ out, int dx, int dy, int dz, float ox, float oy, float oz, float s,
                const float4* __restrict__ P, int nt)
{
    __shared__ float4 tile[LUT_TILE];
    int n = dx * dy * dz;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int x = i % dx, y = (i / dx) % dy, z = i / (dx * dy);
    float fx = ox + (float)x * s, fy = oy + (float)y * s, fz = oz + (float)z * s;
    float best = FLT_MAX;
    for (int base = 0; base < nt; base += LUT_TILE)
    {
        int cnt = min(LUT_TILE, nt - base);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) tile[k] = P[base + k];
        __syncthreads();
        for (int k = 0; k < cnt; ++k)
        {
            float4 p = tile[k];
            float d = fg_sq3(fx - p.x, fy - p.y, fz - p.z);
            best = best < d ? best : d;
        }
    }
    if (i < n) out[i] = best;
}

// dense grid -> corner-packed cells.  Cell (cx, cy, cz) in [0,dx]x[0,dy]x[0,dz] corresponds to
// texel index i = c - 1 and stores T[clamp(i), clamp(i+1)] for the 8 corners, x fastest.
__global__ void __launch_bounds__(256)
k_pack_cells(const float* __restrict__ grid, int dx, int dy, int dz, float4* __restrict__ packed)
{
    size_t ncell = (size_t)(dx + 1) * (dy + 1) * (dz + 1);
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int cx = (int)(c % (dx + 1));
    int cy = (int)((c / (dx + 1)) % (dy + 1));
    int cz = (int)(c / ((size_t)(dx + 1) * (dy + 1)));
    int x0 = max(cx - 1, 0), x1 = min(cx, dx - 1);
    int y0 = max(cy - 1, 0), y1 = min(cy, dy - 1);
    int z0 = max(cz - 1, 0), z1 = min(cz, dz - 1);
    size_t sy = dx, sz = (size_t)dx * dy;
    float4 lo, hi;
    lo.x = grid[x0 + y0 * sy + z0 * sz]; lo.y = grid[x1 + y0 * sy + z0 * sz];
    lo.z = grid[x0 + y1 * sy + z0 * sz]; lo.w = grid[x1 + y1 * sy + z0 * sz];
    hi.x = grid[x0 + y0 * sy + z1 * sz]; hi.y = grid[x1 + y0 * sy + z1 * sz];
    hi.z = grid[x0 + y1 * sy + z1 * sz]; hi.w = grid[x1 + y1 * sy + z1 * sz];
#if FG_BRICKED
    c = fg_brick_cell(cx, cy, cz, (dx + 1 + 3) / 4, (dy + 1 + 3) / 4);
#endif
    packed[2 * c] = lo;
    packed[2 * c + 1] = hi;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

static int build_grid_brute(fgoicp_ctx* c, const float4* d_P)
{
    const LutDev& L = c->lut;
    dim3 grid((L.dx + 32 * LUT_XPT - 1) / (32 * LUT_XPT), (L.dy + 7) / 8, L.dz);
    k_lut_brute<<<grid, 256, 0, c->stream>>>(c->d_grid, L.dx, L.dy, L.dz, c->res, d_P, (int)c->nt);
    FG_CUDA(cudaGetLastError());
    return FGOICP_OK;
}

struct Level { int dx, dy, dz; float ox, oy, oz, s; float* d; };

// Bins the model cloud (LUT-space copy d_P, original copy c->d_model) into the uniform cell grid kept in
// the context.  Cell size: about four points per occupied cell for a surface-like cloud, at least two grid
// nodes wide, at most 256 cells per axis.
static int build_cell_grid(fgoicp_ctx* c, const float4* d_P)
{
    const LutDev& L = c->lut;
    cudaStream_t st = c->stream;
    int nt = (int)c->nt;
    float ext[3] = { L.dx * c->res, L.dy * c->res, L.dz * c->res };
    float max_ext = std::max(ext[0], std::max(ext[1], ext[2]));
    float area = ext[0] * ext[1] + ext[1] * ext[2] + ext[0] * ext[2];     // half the bounding-box surface
    float h = std::sqrt(4.0f * area / (float)nt);
    // FGOICP_NN_CELL_SCALE (experiment knob, default 1): smaller cells mean fewer candidates per far query of the NN
    // search and more (cheap) rows; any size gives the same exact results
    if (const char* e = getenv("FGOICP_NN_CELL_SCALE")) { float f = (float)atof(e); if (f >= 0.125f && f <= 8.0f) h *= f; }
    h = std::max(h, std::max(max_ext / 256.0f, 2.0f * c->res));
    c->cell_h = h; c->cell_inv_h = 1.0f / h;
    c->cnx = std::max(1, (int)std::ceil(ext[0] / h)); c->cny = std::max(1, (int)std::ceil(ext[1] / h)); c->cnz = std::max(1, (int)std::ceil(ext[2] / h));
    int ncell = c->cnx * c->cny * c->cnz;
    int *d_cell_of = nullptr, *d_counts = nullptr, *d_fill = nullptr;
    FG_CUDA(cudaMalloc(&d_cell_of, sizeof(int) * nt));
    FG_CUDA(cudaMalloc(&d_counts, sizeof(int) * ncell));
    FG_CUDA(cudaMalloc(&d_fill, sizeof(int) * ncell));
    FG_CUDA(cudaMalloc(&c->d_cell_start, sizeof(int) * (ncell + 1)));
    FG_CUDA(cudaMalloc(&c->d_cell_P, sizeof(float4) * nt));
    FG_CUDA(cudaMalloc(&c->d_cell_M, sizeof(float4) * nt));
    FG_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * ncell, st));
    FG_CUDA(cudaMemsetAsync(d_fill, 0, sizeof(int) * ncell, st));
    CellGrid g;
    g.start = c->d_cell_start; g.pts = c->d_cell_P;
    g.nx = c->cnx; g.ny = c->cny; g.nz = c->cnz; g.h = h; g.inv_h = c->cell_inv_h;
    k_cell_index<<<(nt + 255) / 256, 256, 0, st>>>(d_P, nt, g, d_cell_of, d_counts);
    k_scan_counts<<<1, 1024, 0, st>>>(d_counts, ncell, c->d_cell_start);
    k_cell_scatter<<<(nt + 255) / 256, 256, 0, st>>>(d_P, c->d_model, nt, d_cell_of, c->d_cell_start, d_fill, c->d_cell_P, c->d_cell_M);
    FG_CUDA(cudaGetLastError());
    FG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_cell_of); cudaFree(d_counts); cudaFree(d_fill);
    // coarse boxes (non-empty blocks of FG_COARSE^3 cells with the bounding box of their points), compacted on the host
    {
        const int Cx = (c->cnx + FG_COARSE - 1) / FG_COARSE, Cy = (c->cny + FG_COARSE - 1) / FG_COARSE, Cz = (c->cnz + FG_COARSE - 1) / FG_COARSE;
        const int nC = Cx * Cy * Cz;
        float4 *d_lo = nullptr, *d_hi = nullptr;
        FG_CUDA(cudaMalloc(&d_lo, sizeof(float4) * nC));
        FG_CUDA(cudaMalloc(&d_hi, sizeof(float4) * nC));
        k_coarse_boxes<<<(nC * 32 + 127) / 128, 128, 0, st>>>(g, Cx, Cy, Cz, d_lo, d_hi);
        FG_CUDA(cudaGetLastError());
        std::vector<float4> hlo(nC), hhi(nC), list;
        FG_CUDA(cudaMemcpyAsync(hlo.data(), d_lo, sizeof(float4) * nC, cudaMemcpyDeviceToHost, st));
        FG_CUDA(cudaMemcpyAsync(hhi.data(), d_hi, sizeof(float4) * nC, cudaMemcpyDeviceToHost, st));
        FG_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_lo); cudaFree(d_hi);
        for (int w = 0; w < nC; ++w)
        {
            int cnt; memcpy(&cnt, &hlo[w].w, 4);
            if (cnt <= 0) continue;
            const unsigned X = (unsigned)(w % Cx), Y = (unsigned)((w / Cx) % Cy), Z = (unsigned)(w / (Cx * Cy));
            const unsigned packed = X | (Y << 10) | (Z << 20);
            float pf; memcpy(&pf, &packed, 4);
            // cells of the block its points actually occupy, per axis (offsets 0..7 inside the block, 3 bits each): the binning
            // function of k_cell_index (one float multiply, floor, clamp) applied to the tight box; it is monotone, so every
            // point of the block lies in [cell(lo), cell(hi)].  A surface crossing an 8^3 block leaves most of its 64 rows empty.
            auto cell = [&](float v, int n) { int q = (int)std::floor(v * c->cell_inv_h); return std::min(std::max(q, 0), n - 1); };
            const int bx0 = (int)X * FG_COARSE, by0 = (int)Y * FG_COARSE, bz0 = (int)Z * FG_COARSE;
            const unsigned tight = (unsigned)(cell(hlo[w].x, c->cnx) - bx0) | ((unsigned)(cell(hhi[w].x, c->cnx) - bx0) << 3)
                                 | ((unsigned)(cell(hlo[w].y, c->cny) - by0) << 6) | ((unsigned)(cell(hhi[w].y, c->cny) - by0) << 9)
                                 | ((unsigned)(cell(hlo[w].z, c->cnz) - bz0) << 12) | ((unsigned)(cell(hhi[w].z, c->cnz) - bz0) << 15);
            float tf; memcpy(&tf, &tight, 4);
            list.push_back(make_float4(hlo[w].x, hlo[w].y, hlo[w].z, pf));
            list.push_back(make_float4(hhi[w].x, hhi[w].y, hhi[w].z, tf));
        }
        c->n_coarse = (int)(list.size() / 2);
        if (c->n_coarse > 0)
        {
            FG_CUDA(cudaMalloc(&c->d_coarse, sizeof(float4) * list.size()));
            FG_CUDA(cudaMemcpy(c->d_coarse, list.data(), sizeof(float4) * list.size(), cudaMemcpyHostToDevice));
        }
    }
    return FGOICP_OK;
}

static int build_grid_hier(fgoicp_ctx* c, const float4* d_P)
{
    const LutDev& L = c->lut;
    cudaStream_t st = c->stream;
    int nt = (int)c->nt;

    CellGrid g;
    g.start = c->d_cell_start; g.pts = c->d_cell_P;
    g.nx = c->cnx; g.ny = c->cny; g.nz = c->cnz; g.h = c->cell_h; g.inv_h = c->cell_inv_h;

    // ---- lattice pyramid: level 0 = the LUT itself; level l+1 = brick centres of level l
    std::vector<Level> lv;
    lv.push_back({ L.dx, L.dy, L.dz, 0.f, 0.f, 0.f, c->res, c->d_grid });
    while ((long long)lv.back().dx * lv.back().dy * lv.back().dz > 4096)
    {
        const Level& f = lv.back();
        Level k;
        k.dx = (f.dx + BRICK - 1) / BRICK; k.dy = (f.dy + BRICK - 1) / BRICK; k.dz = (f.dz + BRICK - 1) / BRICK;
        k.s = f.s * BRICK;
        k.ox = f.ox + 3.5f * f.s; k.oy = f.oy + 3.5f * f.s; k.oz = f.oz + 3.5f * f.s;
        k.d = nullptr;
        FG_CUDA(cudaMalloc(&k.d, sizeof(float) * (size_t)k.dx * k.dy * k.dz));
        lv.push_back(k);
    }
    // top level by brute force
    {
        const Level& t = lv.back();
        int n = t.dx * t.dy * t.dz;
        if (lv.size() == 1)
        {
            int rc = build_grid_brute(c, d_P);
            if (rc) return rc;
        }
        else
        {
            k_lattice_brute<<<(n + 255) / 256, 256, 0, st>>>(t.d, t.dx, t.dy, t.dz, t.ox, t.oy, t.oz, t.s, d_P, nt);
        }
        FG_CUDA(cudaGetLastError());
    }
    for (int l = (int)lv.size() - 2; l >= 0; --l)
    {
        const Level& f = lv[l];
        const Level& k = lv[l + 1];
        int nb = k.dx * k.dy * k.dz;
        k_lut_bricks<<<nb, 256, 0, st>>>(f.d, f.dx, f.dy, f.dz, f.ox, f.oy, f.oz, f.s, k.d, k.dx, k.dy, k.dz, g, l == 0 ? 1 : 0);
        FG_CUDA(cudaGetLastError());
    }
    FG_CUDA(cudaStreamSynchronize(st));
    for (size_t l = 1; l < lv.size(); ++l) cudaFree(lv[l].d);
    return FGOICP_OK;
}

static int build_texture(fgoicp_ctx* c)
{
    const LutDev& L = c->lut;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    FG_CUDA(cudaMalloc3DArray(&c->arr, &desc, make_cudaExtent(L.dx, L.dy, L.dz)));
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(c->d_grid, L.dx * sizeof(float), L.dx, L.dy);
    cp.dstArray = c->arr;
    cp.extent = make_cudaExtent(L.dx, L.dy, L.dz);
    cp.kind = cudaMemcpyDeviceToDevice;
    FG_CUDA(cudaMemcpy3DAsync(&cp, c->stream));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;   // registration.cu:226-228
    td.filterMode = cudaFilterModeLinear;                                                 // :229
    td.readMode = cudaReadModeElementType;                                                // :230
    td.normalizedCoords = 0;                                                              // :231
    FG_CUDA(cudaCreateTextureObject(&c->lut.tex, &rd, &td, nullptr));
    return FGOICP_OK;
}

int fg_icp_prealloc(fgoicp_ctx* c);

extern "C" const char* fgoicp_last_error(void) { return fg::g_last_error.c_str(); }
extern "C" const char* fgoicp_version(void) { return "fgoicp-b200 0.1 (sm_100a)"; }

extern "C" int fgoicp_ctx_create(const float* model_xyz, size_t nt, const float* data_xyz, size_t ns,
                                 const float bbox_min[3], const float bbox_max[3], float lut_resolution,
                                 int device, unsigned flags, fgoicp_ctx** out)
{
    FG_RANGE("fgoicp_ctx_create");
    FG_ARG(out != nullptr, "out is NULL");
    *out = nullptr;
    FG_ARG(model_xyz && data_xyz && bbox_min && bbox_max, "NULL input pointer");
    FG_ARG(nt > 0 && ns > 0, "empty point cloud");
    FG_ARG(nt < (size_t)1 << 31 && ns < (size_t)1 << 31, "point cloud too large");
    FG_ARG(lut_resolution > 0.0f, "lut_resolution must be positive");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
    {
        fg::set_error("no CUDA device available: this library has no CPU fallback");
        return FGOICP_ERR_CUDA;
    }
    FG_ARG(device >= 0 && device < ndev, "device index out of range");
    FG_CUDA(cudaSetDevice(device));

    // L2 fetch granularity is DEVICE-WIDE state shared with everything else in the process, so the library leaves it
    // alone by default (measured on B200: 32-byte fills change nothing for the 32-byte cell gathers -- an L2 miss
    // still moves a 64-byte DRAM atom).  FGOICP_L2_FETCH=32|64|128 sets it for experiments; the previous value is
    // restored by fgoicp_ctx_destroy.
    size_t prev_gran = 0;
    if (const char* eg = getenv("FGOICP_L2_FETCH"))
    {
        size_t gran = (size_t)atoi(eg);
        if ((gran == 32 || gran == 64 || gran == 128) && cudaDeviceGetLimit(&prev_gran, cudaLimitMaxL2FetchGranularity) == cudaSuccess)
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        else prev_gran = 0;
    }

    fgoicp_ctx* c = new fgoicp_ctx();
    c->device = device;
    c->restore_l2_fetch = prev_gran;
    // from here on every failure goes through fgoicp_ctx_destroy (no leaked context, streams or events)
#define FG_TRY0(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { int code__ = fg::cuda_fail(e__, #call, __FILE__, __LINE__); fgoicp_ctx_destroy(c); return code__; } } while (0)
    cudaDeviceProp prop;
    FG_TRY0(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    FG_TRY0(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    FG_TRY0(cudaEventCreate(&c->ev0));
    FG_TRY0(cudaEventCreate(&c->ev1));
#undef FG_TRY0
    c->nt = nt; c->ns = ns; c->res = lut_resolution;
    for (int a = 0; a < 3; ++a) { c->bbox_min[a] = bbox_min[a]; c->bbox_max[a] = bbox_max[a]; }

    // dims = ceil(range / res), scale = 1/res, offset = -bbox_min   (registration.cu:186-201)
    LutDev& L = c->lut;
    L.dx = (int)std::ceil((bbox_max[0] - bbox_min[0]) / lut_resolution);
    L.dy = (int)std::ceil((bbox_max[1] - bbox_min[1]) / lut_resolution);
    L.dz = (int)std::ceil((bbox_max[2] - bbox_min[2]) / lut_resolution);
    if (L.dx < 1 || L.dy < 1 || L.dz < 1 || L.dx >= 2048 || L.dy >= 2048 || L.dz >= 2048)
    {
        // the reference only logs here (registration.cu:191-194) and then fails inside CUDA
        fg::set_error("grid dims out of range [1, 2047]: lower the LUT resolution");
        fgoicp_ctx_destroy(c);
        return FGOICP_ERR_ARG;
    }
    L.scale = 1.0f / lut_resolution;
    L.ox = -bbox_min[0]; L.oy = -bbox_min[1]; L.oz = -bbox_min[2];

    // upload clouds: model as (x,y,z,index), data as (x,y,z,|p|^2 canonical)
    std::vector<float4> hm(nt), hd(ns), hP(nt);
    for (size_t i = 0; i < nt; ++i)
    {
        float x = model_xyz[3 * i], y = model_xyz[3 * i + 1], z = model_xyz[3 * i + 2];
        int idx = (int)i; float idx_bits; memcpy(&idx_bits, &idx, 4);
        hm[i] = make_float4(x, y, z, idx_bits);
        // host shift into LUT space, registration.cu:289-296 (plain fp32 adds)
        hP[i] = make_float4(x + L.ox, y + L.oy, z + L.oz, 0.f);
    }
    // The data points are kept on the device in MORTON order of their coordinates: consecutive lanes then query
    // neighbouring grid cells, which the memory system serves far better than scattered ones (W5: inner searches
    // 111 -> 89 ms, ICP 67 -> 62 ms).  Every result that is a sum over the points is unaffected (fp64 accumulation,
    // rounded once); per-point outputs (fgoicp_nn) and order-dependent tie rules go through d_data_orig.
    std::vector<int> orig(ns);
    for (size_t i = 0; i < ns; ++i) orig[i] = (int)i;
    if (!(flags & FGOICP_BUILD_KEEP_ORDER) && !getenv("FGOICP_KEEP_ORDER") && ns > 1)
    {
        float lo[3] = { data_xyz[0], data_xyz[1], data_xyz[2] }, hi[3] = { data_xyz[0], data_xyz[1], data_xyz[2] };
        for (size_t i = 0; i < ns; ++i)
            for (int a = 0; a < 3; ++a)
            {
                float v = data_xyz[3 * i + a];
                if (v < lo[a]) lo[a] = v;
                if (v > hi[a]) hi[a] = v;
            }
        std::vector<uint32_t> code(ns);
        for (size_t i = 0; i < ns; ++i)
        {
            uint32_t cde = 0;
            for (int a = 0; a < 3; ++a)
            {
                float ext = hi[a] - lo[a];
                float u = ext > 0.f ? (data_xyz[3 * i + a] - lo[a]) / ext : 0.f;
                if (!(u >= 0.f)) u = 0.f;                  // also NaN
                if (u > 1.f) u = 1.f;
                uint32_t q = (uint32_t)(u * 1023.0f);
                for (int b = 0; b < 10; ++b) cde |= ((q >> b) & 1u) << (3 * b + a);
            }
            code[i] = cde;
        }
        std::stable_sort(orig.begin(), orig.end(), [&](int a, int b) { return code[a] < code[b]; });
    }
    for (size_t s = 0; s < ns; ++s)
    {
        size_t i = (size_t)orig[s];
        float x = data_xyz[3 * i], y = data_xyz[3 * i + 1], z = data_xyz[3 * i + 2];
        float r2 = fmaf(z, z, fmaf(x, x, y * y));     // registration.cu:37-39 (SASS: FMUL y,y; FFMA x,x; FFMA z,z)
        hd[s] = make_float4(x, y, z, r2);
    }
    float4* d_P = nullptr;
    int rc = FGOICP_OK;
    auto fail = [&](int code) { if (d_P) cudaFree(d_P); fgoicp_ctx_destroy(c); return code; };
#define FG_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { return fail(fg::cuda_fail(e__, #call, __FILE__, __LINE__)); } } while (0)
    FG_TRY(cudaMalloc(&c->d_model, sizeof(float4) * nt));
    FG_TRY(cudaMalloc(&c->d_data, sizeof(float4) * ns));
    FG_TRY(cudaMalloc(&d_P, sizeof(float4) * nt));
    FG_TRY(cudaMemcpyAsync(c->d_model, hm.data(), sizeof(float4) * nt, cudaMemcpyHostToDevice, c->stream));
    FG_TRY(cudaMemcpyAsync(c->d_data, hd.data(), sizeof(float4) * ns, cudaMemcpyHostToDevice, c->stream));
    FG_TRY(cudaMalloc(&c->d_data_orig, sizeof(int) * ns));
    FG_TRY(cudaMemcpyAsync(c->d_data_orig, orig.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, c->stream));
    FG_TRY(cudaMemcpyAsync(d_P, hP.data(), sizeof(float4) * nt, cudaMemcpyHostToDevice, c->stream));
    size_t cells = (size_t)L.dx * L.dy * L.dz;
    FG_TRY(cudaMalloc(&c->d_grid, sizeof(float) * cells));
    L.grid = c->d_grid;

    FG_TRY(cudaEventRecord(c->ev0, c->stream));
    rc = build_cell_grid(c, d_P);
    if (rc) return fail(rc);
    if (flags & FGOICP_BUILD_BRUTE_LUT) rc = build_grid_brute(c, d_P);
    else rc = build_grid_hier(c, d_P);
    if (rc) return fail(rc);
    if (flags & FGOICP_BUILD_PACKED)
    {
        size_t pcells = (size_t)(L.dx + 1) * (L.dy + 1) * (L.dz + 1);
#if FG_BRICKED
        FG_TRY(cudaMalloc(&c->d_packed, (size_t)((L.dx + 4) / 4) * ((L.dy + 4) / 4) * ((L.dz + 4) / 4) * 64 * 32));
#else
        FG_TRY(cudaMalloc(&c->d_packed, pcells * 32));
#endif
        L.packed = c->d_packed;
        k_pack_cells<<<(unsigned)((pcells + 255) / 256), 256, 0, c->stream>>>(c->d_grid, L.dx, L.dy, L.dz, (float4*)c->d_packed);
        FG_TRY(cudaGetLastError());
    }
    if (flags & FGOICP_BUILD_TEX)
    {
        rc = build_texture(c);
        if (rc) return fail(rc);
    }
    fg_lut_finalise(L);
    FG_TRY(cudaEventRecord(c->ev1, c->stream));
    FG_TRY(cudaStreamSynchronize(c->stream));
    FG_TRY(cudaEventElapsedTime(&c->build_ms, c->ev0, c->ev1));
    cudaFree(d_P);
    d_P = nullptr;
#undef FG_TRY
    c->sampler = c->d_packed ? FGOICP_SAMPLER_PACKED : FGOICP_SAMPLER_GRID;
    if (const char* e = getenv("FGOICP_ICP_MODE")) { int m = atoi(e); c->icp_mode = (m >= 0 && m <= 2) ? m : 0; }
    // buffers of the refinements and of the level driver: allocated here, never inside run()
    rc = fg_icp_prealloc(c);
    if (rc) { fgoicp_ctx_destroy(c); return rc; }
    rc = fg::ensure_scratch(c, (size_t)8 << 20);
    if (rc) { fgoicp_ctx_destroy(c); return rc; }
    // phase-ordered evaluation is the default for the flat bound entry points (bit-identical results, ~1.6x);
    // FGOICP_PHASED=0 selects the plain kernel
    c->phased = c->d_packed != nullptr;
    if (const char* e = getenv("FGOICP_PHASED")) c->phased = atoi(e) != 0;
    *out = c;
    return FGOICP_OK;
}

extern "C" int fgoicp_ctx_destroy(fgoicp_ctx* c)
{
    if (!c) return FGOICP_OK;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->lut.tex) cudaDestroyTextureObject(c->lut.tex);
    if (c->arr) cudaFreeArray(c->arr);
    cudaFree(c->d_model); cudaFree(c->d_data); cudaFree(c->d_data_orig); cudaFree(c->d_grid); cudaFree(c->d_packed);
    cudaFree(c->d_scratch); cudaFree(c->d_work); cudaFree(c->d_nnkey); cudaFree(c->d_icp); cudaFree(c->d_inl); cudaFree(c->d_icp_part); cudaFree(c->d_icp_jobs); cudaFree(c->d_nnmemo); cudaFree(c->d_icp_loop);
    cudaFree(c->d_cell_start); cudaFree(c->d_cell_P); cudaFree(c->d_cell_M); cudaFree(c->d_coarse); cudaFree(c->d_trim); cudaFree(c->d_phase); cudaFree(c->d_rounds);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->restore_l2_fetch) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, c->restore_l2_fetch);
    delete c;
    return FGOICP_OK;
}

extern "C" int fgoicp_ctx_info(const fgoicp_ctx* c, fgoicp_info* o)
{
    FG_ARG(c && o, "NULL pointer");
    o->nt = c->nt; o->ns = c->ns;
    o->dims[0] = c->lut.dx; o->dims[1] = c->lut.dy; o->dims[2] = c->lut.dz;
    o->resolution = c->res; o->scale = c->lut.scale;
    o->offset[0] = c->lut.ox; o->offset[1] = c->lut.oy; o->offset[2] = c->lut.oz;
    o->grid_bytes = (uint64_t)c->lut.dx * c->lut.dy * c->lut.dz * 4;
    o->packed_bytes = c->d_packed ? (uint64_t)(c->lut.dx + 1) * (c->lut.dy + 1) * (c->lut.dz + 1) * 32 : 0;
    o->device = c->device; o->sm_count = c->sm_count; o->sampler = c->sampler;
    o->has_packed = c->d_packed != nullptr; o->has_tex = c->lut.tex != 0;
    o->build_ms = c->build_ms;
    return FGOICP_OK;
}

int fg_rounds_prealloc(fgoicp_ctx* c);
extern "C" int fgoicp_set_trim(fgoicp_ctx* c, float trim_fraction, uint64_t* inliers)
{
    FG_ARG(c, "NULL context");
    FG_ARG(trim_fraction >= 0.0f && trim_fraction < 1.0f, "trim_fraction must be in [0, 1)");
    // inliers kept: ns - floor(float(ns) * rho), fp32 product; all of them = trimming off
    size_t drop = (size_t)((float)c->ns * trim_fraction);
    size_t k = drop >= c->ns ? 1 : c->ns - drop;
    c->trim_k = (k == c->ns) ? 0 : k;
    if (inliers) *inliers = k;
    // trimmed searches run round-synchronously: their per-level scratch is allocated now, not inside run()
    if (c->trim_k > 0) return fg_rounds_prealloc(c);
    return FGOICP_OK;
}

extern "C" int fgoicp_set_sampler(fgoicp_ctx* c, int sampler)
{
    FG_ARG(c, "NULL context");
    FG_ARG(sampler >= 0 && sampler <= 2, "unknown sampler");
    if (sampler == FGOICP_SAMPLER_PACKED && !c->d_packed) { fg::set_error("packed grid was not built"); return FGOICP_ERR_STATE; }
    if (sampler == FGOICP_SAMPLER_TEX && !c->lut.tex) { fg::set_error("texture was not built"); return FGOICP_ERR_STATE; }
    c->sampler = sampler;
    return FGOICP_OK;
}

extern "C" int fgoicp_set_phased(fgoicp_ctx* c, int on)
{
    FG_ARG(c, "NULL context");
    c->phased = on != 0;
    return FGOICP_OK;
}

extern "C" int fgoicp_set_nn_mode(fgoicp_ctx* c, int mode)
{
    FG_ARG(c, "NULL context");
    FG_ARG(mode >= 0 && mode <= 3, "nn mode must be 0 (cell grid), 1 (brute force), 2 (cell grid, coarse boxes for every query) or 3 (cell grid, no coarse boxes)");
    c->nn_mode = mode;
    return FGOICP_OK;
}

extern "C" int fgoicp_set_stream(fgoicp_ctx* c, void* cuda_stream)
{
    FG_ARG(c, "NULL context");
    FG_CUDA(cudaSetDevice(c->device));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return FGOICP_OK;
}

extern "C" int fgoicp_lut_download(fgoicp_ctx* c, float* out, size_t out_floats)
{
    FG_ARG(c && out, "NULL pointer");
    size_t cells = (size_t)c->lut.dx * c->lut.dy * c->lut.dz;
    FG_ARG(out_floats >= cells, "output buffer too small");
    FG_CUDA(cudaSetDevice(c->device));
    FG_CUDA(cudaMemcpyAsync(out, c->d_grid, cells * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FGOICP_OK;
}
