// bounds_phased.cu -- z-phase-ordered variant of the fused bound evaluation (L2 locality).
//
// k_bounds_multi (bounds.cu) is bound by HBM random access: every evaluation gathers one 32-byte cell
// from a 1+ GB grid at an effectively random address, each miss costs ~3 sectors of DRAM traffic, and
// although every cell is used dozens of times per launch, two uses are too far apart in time to meet
// in L2.  This variant reorders the SAME evaluations so that they do meet:
//
//   1. k_phase_bin: per rotation cube, the rotated data points are bucketed by z' = (R p).z into
//      slices of width w = 1/32 (a stable counting sort: point order inside a bucket is ascending index).
//   2. k_bounds_phased: persistent blocks; every warp owns a fixed set of (rotation cube, translation
//      cube) pairs and sweeps a global phase counter phi = 0, 1, 2, ...  At phase phi, pair (r, c)
//      evaluates the points of bucket  b = phi - round(t_c.z / w):  all their queries have
//      q.z = z' + t.z inside ONE z-slab of the grid, the same slab for every warp of the GPU.  A slab
//      is ~20 MB of the corner-packed grid, so the few slabs in flight stay in the 126 MB L2 (cell
//      gathers carry an L2 evict_last hint, the streamed index lists evict_first) and each cell is
//      read from HBM about once per sweep instead of once per use.
//   Warps are not barrier-synchronised: each owns enough pairs (~50) that the work per phase is nearly
//   the same for everybody, so they drift apart by only a phase or two.
//
// Results: per-pair sums are accumulated in fp64 in bucket order (deterministic, no atomics), so they
// equal k_bounds_multi's up to the last bit of the fp64 accumulator, i.e. the same float in practice.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

#define PH_NB      128                 // z' buckets
#define PH_W       0.03125f            // bucket width; leaf translation cubes sit at odd multiples of 2w
#define PH_INV_W   32.0f
#define PH_ZMIN    (-2.0f)             // bucket 0 starts here: |R p| <= sqrt(3) < 2 for data in [-1,1]^3
#define PH_TZOFF   64                  // phase = bucket + round(t.z / w) + PH_TZOFF  (t.z in [-2, 2))
#define PH_PHASES  (PH_NB + 2 * PH_TZOFF)
#define PH_THREADS 256
#define PH_WARPS   8
#define PH_MAXPAIR 64                  // pairs per warp (shared-memory accumulators)
#define PH_ILP     4                   // points per lane per inner iteration

__device__ __forceinline__ int ph_bucket(float z)
{
    int b = (int)floorf((z - PH_ZMIN) * PH_INV_W);
    return min(max(b, 0), PH_NB - 1);
}

// One block per rotation cube: rotation matrix, sin(half-angle), stable z'-bucketing of the data points.
__global__ void __launch_bounds__(PH_THREADS)
k_phase_bin(const float4* __restrict__ data, int ns, const float4* __restrict__ rot, int fix_rot,
            float* __restrict__ Rmats /*[Rn][12]: R(9), sin_half, pad*/,
            unsigned short* __restrict__ order /*[Rn][ns]*/, int* __restrict__ off /*[Rn][PH_NB+1]*/)
{
    __shared__ float sR[9];
    __shared__ int s_cnt[PH_WARPS][PH_NB];      // per-warp bucket counts -> exclusive bases
    __shared__ int s_off[PH_NB + 1];
    const int r = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0)
    {
        float4 rc = rot[r];
        float Rm[9];
        fg_rotation_matrix(rc.x, rc.y, rc.z, Rm);
        for (int k = 0; k < 9; ++k) { sR[k] = Rm[k]; Rmats[12 * r + k] = Rm[k]; }
        Rmats[12 * r + 9] = fix_rot ? 0.0f : fg_rot_sin(rc.w);
    }
    for (int i = threadIdx.x; i < PH_WARPS * PH_NB; i += PH_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = sR[k];
    // each warp owns a contiguous chunk of points and walks it in index order
    const int per = (ns + PH_WARPS - 1) / PH_WARPS;
    const int c0 = w * per, c1 = min(ns, c0 + per);
    for (int base = c0; base < c1; base += 32)
    {
        int i = base + lane;
        int b = -1;
        if (i < c1) { float4 p = data[i]; b = ph_bucket(fg_rotate(R, p.x, p.y, p.z).z); }
        unsigned peers = __match_any_sync(0xffffffffu, b);
        if (b >= 0 && lane == __ffs(peers) - 1) s_cnt[w][b] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan over (bucket, warp): bucket-major so a bucket's points are contiguous, warps in order
    if (threadIdx.x < PH_NB)
    {
        int tot = 0;
        for (int k = 0; k < PH_WARPS; ++k) tot += s_cnt[k][threadIdx.x];
        s_off[threadIdx.x + 1] = tot;
    }
    if (threadIdx.x == 0) s_off[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0) for (int b = 0; b < PH_NB; ++b) s_off[b + 1] += s_off[b];
    __syncthreads();
    if (threadIdx.x < PH_NB)
    {
        int run = s_off[threadIdx.x];
        for (int k = 0; k < PH_WARPS; ++k) { int n = s_cnt[k][threadIdx.x]; s_cnt[k][threadIdx.x] = run; run += n; }
    }
    for (int i = threadIdx.x; i <= PH_NB; i += PH_THREADS) off[(size_t)r * (PH_NB + 1) + i] = s_off[i];
    __syncthreads();
    for (int base = c0; base < c1; base += 32)
    {
        int i = base + lane;
        int b = -1;
        if (i < c1) { float4 p = data[i]; b = ph_bucket(fg_rotate(R, p.x, p.y, p.z).z); }
        unsigned peers = __match_any_sync(0xffffffffu, b);
        if (b >= 0)
        {
            int rank = __popc(peers & ((1u << lane) - 1u));
            order[(size_t)r * ns + s_cnt[w][b] + rank] = (unsigned short)i;
        }
        __syncwarp();
        if (b >= 0 && lane == __ffs(peers) - 1) s_cnt[w][b] += __popc(peers);
        __syncwarp();
    }
}

// 256-bit gather with an L2 evict_last hint (the slab in flight should stay resident)
__device__ __forceinline__ void fg_ld256_keep(const float* p, float (&v)[8])
{
    asm("ld.global.nc.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
        : "l"(p));
}

__device__ __forceinline__ void ph_issue(const LutDev& L, float qx, float qy, float qz, SampleReq& r)
{
    float ux = __fmul_rn(__fadd_rn(qx, L.ox), L.scale);
    float uy = __fmul_rn(__fadd_rn(qy, L.oy), L.scale);
    float uz = __fmul_rn(__fadd_rn(qz, L.oz), L.scale);
    int ix, iy, iz;
    fg_tex_axis(ux, L.dx, ix, r.a);
    fg_tex_axis(uy, L.dy, iy, r.b);
    fg_tex_axis(uz, L.dz, iz, r.c);
    int cx = min(max(ix, -1), L.dx - 1) + 1;
    int cy = min(max(iy, -1), L.dy - 1) + 1;
    int cz = min(max(iz, -1), L.dz - 1) + 1;
    size_t cell = ((size_t)cz * (size_t)(L.dy + 1) + (size_t)cy) * (size_t)(L.dx + 1) + (size_t)cx;
    fg_ld256_keep(L.packed + cell * 8, r.v);
}

__global__ void __launch_bounds__(PH_THREADS)
k_bounds_phased(LutDev L, const float4* __restrict__ data, int ns, int fix_rot,
                const float4* __restrict__ tcubes, int n_pairs, int T,
                const float* __restrict__ Rmats, const unsigned short* __restrict__ order, const int* __restrict__ off,
                float* __restrict__ lb, float* __restrict__ ub, unsigned int* __restrict__ best_ub_bits)
{
    __shared__ double s_acc[PH_WARPS][PH_MAXPAIR][2];
    __shared__ float4 s_tc[PH_WARPS][PH_MAXPAIR];
    __shared__ short s_tzb[PH_WARPS][PH_MAXPAIR];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * PH_WARPS + w, nw = (long long)gridDim.x * PH_WARPS;
    const int pair0 = (int)(gw * n_pairs / nw), pair1 = (int)((gw + 1) * n_pairs / nw);
    const int np = pair1 - pair0;                              // <= PH_MAXPAIR by launch geometry
    for (int k = lane; k < np; k += 32)
    {
        float4 t = tcubes[pair0 + k];
        s_tc[w][k] = t;
        s_tzb[w][k] = (short)(__float2int_rn(t.z * PH_INV_W) + PH_TZOFF);
        s_acc[w][k][0] = 0.0; s_acc[w][k][1] = 0.0;
    }
    __syncwarp();

    int cur_r = -1;
    float R[9], sin_half = 0.f;
    for (int phi = 0; phi < PH_PHASES; ++phi)
    {
        for (int k = 0; k < np; ++k)
        {
            const int b = phi - (int)s_tzb[w][k];
            if (b < 0 || b >= PH_NB) continue;
            const int r = (pair0 + k) / T;
            const int* ro = off + (size_t)r * (PH_NB + 1) + b;
            const int k0 = __ldg(ro), k1 = __ldg(ro + 1);
            if (k0 == k1) continue;
            if (r != cur_r)
            {
                cur_r = r;
#pragma unroll
                for (int j = 0; j < 9; ++j) R[j] = __ldg(Rmats + 12 * r + j);
                sin_half = __ldg(Rmats + 12 * r + 9);
            }
            const float4 t = s_tc[w][k];
            const unsigned short* ord = order + (size_t)r * ns;
            double au = 0.0, al = 0.0;
            for (int j0 = k0 + lane; j0 < k1; j0 += 32 * PH_ILP)
            {
                SampleReq req[PH_ILP];
                float rot_r[PH_ILP];
                bool live[PH_ILP];
#pragma unroll
                for (int u = 0; u < PH_ILP; ++u)
                {
                    int j = j0 + 32 * u;
                    live[u] = j < k1;
                    int idx = live[u] ? (int)__ldcs(ord + j) : 0;          // streamed once per use: evict-first
                    float4 p = __ldg(&data[idx]);
                    float3 rp = fg_rotate(R, p.x, p.y, p.z);
                    rot_r[u] = __fmul_rn(__fadd_rn(p.w, p.w), sin_half);
                    ph_issue(L, __fadd_rn(rp.x, t.x), __fadd_rn(rp.y, t.y), __fadd_rn(rp.z, t.z), req[u]);
                }
#pragma unroll
                for (int u = 0; u < PH_ILP; ++u)
                {
                    float uu, ll;
                    fg_bound_terms(fg_sample_finish<FGOICP_SAMPLER_PACKED>(req[u]), rot_r[u], fix_rot != 0, t.w, uu, ll);
                    if (live[u]) { au += (double)uu; al += (double)ll; }
                }
            }
            au = fg_warp_sum(au); al = fg_warp_sum(al);
            if (lane == 0) { s_acc[w][k][0] += au; s_acc[w][k][1] += al; }
        }
    }
    __syncwarp();
    for (int k = lane; k < np; k += 32)
    {
        float fu = (float)s_acc[w][k][0], fl = (float)s_acc[w][k][1];
        ub[pair0 + k] = fu; lb[pair0 + k] = fl;
        if (best_ub_bits) atomicMin(best_ub_bits, __float_as_uint(fu));
    }
}

// host: returns FGOICP_OK and runs the phased path, or 1 if the problem does not fit it (caller falls back)
int fg_bounds_phased(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, const float4* d_tc, int T,
                     float* d_lb, float* d_ub, float* d_best_ub)
{
    if (!c->d_packed || c->ns > 65535) return 1;
    long long n_pairs = (long long)Rn * T;
    if (n_pairs > (1LL << 30)) return 1;
    // persistent blocks, all co-resident (a second wave would start its sweep out of phase)
    int per_sm = 0;
    FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounds_phased, PH_THREADS, 0));
    if (per_sm < 1) return 1;
    if (const char* e = getenv("FGOICP_PHASED_BPS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
    int blocks = per_sm * c->sm_count;
    if ((n_pairs + (long long)blocks * PH_WARPS - 1) / ((long long)blocks * PH_WARPS) > PH_MAXPAIR) return 1;
    size_t b_R = ((sizeof(float) * 12 * Rn) + 255) & ~(size_t)255;
    size_t b_off = ((sizeof(int) * (PH_NB + 1) * (size_t)Rn) + 255) & ~(size_t)255;
    size_t b_ord = sizeof(unsigned short) * (size_t)Rn * c->ns;
    if (b_R + b_off + b_ord > c->phase_bytes)
    {
        FG_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_phase); c->d_phase = nullptr; c->phase_bytes = 0;
        FG_CUDA(cudaMalloc(&c->d_phase, b_R + b_off + b_ord));
        c->phase_bytes = b_R + b_off + b_ord;
    }
    char* base = (char*)c->d_phase;
    float* d_R = (float*)base;
    int* d_off = (int*)(base + b_R);
    unsigned short* d_ord = (unsigned short*)(base + b_R + b_off);
    k_phase_bin<<<Rn, PH_THREADS, 0, c->stream>>>(c->d_data, (int)c->ns, d_rot, fix_rot, d_R, d_ord, d_off);
    FG_CUDA(cudaGetLastError());
    unsigned int* d_bits = (unsigned int*)d_best_ub;
    if (d_bits) FG_CUDA(cudaMemsetAsync(d_bits, 0x7f, 4, c->stream));   // 0x7f7f7f7f: a huge positive float
    k_bounds_phased<<<blocks, PH_THREADS, 0, c->stream>>>(c->lut, c->d_data, (int)c->ns, fix_rot, d_tc, (int)n_pairs, T,
                                                        d_R, d_ord, d_off, d_lb, d_ub, d_bits);
    FG_CUDA(cudaGetLastError());
    return FGOICP_OK;
}
