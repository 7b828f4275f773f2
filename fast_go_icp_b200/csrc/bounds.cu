// bounds.cu -- fused batched upper/lower bound evaluation over
// (rotation cube x translation cube x data point).
//
// Replaces Registration::compute_sse_error(rnode, tnodes, fix_rot, pool) of the reference
// (fgoicp/registration.cu:88-152): there, every translation cube costs one 32-thread-block
// launch of kernComputeBounds (registration.cu:27-60) writing 2 floats per point to scratch,
// two blocking thrust::reduce calls, plus cudaMalloc/cudaFree per batch.  Here one launch
// covers a whole list of (rotation cube, translation cube) pairs, the per-point terms are
// reduced in registers -> warp shuffles -> shared memory and never touch HBM.
//
// Work decomposition (one thread block = one rotation cube x one slice of the data points):
//   - the block stages up to 32 translation cubes at a time in shared memory;
//   - warps are arranged as Wc cube-groups x Wp point-slices; a warp owns CPW=4 cubes and
//     strides over its points with coalesced float4 loads (x, y, z, |p|^2), rotates each point
//     once and evaluates its 4 cubes -> 4 independent 32-byte gathers in flight per lane;
//   - per-cube sums are accumulated in fp64 (B200 has full-rate-enough FP64: 2 DADD per
//     evaluation), reduced by butterfly shuffles, then across point-slices in a fixed order,
//     so results are deterministic and independent of the launch geometry up to 1 ulp of fp64.
#include "common.cuh"

#include <algorithm>
#include <vector>

#include "bounds_eval.cuh"
#include "trim.cuh"

#define BD_THREADS 256
#define BD_WARPS   (BD_THREADS / 32)

#ifndef BNB_MIN_BLOCKS
#define BNB_MIN_BLOCKS 2      // 2 blocks of 8 warps per SM (<= 128 registers); 1 lets ptxas take 139 and halves the occupancy (inner searches 84 -> 104 ms), 3 spills (108 ms)
#endif
template <int SAMPLER>
__global__ void __launch_bounds__(BD_THREADS, BNB_MIN_BLOCKS)
k_bounds_multi(LutDev L, const float4* __restrict__ data, int ns,
               const float4* __restrict__ rot, const float* __restrict__ Rmats, int fix_rot,
               const float4* __restrict__ tcubes, int T, int S,
               double* __restrict__ partial, float* __restrict__ lb, float* __restrict__ ub,
               unsigned int* __restrict__ best_ub_bits, const int* __restrict__ counts)
{
    __shared__ float sR[9];
    __shared__ float s_sin;
    __shared__ float4 s_tc[BD_CHUNK];
    __shared__ double s_part[BD_WARPS][BD_CPW][2];

    const int r = blockIdx.x / S;
    const int slice = blockIdx.x - r * S;
    // optional per-rotation-cube cube count (<= T): the rest of the row is unused (round-synchronous search)
    const int Tr = counts ? min(T, counts[r]) : T;
    if (Tr <= 0) return;
    if (threadIdx.x == 0)
    {
        float4 rc = rot[r];
        if (Rmats)
        {
            for (int k = 0; k < 9; ++k) sR[k] = Rmats[9 * r + k];
        }
        else
        {
            float Rm[9];
            fg_rotation_matrix(rc.x, rc.y, rc.z, Rm);      // R = I outside the ball (common.hpp:41)
            for (int k = 0; k < 9; ++k) sR[k] = Rm[k];
        }
        s_sin = fix_rot ? 0.0f : fg_rot_sin(rc.w);
    }

    // this block's slice of the data points
    const int per = (ns + S - 1) / S;
    const int p0 = slice * per;
    const int p1 = min(ns, p0 + per);

    for (int c0 = 0; c0 < Tr; c0 += BD_CHUNK)
    {
        const int nch = min(BD_CHUNK, Tr - c0);
        __syncthreads();
        if (threadIdx.x < nch) s_tc[threadIdx.x] = tcubes[(size_t)r * T + c0 + threadIdx.x];
        __syncthreads();

        fg_eval_chunk<SAMPLER, BD_WARPS>(L, data, p0, p1, sR, s_sin, fix_rot != 0, s_tc, nch, s_part);
        __syncthreads();
        if (threadIdx.x < nch)
        {
            int c = threadIdx.x;
            double su, sl;
            fg_eval_gather<BD_WARPS>(s_part, nch, c, su, sl);
            size_t o = (size_t)r * T + c0 + c;
            if (S == 1)
            {
                float fu = (float)su, fl = (float)sl;
                ub[o] = fu; lb[o] = fl;
                if (best_ub_bits) atomicMin(best_ub_bits, __float_as_uint(fu));   // fu >= 0: bit order = value order
            }
            else
            {
                partial[(o * S + slice) * 2 + 0] = su;
                partial[(o * S + slice) * 2 + 1] = sl;
            }
        }
    }
}

// Trimmed bounds (extension): one block per (rotation cube, translation cube) pair.  Both per-point terms of a pair are
// monotone functions of ONE number, the point's signed residual  e_i = sqrt(d2_i) - rot_r_i  (fg_bound_terms:
// ub_i = max(e_i, 0)^2, lb_i = max(fma(span, -sqrt3, e_i), 0)^2, every step non-decreasing in e_i under fp32 rounding),
// so the K smallest ub_i and the K smallest lb_i are the terms of the K smallest e_i -- as multisets, which is all an
// order-independent fp64 sum sees.  The block therefore keeps ONE float per point (the order-preserving integer key
// of e_i) in shared memory, runs ONE exact radix select (trim.cuh) and forms both sums from the selected residuals:
// half the shared memory (4 x ns bytes: five blocks per SM instead of two at ns = 10,000, i.e. 2.5x the gathers in
// flight) and half the select passes of the first version, which selected ub_i and lb_i separately; same bits.
// Unused slots (negative span) are skipped.  Clouds whose keys do not fit shared memory (ns > 51,200) keep them in a
// per-block slice of a global scratch buffer instead (GLOBAL = 1: the blocks are persistent and walk the pairs, the
// select's passes are L2 hits).
#define BT_THREADS 256
__device__ __forceinline__ unsigned int fg_order_key(float x)          // a < b  <=>  key(a) < key(b)  (-0 < +0, NaN on top)
{
    const unsigned int b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fg_order_value(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <int SAMPLER, int GLOBAL>
__global__ void __launch_bounds__(BT_THREADS)
k_bounds_trim(LutDev L, const float4* __restrict__ data, int ns,
              const float4* __restrict__ rot, const float* __restrict__ Rmats, int fix_rot,
              const float4* __restrict__ tcubes, int T, int n_pairs, unsigned int K,
              float* __restrict__ lb, float* __restrict__ ub, unsigned int* __restrict__ best_ub_bits, unsigned int* gscratch)
{
    extern __shared__ unsigned int sv_shared[];   // [ns]: order key of e_i (GLOBAL = 0)
    unsigned int* sv = GLOBAL ? gscratch + (size_t)blockIdx.x * (size_t)ns : sv_shared;
    __shared__ float sR[9];
    __shared__ float s_sin;
    __shared__ unsigned int s_hist[256], s_state[2];
    __shared__ double s_w[32];
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
    {
    const int r = pair / T;
    const float4 t = tcubes[pair];
    if (t.w < 0.0f) continue;
    __syncthreads();                              // the previous pair's residuals and rotation are no longer read
    if (threadIdx.x == 0)
    {
        float4 rc = rot[r];
        if (Rmats) { for (int k = 0; k < 9; ++k) sR[k] = Rmats[9 * r + k]; }
        else
        {
            float Rm[9];
            fg_rotation_matrix(rc.x, rc.y, rc.z, Rm);
            for (int k = 0; k < 9; ++k) sR[k] = Rm[k];
        }
        s_sin = fix_rot ? 0.0f : fg_rot_sin(rc.w);
    }
    __syncthreads();
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = sR[k];
    const float sin_half = s_sin;
    // four points per thread and pass: four independent gathers in flight per lane (one cube per block leaves no
    // other source of memory-level parallelism)
    for (int i0 = threadIdx.x; i0 < ns; i0 += 4 * BT_THREADS)
    {
        SampleReq req[4];
        float rot_r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int i = min(i0 + k * BT_THREADS, ns - 1);
            float4 p = __ldg(&data[i]);
            float3 rp = fg_rotate(R, p.x, p.y, p.z);
            rot_r[k] = __fmul_rn(__fadd_rn(p.w, p.w), sin_half);
            fg_sample_issue<SAMPLER>(L, __fadd_rn(rp.x, t.x), __fadd_rn(rp.y, t.y), __fadd_rn(rp.z, t.z), req[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int i = i0 + k * BT_THREADS;
            if (i < ns) sv[i] = fg_order_key(fg_bound_residual(fg_sample_finish<SAMPLER>(req[k]), rot_r[k], fix_rot != 0));
        }
    }
    __syncthreads();
    const unsigned int* sk = sv;
    const FgSelect sel = fg_block_select([&](int i) { return sk[i]; }, ns, K, s_hist, s_state);
    double pu = 0.0, pl = 0.0;
    for (int i = threadIdx.x; i < ns; i += BT_THREADS)
    {
        const unsigned int key = sk[i];
        if (key < sel.vk_bits)
        {
            float u, l;
            fg_bound_from_residual(fg_order_value(key), t.w, u, l);
            pu += (double)u; pl += (double)l;
        }
    }
    const double tu = fg_block_sum1(pu, s_w);
    const double tl = fg_block_sum1(pl, s_w);
    if (threadIdx.x == 0)
    {
        float uk, lk;
        fg_bound_from_residual(fg_order_value(sel.vk_bits), t.w, uk, lk);
        float fu = (float)(tu + (double)sel.take_eq * (double)uk), fl = (float)(tl + (double)sel.take_eq * (double)lk);
        ub[pair] = fu; lb[pair] = fl;
        if (best_ub_bits) atomicMin(best_ub_bits, __float_as_uint(fu));
    }
    }
}

__global__ void k_bounds_finish(const double* __restrict__ partial, int n, int S,
                                float* __restrict__ lb, float* __restrict__ ub,
                                unsigned int* __restrict__ best_ub_bits)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    double su = 0.0, sl = 0.0;
    for (int s = 0; s < S; ++s) { su += partial[((size_t)o * S + s) * 2]; sl += partial[((size_t)o * S + s) * 2 + 1]; }
    float fu = (float)su, fl = (float)sl;
    ub[o] = fu; lb[o] = fl;
    if (best_ub_bits) atomicMin(best_ub_bits, __float_as_uint(fu));
}

template <int SAMPLER>
__global__ void k_lut_sample(LutDev L, const float* __restrict__ q, size_t n, float* __restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fg_sample<SAMPLER>(L, q[3 * i], q[3 * i + 1], q[3 * i + 2]);
}

__global__ void k_rot_sin(const float* __restrict__ spans, int n, float* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fg_rot_sin(spans[i]);
}

__global__ void k_set_u32(unsigned int* p, unsigned int v) { *p = v; }

// ---------------------------------------------------------------------------------------------

static int pick_slices(const fgoicp_ctx* c, int Rn)
{
    // enough blocks for ~4 per SM; at least 256 points per block
    long long want = 4LL * c->sm_count;
    int S = (int)std::max(1LL, (want + Rn - 1) / Rn);
    int maxS = (int)std::max((size_t)1, c->ns / 256);
    return std::min(S, maxS);
}

struct BoundsLaunch
{
    const float4* d_rot; const float* d_Rmats; int Rn; int fix_rot;
    const float4* d_tc; int T; float* d_lb; float* d_ub; float* d_best_ub; double* d_partial; int S;
    const int* d_counts = nullptr; int no_phased = 0;
};

int fg_bounds_phased(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, const float4* d_tc, int T,
                     float* d_lb, float* d_ub, float* d_best_ub);

static int run_bounds(fgoicp_ctx* c, const BoundsLaunch& b)
{
    FG_RANGE("fgoicp bounds");
    if (c->trim_k > 0)
    {
        // trimmed registration: every pair sums its K smallest terms (one block per pair)
        size_t smem = sizeof(unsigned int) * c->ns;
        const bool global_terms = smem > 200 * 1024 || getenv("FGOICP_TRIM_GLOBAL") != nullptr;   // ns > 51,200: residual keys in HBM / L2
        unsigned int* d_bits = (unsigned int*)b.d_best_ub;
        if (d_bits) k_set_u32<<<1, 1, 0, c->stream>>>(d_bits, 0x7f800000u);
        const int n_pairs = b.Rn * b.T;
        dim3 grid((unsigned)n_pairs);
        unsigned int* d_terms = nullptr;
        if (global_terms)
        {
            grid = dim3((unsigned)std::min(n_pairs, 4 * c->sm_count));
            if (c->trim_bytes < (size_t)grid.x * smem)
            {
                FG_CUDA(cudaStreamSynchronize(c->stream));
                cudaFree(c->d_trim); c->d_trim = nullptr; c->trim_bytes = 0;
                FG_CUDA(cudaMalloc(&c->d_trim, (size_t)4 * c->sm_count * smem));
                c->trim_bytes = (size_t)4 * c->sm_count * smem;
            }
            d_terms = (unsigned int*)c->d_trim;
            smem = 0;
        }
#define FG_LAUNCH_TRIM(SMP)                                                                                              \
        do {                                                                                                             \
            if (global_terms)                                                                                            \
                k_bounds_trim<SMP, 1><<<grid, BT_THREADS, 0, c->stream>>>(c->lut, c->d_data, (int)c->ns, b.d_rot, b.d_Rmats, \
                    b.fix_rot, b.d_tc, b.T, n_pairs, (unsigned int)c->trim_k, b.d_lb, b.d_ub, d_bits, d_terms);          \
            else                                                                                                         \
            {                                                                                                            \
                FG_CUDA(cudaFuncSetAttribute(k_bounds_trim<SMP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                k_bounds_trim<SMP, 0><<<grid, BT_THREADS, smem, c->stream>>>(c->lut, c->d_data, (int)c->ns, b.d_rot, b.d_Rmats, \
                    b.fix_rot, b.d_tc, b.T, n_pairs, (unsigned int)c->trim_k, b.d_lb, b.d_ub, d_bits, nullptr);          \
            }                                                                                                            \
        } while (0)
        switch (c->sampler)
        {
        case FGOICP_SAMPLER_PACKED: FG_LAUNCH_TRIM(FGOICP_SAMPLER_PACKED); break;
        case FGOICP_SAMPLER_TEX:    FG_LAUNCH_TRIM(FGOICP_SAMPLER_TEX); break;
        default:                    FG_LAUNCH_TRIM(FGOICP_SAMPLER_GRID); break;
        }
#undef FG_LAUNCH_TRIM
        FG_CUDA(cudaGetLastError());
        return FGOICP_OK;
    }
    if (c->phased && !b.no_phased && c->sampler == FGOICP_SAMPLER_PACKED && !b.d_Rmats && (long long)b.Rn * b.T >= 4096)
    {
        int rc = fg_bounds_phased(c, b.d_rot, b.Rn, b.fix_rot, b.d_tc, b.T, b.d_lb, b.d_ub, b.d_best_ub);
        if (rc <= 0) return rc;       // done (0) or error (<0); 1 = not applicable, fall through
    }
    unsigned int* d_bits = (unsigned int*)b.d_best_ub;
    if (d_bits) k_set_u32<<<1, 1, 0, c->stream>>>(d_bits, 0x7f800000u);
    dim3 grid((unsigned)(b.Rn * b.S));
#define FG_LAUNCH_BOUNDS(SMP)                                                                        \
    k_bounds_multi<SMP><<<grid, BD_THREADS, 0, c->stream>>>(c->lut, c->d_data, (int)c->ns, b.d_rot,  \
        b.d_Rmats, b.fix_rot, b.d_tc, b.T, b.S, b.d_partial, b.d_lb, b.d_ub, b.S == 1 ? d_bits : nullptr, b.d_counts)
    switch (c->sampler)
    {
    case FGOICP_SAMPLER_PACKED: FG_LAUNCH_BOUNDS(FGOICP_SAMPLER_PACKED); break;
    case FGOICP_SAMPLER_TEX:    FG_LAUNCH_BOUNDS(FGOICP_SAMPLER_TEX); break;
    default:                    FG_LAUNCH_BOUNDS(FGOICP_SAMPLER_GRID); break;
    }
#undef FG_LAUNCH_BOUNDS
    FG_CUDA(cudaGetLastError());
    if (b.S > 1)
    {
        int n = b.Rn * b.T;
        k_bounds_finish<<<(n + 255) / 256, 256, 0, c->stream>>>(b.d_partial, n, b.S, b.d_lb, b.d_ub, d_bits);
        FG_CUDA(cudaGetLastError());
    }
    return FGOICP_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Plain (not phase-ordered) kernel over device-resident cube lists with a per-rotation-cube count; only
// `active` of the Rn rows have work (the launch geometry is sized for those).  d_partial: scratch of
// 2 * Rn * T * S doubles provided by the caller when S > 1 (S from fg_bounds_slices).
int fg_bounds_slices(const fgoicp_ctx* c, int active) { return pick_slices(c, std::max(1, active)); }
int fg_bounds_plain_counts(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, const float4* d_tc, int T,
                           const int* d_counts, int S, double* d_partial, float* d_lb, float* d_ub)
{
    BoundsLaunch b;
    b.d_rot = d_rot; b.d_Rmats = nullptr; b.Rn = Rn; b.fix_rot = fix_rot; b.d_tc = d_tc; b.T = T;
    b.d_lb = d_lb; b.d_ub = d_ub; b.d_best_ub = nullptr; b.d_partial = d_partial; b.S = S;
    b.d_counts = d_counts; b.no_phased = 1;
    return run_bounds(c, b);
}

// host-buffer path: stage through pinned memory, one H2D, one launch, one D2H
static int bounds_host(fgoicp_ctx* c, const float* rot_xyz_span, const float* Rmats, int Rn, int fix_rot,
                       const float* t_xyz_span, int T, float* lb, float* ub)
{
    FG_ARG(c && rot_xyz_span && t_xyz_span && lb && ub, "NULL pointer");
    FG_ARG(Rn > 0 && T > 0, "Rn and T must be positive");
    FG_CUDA(cudaSetDevice(c->device));
    int S = pick_slices(c, Rn);
    size_t n = (size_t)Rn * T;
    size_t b_rot = align256(sizeof(float4) * Rn);
    size_t b_R = align256(Rmats ? sizeof(float) * 9 * Rn : 0);
    size_t b_tc = align256(sizeof(float4) * n);
    size_t b_out = align256(sizeof(float) * n);
    size_t b_part = S > 1 ? align256(sizeof(double) * 2 * n * S) : 0;
    size_t in_bytes = b_rot + b_R + b_tc;
    int rc = fg::ensure_scratch(c, in_bytes + 2 * b_out + b_part);
    if (rc) return rc;
    rc = fg::ensure_pinned(c, in_bytes + 2 * b_out);
    if (rc) return rc;
    char* hp = (char*)c->h_pinned;
    char* dp = (char*)c->d_scratch;
    memcpy(hp, rot_xyz_span, sizeof(float4) * Rn);
    if (Rmats) memcpy(hp + b_rot, Rmats, sizeof(float) * 9 * Rn);
    memcpy(hp + b_rot + b_R, t_xyz_span, sizeof(float4) * n);
    FG_CUDA(cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, c->stream));
    BoundsLaunch b;
    b.d_rot = (const float4*)dp;
    b.d_Rmats = Rmats ? (const float*)(dp + b_rot) : nullptr;
    b.Rn = Rn; b.fix_rot = fix_rot;
    b.d_tc = (const float4*)(dp + b_rot + b_R);
    b.T = T;
    b.d_lb = (float*)(dp + in_bytes);
    b.d_ub = (float*)(dp + in_bytes + b_out);
    b.d_best_ub = nullptr;
    b.d_partial = S > 1 ? (double*)(dp + in_bytes + 2 * b_out) : nullptr;
    b.S = S;
    rc = run_bounds(c, b);
    if (rc) return rc;
    FG_CUDA(cudaMemcpyAsync(hp + in_bytes, dp + in_bytes, 2 * b_out, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(lb, hp + in_bytes, sizeof(float) * n);
    memcpy(ub, hp + in_bytes + b_out, sizeof(float) * n);
    return FGOICP_OK;
}

extern "C" int fgoicp_bounds_batch(fgoicp_ctx* c, const float R[9], float rot_span, int fix_rot,
                                   const float* t_xyz_span, int T, float* lb, float* ub)
{
    FG_ARG(R != nullptr, "R is NULL");
    float rot[4] = { 0.f, 0.f, 0.f, rot_span };
    return bounds_host(c, rot, R, 1, fix_rot, t_xyz_span, T, lb, ub);
}

extern "C" int fgoicp_bounds_multi(fgoicp_ctx* c, const float* rot_xyz_span, int Rn, int fix_rot,
                                   const float* t_xyz_span, int T, float* lb, float* ub)
{
    return bounds_host(c, rot_xyz_span, nullptr, Rn, fix_rot, t_xyz_span, T, lb, ub);
}

extern "C" int fgoicp_bounds_multi_dev(fgoicp_ctx* c, const float* d_rot_xyz_span, int Rn, int fix_rot,
                                       const float* d_t_xyz_span, int T, float* d_lb, float* d_ub,
                                       float* d_best_ub)
{
    FG_ARG(c && d_rot_xyz_span && d_t_xyz_span && d_lb && d_ub, "NULL pointer");
    FG_ARG(Rn > 0 && T > 0, "Rn and T must be positive");
    FG_CUDA(cudaSetDevice(c->device));
    int S = pick_slices(c, Rn);
    BoundsLaunch b;
    b.d_rot = (const float4*)d_rot_xyz_span; b.d_Rmats = nullptr; b.Rn = Rn; b.fix_rot = fix_rot;
    b.d_tc = (const float4*)d_t_xyz_span; b.T = T; b.d_lb = d_lb; b.d_ub = d_ub; b.d_best_ub = d_best_ub;
    b.d_partial = nullptr; b.S = S;
    if (S > 1)
    {
        int rc = fg::ensure_scratch(c, sizeof(double) * 2 * (size_t)Rn * T * S);
        if (rc) return rc;
        b.d_partial = (double*)c->d_scratch;
    }
    return run_bounds(c, b);
}

extern "C" int fgoicp_lut_sample(fgoicp_ctx* c, const float* q_xyz, size_t n, int sampler, float* out_d2)
{
    FG_ARG(c && q_xyz && out_d2, "NULL pointer");
    if (n == 0) return FGOICP_OK;
    FG_ARG(sampler >= 0 && sampler <= 2, "unknown sampler");
    if (sampler == FGOICP_SAMPLER_PACKED && !c->d_packed) { fg::set_error("packed grid was not built"); return FGOICP_ERR_STATE; }
    if (sampler == FGOICP_SAMPLER_TEX && !c->lut.tex) { fg::set_error("texture was not built"); return FGOICP_ERR_STATE; }
    FG_CUDA(cudaSetDevice(c->device));
    size_t b_in = align256(sizeof(float) * 3 * n);
    int rc = fg::ensure_scratch(c, b_in + sizeof(float) * n);
    if (rc) return rc;
    char* dp = (char*)c->d_scratch;
    FG_CUDA(cudaMemcpyAsync(dp, q_xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    unsigned blocks = (unsigned)((n + 255) / 256);
    switch (sampler)
    {
    case FGOICP_SAMPLER_PACKED: k_lut_sample<FGOICP_SAMPLER_PACKED><<<blocks, 256, 0, c->stream>>>(c->lut, (const float*)dp, n, (float*)(dp + b_in)); break;
    case FGOICP_SAMPLER_TEX:    k_lut_sample<FGOICP_SAMPLER_TEX><<<blocks, 256, 0, c->stream>>>(c->lut, (const float*)dp, n, (float*)(dp + b_in)); break;
    default:                    k_lut_sample<FGOICP_SAMPLER_GRID><<<blocks, 256, 0, c->stream>>>(c->lut, (const float*)dp, n, (float*)(dp + b_in)); break;
    }
    FG_CUDA(cudaGetLastError());
    FG_CUDA(cudaMemcpyAsync(out_d2, dp + b_in, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FGOICP_OK;
}

extern "C" int fgoicp_rot_sin(fgoicp_ctx* c, const float* spans, int n, float* out)
{
    FG_ARG(c && spans && out && n > 0, "bad arguments");
    FG_CUDA(cudaSetDevice(c->device));
    int rc = fg::ensure_scratch(c, sizeof(float) * 2 * n);
    if (rc) return rc;
    float* d = (float*)c->d_scratch;
    FG_CUDA(cudaMemcpyAsync(d, spans, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    k_rot_sin<<<(n + 127) / 128, 128, 0, c->stream>>>(d, n, d + n);
    FG_CUDA(cudaGetLastError());
    FG_CUDA(cudaMemcpyAsync(out, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FGOICP_OK;
}
