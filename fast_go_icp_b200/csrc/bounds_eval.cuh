// bounds_eval.cuh -- the (translation-cube chunk x data-point range) evaluation core shared by the
// flat bound kernel (bounds.cu) and the GPU-resident inner branch-and-bound (bnb.cu).
#pragma once
#include "common.cuh"

#define BD_CHUNK   32        // translation cubes staged per pass
#ifndef BD_CPW
#define BD_CPW     4         // cubes per warp; 8 was measured slower (inner searches 84 -> 94 ms); must satisfy BD_CHUNK <= BD_CPW * warps
#endif

// Evaluates nch (<= BD_CHUNK) translation cubes staged in s_tc against data points [p0, p1) with
// NWARPS warps laid out as Wc cube-groups x Wp point-slices.  On return (after the caller's
// __syncthreads) s_part[w][k] holds warp w's fp64 partial sums {ub, lb} for its k-th cube; use
// fg_eval_gather to fold the point-slices in a fixed order.
template <int SAMPLER, int NWARPS>
__device__ __forceinline__ void fg_eval_chunk(const LutDev& L, const float4* __restrict__ data, int p0, int p1,
                                              const float* sR, float sin_half, bool fix_rot,
                                              const float4* s_tc, int nch, double (*s_part)[BD_CPW][2])
{
    static_assert(BD_CHUNK <= BD_CPW * NWARPS, "every cube group of a chunk needs its own warp slot in s_part");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int groups = (nch + BD_CPW - 1) / BD_CPW;
    int Wc = 1; while (Wc < groups) Wc <<= 1;
    if (Wc > NWARPS) Wc = NWARPS;
    const int Wp = NWARPS / Wc;
    const int cg = w % Wc, ps = w / Wc;

    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = sR[k];

    // a warp may own several cube groups when there are more groups than Wc (never for nch <= 4*NWARPS)
    for (int g = cg; g < groups; g += Wc)
    {
        const int cbase = g * BD_CPW;
        float tx[BD_CPW], ty[BD_CPW], tz[BD_CPW], tsp[BD_CPW];
#pragma unroll
        for (int k = 0; k < BD_CPW; ++k)
        {
            float4 t = s_tc[min(cbase + k, nch - 1)];
            tx[k] = t.x; ty[k] = t.y; tz[k] = t.z; tsp[k] = t.w;
        }
        double acc_ub[BD_CPW], acc_lb[BD_CPW];
#pragma unroll
        for (int k = 0; k < BD_CPW; ++k) { acc_ub[k] = 0.0; acc_lb[k] = 0.0; }

        const int i0 = p0 + ps * 32 + lane, stride = Wp * 32;
        float4 p_next = i0 < p1 ? __ldg(&data[i0]) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = i0; i < p1; i += stride)
        {
            float4 p = p_next;
            if (i + stride < p1) p_next = __ldg(&data[i + stride]);     // prefetch: hides the point load behind the gathers
            float3 rp = fg_rotate(R, p.x, p.y, p.z);
            // rot_uncertain_radius = 2 * |p|^2 * sin(half_angle)  (SASS: FADD r,r ; FMUL)
            float rot_r = __fmul_rn(__fadd_rn(p.w, p.w), sin_half);
            // issue all BD_CPW gathers before blending any of them: one DRAM round trip instead of four
            SampleReq req[BD_CPW];
#pragma unroll
            for (int k = 0; k < BD_CPW; ++k)
                fg_sample_issue<SAMPLER>(L, __fadd_rn(rp.x, tx[k]), __fadd_rn(rp.y, ty[k]), __fadd_rn(rp.z, tz[k]), req[k]);
#pragma unroll
            for (int k = 0; k < BD_CPW; ++k)
            {
                float u, l;
                fg_bound_terms(fg_sample_finish<SAMPLER>(req[k]), rot_r, fix_rot, tsp[k], u, l);
                acc_ub[k] += (double)u;
                acc_lb[k] += (double)l;
            }
        }
#pragma unroll
        for (int k = 0; k < BD_CPW; ++k)
        {
            double su = fg_warp_sum(acc_ub[k]);
            double sl = fg_warp_sum(acc_lb[k]);
            if (lane == 0) { s_part[w][k][0] = su; s_part[w][k][1] = sl; }
        }
    }
}

// fold the point-slices of cube c (0 <= c < nch) in fixed order; valid after __syncthreads
template <int NWARPS>
__device__ __forceinline__ void fg_eval_gather(double (*s_part)[BD_CPW][2], int nch, int c, double& su, double& sl)
{
    int groups = (nch + BD_CPW - 1) / BD_CPW;
    int Wc = 1; while (Wc < groups) Wc <<= 1;
    if (Wc > NWARPS) Wc = NWARPS;
    const int Wp = NWARPS / Wc;
    int g = c / BD_CPW, k = c % BD_CPW;
    su = 0.0; sl = 0.0;
    for (int q = 0; q < Wp; ++q) { su += s_part[q * Wc + g][k][0]; sl += s_part[q * Wc + g][k][1]; }
}
