// trim.cuh -- exact "sum of the K smallest" for the trimmed registration (EXTENSION: the reference parses
// `trim` and ignores it, src/utilities.hpp:94, fgoicp/fgoicp.hpp:73; north_star asks for the trimmed-residual sum).
//
// All residuals here are non-negative floats, so their IEEE bit patterns order like the values and an MSB-first
// radix select over the bits finds the K-th smallest value exactly: 4 passes of a 256-bin histogram in shared
// memory.  The trimmed sum is then  sum{ x : x < v_K } + (K - #{x < v_K}) * v_K,  accumulated in fp64 and rounded
// once -- independent of the order the block visits the points in, like every other sum of this library.
#pragma once
#include "common.cuh"

struct FgSelect
{
    unsigned int vk_bits;     // bit pattern of the K-th smallest value
    unsigned int take_eq;     // how many elements equal to v_K belong to the K smallest (>= 1)
};

// Block-wide radix select.  get(i) returns the bit pattern of element i (0 <= i < n); 1 <= K <= n.
// Every thread of the block must call it; s_hist is 256 unsigned ints of shared memory, s_state 2.
template <typename Get>
__device__ __forceinline__ FgSelect fg_block_select(Get get, int n, unsigned int K, unsigned int* s_hist, unsigned int* s_state)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    unsigned int prefix = 0, mask = 0, k = K;
    for (int pass = 0; pass < 4; ++pass)
    {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += nt) s_hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt)
        {
            unsigned int bits = get(i);
            if ((bits & mask) == prefix) atomicAdd(&s_hist[(bits >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0)
        {
            unsigned int cum = 0, b = 0;
            for (; b < 255; ++b)
            {
                unsigned int h = s_hist[b];
                if (cum + h >= k) break;
                cum += h;
            }
            s_state[0] = prefix | (b << shift);
            s_state[1] = k - cum;
        }
        __syncthreads();
        prefix = s_state[0]; k = s_state[1];
        mask |= 0xffu << shift;
        __syncthreads();
    }
    FgSelect r;
    r.vk_bits = prefix; r.take_eq = k;
    return r;
}

// deterministic block sum of one double (fixed tree: lanes by butterfly, warps in index order); result in every thread
__device__ __forceinline__ double fg_block_sum1(double v, double* s_w /*[32]*/)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = fg_warp_sum(v);
    __syncthreads();
    if (lane == 0) s_w[w] = v;
    __syncthreads();
    double a = 0.0;
    for (int q = 0; q < nw; ++q) a += s_w[q];
    return a;
}

// Sum of the K smallest of n non-negative floats reachable through get_bits(i); 1 <= K <= n.
template <typename Get>
__device__ __forceinline__ double fg_block_trimmed_sum(Get get_bits, int n, unsigned int K, unsigned int* s_hist,
                                                       unsigned int* s_state, double* s_w)
{
    FgSelect sel = fg_block_select(get_bits, n, K, s_hist, s_state);
    double part = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        unsigned int bits = get_bits(i);
        if (bits < sel.vk_bits) part += (double)__uint_as_float(bits);
    }
    double below = fg_block_sum1(part, s_w);
    return below + (double)sel.take_eq * (double)__uint_as_float(sel.vk_bits);
}
