// bnb.cu -- the inner R^3 translation branch-and-bound as a GPU-resident best-first search, and
// the per-level driver of the outer SO(3) search.
//
// Replaces FastGoICP::branch_and_bound_R3 of the reference (fgoicp/fgoicp.cpp:102-174), where a
// host std::priority_queue pops <= 32 translation cubes, crosses to the device for their bounds
// (2 cudaMalloc + <= 32 launches + <= 64 blocking reductions + 2 cudaFree per batch) and comes back
// to prune and spawn children.  Here one thread block owns one rotation cube's whole search:
//   - the open list ("pool") lives in shared memory as 64-bit keys
//         [ lb bits : 32 | level : 3 | insertion seq : 13 | iz : 4 | iy : 4 | ix : 4 ]
//     so a plain unsigned sort yields the reference's heap order (smaller lb first, ties -> larger
//     span first; common.hpp:120-127) made total by the insertion sequence number;
//   - every iteration: bitonic sort + compaction of the pool, the reference's stop test, pop of
//     <= 32 cubes, bound evaluation with the same fused evaluator as bounds.cu, update of
//     best_ub / best_error / best_t, pruning, and spawning of 8 children per surviving cube;
//   - nothing returns to the host until the search is over; many rotation cubes run concurrently,
//     which is where the parallelism of a whole outer level comes from;
//   - a rotation cube is owned by a THREAD-BLOCK CLUSTER of 1..8 blocks: every block keeps its own copy of
//     the (tiny) search state and executes the same control flow, but evaluates only its slice of the data
//     points; the per-cube fp64 partial sums are exchanged through distributed shared memory (one
//     cluster.sync per iteration) and folded in rank order, so all blocks see identical bounds.  Few cubes
//     (outer levels 1-2: 8 and <= 64 cubes) get big clusters so the GPU is not idle; many cubes get small
//     clusters, which also evens out the tail (searches differ 100x in length).
// Cube centres are dyadic ( -1 + (2i+1) 2^-level ), so integer coordinates reproduce the reference's
// float arithmetic (t - span/2 + bit*span, fgoicp.cpp:159-163) exactly.
#include "common.cuh"
#include "bounds_eval.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

#define BNB_POOL       8192            // >= 1 + 8 + 64 + 512 + 4096 = 4681 nodes ever pushed
#define BNB_BATCH_MAX  32              // fgoicp.cpp:122
#define BNB_MIN_TSPAN  0.1f            // fgoicp.cpp:155
#define BNB_KEY_MAX    0xffffffffffffffffull

struct BnbOut
{
    float best_ub;
    float best_t[3];
    unsigned long long evals;
    unsigned int batches;
    unsigned int pushed;
};

__device__ __forceinline__ unsigned long long bnb_key(float lb, unsigned level, unsigned seq, unsigned ix, unsigned iy, unsigned iz)
{
    unsigned lo = (level << 25) | (seq << 12) | (iz << 8) | (iy << 4) | ix;
    return ((unsigned long long)__float_as_uint(lb) << 32) | lo;
}

__device__ __forceinline__ float4 bnb_key_cube_m(unsigned long long key)
{
    unsigned lo = (unsigned)key;
    unsigned level = (lo >> 25) & 7u, ix = lo & 15u, iy = (lo >> 4) & 15u, iz = (lo >> 8) & 15u;
    float span = __uint_as_float((127u - level) << 23);                  // 2^-level
    float4 t;
    t.x = __fmaf_rn((float)(2 * ix + 1), span, -1.0f);                   // exact: dyadic
    t.y = __fmaf_rn((float)(2 * iy + 1), span, -1.0f);
    t.z = __fmaf_rn((float)(2 * iz + 1), span, -1.0f);
    t.w = span;
    return t;
}

#ifndef BNB_MIN_BLOCKS
#define BNB_MIN_BLOCKS 2      // 2 blocks of 8 warps per SM (<= 128 registers); 1 lets ptxas take 139 and halves the occupancy (inner searches 84 -> 104 ms), 3 spills (108 ms)
#endif
template <int SAMPLER, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, BNB_MIN_BLOCKS)
k_bnb_r3(LutDev L, const float4* __restrict__ data, int ns, const float4* __restrict__ rot, int fix_rot,
         float best_sse, float sse_threshold, int batch_max, BnbOut* __restrict__ out)
{
    extern __shared__ unsigned long long pool[];          // BNB_POOL keys
    __shared__ float sR[9];
    __shared__ float s_sin;
    __shared__ float4 s_tc[BNB_BATCH_MAX];
    __shared__ unsigned long long s_bkey[BNB_BATCH_MAX];
    __shared__ double s_part[NWARPS][BD_CPW][2];
    __shared__ double s_cpart[2][BNB_BATCH_MAX][2];        // this block's partial sums, double-buffered for DSMEM readers
    __shared__ float s_lb[BNB_BATCH_MAX], s_ub[BNB_BATCH_MAX];
    __shared__ int s_n, s_m, s_nb, s_stop;
    __shared__ float s_best_error, s_best_ub, s_best_t[3];
    __shared__ unsigned int s_seq, s_batches;
    __shared__ unsigned long long s_evals;

    const int NT = NWARPS * 32;
    const int tid = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int r = blockIdx.x / csize;
    // this block's slice of the data points
    const int per = (ns + csize - 1) / csize;
    const int p0 = min(ns, crank * per), p1 = min(ns, p0 + per);
    int parity = 0;

    if (tid == 0)
    {
        float4 rc = rot[r];
        float Rm[9];
        fg_rotation_matrix(rc.x, rc.y, rc.z, Rm);
        for (int k = 0; k < 9; ++k) sR[k] = Rm[k];
        s_sin = fix_rot ? 0.0f : fg_rot_sin(rc.w);
        pool[0] = bnb_key(0.0f, 0, 0, 0, 0, 0);          // root: t = 0, span = 1, lb = 0 (fgoicp.cpp:113)
        s_n = 1; s_seq = 1;
        s_best_error = best_sse;                          // fgoicp.cpp:104
        s_best_ub = FG_INF;                               // fgoicp.cpp:106
        s_best_t[0] = s_best_t[1] = s_best_t[2] = 0.0f;   // fgoicp.cpp:105
        s_evals = 0; s_batches = 0; s_stop = 0;
    }
    __syncthreads();

    while (true)
    {
        // ---- (a) sort the pool ascending; dead entries (KEY_MAX) sink to the end
        const int n = s_n;
        int P = 32; while (P < n) P <<= 1;
        for (int i = n + tid; i < P; i += NT) pool[i] = BNB_KEY_MAX;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1)
            {
                for (int i = tid; i < P; i += NT)
                {
                    int ixj = i ^ j;
                    if (ixj > i)
                    {
                        unsigned long long a = pool[i], b = pool[ixj];
                        bool up = ((i & k) == 0);
                        if ((a > b) == up) { pool[i] = b; pool[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        // ---- (b) live count m = index of the first dead entry
        if (tid == 0) s_m = (pool[P - 1] != BNB_KEY_MAX) ? P : -1;
        __syncthreads();
        for (int i = tid; i < P; i += NT)
        {
            bool live = pool[i] != BNB_KEY_MAX;
            bool prev_live = (i == 0) ? true : (pool[i - 1] != BNB_KEY_MAX);
            if (!live && prev_live) s_m = i;              // exactly one thread at most
        }
        __syncthreads();
        // ---- (c) stop test and (d) batch pop
        if (tid == 0)
        {
            int m = s_m;
            int stop = 0, nb = 0;
            if (m <= 0) stop = 1;
            else
            {
                float top_lb = __uint_as_float((unsigned)(pool[0] >> 32));
                if (__fsub_rn(s_best_error, top_lb) < sse_threshold) stop = 1;      // fgoicp.cpp:120
                else nb = min(m, batch_max);
            }
            s_stop = stop; s_nb = nb;
        }
        __syncthreads();
        if (s_stop) break;
        const int m = s_m, nb = s_nb;
        if (tid < nb)
        {
            unsigned long long key = pool[tid];
            s_bkey[tid] = key;
            unsigned lo = (unsigned)key;
            unsigned level = (lo >> 25) & 7u, ix = lo & 15u, iy = (lo >> 4) & 15u, iz = (lo >> 8) & 15u;
            float span = __uint_as_float((127u - level) << 23);                  // 2^-level
            float4 t;
            t.x = __fmaf_rn((float)(2 * ix + 1), span, -1.0f);                   // exact: dyadic
            t.y = __fmaf_rn((float)(2 * iy + 1), span, -1.0f);
            t.z = __fmaf_rn((float)(2 * iz + 1), span, -1.0f);
            t.w = span;
            s_tc[tid] = t;
            pool[tid] = BNB_KEY_MAX;                                             // popped
        }
        __syncthreads();

        // ---- (e) bounds of the batch (registration.cu:88-152 fused)
        fg_eval_chunk<SAMPLER, NWARPS>(L, data, p0, p1, sR, s_sin, fix_rot != 0, s_tc, nb, s_part);
        __syncthreads();
        if (tid < nb)
        {
            double su, sl;
            fg_eval_gather<NWARPS>(s_part, nb, tid, su, sl);
            s_cpart[parity][tid][0] = su; s_cpart[parity][tid][1] = sl;
        }
        cluster.sync();                                   // every block's partials are visible cluster-wide
        if (tid < nb)
        {
            double su = 0.0, sl = 0.0;
            for (int k = 0; k < csize; ++k)               // rank order: identical totals in every block
            {
                const double* rp = cluster.map_shared_rank(&s_cpart[parity][tid][0], k);
                su += rp[0]; sl += rp[1];
            }
            s_ub[tid] = (float)su; s_lb[tid] = (float)sl;
        }
        parity ^= 1;
        __syncthreads();

        // ---- (f) best of the batch (fgoicp.cpp:139-145): first minimum in pop order
        if (tid == 0)
        {
            int idx_min = 0;
            for (int i = 1; i < nb; ++i) if (s_ub[i] < s_ub[idx_min]) idx_min = i;
            float u = s_ub[idx_min];
            s_best_ub = s_best_ub < u ? s_best_ub : u;
            if (u < s_best_error)
            {
                s_best_error = u;
                s_best_t[0] = s_tc[idx_min].x; s_best_t[1] = s_tc[idx_min].y; s_best_t[2] = s_tc[idx_min].z;
            }
            s_evals += (unsigned long long)nb * (unsigned long long)ns;           // fgoicp.cpp:132 (x ns)
            s_batches += 1;
        }
        __syncthreads();

        // ---- (g) spawn children of surviving cubes (fgoicp.cpp:148-169), in pop order
        const float best_error = s_best_error;
        if (tid < 32)
        {
            bool spawn = false;
            if (tid < nb)
            {
                float lbv = s_lb[tid];
                float span = s_tc[tid].w;
                spawn = !(lbv >= best_error) && !(span < BNB_MIN_TSPAN);
            }
            unsigned mask = __ballot_sync(0xffffffffu, spawn);
            int before = __popc(mask & ((1u << tid) - 1u));
            int total = __popc(mask);
            if (spawn)
            {
                unsigned lo = (unsigned)s_bkey[tid];
                unsigned level = (lo >> 25) & 7u, ix = lo & 15u, iy = (lo >> 4) & 15u, iz = (lo >> 8) & 15u;
                unsigned seq0 = s_seq + 8u * before;
                int base = m + 8 * before;
                float lbv = s_lb[tid];
#pragma unroll
                for (unsigned k = 0; k < 8; ++k)
                    pool[base + k] = bnb_key(lbv, level + 1, seq0 + k, 2 * ix + (k & 1u), 2 * iy + ((k >> 1) & 1u), 2 * iz + ((k >> 2) & 1u));
            }
            __syncwarp();
            if (tid == 0) { s_n = m + 8 * total; s_seq += 8u * total; }
        }
        __syncthreads();

        // ---- (h) prune everything that can no longer be popped (lb >= best_error, fgoicp.cpp:126)
        const unsigned be_bits = __float_as_uint(best_error);
        const int n_new = s_n;
        for (int i = tid; i < n_new; i += NT)
        {
            unsigned long long key = pool[i];
            if (key != BNB_KEY_MAX && (unsigned)(key >> 32) >= be_bits) pool[i] = BNB_KEY_MAX;
        }
        __syncthreads();
    }

    cluster.sync();                                       // nobody exits while a peer may still read its partials
    if (tid == 0 && crank == 0)
    {
        BnbOut o;
        o.best_ub = s_best_ub;
        o.best_t[0] = s_best_t[0]; o.best_t[1] = s_best_t[1]; o.best_t[2] = s_best_t[2];
        o.evals = s_evals; o.batches = s_batches; o.pushed = s_seq;
        out[r] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Same search with the open list KEPT SORTED instead of re-sorted: k_bnb_r3 above runs a full bitonic sort of the pool
// (up to 4096 keys: 78 compare-exchange sweeps, ~20 us) every iteration, which is a quarter of an iteration when a
// level has few rotation cubes and the search is latency-bound (outer levels 1-2; every wave of a multi-GPU run).
// Here the pool is a sorted run in one of two shared-memory buffers:
//   pop        = advance the head by <= 32 keys (they are the smallest);
//   prune      = lb >= best_error is a SUFFIX of a run sorted by lb: one binary search, no sweep;
//   children   = <= 256 new keys, sorted on their own (bitonic over 256) and MERGED with the run into the other buffer:
//                every key finds its final position by one binary search in the other list (keys are unique: the
//                insertion sequence number is part of the key), ~3 us for 4000 keys.
// The pop order is the ascending key order in both kernels, so every decision, bound and evaluation count is the same.
// ---------------------------------------------------------------------------------------------
#define BNBM_POOL  4736                 // >= 4681 nodes ever pushed
#define BNBM_NEW   256                  // children of one batch (8 x 32)

template <int SAMPLER, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
k_bnb_r3m(LutDev L, const float4* __restrict__ data, int ns, const float4* __restrict__ rot, int fix_rot,
          float best_sse, float sse_threshold, int batch_max, BnbOut* __restrict__ out)
{
    extern __shared__ unsigned long long poolm[];         // 2 x BNBM_POOL + 2 x BNBM_NEW keys
    __shared__ float sR[9];
    __shared__ float s_sin;
    __shared__ float4 s_tc[BNB_BATCH_MAX];
    __shared__ unsigned long long s_bkey[BNB_BATCH_MAX];
    __shared__ double s_part[NWARPS][BD_CPW][2];
    __shared__ double s_cpart[2][BNB_BATCH_MAX][2];        // this block's partial sums, double-buffered for DSMEM readers
    __shared__ float s_lb[BNB_BATCH_MAX], s_ub[BNB_BATCH_MAX];
    __shared__ int s_m, s_nb, s_stop, s_nc;
    __shared__ float s_best_error, s_best_ub, s_best_t[3];
    __shared__ unsigned int s_seq, s_batches;
    __shared__ unsigned long long s_evals;

    const int NT = NWARPS * 32;
    const int tid = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int r = blockIdx.x / csize;
    const int per = (ns + csize - 1) / csize;
    const int p0 = min(ns, crank * per), p1 = min(ns, p0 + per);
    int parity = 0;
    unsigned long long* A = poolm;                        // current sorted run: A[head .. head + m)
    unsigned long long* B = poolm + BNBM_POOL;            // merge target
    unsigned long long* const C0 = poolm + 2 * BNBM_POOL; // children of the batch, as spawned; sorted copy behind it
    unsigned long long* C = C0;
    int head = 0;

    if (tid == 0)
    {
        float4 rc = rot[r];
        float Rm[9];
        fg_rotation_matrix(rc.x, rc.y, rc.z, Rm);
        for (int k = 0; k < 9; ++k) sR[k] = Rm[k];
        s_sin = fix_rot ? 0.0f : fg_rot_sin(rc.w);
        A[0] = bnb_key(0.0f, 0, 0, 0, 0, 0);             // root: t = 0, span = 1, lb = 0 (fgoicp.cpp:113)
        s_m = 1; s_seq = 1;
        s_best_error = best_sse;                          // fgoicp.cpp:104
        s_best_ub = FG_INF;                               // fgoicp.cpp:106
        s_best_t[0] = s_best_t[1] = s_best_t[2] = 0.0f;   // fgoicp.cpp:105
        s_evals = 0; s_batches = 0; s_stop = 0;
    }
    __syncthreads();

    while (true)
    {
        // ---- stop test and batch pop: the run is sorted, its head is the smallest lower bound
        if (tid == 0)
        {
            const int m = s_m;
            int stop = 0, nb = 0;
            if (m <= 0) stop = 1;
            else
            {
                float top_lb = __uint_as_float((unsigned)(A[head] >> 32));
                if (__fsub_rn(s_best_error, top_lb) < sse_threshold) stop = 1;      // fgoicp.cpp:120
                else nb = min(m, batch_max);
            }
            s_stop = stop; s_nb = nb;
        }
        __syncthreads();
        if (s_stop) break;
        const int nb = s_nb;
        if (tid < nb)
        {
            unsigned long long key = A[head + tid];
            s_bkey[tid] = key;
            s_tc[tid] = bnb_key_cube_m(key);
        }
        __syncthreads();
        head += nb;                                       // popped (uniform: every thread tracks head)
        int m = s_m - nb;

        // ---- bounds of the batch (registration.cu:88-152 fused)
        fg_eval_chunk<SAMPLER, NWARPS>(L, data, p0, p1, sR, s_sin, fix_rot != 0, s_tc, nb, s_part);
        __syncthreads();
        if (tid < nb)
        {
            double su, sl;
            fg_eval_gather<NWARPS>(s_part, nb, tid, su, sl);
            s_cpart[parity][tid][0] = su; s_cpart[parity][tid][1] = sl;
        }
        cluster.sync();                                   // every block's partials are visible cluster-wide
        if (tid < nb)
        {
            double su = 0.0, sl = 0.0;
            for (int k = 0; k < csize; ++k)               // rank order: identical totals in every block
            {
                const double* rp = cluster.map_shared_rank(&s_cpart[parity][tid][0], k);
                su += rp[0]; sl += rp[1];
            }
            s_ub[tid] = (float)su; s_lb[tid] = (float)sl;
        }
        parity ^= 1;
        __syncthreads();

        // ---- best of the batch (fgoicp.cpp:139-145): first minimum in pop order
        if (tid == 0)
        {
            int idx_min = 0;
            for (int i = 1; i < nb; ++i) if (s_ub[i] < s_ub[idx_min]) idx_min = i;
            float u = s_ub[idx_min];
            s_best_ub = s_best_ub < u ? s_best_ub : u;
            if (u < s_best_error)
            {
                s_best_error = u;
                s_best_t[0] = s_tc[idx_min].x; s_best_t[1] = s_tc[idx_min].y; s_best_t[2] = s_tc[idx_min].z;
            }
            s_evals += (unsigned long long)nb * (unsigned long long)ns;           // fgoicp.cpp:132 (x ns)
            s_batches += 1;
        }
        __syncthreads();

        // ---- children of surviving cubes (fgoicp.cpp:148-169), in pop order, into C
        C = C0;
        const float best_error = s_best_error;
        const unsigned long long cut_key = (unsigned long long)__float_as_uint(best_error) << 32;   // keys >= this are dead
        if (tid < 32)
        {
            bool spawn = false;
            if (tid < nb)
            {
                float lbv = s_lb[tid];
                float span = s_tc[tid].w;
                spawn = !(lbv >= best_error) && !(span < BNB_MIN_TSPAN);
            }
            unsigned mask = __ballot_sync(0xffffffffu, spawn);
            int before = __popc(mask & ((1u << tid) - 1u));
            int total = __popc(mask);
            if (spawn)
            {
                unsigned lo = (unsigned)s_bkey[tid];
                unsigned level = (lo >> 25) & 7u, ix = lo & 15u, iy = (lo >> 4) & 15u, iz = (lo >> 8) & 15u;
                unsigned seq0 = s_seq + 8u * before;
                float lbv = s_lb[tid];
#pragma unroll
                for (unsigned k = 0; k < 8; ++k)
                    C[8 * before + k] = bnb_key(lbv, level + 1, seq0 + k, 2 * ix + (k & 1u), 2 * iy + ((k >> 1) & 1u), 2 * iz + ((k >> 2) & 1u));
            }
            __syncwarp();
            if (tid == 0) { s_nc = 8 * total; s_seq += 8u * total; }
        }
        __syncthreads();
        const int nc = s_nc;

        // ---- prune the run: everything with lb >= best_error is a suffix (fgoicp.cpp:126)
        {
            int lo = 0, hi = m;                           // first index whose key >= cut_key
            while (lo < hi) { int mid = (lo + hi) >> 1; if (A[head + mid] < cut_key) lo = mid + 1; else hi = mid; }
            m = lo;
        }
        if (nc > 0)
        {
            // sort the children by rank: keys are unique, so a key's position is the number of smaller keys -- nc broadcast
            // reads per thread and ONE barrier instead of the 36 compare-exchange sweeps (each with its own barrier) of a
            // bitonic network over 256 keys, which were ~10 us of every iteration of the latency-bound coarse levels
            unsigned long long* Cs = C + BNBM_NEW;            // sorted children
            for (int i = tid; i < nc; i += NT)
            {
                const unsigned long long key = C[i];
                int rank = 0;
                for (int j = 0; j < nc; ++j) rank += C[j] < key;
                Cs[rank] = key;
            }
            __syncthreads();
            C = Cs;
            // children at or above the cut are dead too (cannot happen for a freshly spawned child -- its lb is its
            // parent's, which was below best_error -- but the rule is applied uniformly)
            int ncl;
            {
                int lo = 0, hi = nc;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (C[mid] < cut_key) lo = mid + 1; else hi = mid; }
                ncl = lo;
            }
            // merge A[head .. head + m) and C[0 .. ncl) into B by rank
            for (int i = tid; i < m; i += NT)
            {
                const unsigned long long a = A[head + i];
                int lo = 0, hi = ncl;                     // children smaller than a
                while (lo < hi) { int mid = (lo + hi) >> 1; if (C[mid] < a) lo = mid + 1; else hi = mid; }
                B[i + lo] = a;
            }
            for (int i = tid; i < ncl; i += NT)
            {
                const unsigned long long c = C[i];
                int lo = 0, hi = m;                       // run keys smaller than c
                while (lo < hi) { int mid = (lo + hi) >> 1; if (A[head + mid] < c) lo = mid + 1; else hi = mid; }
                B[i + lo] = c;
            }
            __syncthreads();
            unsigned long long* T = A; A = B; B = T;
            head = 0;
            m += ncl;
        }
        if (tid == 0) s_m = m;
        __syncthreads();
    }

    cluster.sync();                                       // nobody exits while a peer may still read its partials
    if (tid == 0 && crank == 0)
    {
        BnbOut o;
        o.best_ub = s_best_ub;
        o.best_t[0] = s_best_t[0]; o.best_t[1] = s_best_t[1]; o.best_t[2] = s_best_t[2];
        o.evals = s_evals; o.batches = s_batches; o.pushed = s_seq;
        out[r] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Round-synchronous form of the same search, for levels with many rotation cubes.
//
// k_bnb_r3 above lets every rotation cube run ahead on its own, which makes the cube-bound gathers of
// the ~450 resident blocks land all over the 1 GB grid (HBM random-access bound, ~6e10 evals/s).  Here all
// searches of the level advance one iteration per ROUND:
//   k_bnbr_round  (one block per rotation cube): fold the bounds of the previous batch into the search
//                 (best update, child spawning, pruning), sort, stop test, pop the next <= 32 cubes;
//   bounds        of ALL popped cubes of the level in one launch of the z-phase-ordered kernel
//                 (bounds_phased.cu, 2.4x the throughput of unordered gathers), or of the plain kernel
//                 when only a few searches are still running.
// The per-cube search logic, its order of operations and the bound values are exactly those of k_bnb_r3,
// so every rotation cube takes the same decisions and spends the same number of evaluations (tested).
// The open lists live in HBM between rounds (<= 4736 keys = 37 KB per rotation cube).
// ---------------------------------------------------------------------------------------------
#define BNBR_POOL 4736                  // >= 4681 nodes ever pushed; multiple of 32
#define BNBR_THREADS 256

struct BnbrMeta
{
    int n;                              // live keys stored for the next round
    unsigned int seq;
    float best_error, best_ub, best_t[3];
    int nb;                             // cubes popped this round (bounds pending)
    int done;
    unsigned int batches;
    unsigned long long evals;
};

__global__ void k_bnbr_init(BnbrMeta* __restrict__ meta, unsigned long long* __restrict__ pools, int Rn, float best_sse,
                            int* __restrict__ counts, float4* __restrict__ tcubes)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Rn) return;
    BnbrMeta m;
    m.n = 1; m.seq = 1;
    m.best_error = best_sse;                                  // fgoicp.cpp:104
    m.best_ub = FG_INF;                                       // fgoicp.cpp:106
    m.best_t[0] = m.best_t[1] = m.best_t[2] = 0.0f;           // fgoicp.cpp:105
    m.nb = 0; m.done = 0; m.batches = 0; m.evals = 0;
    meta[r] = m;
    pools[(size_t)r * BNBR_POOL] = bnb_key(0.0f, 0, 0, 0, 0, 0);   // root: t = 0, span = 1, lb = 0 (fgoicp.cpp:113)
    counts[r] = 0;
    for (int i = 0; i < BNB_BATCH_MAX; ++i) tcubes[(size_t)r * BNB_BATCH_MAX + i] = make_float4(0.f, 0.f, 0.f, -1.0f);
}

__device__ __forceinline__ float4 bnb_key_cube(unsigned long long key)
{
    unsigned lo = (unsigned)key;
    unsigned level = (lo >> 25) & 7u, ix = lo & 15u, iy = (lo >> 4) & 15u, iz = (lo >> 8) & 15u;
    float span = __uint_as_float((127u - level) << 23);                  // 2^-level
    float4 t;
    t.x = __fmaf_rn((float)(2 * ix + 1), span, -1.0f);                   // exact: dyadic
    t.y = __fmaf_rn((float)(2 * iy + 1), span, -1.0f);
    t.z = __fmaf_rn((float)(2 * iz + 1), span, -1.0f);
    t.w = span;
    return t;
}

__global__ void __launch_bounds__(BNBR_THREADS)
k_bnbr_round(BnbrMeta* __restrict__ meta, unsigned long long* __restrict__ pools, unsigned long long* __restrict__ bkeys,
             float4* __restrict__ tcubes, int* __restrict__ counts, const float* __restrict__ lb, const float* __restrict__ ub,
             float sse_threshold, int ns, unsigned int* __restrict__ ctl /*[2]: active searches, popped cubes*/)
{
    extern __shared__ unsigned long long pool[];          // BNB_POOL keys
    __shared__ BnbrMeta M;
    __shared__ int s_m, s_stop, s_nb;
    const int NT = BNBR_THREADS;
    const int tid = threadIdx.x, r = blockIdx.x;
    if (tid == 0) M = meta[r];
    __syncthreads();
    if (M.done) return;
    unsigned long long* gpool = pools + (size_t)r * BNBR_POOL;
    unsigned long long* gkeys = bkeys + (size_t)r * BNB_BATCH_MAX;
    float4* gtc = tcubes + (size_t)r * BNB_BATCH_MAX;
    const int n0 = M.n, nb_prev = M.nb;
    for (int i = tid; i < n0; i += NT) pool[i] = gpool[i];
    __syncthreads();

    if (nb_prev > 0)
    {
        const float* rub = ub + (size_t)r * BNB_BATCH_MAX;
        const float* rlb = lb + (size_t)r * BNB_BATCH_MAX;
        // ---- (f) best of the batch (fgoicp.cpp:139-145): first minimum in pop order
        if (tid == 0)
        {
            int idx_min = 0;
            for (int i = 1; i < nb_prev; ++i) if (rub[i] < rub[idx_min]) idx_min = i;
            float u = rub[idx_min];
            M.best_ub = M.best_ub < u ? M.best_ub : u;
            if (u < M.best_error)
            {
                float4 t = bnb_key_cube(gkeys[idx_min]);
                M.best_error = u;
                M.best_t[0] = t.x; M.best_t[1] = t.y; M.best_t[2] = t.z;
            }
            M.evals += (unsigned long long)nb_prev * (unsigned long long)ns;      // fgoicp.cpp:132 (x ns)
            M.batches += 1;
        }
        __syncthreads();
        // ---- (g) spawn children of surviving cubes (fgoicp.cpp:148-169), in pop order
        const float best_error = M.best_error;
        if (tid < 32)
        {
            bool spawn = false;
            unsigned long long key = 0;
            float lbv = 0.f;
            if (tid < nb_prev)
            {
                key = gkeys[tid];
                lbv = rlb[tid];
                float span = bnb_key_cube(key).w;
                spawn = !(lbv >= best_error) && !(span < BNB_MIN_TSPAN);
            }
            unsigned mask = __ballot_sync(0xffffffffu, spawn);
            int before = __popc(mask & ((1u << tid) - 1u));
            int total = __popc(mask);
            if (spawn)
            {
                unsigned lo = (unsigned)key;
                unsigned level = (lo >> 25) & 7u, ix = lo & 15u, iy = (lo >> 4) & 15u, iz = (lo >> 8) & 15u;
                unsigned seq0 = M.seq + 8u * before;
                int base = n0 + 8 * before;
#pragma unroll
                for (unsigned k = 0; k < 8; ++k)
                    pool[base + k] = bnb_key(lbv, level + 1, seq0 + k, 2 * ix + (k & 1u), 2 * iy + ((k >> 1) & 1u), 2 * iz + ((k >> 2) & 1u));
            }
            __syncwarp();
            if (tid == 0) { M.n = n0 + 8 * total; M.seq += 8u * total; }
        }
        __syncthreads();
        // ---- (h) prune everything that can no longer be popped (lb >= best_error, fgoicp.cpp:126)
        const unsigned be_bits = __float_as_uint(best_error);
        for (int i = tid; i < M.n; i += NT)
        {
            unsigned long long key = pool[i];
            if (key != BNB_KEY_MAX && (unsigned)(key >> 32) >= be_bits) pool[i] = BNB_KEY_MAX;
        }
        __syncthreads();
    }

    // ---- (a) sort the pool ascending; dead entries (KEY_MAX) sink to the end
    const int n = M.n;
    int P = 32; while (P < n) P <<= 1;
    for (int i = n + tid; i < P; i += NT) pool[i] = BNB_KEY_MAX;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int i = tid; i < P; i += NT)
            {
                int ixj = i ^ j;
                if (ixj > i)
                {
                    unsigned long long a = pool[i], b = pool[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { pool[i] = b; pool[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // ---- (b) live count m = index of the first dead entry
    if (tid == 0) s_m = (pool[P - 1] != BNB_KEY_MAX) ? P : -1;
    __syncthreads();
    for (int i = tid; i < P; i += NT)
    {
        bool live = pool[i] != BNB_KEY_MAX;
        bool prev_live = (i == 0) ? true : (pool[i - 1] != BNB_KEY_MAX);
        if (!live && prev_live) s_m = i;              // exactly one thread at most
    }
    __syncthreads();
    // ---- (c) stop test and (d) batch pop
    if (tid == 0)
    {
        int m = s_m;
        int stop = 0, nb = 0;
        if (m <= 0) stop = 1;
        else
        {
            float top_lb = __uint_as_float((unsigned)(pool[0] >> 32));
            if (__fsub_rn(M.best_error, top_lb) < sse_threshold) stop = 1;      // fgoicp.cpp:120
            else nb = min(m, BNB_BATCH_MAX);
        }
        s_stop = stop; s_nb = nb;
    }
    __syncthreads();
    const int m = s_m, nb = s_nb;
    if (tid < BNB_BATCH_MAX)
    {
        float4 t = make_float4(0.f, 0.f, 0.f, -1.0f);      // unused slot
        if (tid < nb)
        {
            unsigned long long key = pool[tid];
            gkeys[tid] = key;
            t = bnb_key_cube(key);
        }
        if (tid < nb || tid < nb_prev) gtc[tid] = t;      // also clears the slots the previous batch used
    }
    // keep the survivors [nb, m) for the next round
    if (!s_stop) for (int i = nb + tid; i < m; i += NT) gpool[i - nb] = pool[i];
    if (tid == 0)
    {
        M.nb = nb; M.n = s_stop ? 0 : m - nb; M.done = s_stop;
        meta[r] = M;
        counts[r] = nb;
        if (!s_stop) { atomicAdd(&ctl[0], 1u); atomicAdd(&ctl[1], (unsigned)nb); }
    }
}

// ---------------------------------------------------------------------------------------------

template <int SAMPLER, int NWARPS>
static int launch_bnb_t(fgoicp_ctx* c, const float4* d_rot, int Rn, int csize, int fix_rot, float best_sse, float thr, BnbOut* d_out)
{
    size_t smem = sizeof(unsigned long long) * BNB_POOL;
    FG_CUDA(cudaFuncSetAttribute(k_bnb_r3<SAMPLER, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) FG_CUDA(cudaFuncSetAttribute(k_bnb_r3<SAMPLER, NWARPS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(Rn * csize));
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int ns = (int)c->ns, bm = BNB_BATCH_MAX;
    FG_CUDA(cudaLaunchKernelEx(&cfg, k_bnb_r3<SAMPLER, NWARPS>, c->lut, (const float4*)c->d_data, ns, d_rot, fix_rot,
                               best_sse, thr, bm, d_out));
    return FGOICP_OK;
}

template <int NWARPS>
static int launch_bnb_w(fgoicp_ctx* c, const float4* d_rot, int Rn, int csize, int fix_rot, float best_sse, float thr, BnbOut* d_out)
{
    switch (c->sampler)
    {
    case FGOICP_SAMPLER_PACKED: return launch_bnb_t<FGOICP_SAMPLER_PACKED, NWARPS>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    case FGOICP_SAMPLER_TEX:    return launch_bnb_t<FGOICP_SAMPLER_TEX, NWARPS>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    default:                    return launch_bnb_t<FGOICP_SAMPLER_GRID, NWARPS>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    }
}

// sorted-run kernel (k_bnb_r3m)
template <int SAMPLER, int NWARPS, int MINB>
static int launch_bnbm_t(fgoicp_ctx* c, const float4* d_rot, int Rn, int csize, int fix_rot, float best_sse, float thr, BnbOut* d_out)
{
    size_t smem = sizeof(unsigned long long) * (2 * BNBM_POOL + 2 * BNBM_NEW);
    FG_CUDA(cudaFuncSetAttribute(k_bnb_r3m<SAMPLER, NWARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) FG_CUDA(cudaFuncSetAttribute(k_bnb_r3m<SAMPLER, NWARPS, MINB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(Rn * csize));
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int ns = (int)c->ns, bm = BNB_BATCH_MAX;
    FG_CUDA(cudaLaunchKernelEx(&cfg, k_bnb_r3m<SAMPLER, NWARPS, MINB>, c->lut, (const float4*)c->d_data, ns, d_rot, fix_rot,
                               best_sse, thr, bm, d_out));
    return FGOICP_OK;
}

template <int NWARPS, int MINB>
static int launch_bnbm_w(fgoicp_ctx* c, const float4* d_rot, int Rn, int csize, int fix_rot, float best_sse, float thr, BnbOut* d_out)
{
    switch (c->sampler)
    {
    case FGOICP_SAMPLER_PACKED: return launch_bnbm_t<FGOICP_SAMPLER_PACKED, NWARPS, MINB>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    case FGOICP_SAMPLER_TEX:    return launch_bnbm_t<FGOICP_SAMPLER_TEX, NWARPS, MINB>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    default:                    return launch_bnbm_t<FGOICP_SAMPLER_GRID, NWARPS, MINB>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    }
}

// Cluster size: enough blocks for ~8 resident waves' worth of granularity, at most 8 (portable limit),
// at least 1; never so large that a block gets fewer than 256 points.  FGOICP_BNB_CLUSTER overrides.
// Kernel: the sorted-run kernel k_bnb_r3m (default; FGOICP_BNB_KERNEL=bitonic selects k_bnb_r3), with 16 warps per
// block when the launch cannot fill the GPU anyway (few rotation cubes: latency, not throughput, is what counts).
static int launch_bnb(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, float best_sse, float thr, BnbOut* d_out)
{
    int slots = 3 * c->sm_count;                       // 8-warp blocks resident at once (register-limited)
    int want = (8 * slots + Rn - 1) / Rn;
    int csize = 1;
    int cmax = 8;
    if (const char* e = getenv("FGOICP_BNB_CLUSTER_MAX")) cmax = atoi(e);
    while (csize < cmax && csize < want) csize <<= 1;
    while (csize > 1 && (int)c->ns / csize < 256) csize >>= 1;
    if (const char* e = getenv("FGOICP_BNB_CLUSTER")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) csize = v; }
    static const bool bitonic = getenv("FGOICP_BNB_KERNEL") && !strcmp(getenv("FGOICP_BNB_KERNEL"), "bitonic");
    if (bitonic) return launch_bnb_w<8>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    int warps = ((long long)Rn * csize <= c->sm_count && (int)c->ns / csize >= 512) ? 16 : 8;
    if (const char* e = getenv("FGOICP_BNB_WARPS")) { int v = atoi(e); if (v == 8 || v == 16) warps = v; }
    if (warps == 16) return launch_bnbm_w<16, 1>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
    return launch_bnbm_w<8, BNB_MIN_BLOCKS>(c, d_rot, Rn, csize, fix_rot, best_sse, thr, d_out);
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void k_bnbr_export(const BnbrMeta* __restrict__ meta, int Rn, BnbOut* __restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Rn) return;
    BnbrMeta m = meta[r];
    BnbOut o;
    o.best_ub = m.best_ub;
    o.best_t[0] = m.best_t[0]; o.best_t[1] = m.best_t[1]; o.best_t[2] = m.best_t[2];
    o.evals = m.evals; o.batches = m.batches; o.pushed = m.seq;
    out[r] = o;
}

int fg_phased_prepare(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, int T);
int fg_phased_eval(fgoicp_ctx* c, int Rn, const float4* d_tc, int T, float* d_lb, float* d_ub, unsigned int* d_best_bits);
int fg_phased_max_cubes(const fgoicp_ctx* c);
int fg_bounds_slices(const fgoicp_ctx* c, int active);
int fg_bounds_plain_counts(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, const float4* d_tc, int T,
                           const int* d_counts, int S, double* d_partial, float* d_lb, float* d_ub);

// Scratch of the round-synchronous schedule for a whole level of the outer search, allocated when the schedule is switched
// on (fgoicp_set_trim / fgoicp_set_bnb_mode) so that run() allocates nothing.  Best effort: bnb_rounds allocates what it
// needs if this did not happen.
int fg_rounds_prealloc(fgoicp_ctx* c)
{
    const int T = BNB_BATCH_MAX, Rn = 4096;
    const int S_max = std::max(1, std::min(16, (int)(c->ns / 256)));
    size_t want = sizeof(BnbrMeta) * (size_t)Rn + sizeof(unsigned long long) * (size_t)Rn * BNBR_POOL + sizeof(unsigned long long) * (size_t)Rn * T
                + sizeof(float4) * (size_t)Rn * T + sizeof(int) * (size_t)Rn + 2 * sizeof(float) * (size_t)Rn * T
                + sizeof(double) * 2 * (size_t)Rn * T * S_max + 16 * 256;
    if (want <= c->rounds_bytes) return FGOICP_OK;
    FG_CUDA(cudaSetDevice(c->device));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_rounds); c->d_rounds = nullptr; c->rounds_bytes = 0;
    if (cudaMalloc(&c->d_rounds, want) == cudaSuccess) c->rounds_bytes = want;
    else cudaGetLastError();
    size_t smem = sizeof(unsigned long long) * BNB_POOL;
    FG_CUDA(cudaFuncSetAttribute(k_bnbr_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return fg::ensure_pinned(c, 4096);
}

// All Rn searches of a level, one iteration per round (see k_bnbr_round).  Results land in d_out exactly as
// launch_bnb leaves them.
static int bnb_rounds(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, float best_sse, float thr, BnbOut* d_out)
{
    const int T = BNB_BATCH_MAX;
    const int S_max = std::max(1, std::min(16, (int)(c->ns / 256)));
    size_t b_meta = align256(sizeof(BnbrMeta) * (size_t)Rn);
    size_t b_pool = align256(sizeof(unsigned long long) * (size_t)Rn * BNBR_POOL);
    size_t b_keys = align256(sizeof(unsigned long long) * (size_t)Rn * T);
    size_t b_tc = align256(sizeof(float4) * (size_t)Rn * T);
    size_t b_cnt = align256(sizeof(int) * (size_t)Rn);
    size_t b_out = align256(sizeof(float) * (size_t)Rn * T);
    size_t b_ctl = 256;
    size_t b_part = align256(sizeof(double) * 2 * (size_t)Rn * T * S_max);
    size_t need = b_meta + b_pool + b_keys + b_tc + b_cnt + 2 * b_out + b_ctl + b_part;
    if (need > c->rounds_bytes)
    {
        // Sized once for a whole level of the outer search (<= 4096 rotation cubes: ~45 KB each, ~190 MB), not for this call:
        // the waves of a level grow (32, 128, 512, the rest, then all of them for the lower bounds), and a cudaFree + cudaMalloc
        // per growth inside run() was measured to stall the host for 200-550 ms now and then when another process on the
        // box is in the allocator too (the sporadic outliers of the trimmed searches, profiles/in_search_r02.md).
        const size_t per_cube = need / (size_t)Rn + 1;
        const size_t want = std::max(need, per_cube * 4096 + 4096);
        FG_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_rounds); c->d_rounds = nullptr; c->rounds_bytes = 0;
        if (cudaMalloc(&c->d_rounds, want) == cudaSuccess) c->rounds_bytes = want;
        else
        {
            cudaGetLastError();                          // not enough memory for the generous size: exactly what this call needs
            FG_CUDA(cudaMalloc(&c->d_rounds, need));
            c->rounds_bytes = need;
        }
    }
    int rc = fg::ensure_pinned(c, 4096);
    if (rc) return rc;
    char* p = (char*)c->d_rounds;
    BnbrMeta* d_meta = (BnbrMeta*)p; p += b_meta;
    unsigned long long* d_pool = (unsigned long long*)p; p += b_pool;
    unsigned long long* d_keys = (unsigned long long*)p; p += b_keys;
    float4* d_tc = (float4*)p; p += b_tc;
    int* d_cnt = (int*)p; p += b_cnt;
    float* d_lb = (float*)p; p += b_out;
    float* d_ub = (float*)p; p += b_out;
    unsigned int* d_ctl = (unsigned int*)p; p += b_ctl;
    double* d_part = (double*)p;

    int min_pairs = 12000;                             // fewer popped cubes than this: plain kernel (the sweep has a ~1 ms floor)
    if (const char* e = getenv("FGOICP_BNBR_MIN_PAIRS")) min_pairs = atoi(e);
    bool can_phase = c->trim_k == 0 && c->phased && c->sampler == FGOICP_SAMPLER_PACKED && Rn <= fg_phased_max_cubes(c);
    bool prepared = false;

    size_t smem = sizeof(unsigned long long) * BNB_POOL;
    FG_CUDA(cudaFuncSetAttribute(k_bnbr_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bnbr_init<<<(Rn + 127) / 128, 128, 0, c->stream>>>(d_meta, d_pool, Rn, best_sse, d_cnt, d_tc);
    FG_CUDA(cudaGetLastError());
    volatile unsigned int* h_ctl = (volatile unsigned int*)((char*)c->h_pinned + 2048);
    const bool log_rounds = getenv("FGOICP_BNBR_LOG") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    // a search pops at most 4681 cubes, at least one per round
    bool drained = false;
    // Rounds that cannot use the phase-ordered kernel (trimmed bounds, other samplers) need nothing from the host between
    // rounds -- finished searches return at once, their slots carry negative spans -- so they are enqueued `burst` at a time
    // and the host looks at the counters once per burst: 8x fewer synchronisations, i.e. 8x less exposure to the
    // scheduling of the host thread (the per-round round trip was where the 300-1500 ms outliers of round 1 came from).
    int burst = can_phase ? 1 : 8;
    if (const char* e = getenv("FGOICP_BNBR_BURST")) burst = std::max(1, atoi(e));
    int active_bound = Rn;                             // searches that may still be running (never grows)
    for (int round = 0; round < 4700 && !drained; round += burst)
    {
        for (int b = 0; b < burst; ++b)
        {
            FG_CUDA(cudaMemsetAsync(d_ctl, 0, 8, c->stream));
            k_bnbr_round<<<Rn, BNBR_THREADS, smem, c->stream>>>(d_meta, d_pool, d_keys, d_tc, d_cnt, d_lb, d_ub, thr, (int)c->ns, d_ctl);
            FG_CUDA(cudaGetLastError());
            if (burst > 1)
            {
                int S = std::min(S_max, fg_bounds_slices(c, active_bound));
                rc = fg_bounds_plain_counts(c, d_rot, Rn, fix_rot, d_tc, T, d_cnt, S, d_part, d_lb, d_ub);
                if (rc) return rc;
            }
        }
        FG_CUDA(cudaMemcpyAsync((void*)h_ctl, d_ctl, 8, cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        const int active = (int)h_ctl[0], pairs = (int)h_ctl[1];
        if (log_rounds)
        {
            auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "[bnbr] Rn %d round %d active %d pairs %d  +%.1f us\n", Rn, round + burst - 1, active, pairs,
                    std::chrono::duration<double, std::micro>(now - t_prev).count());
            t_prev = now;
        }
        if (active == 0) { drained = true; break; }
        active_bound = std::min(active_bound, active);
        if (burst > 1) continue;                       // bounds of the burst's rounds are already enqueued
        bool phased = can_phase && pairs >= min_pairs;
        if (phased)
        {
            if (!prepared)
            {
                rc = fg_phased_prepare(c, d_rot, Rn, fix_rot, T);
                if (rc < 0) return rc;
                if (rc > 0) { can_phase = false; phased = false; }
                prepared = rc == 0;
            }
            if (phased)
            {
                rc = fg_phased_eval(c, Rn, d_tc, T, d_lb, d_ub, nullptr);
                if (rc < 0) return rc;
                if (rc > 0) { can_phase = false; phased = false; }
            }
        }
        if (!phased)
        {
            int S = std::min(S_max, fg_bounds_slices(c, active));
            rc = fg_bounds_plain_counts(c, d_rot, Rn, fix_rot, d_tc, T, d_cnt, S, d_part, d_lb, d_ub);
            if (rc) return rc;
        }
    }
    // never hand unfinished bounds to the pruning logic above
    if (!drained) { fg::set_error("inner search did not terminate (round cap reached with live searches)"); return FGOICP_ERR_STATE; }
    k_bnbr_export<<<(Rn + 127) / 128, 128, 0, c->stream>>>(d_meta, Rn, d_out);
    FG_CUDA(cudaGetLastError());
    return FGOICP_OK;
}

// host cubes -> device, run Rn searches, results to host vector
static int bnb_batch_host(fgoicp_ctx* c, const float* rot_xyz_span, int Rn, int fix_rot, float best_sse,
                          float sse_threshold, std::vector<BnbOut>& res, float* ms)
{
    FG_RANGE("fgoicp inner searches");
    size_t b_rot = align256(sizeof(float4) * Rn), b_out = align256(sizeof(BnbOut) * Rn);
    int rc = fg::ensure_scratch(c, b_rot + b_out);
    if (rc) return rc;
    rc = fg::ensure_pinned(c, b_rot + b_out);
    if (rc) return rc;
    char* hp = (char*)c->h_pinned;
    char* dp = (char*)c->d_scratch;
    memcpy(hp, rot_xyz_span, sizeof(float4) * Rn);
    FG_CUDA(cudaMemcpyAsync(dp, hp, sizeof(float4) * Rn, cudaMemcpyHostToDevice, c->stream));
    FG_CUDA(cudaEventRecord(c->ev0, c->stream));
    // Default: one persistent block (cluster) per search.  The round-synchronous schedule (all searches in
    // lock step, each round's cubes through the phase-ordered kernel) is opt-in (fgoicp_set_bnb_mode(2) or
    // FGOICP_BNBR_MIN_CUBES=<n>): measured on W5 it evaluates a 48k-cube round 1.6x faster but the search is a
    // chain of ~30 rounds of <= 4096 cubes for most waves, where the sweep's ~1 ms floor and the host
    // round trip per round cancel the gain (run() 211 ms either way).
    int rounds_min = 0x7fffffff;
    if (const char* e = getenv("FGOICP_BNBR_MIN_CUBES")) rounds_min = atoi(e);
    // trimmed bounds need all residuals of a cube in one place: only the round-synchronous schedule evaluates them
    bool rounds = c->trim_k > 0 || c->bnb_mode == 2 || (c->bnb_mode == 0 && Rn >= rounds_min && c->phased && c->sampler == FGOICP_SAMPLER_PACKED);
    if (rounds) rc = bnb_rounds(c, (const float4*)dp, Rn, fix_rot, best_sse, sse_threshold, (BnbOut*)(dp + b_rot));
    else rc = launch_bnb(c, (const float4*)dp, Rn, fix_rot, best_sse, sse_threshold, (BnbOut*)(dp + b_rot));
    if (rc) return rc;
    FG_CUDA(cudaEventRecord(c->ev1, c->stream));
    FG_CUDA(cudaMemcpyAsync(hp + b_rot, dp + b_rot, sizeof(BnbOut) * Rn, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    res.resize(Rn);
    memcpy(res.data(), hp + b_rot, sizeof(BnbOut) * Rn);
    if (ms) FG_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return FGOICP_OK;
}

extern "C" int fgoicp_bnb_r3_batch(fgoicp_ctx* c, const float* rot_xyz_span, int Rn, int fix_rot,
                                   float best_sse, float sse_threshold,
                                   float* best_ub, float* best_t, uint64_t* evals)
{
    FG_ARG(c && rot_xyz_span && best_ub, "NULL pointer");
    FG_ARG(Rn > 0, "Rn must be positive");
    FG_CUDA(cudaSetDevice(c->device));
    std::vector<BnbOut> res;
    int rc = bnb_batch_host(c, rot_xyz_span, Rn, fix_rot, best_sse, sse_threshold, res, nullptr);
    if (rc) return rc;
    for (int i = 0; i < Rn; ++i)
    {
        best_ub[i] = res[i].best_ub;
        if (best_t) { best_t[3 * i] = res[i].best_t[0]; best_t[3 * i + 1] = res[i].best_t[1]; best_t[3 * i + 2] = res[i].best_t[2]; }
        if (evals) evals[i] = res[i].evals;
    }
    return FGOICP_OK;
}

extern "C" int fgoicp_set_bnb_mode(fgoicp_ctx* c, int mode)
{
    FG_ARG(c, "NULL context");
    FG_ARG(mode >= 0 && mode <= 2, "mode must be 0 (auto), 1 (persistent) or 2 (round-synchronous)");
    c->bnb_mode = mode;
    if (mode == 2) return fg_rounds_prealloc(c);
    return FGOICP_OK;
}

extern "C" int fgoicp_bnb_r3(fgoicp_ctx* c, const float rot_xyz_span[4], int fix_rot,
                             float best_sse, float sse_threshold,
                             float* best_ub, float best_t[3], uint64_t* evals)
{
    return fgoicp_bnb_r3_batch(c, rot_xyz_span, 1, fix_rot, best_sse, sse_threshold, best_ub, best_t, evals);
}

int fg_icp_run_batch(fgoicp_ctx* c, const float* R0s, const float* t0s, int n, int max_iter, float thr,
                     float* sse, float* R, float* t, int* iters);

// Rotation(x, y, z) on the host: same unfused fp32 arithmetic as the reference (common.hpp:37-57)
static void host_rotation(float x, float y, float z, float* R)
{
    float r = x * x + y * y + z * z;
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    if (r > 1.0f) return;
    float ww = 1.0f - r, w = sqrtf(ww);
    float wx = w * x, xx = x * x, wy = w * y, xy = x * y, yy = y * y, wz = w * z, xz = x * z, yz = y * z, zz = z * z;
    R[0] = ww + xx - yy - zz; R[1] = 2 * (xy - wz);     R[2] = 2 * (xz + wy);
    R[3] = 2 * (xy + wz);     R[4] = ww - xx + yy - zz; R[5] = 2 * (yz - wx);
    R[6] = 2 * (xz - wy);     R[7] = 2 * (yz + wx);     R[8] = ww - xx - yy + zz;
}

extern "C" int fgoicp_so3_level_ub(fgoicp_ctx* c, const float* cubes, int n,
                                   float best_sse, float sse_threshold,
                                   float* ub, float* bt,
                                   float* io_best_sse, float io_best_R[9], float io_best_t[3],
                                   fgoicp_level_stats* stats)
{
    FG_RANGE("fgoicp_so3_level_ub");
    FG_ARG(c && io_best_sse && io_best_R && io_best_t, "NULL pointer");
    FG_ARG(n >= 0, "n must be non-negative");
    if (stats) { memset(stats, 0, sizeof(*stats)); stats->best_icp_index = -1; }
    if (n == 0) return FGOICP_OK;
    FG_ARG(cubes && ub && bt, "NULL pointer");
    FG_CUDA(cudaSetDevice(c->device));
    std::vector<BnbOut> res;
    float ms = 0.f;
    int rc = bnb_batch_host(c, cubes, n, 1, best_sse, sse_threshold, res, &ms);       // fgoicp.cpp:69
    if (rc) return rc;
    uint64_t evals = 0;
    for (int i = 0; i < n; ++i)
    {
        ub[i] = res[i].best_ub;
        bt[3 * i] = res[i].best_t[0]; bt[3 * i + 1] = res[i].best_t[1]; bt[3 * i + 2] = res[i].best_t[2];
        evals += res[i].evals;
    }
    // ICP on promising cubes (fgoicp.cpp:74-88).  The trigger compares against the LEVEL-START
    // best_sse so that the set of refined cubes does not depend on how the level was sharded.
    uint32_t n_icp = 0, icp_iters = 0;
    int best_idx = -1;
    cudaEvent_t e0 = c->ev0, e1 = c->ev1;
    FG_CUDA(cudaEventRecord(e0, c->stream));
    std::vector<int> who;
    for (int i = 0; i < n; ++i)
        if ((double)ub[i] < (double)best_sse * 1.8) who.push_back(i);
    if (!who.empty())
    {
        const int m = (int)who.size();
        std::vector<float> R0(9 * (size_t)m), t0(3 * (size_t)m), e(m), R(9 * (size_t)m), t(3 * (size_t)m);
        std::vector<int> it(m);
        for (int k = 0; k < m; ++k)
        {
            int i = who[k];
            host_rotation(cubes[4 * i], cubes[4 * i + 1], cubes[4 * i + 2], &R0[9 * k]);
            memcpy(&t0[3 * k], bt + 3 * i, 3 * sizeof(float));
        }
        // all promising cubes refine concurrently (independent of one another)          fgoicp.cpp:76
        rc = fg_icp_run_batch(c, R0.data(), t0.data(), m, 100, (float)0.005, e.data(), R.data(), t.data(), it.data());
        if (rc) return rc;
        for (int k = 0; k < m; ++k)
        {
            ++n_icp; icp_iters += (uint32_t)it[k];
            if (e[k] < *io_best_sse)                                                      // fgoicp.cpp:79-84, ascending cube order
            {
                *io_best_sse = e[k];
                best_idx = who[k];
                memcpy(io_best_R, &R[9 * k], sizeof(float) * 9);
                memcpy(io_best_t, &t[3 * k], sizeof(float) * 3);
            }
        }
    }
    FG_CUDA(cudaEventRecord(e1, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    float ms_icp = 0.f;
    FG_CUDA(cudaEventElapsedTime(&ms_icp, e0, e1));
    if (stats) { stats->evals = evals; stats->n_icp = n_icp; stats->icp_iters = icp_iters; stats->ms_bnb_ub = ms; stats->ms_icp = ms_icp; stats->best_icp_index = best_idx; }
    return FGOICP_OK;
}

extern "C" int fgoicp_so3_level_lb(fgoicp_ctx* c, const float* cubes, int n,
                                   float best_sse, float sse_threshold,
                                   float* lb, fgoicp_level_stats* stats)
{
    FG_RANGE("fgoicp_so3_level_lb");
    FG_ARG(c, "NULL context");
    FG_ARG(n >= 0, "n must be non-negative");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n == 0) return FGOICP_OK;
    FG_ARG(cubes && lb, "NULL pointer");
    FG_CUDA(cudaSetDevice(c->device));
    std::vector<BnbOut> res;
    float ms = 0.f;
    int rc = bnb_batch_host(c, cubes, n, 0, best_sse, sse_threshold, res, &ms);       // fgoicp.cpp:90
    if (rc) return rc;
    uint64_t evals = 0;
    for (int i = 0; i < n; ++i) { lb[i] = res[i].best_ub; evals += res[i].evals; }
    if (stats) { stats->evals = evals; stats->ms_bnb_lb = ms; }
    return FGOICP_OK;
}
