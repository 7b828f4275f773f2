// bounds_phased.cu -- z-phase-ordered variant of the fused bound evaluation (L2 locality).
//
// k_bounds_multi (bounds.cu) is bound by HBM random access: every evaluation gathers one 32-byte cell
// from a 1+ GB grid at an effectively random address, each miss costs ~3 sectors of DRAM traffic, and
// although every cell is used dozens of times per launch, two uses are too far apart in time to meet
// in L2.  This variant reorders the SAME evaluations so that they do meet:
//
//   1. k_phase_bin: per rotation cube, the rotated data points R p (with their rotation-uncertainty radius
//      in .w) are bucketed by z' = (R p).z into slices of width w = 1/16 and WRITTEN OUT in bucket order
//      (a stable counting sort: point order inside a bucket is ascending index).  The main kernel then
//      streams them sequentially: no index indirection, no rotation, nothing between a bucket's offsets
//      and its gathers but one coalesced load.
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

// Bucket width w = 1/PH_INV_W (a power of two, so leaf translation cubes -- odd multiples of 1/16 -- shift
// buckets by whole numbers).  z' in [-2, 2): |R p| <= sqrt(3) < 2 for data in [-1,1]^3.
#ifndef PH_INV_W_I
#define PH_INV_W_I 16
#endif
#define PH_INV_W   ((float)PH_INV_W_I)
#define PH_NB      (4 * PH_INV_W_I)    // z' buckets
#define PH_ZMIN    (-2.0f)
#define PH_TZOFF   (2 * PH_INV_W_I)    // phase = bucket + round(t.z / w) + PH_TZOFF  (t.z in [-2, 2))
#define PH_PHASES  (PH_NB + 2 * PH_TZOFF)
#ifndef PH_EVICT_LAST
#define PH_EVICT_LAST 0
#endif
#define PH_THREADS 256
#define PH_WARPS   8

__device__ __forceinline__ int ph_bucket(float z)
{
    int b = (int)floorf((z - PH_ZMIN) * PH_INV_W);
    return min(max(b, 0), PH_NB - 1);
}

// One block per rotation cube: rotation matrix, sin(half-angle), stable z'-bucketing of the data points.
__global__ void __launch_bounds__(PH_THREADS)
k_phase_bin(const float4* __restrict__ data, int ns, const float4* __restrict__ rot, int fix_rot,
            float4* __restrict__ P /*[Rn][ns]: (R p, rot_r) in bucket order*/, int* __restrict__ off /*[Rn][PH_NB+1]*/)
{
    __shared__ float sR[10];
    __shared__ int s_cnt[PH_WARPS][PH_NB];      // per-warp bucket counts -> exclusive bases
    __shared__ int s_off[PH_NB + 1];
    const int r = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0)
    {
        float4 rc = rot[r];
        float Rm[9];
        fg_rotation_matrix(rc.x, rc.y, rc.z, Rm);
        for (int k = 0; k < 9; ++k) sR[k] = Rm[k];
        sR[9] = fix_rot ? 0.0f : fg_rot_sin(rc.w);
    }
    for (int i = threadIdx.x; i < PH_WARPS * PH_NB; i += PH_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = sR[k];
    const float sin_half = sR[9];
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
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < c1)
        {
            float4 p = data[i];
            float3 rp = fg_rotate(R, p.x, p.y, p.z);
            b = ph_bucket(rp.z);
            // rot_uncertain_radius = 2 * |p|^2 * sin(half_angle)  (SASS: FADD r,r ; FMUL); 0 when fix_rot
            out = make_float4(rp.x, rp.y, rp.z, __fmul_rn(__fadd_rn(p.w, p.w), sin_half));
        }
        unsigned peers = __match_any_sync(0xffffffffu, b);
        if (b >= 0)
        {
            int rank = __popc(peers & ((1u << lane) - 1u));
            P[(size_t)r * ns + s_cnt[w][b] + rank] = out;
        }
        __syncwarp();
        if (b >= 0 && lane == __ffs(peers) - 1) s_cnt[w][b] += __popc(peers);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// z-groups: translation cubes of one rotation cube that share t.z visit the same bucket in the same
// phase, so the bucket's points are loaded and rotated ONCE for up to PH_G cubes (children of one parent
// cube come as 2 z-values x 4 cubes).  One warp per rotation cube builds its groups.
// ---------------------------------------------------------------------------------------------
#ifndef PH_G
#define PH_G 4
#endif

struct PhGroup
{
    int r;                  // rotation cube
    short tzb;              // phase offset of the group
    short cnt;              // cubes in the group (1..PH_G)
    int pair[PH_G];         // flattened pair indices r*T + c
};

__global__ void __launch_bounds__(32)
k_phase_groups(const float4* __restrict__ tcubes, int T, PhGroup* __restrict__ gpad /*[Rn][T]*/, int* __restrict__ gcount /*[Rn]*/)
{
    const int r = blockIdx.x, lane = threadIdx.x;
    __shared__ int s_tzb[32], s_cube[32];
    int n_groups = 0;
    // T may exceed 32: process the cube list in chunks of 32 (groups never span chunks)
    for (int c0 = 0; c0 < T; c0 += 32)
    {
        int c = c0 + lane;
        bool valid = false;
        int tzb = 0x7fff;   // sentinel > any real phase offset
        if (c < T)
        {
            float4 t = tcubes[(size_t)r * T + c];
            valid = t.w >= 0.0f;                      // negative span marks an unused slot
            if (valid) tzb = __float2int_rn(t.z * PH_INV_W) + PH_TZOFF;
        }
        // stable rank sort by tzb (invalid slots last)
        int rank = 0;
        for (int j = 0; j < 32; ++j)
        {
            int tj = __shfl_sync(0xffffffffu, tzb, j);
            rank += (tj < tzb) || (tj == tzb && j < lane);
        }
        s_tzb[rank] = tzb; s_cube[rank] = c;
        __syncwarp();
        int my_t = s_tzb[lane];
        bool my_valid = my_t != 0x7fff;
        // position inside the run of equal tzb
        int run_start = lane;
        while (run_start > 0 && s_tzb[run_start - 1] == my_t) --run_start;
        bool leader = my_valid && ((lane - run_start) % PH_G == 0);
        unsigned lead_mask = __ballot_sync(0xffffffffu, leader);
        if (leader)
        {
            int g = n_groups + __popc(lead_mask & ((1u << lane) - 1u));
            PhGroup grp;
            grp.r = r; grp.tzb = (short)my_t;
            int cnt = 0;
            for (int k = 0; k < PH_G; ++k)
            {
                int q = lane + k;
                bool in = q < 32 && s_tzb[q] == my_t && (q - run_start) / PH_G == (lane - run_start) / PH_G;
                grp.pair[k] = in ? r * T + s_cube[q] : -1;
                cnt += in;
            }
            grp.cnt = (short)cnt;
            gpad[(size_t)r * T + g] = grp;
        }
        n_groups += __popc(lead_mask);
        __syncwarp();
    }
    if (lane == 0) gcount[r] = n_groups;
}

// single-block exclusive scan: base[0..n] from count[0..n-1]
__global__ void __launch_bounds__(1024) k_phase_scan(const int* __restrict__ cnt, int n, int* __restrict__ base)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += 1024)
    {
        int i = b0 + threadIdx.x;
        int v = i < n ? cnt[i] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0)
        {
            int sv = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, sv, o); if (lane >= o) sv += t; }
            warp_sums[lane] = sv;
        }
        __syncthreads();
        int prefix = carry + (w > 0 ? warp_sums[w - 1] : 0) + incl - v;
        if (i < n) base[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) base[n] = carry;
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
#if PH_EVICT_LAST
    int ix, iy, iz;
    fg_axis(qx, L.ox, L.scale256, L.vmx, ix, r.a);
    fg_axis(qy, L.oy, L.scale256, L.vmy, iy, r.b);
    fg_axis(qz, L.oz, L.scale256, L.vmz, iz, r.c);
    fg_ld256_keep(fg_packed_cell(L, ix, iy, iz), r.v);
#else
    fg_sample_issue<FGOICP_SAMPLER_PACKED>(L, qx, qy, qz, r);
#endif
}

#define PH_MAXGRP 32                  // groups per warp and sweep (one per lane; shared-memory accumulators)

// One (group, bucket) pass: CNT translation cubes against the bucket's rotated points Pr[k0 .. k1).
// p_head is this lane's first point, loaded by the caller while the previous group was being evaluated.
template <int CNT>
__device__ __forceinline__ void ph_run(const LutDev& L, const float4* __restrict__ Pr, int k0, int k1, int lane,
                                       float4 p_head, const float4* tc, double (*acc)[2])
{
    float tx[CNT], ty[CNT], tz[CNT], tsp[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) { float4 t = tc[j]; tx[j] = t.x; ty[j] = t.y; tz[j] = t.z; tsp[j] = t.w; }
    double au[CNT], al[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) { au[j] = 0.0; al[j] = 0.0; }
    float4 p_next = p_head;
    for (int j0 = k0 + lane; j0 < k1; j0 += 32)
    {
        float4 p = p_next;
        if (j0 + 32 < k1) p_next = __ldcs(&Pr[j0 + 32]);                // next point: a plain sequential stream
        SampleReq req[CNT];
#pragma unroll
        for (int j = 0; j < CNT; ++j)
            ph_issue(L, __fadd_rn(p.x, tx[j]), __fadd_rn(p.y, ty[j]), __fadd_rn(p.z, tz[j]), req[j]);
#pragma unroll
        for (int j = 0; j < CNT; ++j)
        {
            float uu, ll;
            fg_bound_terms(fg_sample_finish<FGOICP_SAMPLER_PACKED>(req[j]), p.w, false, tsp[j], uu, ll);
            au[j] += (double)uu; al[j] += (double)ll;
        }
    }
#pragma unroll
    for (int j = 0; j < CNT; ++j)
    {
        double su = fg_warp_sum(au[j]), sl = fg_warp_sum(al[j]);
        if (lane == 0) { acc[j][0] += su; acc[j][1] += sl; }
    }
}

#ifndef PH_MIN_BLOCKS
#define PH_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(PH_THREADS, PH_MIN_BLOCKS)
k_bounds_phased(LutDev L, int ns,
                const float4* __restrict__ tcubes, int Rn, int T,
                const PhGroup* __restrict__ gpad, const int* __restrict__ gbase,
                const float4* __restrict__ P, const int* __restrict__ off,
                float* __restrict__ lb, float* __restrict__ ub, unsigned int* __restrict__ best_ub_bits,
                int* __restrict__ phase_done, int pace_lag, int pf_dist)
{
    __shared__ double s_acc[PH_WARPS][PH_MAXGRP][PH_G][2];
    __shared__ float4 s_tc[PH_WARPS][PH_MAXGRP][PH_G];
    __shared__ int s_pair[PH_WARPS][PH_MAXGRP][PH_G];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_groups = gbase[Rn];
    const long long gw = (long long)blockIdx.x * PH_WARPS + w, nw = (long long)gridDim.x * PH_WARPS;
    const int gfirst = (int)(gw * n_groups / nw), glast = (int)((gw + 1) * n_groups / nw);
    // warps that own at least one group (the others never enter the sweep, so pacing must not wait for them)
    const int nw_paced = (int)min((long long)n_groups, nw);
    // normally one sweep; a warp that owns more than PH_MAXGRP groups sweeps again for the rest (correct,
    // merely out of phase with the others)
    for (int g0 = gfirst; g0 < glast; g0 += PH_MAXGRP)
    {
    const int ng = min(PH_MAXGRP, glast - g0);
    // lane k holds group g0 + k: rotation cube (binary search in gbase), phase offset, cube count; the cubes
    // themselves and the accumulators live in shared memory
    int my_r = 0, my_tzb = 0, my_cnt = 0;
    if (lane < ng)
    {
        int g = g0 + lane;
        int lo = 0, hi = Rn;                                    // largest r with gbase[r] <= g
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (gbase[mid] <= g) lo = mid; else hi = mid; }
        PhGroup grp = gpad[(size_t)lo * T + (g - gbase[lo])];
        my_r = grp.r; my_tzb = grp.tzb; my_cnt = grp.cnt;
#pragma unroll
        for (int j = 0; j < PH_G; ++j)
        {
            s_pair[w][lane][j] = grp.pair[j];
            s_tc[w][lane][j] = __ldg(&tcubes[grp.pair[j < my_cnt ? j : 0]]);
            s_acc[w][lane][j][0] = 0.0; s_acc[w][lane][j][1] = 0.0;
        }
    }
    __syncwarp();

    const bool paced = pace_lag > 0 && g0 == gfirst;           // only the first sweep is in step with the others
    // Pacing is an optimisation, never a dependency: a warp spends at most ~1 ms of polls per launch waiting for
    // the others (e.g. if the blocks were not all co-resident after all) and then simply stops waiting.
    int spin_budget = 8000;
    for (int phi = 0; phi < PH_PHASES; ++phi)
    {
        // Soft pacing: do not run more than pace_lag phases ahead of the slowest warp, so the slabs in flight
        // stay within L2.  Never blocks for long: the wait gives up after a bounded number of polls.
        if (paced && phi >= pace_lag && spin_budget > 0)
        {
            if (lane == 0)
            {
                const volatile int* flag = phase_done + (phi - pace_lag);
#pragma unroll 1
                while (spin_budget > 0 && *flag < nw_paced) { __nanosleep(100); --spin_budget; }
            }
            // keep the budget warp-uniform: the branch above must be taken (or not) by the whole warp
            spin_budget = __shfl_sync(0xffffffffu, spin_budget, 0);
        }
        // Optional slab prefetch (off by default: measured neutral): every warp pulls its share of the
        // packed-grid layers that phase phi + pf_dist adds into L2 with bulk prefetches.
        if (pf_dist > 0 && g0 == gfirst)
        {
            const int psi = phi + pf_dist;
            auto cz_hi = [&](int ps) {
                float zq = PH_ZMIN + ((float)(ps - PH_TZOFF) + 1.5f) * (1.0f / PH_INV_W);
                int c = (int)floorf((zq + L.oz) * L.scale - 0.5f) + 3;           // +1 cell offset, +2 margin
                return min(max(c, 0), L.dz + 1);
            };
            const int c0 = phi == 0 ? 0 : cz_hi(psi - 1), c1 = cz_hi(psi);
            if (c1 > c0)
            {
                const size_t layer = (size_t)L.dx1 * (size_t)L.dy1 * 32;
                const size_t total = (size_t)(c1 - c0) * layer;
                size_t piece = (total / (size_t)(nw * 32) + 127) & ~(size_t)127;
                size_t o = ((size_t)gw * 32 + lane) * piece;
                if (o < total)
                {
                    size_t n = min(piece, total - o);                             // layer is a multiple of 32 B
                    const char* src = (const char*)L.packed + (size_t)c0 * layer + o;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"((unsigned)n) : "memory");
                }
            }
        }
        // all groups of the warp look up their bucket of this phase at once (one memory round trip, not ng)
        int k0 = 0, k1 = 0;
        {
            const int b = phi - my_tzb;
            if (lane < ng && b >= 0 && b < PH_NB)
            {
                const int* ro = off + (size_t)my_r * (PH_NB + 1) + b;
                k0 = __ldg(ro); k1 = __ldg(ro + 1);
            }
        }
        unsigned act = __ballot_sync(0xffffffffu, k1 > k0);
        if (act)
        {
            int k = __ffs(act) - 1; act &= act - 1;
            int ck0 = __shfl_sync(0xffffffffu, k0, k), ck1 = __shfl_sync(0xffffffffu, k1, k);
            const float4* Pr = P + (size_t)__shfl_sync(0xffffffffu, my_r, k) * ns;
            float4 p_head = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ck0 + lane < ck1) p_head = __ldcs(&Pr[ck0 + lane]);
            while (true)
            {
                // first points of the NEXT active group: in flight while this one is evaluated
                int kn = -1, nk0 = 0, nk1 = 0;
                const float4* Pn = Pr;
                float4 p_head_n = make_float4(0.f, 0.f, 0.f, 0.f);
                if (act)
                {
                    kn = __ffs(act) - 1; act &= act - 1;
                    nk0 = __shfl_sync(0xffffffffu, k0, kn); nk1 = __shfl_sync(0xffffffffu, k1, kn);
                    Pn = P + (size_t)__shfl_sync(0xffffffffu, my_r, kn) * ns;
                    if (nk0 + lane < nk1) p_head_n = __ldcs(&Pn[nk0 + lane]);
                }
                const int cnt = __shfl_sync(0xffffffffu, my_cnt, k);
                switch (cnt)                                    // warp-uniform
                {
                case 1: ph_run<1>(L, Pr, ck0, ck1, lane, p_head, s_tc[w][k], s_acc[w][k]); break;
                case 2: ph_run<2>(L, Pr, ck0, ck1, lane, p_head, s_tc[w][k], s_acc[w][k]); break;
                case 3: ph_run<3>(L, Pr, ck0, ck1, lane, p_head, s_tc[w][k], s_acc[w][k]); break;
                default: ph_run<PH_G>(L, Pr, ck0, ck1, lane, p_head, s_tc[w][k], s_acc[w][k]); break;
                }
                if (kn < 0) break;
                k = kn; ck0 = nk0; ck1 = nk1; Pr = Pn; p_head = p_head_n;
            }
        }
        if (paced && lane == 0) atomicAdd(phase_done + phi, 1);
    }
    __syncwarp();
    for (int k = lane; k < ng * PH_G; k += 32)
    {
        int gi = k / PH_G, j = k % PH_G;
        int pr = s_pair[w][gi][j];
        if (pr < 0) continue;
        float fu = (float)s_acc[w][gi][j][0], fl = (float)s_acc[w][gi][j][1];
        ub[pr] = fu; lb[pr] = fl;
        if (best_ub_bits) atomicMin(best_ub_bits, __float_as_uint(fu));
    }
    __syncwarp();
    }   // sweeps
}

// ---------------------------------------------------------------------------------------------
// host side.  Two steps so that callers evaluating MANY cube lists against the SAME rotation cubes (the
// round-synchronous inner search, bnb_rounds.cu) bin the points once:
//   fg_phased_prepare : scratch layout for (Rn, T) + k_phase_bin (rotated points in bucket order)
//   fg_phased_eval    : z-groups of the given translation cubes + the phase-ordered sweep
// Both return FGOICP_OK, a negative error, or 1 = "does not fit this path" (caller uses the plain kernel).
// ---------------------------------------------------------------------------------------------
struct PhLayout
{
    int Rn, T, blocks;
    int* d_off; PhGroup* d_grp; int* d_gcount; int* d_gbase; int* d_pace; float4* d_P;
};

static size_t ph_scratch_cap()
{
    size_t cap = (size_t)2 << 30;
    if (const char* e = getenv("FGOICP_PHASED_SCRATCH_MB")) cap = (size_t)atoll(e) << 20;
    return cap;
}

// largest number of rotation cubes whose bucket-ordered points (16 B per cube and point) fit the scratch cap
int fg_phased_max_cubes(const fgoicp_ctx* c)
{
    return (int)std::min<size_t>((size_t)1 << 30, std::max<size_t>(1, ph_scratch_cap() / (sizeof(float4) * c->ns)));
}

static int ph_layout(fgoicp_ctx* c, int Rn, int T, PhLayout& L)
{
    if (!c->d_packed) return 1;
    if ((long long)Rn * T > (1LL << 30) || Rn > fg_phased_max_cubes(c)) return 1;
    // persistent blocks, all co-resident (a second wave would start its sweep out of phase)
    int per_sm = 0;
    FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounds_phased, PH_THREADS, 0));
    if (per_sm < 1) return 1;
    if (const char* e = getenv("FGOICP_PHASED_BPS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t b_off = al(sizeof(int) * (PH_NB + 1) * (size_t)Rn);
    size_t b_grp = al(sizeof(PhGroup) * (size_t)Rn * T);
    size_t b_cnt = al(sizeof(int) * (size_t)(Rn + 1)) * 2;
    size_t b_pace = al(sizeof(int) * PH_PHASES);
    size_t b_P = sizeof(float4) * (size_t)Rn * c->ns;
    size_t need = b_off + b_grp + b_cnt + b_pace + b_P;
    if (need > c->phase_bytes)
    {
        FG_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_phase); c->d_phase = nullptr; c->phase_bytes = 0;
        FG_CUDA(cudaMalloc(&c->d_phase, need));
        c->phase_bytes = need;
    }
    char* base = (char*)c->d_phase;
    L.Rn = Rn; L.T = T; L.blocks = per_sm * c->sm_count;
    L.d_off = (int*)base;
    L.d_grp = (PhGroup*)(base + b_off);
    L.d_gcount = (int*)(base + b_off + b_grp);
    L.d_gbase = (int*)(base + b_off + b_grp + b_cnt / 2);
    L.d_pace = (int*)(base + b_off + b_grp + b_cnt);
    L.d_P = (float4*)(base + b_off + b_grp + b_cnt + b_pace);
    return FGOICP_OK;
}

int fg_phased_prepare(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, int T)
{
    PhLayout L;
    int rc = ph_layout(c, Rn, T, L);
    if (rc) return rc;
    k_phase_bin<<<Rn, PH_THREADS, 0, c->stream>>>(c->d_data, (int)c->ns, d_rot, fix_rot, L.d_P, L.d_off);
    FG_CUDA(cudaGetLastError());
    return FGOICP_OK;
}

// d_tc[Rn][T] (slots with a negative span are unused) -> d_lb, d_ub [Rn][T]; (Rn, T) as prepared.
// d_best_bits (optional) is atomicMin'ed with the float bits of every ub; the caller initialises it.
int fg_phased_eval(fgoicp_ctx* c, int Rn, const float4* d_tc, int T, float* d_lb, float* d_ub, unsigned int* d_best_bits)
{
    PhLayout L;
    int rc = ph_layout(c, Rn, T, L);
    if (rc) return rc;
    int pace = 2, pf = 0;
    if (const char* e = getenv("FGOICP_PHASED_LAG")) pace = atoi(e);
    if (const char* e = getenv("FGOICP_PHASED_PF")) pf = atoi(e);
    k_phase_groups<<<Rn, 32, 0, c->stream>>>(d_tc, T, L.d_grp, L.d_gcount);
    k_phase_scan<<<1, 1024, 0, c->stream>>>(L.d_gcount, Rn, L.d_gbase);
    FG_CUDA(cudaMemsetAsync(L.d_pace, 0, sizeof(int) * PH_PHASES, c->stream));
    k_bounds_phased<<<L.blocks, PH_THREADS, 0, c->stream>>>(c->lut, (int)c->ns, d_tc, Rn, T, L.d_grp, L.d_gbase, L.d_P, L.d_off,
                                                          d_lb, d_ub, d_best_bits, L.d_pace, pace,
                                                          (long long)Rn * T < 4096 ? 0 : pf);
    FG_CUDA(cudaGetLastError());
    return FGOICP_OK;
}

int fg_bounds_phased(fgoicp_ctx* c, const float4* d_rot, int Rn, int fix_rot, const float4* d_tc, int T,
                     float* d_lb, float* d_ub, float* d_best_ub)
{
    if (!c->d_packed) return 1;
    // the bucket-ordered rotated points take 16 B per (rotation cube, point): if the scratch cap is smaller,
    // sweep the rotation cubes in several launches
    const int Rc = std::min(Rn, fg_phased_max_cubes(c));
    if ((long long)Rc * T > (1LL << 30)) return 1;
    unsigned int* d_bits = (unsigned int*)d_best_ub;
    if (d_bits) FG_CUDA(cudaMemsetAsync(d_bits, 0x7f, 4, c->stream));   // 0x7f7f7f7f: a huge positive float
    for (int r0 = 0; r0 < Rn; r0 += Rc)
    {
        const int rn = std::min(Rc, Rn - r0);
        int rc = fg_phased_prepare(c, d_rot + r0, rn, fix_rot, T);
        if (rc) return rc;
        rc = fg_phased_eval(c, rn, d_tc + (size_t)r0 * T, T, d_lb + (size_t)r0 * T, d_ub + (size_t)r0 * T, d_bits);
        if (rc) return rc;
    }
    return FGOICP_OK;
}
