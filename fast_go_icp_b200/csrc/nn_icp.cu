// nn_icp.cu -- exact nearest-neighbour search, exact SSE and the nested ICP refinement.
//
// Replaces, from the reference:
//   kernComputeClosestError + brute_force_find_nearest_neighbor + thrust::reduce
//       (fgoicp/registration.cu:14-25, 62-86, 154-174)                     -> fgoicp_sse
//   kernFindNearestNeighbor (fgoicp/icp3d.cu:11-28)                         -> NN with rooted compare
//   IterativeClosestPoint3D::{ctor, run, procrustes} (fgoicp/icp3d.cu:55-172) -> fgoicp_icp
// The reference allocates six buffers and uploads both clouds per ICP instance, launches six
// small kernels and four blocking reductions per iteration and does the 3x3 SVD on the host.
// Here the clouds stay resident, the loop state lives in device memory, the SVD runs in a device
// thread, and the host only polls a "done" flag.
//
// Tie rules kept bit-exact:
//   squared compare (K5): strict <, ascending j  -> lowest index among equal d2 wins;
//   rooted compare  (K7): strict >, on sqrtf(d2) -> lowest index among equal sqrtf(d2) wins.
// Both are order-independent once phrased as a lexicographic min over (value, index), which is
// what lets the model cloud be split across thread blocks and merged with a 64-bit atomicMin.
#include "common.cuh"
#include "svd3_device.cuh"
#include "trim.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#define NN_THREADS 256
#define NN_QPT     2                       // queries per thread
#define NN_TILE    1024                    // model points staged in shared memory

struct IcpState
{
    float R[9], t[3];            // current pose (must stay first: kernels read pose = (R, t))
    float lastR[9], lastT[3];
    float Rd[9], td[3];          // last increment
    float abar[3], bbar[3];
    float sse, last_sse;
    float thr;
    int iter, max_iter, done;
    float out_sse, outR[9], outT[3];
    int out_iters;
    double sums[16];
};

// One 512-byte block per concurrent ICP instance (the ICPs of all promising cubes of a level run together).
struct IcpInst
{
    IcpState st;
    float pose0[12];             // seed pose (R0 column-major, t0)
    int job;                     // index of the refinement this slot is running (-1: none)
    int fresh;                   // 1: seeded by the last loop head; the next loop body starts with W = pose0 * data
};

// Slot pool with a device-side job queue: a batch of n refinements runs on S <= n instance slots.  The loop head of
// a slot that has just finished publishes its result and seeds the slot with the next pending job, so every slot
// stays busy until the queue is empty and the host only polls `finished` (no per-job host round trip, and a long
// refinement no longer holds 63 finished neighbours hostage).
struct IcpQueue { int n_jobs, next, finished, pad; };
struct IcpResult { float sse, R[9], t[3]; int iters; };
#define ICP_INST_BYTES 512
static_assert(sizeof(IcpInst) <= ICP_INST_BYTES, "IcpInst must fit its slot");
#define ICP_MAX_BATCH 64

enum { SRC_DATA = 0, SRC_WORK = 1 };                               // query cloud: shared data cloud / instance working copy
enum { POSE_NONE = 0, POSE_SEED = 1, POSE_CUR = 2, POSE_INC = 3 }; // transform applied to the source points first

__device__ __forceinline__ IcpInst* fg_inst(char* base, int k) { return (IcpInst*)(base + (size_t)k * ICP_INST_BYTES); }
__device__ __forceinline__ const float* fg_pose(IcpInst* in, int sel)
{
    return sel == POSE_SEED ? in->pose0 : (sel == POSE_CUR ? in->st.R : (sel == POSE_INC ? in->st.Rd : nullptr));
}

// smallest float x with sqrtf(x) == s  (so that  sqrtf(d) < s  <=>  d < lo(s))
__device__ __forceinline__ float fg_sqrt_preimage_lo(float s)
{
    float x = __fmul_rn(s, s);
    for (int it = 0; it < 4; ++it)
    {
        if (!(x > 0.0f)) break;
        float xm = __uint_as_float(__float_as_uint(x) - 1u);
        if (__fsqrt_rn(xm) == s) x = xm; else break;
    }
    for (int it = 0; it < 4; ++it)
    {
        if (__fsqrt_rn(x) < s) x = __uint_as_float(__float_as_uint(x) + 1u); else break;
    }
    return x;
}

// One block: NN_THREADS*NN_QPT queries against one chunk of the model cloud.
template <int ROOTED>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_brute(const float4* __restrict__ model, int nt, int chunk,
           const float4* __restrict__ data, const float4* __restrict__ work_base, int ns, char* inst_base,
           int src_sel, int pose_sel, unsigned long long* __restrict__ keys_base, int check_done)
{
    __shared__ float4 tile[NN_TILE];
    IcpInst* inst = fg_inst(inst_base, blockIdx.z);
    if (check_done && inst->st.done) return;
    const float4* src = src_sel == SRC_DATA ? data : work_base + (size_t)blockIdx.z * ns;
    const float* pose = fg_pose(inst, pose_sel);
    unsigned long long* keys = keys_base + (size_t)blockIdx.z * ns;

    float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT];
    float best[NN_QPT], thr[NN_QPT];
    int bi[NN_QPT];
    int q0 = blockIdx.x * (NN_THREADS * NN_QPT) + threadIdx.x;
    float R[9], t[3];
    if (pose)
    {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = pose[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = pose[9 + k];
    }
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k)
    {
        int i = min(q0 + k * NN_THREADS, ns - 1);
        float4 p = src[i];
        if (pose)
        {
            // query = R * p + t   (registration.cu:20; SASS: FMUL, FFMA, FFMA, FADD)
            float3 rp = fg_rotate(R, p.x, p.y, p.z);
            qx[k] = __fadd_rn(rp.x, t[0]); qy[k] = __fadd_rn(rp.y, t[1]); qz[k] = __fadd_rn(rp.z, t[2]);
        }
        else { qx[k] = p.x; qy[k] = p.y; qz[k] = p.z; }
        best[k] = FG_INF;                                        // M_INF seed (registration.cu:164, icp3d.cu:16)
        thr[k] = ROOTED ? fg_sqrt_preimage_lo(FG_INF) : FG_INF;
        bi[k] = -1;
    }

    int j0 = blockIdx.y * chunk;
    int j1 = min(nt, j0 + chunk);
    for (int base = j0; base < j1; base += NN_TILE)
    {
        int cnt = min(NN_TILE, j1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += NN_THREADS) tile[i] = model[base + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j)
        {
            float4 m = tile[j];
#pragma unroll
            for (int k = 0; k < NN_QPT; ++k)
            {
                float d = fg_sq3(__fsub_rn(qx[k], m.x), __fsub_rn(qy[k], m.y), __fsub_rn(qz[k], m.z));
                if (d < thr[k])
                {
                    if (ROOTED)
                    {
                        float s = __fsqrt_rn(d);                 // glm::distance (icp3d.cu:20)
                        best[k] = s;
                        thr[k] = fg_sqrt_preimage_lo(s);
                    }
                    else { best[k] = d; thr[k] = d; }
                    bi[k] = __float_as_int(m.w);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k)
    {
        int i = q0 + k * NN_THREADS;
        if (i < ns && bi[k] >= 0)
        {
            unsigned long long key = ((unsigned long long)__float_as_uint(best[k]) << 32) | (unsigned int)bi[k];
            atomicMin(&keys[i], key);
        }
    }
}

// largest float x with sqrtf(x) == s
__device__ __forceinline__ float fg_sqrt_preimage_hi(float s)
{
    float x = __fmul_rn(s, s);
    for (int it = 0; it < 4; ++it)
    {
        float xp = __uint_as_float(__float_as_uint(x) + 1u);
        if (__fsqrt_rn(xp) == s) x = xp; else break;
    }
    for (int it = 0; it < 4; ++it)
    {
        if (__fsqrt_rn(x) > s && x > 0.0f) x = __uint_as_float(__float_as_uint(x) - 1u); else break;
    }
    return x;
}

// squared search radius once a candidate at squared distance d2 is known: its distance plus the winner-memo margin
__device__ __forceinline__ float fg_shrink(float d2, float margin)
{
    if (margin <= 0.0f) return d2 * 1.00002f + 1e-12f;
    const float r = sqrtf(d2) + margin;
    return r * r * 1.00002f + 1e-12f;
}

// Exact NN through the uniform cell grid, one WARP per query.
//   1. The dense distance grid gives a tight search radius for free: the grid node n nearest to q has a
//      model point within sqrt(T[n]), so the NN of q lies within U = sqrt(T[n]) + |q - x_n|.
//   2. Only the (y, z) rows of cells cutting the ball B(q, U) are visited, 32 rows at a time (one per lane);
//      inside a row the chord of the ball selects a run of cells, which is ONE contiguous range of the
//      cell-sorted point array.  A query next to the surface needs 9-16 rows (one pass); a far query has
//      its hundreds of rows spread over the lanes instead of walked serially.
//   3. Candidates are compared as a lexicographic (value, original index) minimum, value = d2 (K5 rule)
//      or sqrtf(d2) (K7 rule), so the visiting order does not matter and ties resolve to the lowest index
//      exactly as the reference's ascending scan does.  Lanes share the shrinking radius after every pass
//      and merge their winners with a shuffle reduction on the packed 64-bit key.
// All pruning tests carry a relative slack: they may visit too much, never too little.
#define NNG_WARPS 4
#ifndef NN_LPQ
#define NN_LPQ 32           // lanes per query (32, 16 or 8); measured on W5: 32 -> 56 ms, 16 -> 71 ms, 8 -> 93 ms of NN time per run()
#endif
#ifndef FG_NN_MARGIN
#define FG_NN_MARGIN 2.5e-4f  // winner-memo scan margin in normalised units (0 disables the memo; env FGOICP_NN_MARGIN overrides)
#endif
#ifndef FG_NN_RHO_SLACK
#define FG_NN_RHO_SLACK 1e-5f // relative slack of the winner memo's clearance (the distance formula rounds at ~1.2e-7 relative); 2e-6 / 5e-7 measured: 9-10 % fewer rooted rescans, ICP -1 % (profiles/nn_scan_r02.md)
#endif
#ifndef NN_RPL
#define NN_RPL 1            // rows per lane and pass of the cell-grid search; measured on the dragon pair (W3, ICP ms): 1 -> 3688, 4 -> 3873, 8 -> 4101
#endif
#ifndef NN_FAST_ROOTED
#define NN_FAST_ROOTED 1      // 0: always the exact rooted scan (measured: ICP 86 -> 71 ms on W5 with 1)
#endif
// barrier over one group of four warps (128 threads) of the block: named barriers 1.. (0 is __syncthreads)
__device__ __forceinline__ void fg_group_barrier(int group)
{
    asm volatile("bar.sync %0, 128;" :: "r"(group + 1) : "memory");
}

// Cooperative scan of ONE query by the `nparts` warps of a group (persistent ICP kernel, heavy queries): warp `part`
// takes every nparts-th run of coarse boxes, and the warps merge winner, runner-up and radius through the group's
// shared-memory slots.  nparts == 1: the plain single-warp scan.
struct NnCoop
{
    int part = 0, nparts = 1, group = 0;
    unsigned long long* s_key = nullptr;      // [nparts]
    float* s_f = nullptr;                     // [nparts]
};

// The exact search of ONE query by one team of NN_LPQ lanes (a whole warp by default): steps 1-3 above.  `prev` is the
// warm start -- the winner of the previous pass of the ICP loop (0xffffffff: none), taken at a position a rounding error
// (squared pass -> next rooted pass) or one ICP increment (rooted -> squared pass) away: its distance is an upper bound on
// the nearest distance, usually the nearest distance itself, and shrinks the ball from "node distance + 2 x offset to the
// node" to just that.  Only the radius changes: the scan still visits every point inside it, so the result (tie rule
// included) is the same.  Returns the packed winner (value bits << 32 | index), all ones if the cloud is empty;
// `rho` receives the proven clearance of the winner (winner memo) when want_rho is set.  Shared by k_nn_grid and by the
// persistent ICP loop kernel, so both produce the same bits by construction.
template <int ROOTED>
__device__ __forceinline__ unsigned long long fg_nn_scan(const CellGrid& g, const LutDev& L, float res, float qx, float qy, float qz,
                                                         unsigned int prev, const float4* __restrict__ model_by_index, float margin,
                                                         bool want_rho, int lane, unsigned int team_mask, float& rho,
                                                         const NnCoop coop = NnCoop())
{
    // minimum over the cooperating warps of a value that is uniform inside each of them (identity when alone)
    auto coop_min_u64 = [&](unsigned long long v)
    {
        if (coop.nparts == 1) return v;
        if (lane == 0) coop.s_key[coop.part] = v;
        fg_group_barrier(coop.group);
        unsigned long long r = coop.s_key[0];
        for (int p = 1; p < coop.nparts; ++p) { const unsigned long long o = coop.s_key[p]; r = o < r ? o : r; }
        fg_group_barrier(coop.group);
        return r;
    };
    auto coop_min_f = [&](float v)
    {
        if (coop.nparts == 1) return v;
        if (lane == 0) coop.s_f[coop.part] = v;
        fg_group_barrier(coop.group);
        float r = coop.s_f[0];
        for (int p = 1; p < coop.nparts; ++p) r = fminf(r, coop.s_f[p]);
        fg_group_barrier(coop.group);
        return r;
    };
    // LUT-space position (binning frame) and the nearest grid node
    float lx = qx + L.ox, ly = qy + L.oy, lz = qz + L.oz;
    int nx = min(max(__float2int_rn(lx * L.scale), 0), L.dx - 1);
    int ny = min(max(__float2int_rn(ly * L.scale), 0), L.dy - 1);
    int nz = min(max(__float2int_rn(lz * L.scale), 0), L.dz - 1);
    float Tn = __ldg(L.grid + ((size_t)nz * L.dy + ny) * L.dx + nx);
    float ex = lx - (float)nx * res, ey = ly - (float)ny * res, ez = lz - (float)nz * res;
    float U = (sqrtf(Tn) + sqrtf(ex * ex + ey * ey + ez * ez)) * 1.0001f + 1e-6f;
    // warm start: see above
    if (model_by_index)
    {
        if (prev != 0xffffffffu)
        {
            float4 m = __ldg(model_by_index + prev);
            float d0 = sqrtf(fg_sq3(__fsub_rn(qx, m.x), __fsub_rn(qy, m.y), __fsub_rn(qz, m.z))) * 1.0001f + 1e-6f + margin;
            U = fminf(U, d0);
        }
    }
    float U2 = U * U;

    // Rooted rule, fast path (NN_FAST_ROOTED): first j minimising sqrtf(d2) (icp3d.cu:17-26) differs from the
    // squared rule only when another point's d2 rounds to the same square root as the winner's, i.e. lies in
    // (d2_min, hi(sqrtf(d2_min))], a window of about one ulp.  Pass 1 searches with the cheap squared compare and
    // tracks the runner-up distance; only if that falls in the window is the query redone with the exact rule.
    const float U2_0 = U2;
    unsigned long long key = 0xffffffffffffffffull;
    rho = 0.0f;                                                    // proven clearance of the winner (winner memo)
#pragma unroll 1
    for (int exact = (ROOTED && NN_FAST_ROOTED) ? 0 : 1; exact < 2; ++exact)
    {
    U2 = U2_0;
    rho = 0.0f;
    float best = FG_INF, thr_lo = (ROOTED && exact) ? fg_sqrt_preimage_lo(FG_INF) : FG_INF, thr_hi = thr_lo;
    float second = FG_INF;
    bool tie = false;                                              // another point at exactly the lane's best distance
    int best_idx = 0x7fffffff;

    const float h = g.h, inv_h = g.inv_h;
    const int cz0 = min(max((int)floorf((lz - U) * inv_h), 0), g.nz - 1);
    const int cz1 = min(max((int)floorf((lz + U) * inv_h), 0), g.nz - 1);
    const int cy0 = min(max((int)floorf((ly - U) * inv_h), 0), g.ny - 1);
    const int cy1 = min(max((int)floorf((ly + U) * inv_h), 0), g.ny - 1);

    // one candidate: its distance, and -- only if it lies inside the current ball -- the winner / runner-up bookkeeping
    auto consider = [&](const float4 m)
    {
        const float d = fg_sq3(__fsub_rn(qx, m.x), __fsub_rn(qy, m.y), __fsub_rn(qz, m.z));
        // outside the current ball: cannot win, and the clearance of the winner memo never looks beyond the ball (every
        // radius the ball shrinks to stays above the distances that matter below)
        if (d <= U2)
        {
            const int idx = __float_as_int(m.w);
            if (ROOTED && !(NN_FAST_ROOTED && !exact))
            {
                if (d < thr_lo)
                {
                    float s = __fsqrt_rn(d);
                    best = s; best_idx = idx;
                    thr_lo = fg_sqrt_preimage_lo(s); thr_hi = fg_sqrt_preimage_hi(s);
                    U2 = fminf(U2, fg_shrink(thr_hi, margin));
                }
                else if (d <= thr_hi && idx < best_idx) best_idx = idx;
            }
            else
            {
                // squared compare that also keeps the runner-up distance (rooted fast path: decides below whether the
                // rooted rule could pick another index; both: clearance of the winner memo)
                if (d < best) { second = best; best = d; best_idx = idx; U2 = fminf(U2, fg_shrink(d, margin)); }
                else if (d == best) { best_idx = min(best_idx, idx); tie = true; }
                else second = fminf(second, d);
            }
        }
    };

    // All rows (y, z) of the cell box [z0, z1] x [y0, y1] that cut the ball, their chords clipped to the cell columns
    // [xlo, xhi]; NN_RPL rows per lane and pass (the two cell-range lookups of all of them are in flight together;
    // measured neutral to slightly negative, see NN_RPL above: the row lookups are not what far queries wait for).
    auto scan_rows = [&](const int z0, const int z1, const int y0, const int y1, const int xlo, const int xhi)
    {
    const int ny_rows = y1 - y0 + 1;
    const int n_rows = ny_rows * (z1 - z0 + 1);
    // row -> (z, y) without an integer division (6.8 % of the ICP loop kernel's instructions, ncu source page of round 2): the
    // quotient is at most the number of cell layers (<= 256), so the float product is within 1e-4 of the true quotient of
    // (row + 0.5) / ny_rows, which is at least 0.5 / ny_rows >= 2e-3 away from an integer; one correction step makes it unconditional
    const float inv_ny_rows = 1.0f / (float)ny_rows;
    for (int row0 = 0; row0 < n_rows; row0 += NN_LPQ * NN_RPL)
    {
        int rb[NN_RPL], re[NN_RPL];
#pragma unroll
        for (int r = 0; r < NN_RPL; ++r)
        {
            rb[r] = 0; re[r] = 0;
            const int row = row0 + r * NN_LPQ + lane;
            if (row < n_rows)
            {
                int rz = (int)(((float)row + 0.5f) * inv_ny_rows);
                rz += ((rz + 1) * ny_rows <= row) - (rz * ny_rows > row);
                int cz = z0 + rz, cy = y0 + (row - rz * ny_rows);
                // edge cells also hold points clamped into them: their slab extends to infinity
                float zlo = cz == 0 ? -FG_INF : (float)cz * h, zhi = cz == g.nz - 1 ? FG_INF : (float)(cz + 1) * h;
                float ylo = cy == 0 ? -FG_INF : (float)cy * h, yhi = cy == g.ny - 1 ? FG_INF : (float)(cy + 1) * h;
                float dz = fmaxf(fmaxf(zlo - lz, lz - zhi), 0.0f);
                float dy = fmaxf(fmaxf(ylo - ly, ly - yhi), 0.0f);
                float dyz2 = dz * dz + dy * dy;
                if (dyz2 <= U2)
                {
                    float wx = sqrtf(U2 - dyz2) * 1.00001f + 1e-6f;
                    int cx0 = max(min(max((int)floorf((lx - wx) * inv_h), 0), g.nx - 1), xlo);
                    int cx1 = min(min(max((int)floorf((lx + wx) * inv_h), 0), g.nx - 1), xhi);
                    if (cx0 <= cx1)
                    {
                        int c0 = (cz * g.ny + cy) * g.nx;
                        rb[r] = __ldg(g.start + c0 + cx0); re[r] = __ldg(g.start + c0 + cx1 + 1);
                    }
                }
            }
        }
        // Candidates of the team's rows.  A far query grazes the surface: the cells its ball cuts hold hundreds to
        // thousands of points, nearly all just outside the ball and owned by a handful of rows; walking each row's range
        // in its own lane left 4-9 of 32 lanes active (ncu: profiles/nn_grid_w3_r01.md).
        //   * LONG rows (>= NN_LPQ candidates: where a grazing ball has its points) are walked by the whole team, one
        //     row at a time, lanes striding the row's contiguous range: nothing per candidate but its load, its distance
        //     and one compare (round 2: the flat dealing below spent 45 % of the kernel's instructions on finding each
        //     candidate's row, profiles/nn_scan_r02.md);
        //   * the SHORT rows that remain are concatenated (prefix sum over the lanes) and dealt evenly: flat position t
        //     belongs to the first lane whose inclusive prefix exceeds t, found by a five-step binary search over
        //     shuffles.
        // Which lane sees which candidate does not matter: winners merge as a lexicographic minimum over (value, index).
#pragma unroll
        for (int r = 0; r < NN_RPL; ++r)
        {
            int cnt = re[r] - rb[r];
            unsigned int longm = __ballot_sync(team_mask, cnt >= NN_LPQ);
            if (NN_LPQ != 32) longm = (longm & team_mask) >> ((threadIdx.x & 31) & ~(NN_LPQ - 1));
            while (longm)
            {
                const int j = __ffs(longm) - 1;
                longm &= longm - 1;
                const int b0 = __shfl_sync(team_mask, rb[r], j, NN_LPQ);
                const int n0 = __shfl_sync(team_mask, cnt, j, NN_LPQ);
                // four candidates per lane in flight: a grazing ball walks thousands of them and a single warp would
                // otherwise wait out one L1 / L2 round trip per 32 candidates
                const float4 nowhere = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.0f);      // beyond any ball: dropped by `consider`
                for (int k = lane; k < n0; k += 4 * NN_LPQ)
                {
                    const float4 m0 = __ldg(g.pts + b0 + k);
                    const float4 m1 = k + NN_LPQ < n0 ? __ldg(g.pts + b0 + k + NN_LPQ) : nowhere;
                    const float4 m2 = k + 2 * NN_LPQ < n0 ? __ldg(g.pts + b0 + k + 2 * NN_LPQ) : nowhere;
                    const float4 m3 = k + 3 * NN_LPQ < n0 ? __ldg(g.pts + b0 + k + 3 * NN_LPQ) : nowhere;
                    consider(m0); consider(m1); consider(m2); consider(m3);
                }
            }
            if (cnt >= NN_LPQ) cnt = 0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < NN_LPQ; o <<= 1)
            {
                const int v = __shfl_up_sync(team_mask, incl, o, NN_LPQ);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(team_mask, incl, NN_LPQ - 1, NN_LPQ);
            const int excl = incl - cnt;
            for (int t0 = 0; t0 < total; t0 += NN_LPQ)
            {
                const int tt = t0 + lane;
                int owner = 0;
#pragma unroll
                for (int step = NN_LPQ / 2; step > 0; step >>= 1)
                {
                    const int v = __shfl_sync(team_mask, incl, owner + step - 1, NN_LPQ);
                    if (v <= tt) owner += step;
                }
                const int ob = __shfl_sync(team_mask, rb[r], owner, NN_LPQ);
                const int oe = __shfl_sync(team_mask, excl, owner, NN_LPQ);
                if (tt < total) consider(__ldg(g.pts + ob + (tt - oe)));
            }
        }
        // share the tightest radius before the next pass
#pragma unroll
        for (int o = NN_LPQ / 2; o > 0; o >>= 1) U2 = fminf(U2, __shfl_xor_sync(team_mask, U2, o, NN_LPQ));
    }
    };

    // Far queries.  The ball of a query at distance D from the surface is empty but for a small patch where it touches
    // the surface, yet proving that costs pi (D / h)^2 row look-ups -- thousands for the data points of a partial-
    // overlap scan that have no counterpart in the model, which is where the refinements of such pairs spend their time
    // (the 6 % far queries of the dragon pair took 90 % of it).  For them the rows are not enumerated over the whole
    // ball: the lanes test the bounding boxes of the NON-EMPTY coarse blocks (FG_COARSE^3 cells; a few hundred for a
    // surface) against the ball, and only the rows of the few blocks that reach into it are scanned.  Exact all the
    // same: every model point lies inside the box of its block, so a point inside the ball makes its block survive
    // (the ball radius carries its usual slack, far above the rounding of the box distance).
    const int n_rows_ball = (cy1 - cy0 + 1) * (cz1 - cz0 + 1);
    if (g.n_coarse > 0 && (coop.nparts > 1 || n_rows_ball > g.coarse_min_rows + (g.n_coarse >> 3)))
    {
        for (int k0 = coop.part * NN_LPQ; k0 < g.n_coarse; k0 += NN_LPQ * coop.nparts)
        {
            const int k = k0 + lane;
            float lb2 = FG_INF;
            unsigned int box = 0, tight = 0;
            if (k < g.n_coarse)
            {
                const float4 lo = __ldg(g.coarse + 2 * k), hi = __ldg(g.coarse + 2 * k + 1);
                const float dx = fmaxf(fmaxf(lo.x - lx, lx - hi.x), 0.0f);
                const float dy = fmaxf(fmaxf(lo.y - ly, ly - hi.y), 0.0f);
                const float dz = fmaxf(fmaxf(lo.z - lz, lz - hi.z), 0.0f);
                lb2 = dx * dx + dy * dy + dz * dz;
                box = __float_as_uint(lo.w);
                tight = __float_as_uint(hi.w);                     // occupied cell range inside the block, 3 bits per bound (ctx.cu)
            }
            unsigned int live = __ballot_sync(team_mask, lb2 <= U2);
            if (NN_LPQ != 32) live = (live & team_mask) >> ((threadIdx.x & 31) & ~(NN_LPQ - 1));
            while (live)
            {
                const int j = __ffs(live) - 1;
                live &= live - 1;
                const float lbj = __shfl_sync(team_mask, lb2, j, NN_LPQ);
                if (lbj > U2) continue;                            // the ball has shrunk meanwhile (U2 is team-uniform here)
                const unsigned int bj = __shfl_sync(team_mask, box, j, NN_LPQ);
                const unsigned int tj = __shfl_sync(team_mask, tight, j, NN_LPQ);
                const int X = (int)(bj & 1023u) * FG_COARSE, Y = (int)((bj >> 10) & 1023u) * FG_COARSE, Z = (int)(bj >> 20) * FG_COARSE;
                // only the rows and columns of the block that hold points (a surface leaves most of its 64 rows empty)
                const int z0 = max(Z + (int)((tj >> 12) & 7u), cz0), z1 = min(Z + (int)((tj >> 15) & 7u), cz1);
                const int y0 = max(Y + (int)((tj >> 6) & 7u), cy0), y1 = min(Y + (int)((tj >> 9) & 7u), cy1);
                if (z0 <= z1 && y0 <= y1) scan_rows(z0, z1, y0, y1, X + (int)(tj & 7u), X + (int)((tj >> 3) & 7u));
            }
        }
    }
    else scan_rows(cz0, cz1, cy0, cy1, 0, g.nx - 1);
    key = 0xffffffffffffffffull;
    if (best_idx != 0x7fffffff) key = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned int)best_idx;
#pragma unroll
    for (int o = NN_LPQ / 2; o > 0; o >>= 1)
    {
        unsigned long long other = __shfl_xor_sync(team_mask, key, o, NN_LPQ);
        key = other < key ? other : key;
    }
    key = coop_min_u64(key);
    if (coop.nparts > 1) U2 = coop_min_f(U2);      // every point inside the smallest of the warps' balls was seen by its owner
    if (want_rho && (!ROOTED || !exact) && key != 0xffffffffffffffffull)
    {
        // Clearance: every point other than the winner is either a candidate this scan has seen (runner-up distance
        // `other`, equal to the winner's on a tie) or lies outside the final ball (all pruning is conservative by
        // ~1e-5, taken as 1e-4 here).  Half the gap, less a slack far above the rounding of the distance formula.
        const unsigned int win_idx = (unsigned int)(key & 0xffffffffull);
        const float wd2 = __uint_as_float((unsigned int)(key >> 32));
        float other = ((unsigned int)best_idx != win_idx || tie) ? best : second;
#pragma unroll
        for (int o = NN_LPQ / 2; o > 0; o >>= 1) other = fminf(other, __shfl_xor_sync(team_mask, other, o, NN_LPQ));
        other = coop_min_f(other);
        const float ds = fminf(sqrtf(other), sqrtf(U2) * 0.9999f);
        rho = 0.5f * (ds - sqrtf(wd2)) - (FG_NN_RHO_SLACK * ds + 1e-7f);
    }
    if (ROOTED && NN_FAST_ROOTED && !exact)
    {
        if (key == 0xffffffffffffffffull) break;
        const float wbest = __uint_as_float((unsigned int)(key >> 32));
        float cand = best > wbest ? best : second;                 // smallest d2 strictly above the winner's, warp-wide
#pragma unroll
        for (int o = NN_LPQ / 2; o > 0; o >>= 1) cand = fminf(cand, __shfl_xor_sync(team_mask, cand, o, NN_LPQ));
        cand = coop_min_f(cand);
        const float s = __fsqrt_rn(wbest);
        if (cand > fg_sqrt_preimage_hi(s))
        {
            key = ((unsigned long long)__float_as_uint(s) << 32) | (key & 0xffffffffull);
            break;                                                 // the usual case: no near-tie, done
        }
    }
    }   // exact
    return key;
}

// Rows of cells the ball of a query would span (the bounding square of its (y, z) projection, as fg_nn_scan sizes it):
// what the persistent ICP kernel uses to tell the heavy queries -- points near the centre of a closed model, equidistant
// from most of its surface, whose scan walks a large share of the cloud -- from the rest.
__device__ __forceinline__ int fg_nn_ball_rows(const CellGrid& g, const LutDev& L, float res, float qx, float qy, float qz,
                                               unsigned int prev, const float4* __restrict__ model_by_index, float margin)
{
    float lx = qx + L.ox, ly = qy + L.oy, lz = qz + L.oz;
    int nx = min(max(__float2int_rn(lx * L.scale), 0), L.dx - 1);
    int ny = min(max(__float2int_rn(ly * L.scale), 0), L.dy - 1);
    int nz = min(max(__float2int_rn(lz * L.scale), 0), L.dz - 1);
    float Tn = __ldg(L.grid + ((size_t)nz * L.dy + ny) * L.dx + nx);
    float ex = lx - (float)nx * res, ey = ly - (float)ny * res, ez = lz - (float)nz * res;
    float U = (sqrtf(Tn) + sqrtf(ex * ex + ey * ey + ez * ez)) * 1.0001f + 1e-6f;
    if (prev != 0xffffffffu)
    {
        float4 m = __ldg(model_by_index + prev);
        U = fminf(U, sqrtf(fg_sq3(__fsub_rn(qx, m.x), __fsub_rn(qy, m.y), __fsub_rn(qz, m.z))) * 1.0001f + 1e-6f + margin);
    }
    const float inv_h = g.inv_h;
    const int cz0 = min(max((int)floorf((lz - U) * inv_h), 0), g.nz - 1), cz1 = min(max((int)floorf((lz + U) * inv_h), 0), g.nz - 1);
    const int cy0 = min(max((int)floorf((ly - U) * inv_h), 0), g.ny - 1), cy1 = min(max((int)floorf((ly + U) * inv_h), 0), g.ny - 1);
    return (cy1 - cy0 + 1) * (cz1 - cz0 + 1);
}

template <int ROOTED>
__global__ void __launch_bounds__(NNG_WARPS * 32)
k_nn_grid(CellGrid g, LutDev L, float res, const float4* __restrict__ data, const float4* __restrict__ work_base,
          int ns, char* inst_base, int src_sel, int pose_sel, unsigned long long* __restrict__ keys_base, int check_done,
          const float4* __restrict__ model_by_index, float4* __restrict__ memo_base, float margin)
{
    IcpInst* inst = fg_inst(inst_base, blockIdx.y);
    if (check_done && inst->st.done) return;
    const float4* src = src_sel == SRC_DATA ? data : work_base + (size_t)blockIdx.y * ns;
    const float* pose = fg_pose(inst, pose_sel);
    unsigned long long* keys = keys_base + (size_t)blockIdx.y * ns;
    float4* memo = memo_base ? memo_base + (size_t)blockIdx.y * ns : nullptr;
    // NN_LPQ lanes per query (a whole warp by default; smaller teams put more queries in flight but were measured
    // slower: the rows of a query are better spread over 32 lanes).  Teams of one warp never talk to each other:
    // every shuffle is confined to the team's lanes.
    const int lane = threadIdx.x & (NN_LPQ - 1);
    const unsigned int team_mask = NN_LPQ == 32 ? 0xffffffffu : (((1u << NN_LPQ) - 1u) << ((threadIdx.x & 31) & ~(NN_LPQ - 1)));
    const int i = (blockIdx.x * NNG_WARPS * 32 + threadIdx.x) / NN_LPQ;
    if (i >= ns) return;
    float4 p = src[i];
    float qx = p.x, qy = p.y, qz = p.z;
    if (pose)
    {
        float R[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = pose[k];
        float3 rp = fg_rotate(R, p.x, p.y, p.z);
        qx = __fadd_rn(rp.x, pose[9]); qy = __fadd_rn(rp.y, pose[10]); qz = __fadd_rn(rp.z, pose[11]);
    }
    // Winner memo (ICP loop).  The last full scan of this point, done at position memo.xyz, proved a clearance
    // memo.w: the winner stays the unique nearest point (by a margin far above fp32 rounding, so under both tie
    // rules) for every query within that distance of the scan position.  The two searches of an ICP iteration --
    // composed pose on the original point (icp3d.cu:103) and the incrementally moved working copy (icp3d.cu:146) --
    // sit a rounding error apart, and late iterations move points by less than the gap to the runner-up: such
    // queries need one distance evaluation instead of a scan.  Only provably unchanged winners take this path.
    if (memo && model_by_index)
    {
        const float4 mm = memo[i];
        const unsigned int prev = (unsigned int)(keys[i] & 0xffffffffull);
        if (prev != 0xffffffffu && mm.w > 0.0f)
        {
            const float ex = qx - mm.x, ey = qy - mm.y, ez = qz - mm.z;
            const float moved = sqrtf(ex * ex + ey * ey + ez * ez) * 1.0001f + 1e-9f;
            if (moved < mm.w)
            {
                const float4 m = __ldg(model_by_index + prev);
                float d = fg_sq3(__fsub_rn(qx, m.x), __fsub_rn(qy, m.y), __fsub_rn(qz, m.z));
                if (ROOTED) d = __fsqrt_rn(d);
                if (lane == 0) keys[i] = ((unsigned long long)__float_as_uint(d) << 32) | prev;
                return;
            }
        }
    }
    const unsigned int prev = model_by_index ? (unsigned int)(keys[i] & 0xffffffffull) : 0xffffffffu;
    float rho;
    const unsigned long long key = fg_nn_scan<ROOTED>(g, L, res, qx, qy, qz, prev, model_by_index, margin, memo != nullptr, lane, team_mask, rho);
    if (lane == 0)
    {
        keys[i] = key;
        if (memo) memo[i] = make_float4(qx, qy, qz, rho);
    }
}

// keys -> (idx, d2 of the winner recomputed with the canonical formula)
__global__ void k_nn_finish(const unsigned long long* __restrict__ keys, const float4* __restrict__ model,
                            const float4* __restrict__ src, int ns, const float* __restrict__ pose,
                            int* __restrict__ idx, float* __restrict__ d2, const int* __restrict__ orig)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const int o = orig[i];       // outputs go to the caller's point order (the device keeps the data cloud in Morton order)
    unsigned int j = (unsigned int)(keys[i] & 0xffffffffull);
    int ji = (j == 0xffffffffu) ? -1 : (int)j;
    if (idx) idx[o] = ji;
    if (d2)
    {
        if (ji < 0) { d2[o] = FG_INF; return; }
        float4 p = src[i];
        float qx = p.x, qy = p.y, qz = p.z;
        if (pose)
        {
            float R[9];
            for (int k = 0; k < 9; ++k) R[k] = pose[k];
            float3 rp = fg_rotate(R, p.x, p.y, p.z);
            qx = __fadd_rn(rp.x, pose[9]); qy = __fadd_rn(rp.y, pose[10]); qz = __fadd_rn(rp.z, pose[11]);
        }
        float4 m = model[ji];
        d2[o] = fg_sq3(__fsub_rn(qx, m.x), __fsub_rn(qy, m.y), __fsub_rn(qz, m.z));
    }
}

// deterministic single-block sum helpers ---------------------------------------------------------
template <int N>
__device__ __forceinline__ void fg_block_sum(double (&v)[N], double* out /* shared or global, N values */)
{
    __shared__ double s_w[32][N];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = fg_warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < N; ++k) s_w[w][k] = v[k];
    __syncthreads();
    if (threadIdx.x < N)
    {
        double a = 0.0;
        for (int q = 0; q < nw; ++q) a += s_w[q][threadIdx.x];
        out[threadIdx.x] = a;
    }
    __syncthreads();
}

// SSE = sum of the winning squared distances (keys carry d2 bits in the high word), fp64 -> float.
// K > 0: trimmed SSE, the K smallest distances only (trim.cuh).
__global__ void __launch_bounds__(1024)
k_sse_reduce(const unsigned long long* __restrict__ keys_base, int ns, char* inst_base, int check_done, float* out,
             unsigned int K)
{
    IcpInst* inst = fg_inst(inst_base, blockIdx.x);
    if (check_done && inst->st.done) return;
    const unsigned long long* keys = keys_base + (size_t)blockIdx.x * ns;
    __shared__ double s_out[1];
    __shared__ unsigned int s_hist[256], s_state[2];
    __shared__ double s_w[32];
    double total;
    if (K > 0 && K < (unsigned int)ns)
    {
        total = fg_block_trimmed_sum([&](int i) { return (unsigned int)(keys[i] >> 32); }, ns, K, s_hist, s_state, s_w);
    }
    else
    {
        double v[1] = { 0.0 };
        for (int i = threadIdx.x; i < ns; i += blockDim.x)
            v[0] += (double)__uint_as_float((unsigned int)(keys[i] >> 32));
        fg_block_sum<1>(v, s_out);
        total = s_out[0];
    }
    if (threadIdx.x == 0)
    {
        float sse = (float)total;
        inst->st.sse = sse;
        if (out) out[blockIdx.x] = sse;
    }
}

// Trimmed ICP: inliers of the Procrustes step = the K correspondences with the smallest (rooted) distance, ties
// taken in the caller's point order.  One block per instance: radix select of the K-th value, then (only if some
// but not all points exactly at that value are in) a second select over their original indices.
__global__ void __launch_bounds__(1024)
k_icp_select(const unsigned long long* __restrict__ keys_base, int ns, char* inst_base, unsigned int K,
             unsigned char* __restrict__ inl_base, const int* __restrict__ orig)
{
    IcpInst* inst = fg_inst(inst_base, blockIdx.x);
    if (inst->st.done) return;
    const unsigned long long* keys = keys_base + (size_t)blockIdx.x * ns;
    unsigned char* inl = inl_base + (size_t)blockIdx.x * ns;
    __shared__ unsigned int s_hist[256], s_state[2];
    FgSelect sel = fg_block_select([&](int i) { return (unsigned int)(keys[i] >> 32); }, ns, K, s_hist, s_state);
    // ties at the K-th value are taken in the CALLER's point order (the oracle's rule): among the points exactly at
    // v_K keep the take_eq ones with the smallest original index -- a second select, over those indices
    unsigned int thr_orig = 0xffffffffu;
    {
        unsigned int n_eq_local = 0;
        for (int i = threadIdx.x; i < ns; i += blockDim.x) n_eq_local += (unsigned int)(keys[i] >> 32) == sel.vk_bits;
        __shared__ unsigned int s_neq;
        if (threadIdx.x == 0) s_neq = 0;
        __syncthreads();
        if (n_eq_local) atomicAdd(&s_neq, n_eq_local);
        __syncthreads();
        if (s_neq > sel.take_eq)                                  // uniform over the block
        {
            FgSelect so = fg_block_select([&](int i) { return (unsigned int)(keys[i] >> 32) == sel.vk_bits ? (unsigned int)orig[i] : 0xffffffffu; },
                                          ns, sel.take_eq, s_hist, s_state);
            thr_orig = so.vk_bits;
        }
    }
    for (int i = threadIdx.x; i < ns; i += blockDim.x)
    {
        unsigned int bits = (unsigned int)(keys[i] >> 32);
        inl[i] = (bits < sel.vk_bits || (bits == sel.vk_bits && (unsigned int)orig[i] <= thr_orig)) ? 1 : 0;
    }
}

// ---- ICP -----------------------------------------------------------------------------------

__device__ __forceinline__ void fg_icp_seed(IcpInst* in, const float* __restrict__ seed, int job, int max_iter, float thr)
{
    IcpState* st = &in->st;
    for (int k = 0; k < 12; ++k) in->pose0[k] = seed[k];
    for (int k = 0; k < 9; ++k) { st->R[k] = seed[k]; st->lastR[k] = seed[k]; }
    for (int k = 0; k < 3; ++k) { st->t[k] = seed[9 + k]; st->lastT[k] = seed[9 + k]; }
    st->sse = FG_INF; st->last_sse = __fmul_rn(2.0f, FG_INF);      // icp3d.cu:89-90
    st->thr = thr; st->iter = 0; st->max_iter = max_iter; st->done = 0;
    st->out_iters = 0;
    in->job = job; in->fresh = 1;
}

// W_i = R * p_i + t  (icp3d.cu:85 with the seed pose; :100 with the increment, in place)
__global__ void k_icp_transform(const float4* __restrict__ data, float4* work_base, int ns, char* inst_base,
                                int src_sel, int pose_sel, int check_done)
{
    IcpInst* inst = fg_inst(inst_base, blockIdx.y);
    if (check_done && inst->st.done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    float4* work = work_base + (size_t)blockIdx.y * ns;
    const float* pose = fg_pose(inst, pose_sel);
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = pose[k];
    float4 p = src_sel == SRC_DATA ? data[i] : work[i];
    float3 rp = fg_rotate(R, p.x, p.y, p.z);
    work[i] = make_float4(__fadd_rn(rp.x, pose[9]), __fadd_rn(rp.y, pose[10]), __fadd_rn(rp.z, pose[11]), p.w);
}

// loop head: while (iter++ < max_iter && (last_sse - sse) > thr * last_sse)   (icp3d.cu:94-98)
__device__ __forceinline__ void fg_icp_loop_head(IcpState* st)
{
    bool go = (st->iter++ < st->max_iter) &&
              (__fsub_rn(st->last_sse, st->sse) > __fmul_rn(st->thr, st->last_sse));
    if (!go)
    {
        st->done = 1;
        st->out_iters = st->iter - 1;
        // return sse < last_sse ? (sse, R, t) : (last_sse, last_R, last_t)   (icp3d.cu:106-107)
        bool cur = st->sse < st->last_sse;
        st->out_sse = cur ? st->sse : st->last_sse;
        for (int k = 0; k < 9; ++k) st->outR[k] = cur ? st->R[k] : st->lastR[k];
        for (int k = 0; k < 3; ++k) st->outT[k] = cur ? st->t[k] : st->lastT[k];
        return;
    }
    st->last_sse = st->sse;
    for (int k = 0; k < 9; ++k) st->lastR[k] = st->R[k];
    for (int k = 0; k < 3; ++k) st->lastT[k] = st->t[k];
}

// First jobs of a batch: slot k runs job k.
__global__ void k_icp_assign(char* inst_base, const float* __restrict__ jobs, int max_iter, float thr)
{
    IcpInst* in = fg_inst(inst_base, blockIdx.x);
    fg_icp_seed(in, jobs + 12 * blockIdx.x, (int)blockIdx.x, max_iter, thr);
    fg_icp_loop_head(&in->st);                                   // loop head of iteration 1
}

// Start of every loop body: slots seeded by the last loop head get W = pose0 * data (icp3d.cu:85) and an empty
// winner list (no warm start for the first search).
__global__ void k_icp_prepare(const float4* __restrict__ data, float4* work_base, unsigned long long* keys_base, int ns,
                              char* inst_base)
{
    IcpInst* inst = fg_inst(inst_base, blockIdx.y);
    if (!inst->fresh || inst->st.done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const float* pose = inst->pose0;
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = pose[k];
    float4 p = data[i];
    float3 rp = fg_rotate(R, p.x, p.y, p.z);
    work_base[(size_t)blockIdx.y * ns + i] = make_float4(__fadd_rn(rp.x, pose[9]), __fadd_rn(rp.y, pose[10]), __fadd_rn(rp.z, pose[11]), p.w);
    keys_base[(size_t)blockIdx.y * ns + i] = 0xffffffffffffffffull;
}

__device__ __forceinline__ void fg_icp_publish(IcpInst* in, const float* __restrict__ jobs, IcpResult* __restrict__ results,
                                               IcpQueue* q, int max_iter, float thr);

// End of every loop body: loop head of the next iteration; a slot whose refinement has ended publishes the result
// and takes the next pending job.
__device__ __forceinline__ void fg_icp_next_job(IcpInst* in, const float* __restrict__ jobs, IcpResult* __restrict__ results,
                                                IcpQueue* q, int max_iter, float thr)
{
    IcpState* st = &in->st;
    in->fresh = 0;
    if (!st->done) fg_icp_loop_head(st);
    fg_icp_publish(in, jobs, results, q, max_iter, thr);
}

// a slot whose refinement has ended publishes its result and takes the next pending job (repeatedly, should a job end
// at its first loop head: max_iter == 0)
__device__ __forceinline__ void fg_icp_publish(IcpInst* in, const float* __restrict__ jobs, IcpResult* __restrict__ results,
                                               IcpQueue* q, int max_iter, float thr)
{
    IcpState* st = &in->st;
    while (st->done && in->job >= 0)
    {
        IcpResult* r = results + in->job;
        r->sse = st->out_sse; r->iters = st->out_iters;
        for (int k = 0; k < 9; ++k) r->R[k] = st->outR[k];
        for (int k = 0; k < 3; ++k) r->t[k] = st->outT[k];
        __threadfence();
        atomicAdd(&q->finished, 1);
        int j = atomicAdd(&q->next, 1);
        if (j >= q->n_jobs) { in->job = -1; break; }
        fg_icp_seed(in, jobs + 12 * j, j, max_iter, thr);
        fg_icp_loop_head(st);                                    // loop head of iteration 1 (ends at once if max_iter == 0)
    }
}

__global__ void k_icp_next(char* inst_base, const float* __restrict__ jobs, IcpResult* __restrict__ results,
                           IcpQueue* q, int max_iter, float thr)
{
    fg_icp_next_job(fg_inst(inst_base, blockIdx.x), jobs, results, q, max_iter, thr);
}

// ---- Procrustes step over ICP_NB blocks per instance ------------------------------------------------------
// Each block sums its slice of the points in fp64; the partials go to HBM and the block that arrives LAST
// (a counter per instance) folds them in block order, so the result is deterministic and the step no longer
// runs at the pace of one thread block (57 -> ~10 us per call for 10,000 points).
#define ICP_NB 8
#define ICP_BT 256
#define ICP_PART 16                     // doubles per block slot

template <int N>
__device__ __forceinline__ bool fg_grid_sum(double (&v)[N], double* part, unsigned int* counter, double* s_total)
{
    __shared__ double s_blk[N];
    __shared__ int s_last;
    fg_block_sum<N>(v, s_blk);
    if (threadIdx.x < N) part[blockIdx.x * ICP_PART + threadIdx.x] = s_blk[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x < N)
    {
        double a = 0.0;
        for (unsigned int q = 0; q < gridDim.x; ++q) a += __ldcg(&part[q * ICP_PART + threadIdx.x]);
        s_total[threadIdx.x] = a;
    }
    if (threadIdx.x == 0) *counter = 0;                     // ready for the next call
    __syncthreads();
    return true;
}

// centroids of the working cloud and of its correspondences (icp3d.cu:150-156)
__global__ void __launch_bounds__(ICP_BT)
k_icp_centroids(const float4* __restrict__ work_base, const unsigned long long* __restrict__ keys_base,
                const float4* __restrict__ model, int ns, char* inst_base,
                const unsigned char* __restrict__ inl_base, int n_in, double* __restrict__ part_base,
                unsigned int* __restrict__ counters)
{
    const int inst = blockIdx.y;
    IcpState* st = &fg_inst(inst_base, inst)->st;
    if (st->done) return;
    const float4* W = work_base + (size_t)inst * ns;
    const unsigned long long* keys = keys_base + (size_t)inst * ns;
    __shared__ double s_out[6];
    const unsigned char* inl = inl_base ? inl_base + (size_t)inst * ns : nullptr;
    const int per = (ns + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(ns, i0 + per);
    double v[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x)
    {
        if (inl && !inl[i]) continue;
        float4 a = W[i];
        float4 b = model[(unsigned int)(keys[i] & 0xffffffffull)];
        v[0] += (double)a.x; v[1] += (double)a.y; v[2] += (double)a.z;
        v[3] += (double)b.x; v[4] += (double)b.y; v[5] += (double)b.z;
    }
    if (!fg_grid_sum<6>(v, part_base + (size_t)inst * ICP_NB * ICP_PART, counters + 2 * inst, s_out)) return;
    if (threadIdx.x < 3)
    {
        st->abar[threadIdx.x] = __fdiv_rn((float)s_out[threadIdx.x], (float)n_in);
        st->bbar[threadIdx.x] = __fdiv_rn((float)s_out[3 + threadIdx.x], (float)n_in);
    }
}

// closest rotation of the summed cross-covariance, increment and composed pose (icp3d.cu:164-172, 101-102);
// one thread per instance
__device__ __noinline__ void fg_icp_pose_update(IcpState* st, const double* sums9, const float* ab, const float* bb)
{
    float ABt[9], Rd[9], td[3], Rn[9], tn[3];
    for (int k = 0; k < 9; ++k) ABt[k] = (float)sums9[k];
    fg_closest_rotation(ABt, Rd);
    // host-side glm arithmetic in the reference: unfused, left to right
    for (int r = 0; r < 3; ++r)
    {
        float ra = __fadd_rn(__fadd_rn(__fmul_rn(Rd[r], ab[0]), __fmul_rn(Rd[3 + r], ab[1])), __fmul_rn(Rd[6 + r], ab[2]));
        td[r] = __fsub_rn(bb[r], ra);                                                   // icp3d.cu:169
    }
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            Rn[c * 3 + r] = __fadd_rn(__fadd_rn(__fmul_rn(Rd[r], st->R[c * 3]), __fmul_rn(Rd[3 + r], st->R[c * 3 + 1])),
                                      __fmul_rn(Rd[6 + r], st->R[c * 3 + 2]));          // R = R_ * R  (icp3d.cu:101)
    for (int r = 0; r < 3; ++r)
    {
        float rt = __fadd_rn(__fadd_rn(__fmul_rn(Rd[r], st->t[0]), __fmul_rn(Rd[3 + r], st->t[1])), __fmul_rn(Rd[6 + r], st->t[2]));
        tn[r] = __fadd_rn(rt, td[r]);                                                   // t = R_ * t + t_  (icp3d.cu:102)
    }
    for (int k = 0; k < 9; ++k) { st->Rd[k] = Rd[k]; st->R[k] = Rn[k]; }
    for (int k = 0; k < 3; ++k) { st->td[k] = td[k]; st->t[k] = tn[k]; }
}

// cross-covariance of the centred clouds, closest rotation, pose update (icp3d.cu:158-172, 101-102)
__global__ void __launch_bounds__(ICP_BT)
k_icp_procrustes(const float4* __restrict__ work_base, const unsigned long long* __restrict__ keys_base,
                 const float4* __restrict__ model, int ns, char* inst_base, const unsigned char* __restrict__ inl_base,
                 double* __restrict__ part_base, unsigned int* __restrict__ counters)
{
    const int inst = blockIdx.y;
    IcpState* st = &fg_inst(inst_base, inst)->st;
    if (st->done) return;
    const float4* W = work_base + (size_t)inst * ns;
    const unsigned long long* keys = keys_base + (size_t)inst * ns;
    __shared__ double s_out[9];
    float ab[3] = { st->abar[0], st->abar[1], st->abar[2] };
    float bb[3] = { st->bbar[0], st->bbar[1], st->bbar[2] };
    const unsigned char* inl = inl_base ? inl_base + (size_t)inst * ns : nullptr;
    const int per = (ns + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(ns, i0 + per);
    double v[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x)
    {
        if (inl && !inl[i]) continue;
        float4 w4 = W[i];
        float4 m4 = model[(unsigned int)(keys[i] & 0xffffffffull)];
        float a[3] = { __fsub_rn(w4.x, ab[0]), __fsub_rn(w4.y, ab[1]), __fsub_rn(w4.z, ab[2]) };   // icp3d.cu:43
        float b[3] = { __fsub_rn(m4.x, bb[0]), __fsub_rn(m4.y, bb[1]), __fsub_rn(m4.z, bb[2]) };
        // glm::outerProduct(a, b)[c][r] = a[r] * b[c]   (icp3d.cu:51)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r)
                v[c * 3 + r] += (double)__fmul_rn(a[r], b[c]);
    }
    if (!fg_grid_sum<9>(v, part_base + (size_t)inst * ICP_NB * ICP_PART, counters + 2 * inst + 1, s_out)) return;
    if (threadIdx.x == 0) fg_icp_pose_update(st, s_out, ab, bb);
}

// ---- the whole ICP loop in ONE persistent cooperative kernel ---------------------------------------------------
// The launch chain above costs nine launches per iteration, each a whole-GPU ramp-up for a few microseconds of work,
// plus a host poll every eight iterations; a search on partial-overlap scans runs ~46,000 such iterations (W3), and on
// W5 the refinements of the two coarse levels are pure latency.  k_icp_loop keeps every block of the GPU resident for
// the whole batch and walks the same stages separated by grid-wide barriers (cooperative launch):
//   S1  thread per (slot, point): fresh slots get W = pose0 * data; the winner memo is tested by ONE THREAD per query
//       (the chain spent a whole warp on it); queries it cannot answer go to a miss list          (icp3d.cu:85, 146)
//   S2  warp per miss, dealt dynamically: the exact rooted search (fg_nn_scan, the same code as k_nn_grid)
//   S3  centroids: fp64 partial sums per (slot, 512-point chunk)                                   (icp3d.cu:150-156)
//   S4  cross-covariance partials; every block folds the centroid partials itself, in chunk order  (icp3d.cu:158-163)
//   S5  one thread per slot: fold, closest rotation (fp64 Jacobi SVD), pose update                 (icp3d.cu:164-172, 101-102)
//   S6  W = Rd * W + td, query = R * data + t, memo test for the squared search                    (icp3d.cu:100, 103)
//   S7  warp per miss: exact squared search
//   S8  SSE partials per (slot, chunk);  S9  one thread per slot: fold, loop head, publish / next job (icp3d.cu:94-98)
// The chunk partition depends on ns alone and partials are folded in chunk order, so results do not depend on the
// grid size, the slot a refinement ran in or what its neighbours did; per-point values come from the same device
// functions as the chain, and every sum is fp64 of fp32 terms rounded once.
// The host launches it once per batch and synchronises once: no polls, no allocation, ~9 barriers per iteration.
namespace cgrp = cooperative_groups;

#define ICPL_CHUNK   512                // points per reduction item
#define ICPL_PART    32                 // doubles per (slot, chunk): [0,6) centroids, [8,17) cross-covariance, [24] SSE

struct IcpLoopCtl
{
    unsigned int n_miss_a, next_a;      // rooted search: misses appended / handed out
    unsigned int n_miss_b, next_b;      // squared search
    int error;                          // 1: iteration guard hit
    unsigned int iterations;            // loop trips (diagnostics)
    unsigned int scans_a, scans_b;      // full scans run (diagnostics: the rest were memo hits)
    unsigned int heavy, pad;            // scans done cooperatively by a group of four warps
    unsigned long long stage_ns[10];    // time block 0 spent in each stage incl. the barrier that ends it (diagnostics)
};

__device__ __forceinline__ unsigned long long fg_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// out-of-line form of the scan: its registers are allocated on their own instead of on top of the loop kernel's state
template <int ROOTED>
__device__ __noinline__ unsigned long long fg_nn_scan_call(const CellGrid* g, const LutDev* L, float res, float qx, float qy, float qz,
                                                           unsigned int prev, const float4* __restrict__ model, float margin, int lane, float* rho)
{
    float r;
    const unsigned long long key = fg_nn_scan<ROOTED>(*g, *L, res, qx, qy, qz, prev, model, margin, margin > 0.0f, lane, 0xffffffffu, r);
    *rho = r;
    return key;
}

struct IcpLoopArgs
{
    CellGrid g; LutDev L; float res;
    const float4* data; const float4* model; float4* work; unsigned long long* keys; float4* memo;
    char* inst; int ns, S;
    const float* jobs; IcpResult* results; IcpQueue* q; int max_iter; float thr;
    double* part; unsigned int* miss; IcpLoopCtl* ctl; unsigned int* arrive;   // arrive: [S][2] arrival counters (cross-covariance, SSE)
    float margin; long long guard_max;
    int heavy_rows;              // a query whose ball spans at least this many rows of cells is scanned by four warps together
};

// winner memo, one thread per query (see k_nn_grid): true = the previous winner provably still wins, key rewritten
template <int ROOTED>
__device__ __forceinline__ bool fg_memo_hit(float qx, float qy, float qz, const float4 mm, unsigned long long old_key,
                                            const float4* __restrict__ model, unsigned long long& new_key)
{
    const unsigned int prev = (unsigned int)(old_key & 0xffffffffull);
    if (prev == 0xffffffffu || !(mm.w > 0.0f)) return false;
    const float ex = qx - mm.x, ey = qy - mm.y, ez = qz - mm.z;
    const float moved = sqrtf(ex * ex + ey * ey + ez * ez) * 1.0001f + 1e-9f;
    if (!(moved < mm.w)) return false;
    const float4 m = __ldg(model + prev);
    float d = fg_sq3(__fsub_rn(qx, m.x), __fsub_rn(qy, m.y), __fsub_rn(qz, m.z));
    if (ROOTED) d = __fsqrt_rn(d);
    new_key = ((unsigned long long)__float_as_uint(d) << 32) | prev;
    return true;
}

// append the items of the warp's lanes whose `miss` is set to the list (one atomic per warp)
__device__ __forceinline__ void fg_miss_append(bool miss, unsigned int item, unsigned int* counter, unsigned int* list, int lane)
{
    const unsigned int m = __ballot_sync(0xffffffffu, miss);
    if (m == 0) return;
    unsigned int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(counter, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (miss) list[base + __popc(m & ((1u << lane) - 1u))] = item;
}

template <int ICPL_THREADS, int MINB, int OUTLINE>
__global__ void __launch_bounds__(ICPL_THREADS, MINB)
k_icp_loop(IcpLoopArgs a)
{
    cgrp::grid_group grid = cgrp::this_grid();
    unsigned long long t_stage = fg_globaltimer();
#define ICPL_STAGE(k) do { if (gtid == 0) { const unsigned long long now__ = fg_globaltimer(); a.ctl->stage_ns[k] += now__ - t_stage; t_stage = now__; } } while (0)
    const int tid = threadIdx.x, lane = tid & 31;
    const int gthreads = (int)gridDim.x * ICPL_THREADS, gtid = (int)blockIdx.x * ICPL_THREADS + tid;
    const int ns = a.ns, S = a.S;
    const int NC = (ns + ICPL_CHUNK - 1) / ICPL_CHUNK;
    const int items = S * ns;
    __shared__ double s_out[16];
    __shared__ float s_ab[6];
    __shared__ int s_last;
    __shared__ unsigned int s_chunk[ICPL_THREADS / 128];          // first miss of the four a group of four warps works on
    __shared__ float4 s_hq[ICPL_THREADS / 128][4];                // heavy queries of a group: position, warm-start index bits
    __shared__ int s_hflag[ICPL_THREADS / 128][4];
    __shared__ unsigned long long s_ckey[ICPL_THREADS / 128][4];  // cooperative scan: per-warp winners / minima
    __shared__ float s_cf[ICPL_THREADS / 128][4];
    const int grp = tid >> 7;

    // first jobs of the batch: slot k runs job k (k_icp_assign)
    if (gtid < S)
    {
        IcpInst* in = fg_inst(a.inst, gtid);
        fg_icp_seed(in, a.jobs + 12 * gtid, gtid, a.max_iter, a.thr);
        fg_icp_loop_head(&in->st);                               // loop head of iteration 1
        fg_icp_publish(in, a.jobs, a.results, a.q, a.max_iter, a.thr);   // (ends at once when max_iter == 0)
    }

    for (long long it = 0;; ++it)
    {
        grid.sync(); ICPL_STAGE(0);                                             // slot states of the prologue / of S9 are visible
        {
            const volatile IcpQueue* vq = a.q;
            if (vq->finished >= vq->n_jobs) break;               // uniform: nobody writes the queue before the next S9
        }
        if (it >= a.guard_max) { if (gtid == 0) a.ctl->error = 1; break; }

        // ---- S1: working copies of fresh slots, memo test of the rooted search
        for (int base = gtid - lane; base < items; base += gthreads)
        {
            const int item = base + lane;
            bool miss = false;
            if (item < items)
            {
                const int slot = item / ns, i = item - slot * ns;
                IcpInst* in = fg_inst(a.inst, slot);
                if (!in->st.done)
                {
                    const size_t o = (size_t)slot * ns + i;
                    if (in->fresh)
                    {
                        const float* pose = in->pose0;
                        float R[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) R[k] = pose[k];
                        const float4 p = a.data[i];
                        const float3 rp = fg_rotate(R, p.x, p.y, p.z);
                        a.work[o] = make_float4(__fadd_rn(rp.x, pose[9]), __fadd_rn(rp.y, pose[10]), __fadd_rn(rp.z, pose[11]), p.w);
                        a.keys[o] = 0xffffffffffffffffull;
                        miss = true;
                    }
                    else
                    {
                        const float4 w = __ldcg(a.work + o);
                        unsigned long long nk;
                        if (a.margin > 0.0f && fg_memo_hit<1>(w.x, w.y, w.z, __ldcg(a.memo + o), __ldcg(a.keys + o), a.model, nk)) a.keys[o] = nk;
                        else miss = true;
                    }
                }
            }
            fg_miss_append(miss, (unsigned int)item, &a.ctl->n_miss_a, a.miss, lane);
        }
        grid.sync(); ICPL_STAGE(1);

        // ---- S2: exact rooted search of the misses, one warp per query, dealt dynamically in small runs
        {
            // Misses are dealt to GROUPS OF FOUR WARPS, four consecutive list entries at a time (one each): neighbouring
            // queries (the data cloud is in Morton order) are searched at the same time on the same SM and share the cell
            // rows and candidate points they pull through its L1 -- the locality and the granularity of a launch with one
            // warp per query and four warps per block.  A group synchronises on its own named barrier, never the block:
            // scans differ 100x in length, and a barrier over 16 warps per chunk (or a run of consecutive misses per
            // warp: 32 unrelated neighbourhoods per SM thrash the L1) measured 1.3-1.6x slower on the dragon pair.
            const unsigned int n_miss = *(volatile unsigned int*)&a.ctl->n_miss_a;
            while (true)
            {
                fg_group_barrier(grp);
                if ((tid & 127) == 0) s_chunk[grp] = atomicAdd(&a.ctl->next_a, 4u);
                fg_group_barrier(grp);
                const unsigned int m0 = s_chunk[grp];
                if (m0 >= n_miss) break;
                const int wig = (tid >> 5) & 3;
                const unsigned int m = m0 + (unsigned int)wig;
                size_t o = 0;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                unsigned int prev = 0xffffffffu;
                bool heavy = false;
                if (m < n_miss)
                {
                    o = (size_t)__ldcg(a.miss + m);
                    w = __ldcg(a.work + o);
                    prev = (unsigned int)(__ldcg(a.keys + o) & 0xffffffffull);
                    heavy = a.g.n_coarse > 0 && fg_nn_ball_rows(a.g, a.L, a.res, w.x, w.y, w.z, prev, a.model, a.margin) >= a.heavy_rows;
                    if (!heavy)
                    {
                        float rho;
                        const unsigned long long key = OUTLINE
                            ? fg_nn_scan_call<1>(&a.g, &a.L, a.res, w.x, w.y, w.z, prev, a.model, a.margin, lane, &rho)
                            : fg_nn_scan<1>(a.g, a.L, a.res, w.x, w.y, w.z, prev, a.model, a.margin, a.margin > 0.0f, lane, 0xffffffffu, rho);
                        if (lane == 0)
                        {
                            a.keys[o] = key;
                            if (a.margin > 0.0f) a.memo[o] = make_float4(w.x, w.y, w.z, rho);
                        }
                    }
                }
                // heavy queries: all four warps of the group scan them together, one after the other
                if (lane == 0) { s_hq[grp][wig] = make_float4(w.x, w.y, w.z, __uint_as_float(prev)); s_hflag[grp][wig] = heavy ? 1 : 0; }
                fg_group_barrier(grp);
                for (int h = 0; h < 4; ++h)
                {
                    if (!s_hflag[grp][h]) continue;                               // uniform over the group
                    const float4 hq = s_hq[grp][h];
                    NnCoop coop;
                    coop.part = wig; coop.nparts = 4; coop.group = grp; coop.s_key = s_ckey[grp]; coop.s_f = s_cf[grp];
                    float rho;
                    const unsigned long long key = fg_nn_scan<1>(a.g, a.L, a.res, hq.x, hq.y, hq.z, __float_as_uint(hq.w), a.model, a.margin,
                                                                 a.margin > 0.0f, lane, 0xffffffffu, rho, coop);
                    if (h == wig && lane == 0)
                    {
                        a.keys[o] = key;
                        if (a.margin > 0.0f) a.memo[o] = make_float4(hq.x, hq.y, hq.z, rho);
                        atomicAdd(&a.ctl->heavy, 1u);
                    }
                }
            }
            if (gtid == 0) { a.ctl->n_miss_b = 0; a.ctl->next_b = 0; a.ctl->scans_a += n_miss; }   // the squared search's list is idle here
        }
        grid.sync(); ICPL_STAGE(2);

        // ---- S3: centroid partials of the working cloud and of its correspondences (icp3d.cu:150-156)
        for (int it2 = blockIdx.x; it2 < S * NC; it2 += gridDim.x)
        {
            const int slot = it2 / NC, c = it2 - slot * NC;
            if (fg_inst(a.inst, slot)->st.done) continue;
            const int i1 = min(ns, (c + 1) * ICPL_CHUNK);
            double v[6] = { 0, 0, 0, 0, 0, 0 };
            for (int i = c * ICPL_CHUNK + tid; i < i1; i += ICPL_THREADS)
            {
                const size_t o = (size_t)slot * ns + i;
                const float4 w4 = __ldcg(a.work + o);
                const float4 m4 = __ldg(a.model + (unsigned int)(__ldcg(a.keys + o) & 0xffffffffull));
                v[0] += (double)w4.x; v[1] += (double)w4.y; v[2] += (double)w4.z;
                v[3] += (double)m4.x; v[4] += (double)m4.y; v[5] += (double)m4.z;
            }
            fg_block_sum<6>(v, s_out);
            if (tid < 6) a.part[(size_t)it2 * ICPL_PART + tid] = s_out[tid];
            __syncthreads();
        }
        grid.sync(); ICPL_STAGE(3);

        // ---- S4: cross-covariance partials of the centred clouds (icp3d.cu:158-163)
        for (int it2 = blockIdx.x; it2 < S * NC; it2 += gridDim.x)
        {
            const int slot = it2 / NC, c = it2 - slot * NC;
            IcpState* st = &fg_inst(a.inst, slot)->st;
            if (st->done) continue;
            if (tid < 6)
            {
                double acc = 0.0;
                for (int q = 0; q < NC; ++q) acc += __ldcg(a.part + ((size_t)slot * NC + q) * ICPL_PART + tid);     // chunk order
                s_ab[tid] = __fdiv_rn((float)acc, (float)ns);
            }
            __syncthreads();
            const float ab[3] = { s_ab[0], s_ab[1], s_ab[2] }, bb[3] = { s_ab[3], s_ab[4], s_ab[5] };
            if (c == 0 && tid < 3) { st->abar[tid] = ab[tid]; st->bbar[tid] = bb[tid]; }
            const int i1 = min(ns, (c + 1) * ICPL_CHUNK);
            double v[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            for (int i = c * ICPL_CHUNK + tid; i < i1; i += ICPL_THREADS)
            {
                const size_t o = (size_t)slot * ns + i;
                const float4 w4 = __ldcg(a.work + o);
                const float4 m4 = __ldg(a.model + (unsigned int)(__ldcg(a.keys + o) & 0xffffffffull));
                const float av[3] = { __fsub_rn(w4.x, ab[0]), __fsub_rn(w4.y, ab[1]), __fsub_rn(w4.z, ab[2]) };   // icp3d.cu:43
                const float bv[3] = { __fsub_rn(m4.x, bb[0]), __fsub_rn(m4.y, bb[1]), __fsub_rn(m4.z, bb[2]) };
                // glm::outerProduct(a, b)[c][r] = a[r] * b[c]   (icp3d.cu:51)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        v[cc * 3 + r] += (double)__fmul_rn(av[r], bv[cc]);
            }
            fg_block_sum<9>(v, s_out);
            if (tid < 9) a.part[(size_t)it2 * ICPL_PART + 8 + tid] = s_out[tid];
            // ---- S5, by whichever block finishes the slot's partials LAST (no grid barrier in between): fold in chunk
            // order, closest rotation and pose update in one thread (icp3d.cu:164-172, 101-102)
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = atomicAdd(a.arrive + 2 * slot, 1u) == (unsigned int)(NC - 1);
            __syncthreads();
            if (s_last)
            {
                __threadfence();
                if (tid < 9)
                {
                    double acc = 0.0;
                    for (int q = 0; q < NC; ++q) acc += __ldcg(a.part + ((size_t)slot * NC + q) * ICPL_PART + 8 + tid);
                    s_out[tid] = acc;
                }
                __syncthreads();
                if (tid == 0)
                {
                    fg_icp_pose_update(st, s_out, ab, bb);           // ab, bb: folded by this block from the same partials
                    a.arrive[2 * slot] = 0;
                }
            }
            __syncthreads();
        }
        grid.sync(); ICPL_STAGE(4);

        // ---- S6: W = Rd * W + td (icp3d.cu:100); query of the SSE search = R * data + t (icp3d.cu:103); memo test
        for (int base = gtid - lane; base < items; base += gthreads)
        {
            const int item = base + lane;
            bool miss = false;
            if (item < items)
            {
                const int slot = item / ns, i = item - slot * ns;
                IcpInst* in = fg_inst(a.inst, slot);
                if (!in->st.done)
                {
                    const size_t o = (size_t)slot * ns + i;
                    const IcpState* st = &in->st;
                    float Rd[9], R[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) { Rd[k] = st->Rd[k]; R[k] = st->R[k]; }
                    const float4 w = __ldcg(a.work + o);
                    const float3 rw = fg_rotate(Rd, w.x, w.y, w.z);
                    a.work[o] = make_float4(__fadd_rn(rw.x, st->td[0]), __fadd_rn(rw.y, st->td[1]), __fadd_rn(rw.z, st->td[2]), w.w);
                    const float4 p = a.data[i];
                    const float3 rp = fg_rotate(R, p.x, p.y, p.z);
                    const float qx = __fadd_rn(rp.x, st->t[0]), qy = __fadd_rn(rp.y, st->t[1]), qz = __fadd_rn(rp.z, st->t[2]);
                    unsigned long long nk;
                    if (a.margin > 0.0f && fg_memo_hit<0>(qx, qy, qz, __ldcg(a.memo + o), __ldcg(a.keys + o), a.model, nk)) a.keys[o] = nk;
                    else miss = true;
                }
            }
            fg_miss_append(miss, (unsigned int)item, &a.ctl->n_miss_b, a.miss, lane);
        }
        grid.sync(); ICPL_STAGE(6);

        // ---- S7: exact squared search of the misses
        {
            const unsigned int n_miss = *(volatile unsigned int*)&a.ctl->n_miss_b;
            while (true)
            {
                fg_group_barrier(grp);
                if ((tid & 127) == 0) s_chunk[grp] = atomicAdd(&a.ctl->next_b, 4u);
                fg_group_barrier(grp);
                const unsigned int m0 = s_chunk[grp];
                if (m0 >= n_miss) break;
                const int wig = (tid >> 5) & 3;
                const unsigned int m = m0 + (unsigned int)wig;
                size_t o = 0;
                float qx = 0.f, qy = 0.f, qz = 0.f;
                unsigned int prev = 0xffffffffu;
                bool heavy = false;
                if (m < n_miss)
                {
                    const unsigned int item = __ldcg(a.miss + m);
                    const int slot = (int)(item / (unsigned int)ns), i = (int)(item - (unsigned int)slot * (unsigned int)ns);
                    o = (size_t)item;
                    const IcpState* st = &fg_inst(a.inst, slot)->st;
                    float R[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) R[k] = st->R[k];
                    const float4 p = a.data[i];
                    const float3 rp = fg_rotate(R, p.x, p.y, p.z);
                    qx = __fadd_rn(rp.x, st->t[0]); qy = __fadd_rn(rp.y, st->t[1]); qz = __fadd_rn(rp.z, st->t[2]);
                    prev = (unsigned int)(__ldcg(a.keys + o) & 0xffffffffull);
                    heavy = a.g.n_coarse > 0 && fg_nn_ball_rows(a.g, a.L, a.res, qx, qy, qz, prev, a.model, a.margin) >= a.heavy_rows;
                    if (!heavy)
                    {
                        float rho;
                        const unsigned long long key = OUTLINE
                            ? fg_nn_scan_call<0>(&a.g, &a.L, a.res, qx, qy, qz, prev, a.model, a.margin, lane, &rho)
                            : fg_nn_scan<0>(a.g, a.L, a.res, qx, qy, qz, prev, a.model, a.margin, a.margin > 0.0f, lane, 0xffffffffu, rho);
                        if (lane == 0)
                        {
                            a.keys[o] = key;
                            if (a.margin > 0.0f) a.memo[o] = make_float4(qx, qy, qz, rho);
                        }
                    }
                }
                if (lane == 0) { s_hq[grp][wig] = make_float4(qx, qy, qz, __uint_as_float(prev)); s_hflag[grp][wig] = heavy ? 1 : 0; }
                fg_group_barrier(grp);
                for (int h = 0; h < 4; ++h)
                {
                    if (!s_hflag[grp][h]) continue;                               // uniform over the group
                    const float4 hq = s_hq[grp][h];
                    NnCoop coop;
                    coop.part = wig; coop.nparts = 4; coop.group = grp; coop.s_key = s_ckey[grp]; coop.s_f = s_cf[grp];
                    float rho;
                    const unsigned long long key = fg_nn_scan<0>(a.g, a.L, a.res, hq.x, hq.y, hq.z, __float_as_uint(hq.w), a.model, a.margin,
                                                                 a.margin > 0.0f, lane, 0xffffffffu, rho, coop);
                    if (h == wig && lane == 0)
                    {
                        a.keys[o] = key;
                        if (a.margin > 0.0f) a.memo[o] = make_float4(hq.x, hq.y, hq.z, rho);
                        atomicAdd(&a.ctl->heavy, 1u);
                    }
                }
            }
            if (gtid == 0) { a.ctl->n_miss_a = 0; a.ctl->next_a = 0; a.ctl->scans_b += n_miss; a.ctl->iterations += 1; }
        }
        grid.sync(); ICPL_STAGE(7);

        // ---- S8: SSE partials (keys carry d2 bits in the high word)
        for (int it2 = blockIdx.x; it2 < S * NC; it2 += gridDim.x)
        {
            const int slot = it2 / NC, c = it2 - slot * NC;
            if (fg_inst(a.inst, slot)->st.done) continue;
            const int i1 = min(ns, (c + 1) * ICPL_CHUNK);
            double v[1] = { 0.0 };
            for (int i = c * ICPL_CHUNK + tid; i < i1; i += ICPL_THREADS)
                v[0] += (double)__uint_as_float((unsigned int)(__ldcg(a.keys + (size_t)slot * ns + i) >> 32));
            fg_block_sum<1>(v, s_out);
            if (tid == 0)
            {
                a.part[(size_t)it2 * ICPL_PART + 24] = s_out[0];
                // ---- S9, by the block that finishes the slot's partials last: SSE in chunk order, loop head of the next
                // iteration, publish / next job (k_sse_reduce + k_icp_next)
                __threadfence();
                if (atomicAdd(a.arrive + 2 * slot + 1, 1u) == (unsigned int)(NC - 1))
                {
                    __threadfence();
                    IcpInst* in = fg_inst(a.inst, slot);
                    double acc = 0.0;
                    for (int q = 0; q < NC; ++q) acc += __ldcg(a.part + ((size_t)slot * NC + q) * ICPL_PART + 24);
                    in->st.sse = (float)acc;
                    a.arrive[2 * slot + 1] = 0;
                    fg_icp_next_job(in, a.jobs, a.results, a.q, a.max_iter, a.thr);
                }
            }
            __syncthreads();
        }
        // slots that were idle in this trip (no job left) need nothing: their state does not change any more
    }
}
#undef ICPL_STAGE

// ---------------------------------------------------------------------------------------------

static size_t icp_loop_part_bytes(const fgoicp_ctx* c, int n)
{
    const size_t nc = (c->ns + ICPL_CHUNK - 1) / ICPL_CHUNK;
    return sizeof(double) * ICPL_PART * nc * (size_t)n;
}
#define ICPL_HEAD 4096                  // control block (first 1 KB) + arrival counters [slots][2]
static size_t icp_loop_bytes(const fgoicp_ctx* c, int n) { return ICPL_HEAD + icp_loop_part_bytes(c, n) + sizeof(unsigned int) * c->ns * (size_t)n; }

// per-context ICP buffers sized for `n` concurrent instances
int fg_ensure_icp_capacity(fgoicp_ctx* c, int n)
{
    if (n <= c->icp_capacity) return FGOICP_OK;
    // size for a whole batch at once when that is cheap (25 bytes per instance and data point): the capacity then
    // never changes during a search (no cudaFree / cudaMalloc between levels)
    if ((size_t)ICP_MAX_BATCH * c->ns * 41 <= ((size_t)512 << 20)) n = std::max(n, ICP_MAX_BATCH);
    FG_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_work); cudaFree(c->d_nnkey); cudaFree(c->d_icp); cudaFree(c->d_inl); cudaFree(c->d_icp_part); cudaFree(c->d_nnmemo); cudaFree(c->d_icp_loop);
    c->d_work = nullptr; c->d_nnkey = nullptr; c->d_icp = nullptr; c->d_inl = nullptr; c->d_icp_part = nullptr; c->d_nnmemo = nullptr; c->d_icp_loop = nullptr; c->icp_capacity = 0;
    FG_CUDA(cudaMalloc(&c->d_work, sizeof(float4) * c->ns * n));
    FG_CUDA(cudaMalloc(&c->d_nnmemo, sizeof(float4) * c->ns * n));
    FG_CUDA(cudaMalloc(&c->d_nnkey, sizeof(unsigned long long) * c->ns * n));
    FG_CUDA(cudaMalloc(&c->d_icp, (size_t)ICP_INST_BYTES * n + 256));
    FG_CUDA(cudaMalloc(&c->d_inl, c->ns * (size_t)n));
    // Procrustes partial sums (ICP_NB slots of ICP_PART doubles per instance) followed by 2 arrival counters per instance
    size_t part_bytes = sizeof(double) * ICP_NB * ICP_PART * (size_t)n;
    FG_CUDA(cudaMalloc(&c->d_icp_part, part_bytes + sizeof(unsigned int) * 2 * (size_t)n));
    FG_CUDA(cudaMemsetAsync((char*)c->d_icp_part + part_bytes, 0, sizeof(unsigned int) * 2 * (size_t)n, c->stream));
    // persistent loop kernel: control block | partial sums [n][chunks][ICPL_PART] | miss list [n][ns]
    FG_CUDA(cudaMalloc(&c->d_icp_loop, icp_loop_bytes(c, n)));
    c->icp_capacity = n;
    return FGOICP_OK;
}

// coarse boxes of the cell grid for the far-query path of fg_nn_scan (FGOICP_NN_COARSE=0 turns it off,
// FGOICP_NN_COARSE_MIN_ROWS=<n> moves the switch-over: test hooks, results are exact either way)
static void fg_set_coarse(const fgoicp_ctx* c, CellGrid& g)
{
    static const int on = getenv("FGOICP_NN_COARSE") ? atoi(getenv("FGOICP_NN_COARSE")) : 1;
    static const int min_rows = getenv("FGOICP_NN_COARSE_MIN_ROWS") ? atoi(getenv("FGOICP_NN_COARSE_MIN_ROWS")) : 320;
    const bool use = (on && c->nn_mode != 3) || c->nn_mode == 2;
    g.coarse = use ? c->d_coarse : nullptr;
    g.n_coarse = use ? c->n_coarse : 0;
    g.coarse_min_rows = c->nn_mode == 2 ? -(1 << 30) : min_rows;
}

static void nn_geometry(const fgoicp_ctx* c, dim3& grid, int& chunk, int n_inst)
{
    int qtiles = (int)((c->ns + NN_THREADS * NN_QPT - 1) / (NN_THREADS * NN_QPT));
    int want_chunks = std::max(1, (4 * c->sm_count + qtiles * n_inst - 1) / (qtiles * n_inst));
    chunk = (int)((c->nt + want_chunks - 1) / want_chunks);
    chunk = ((chunk + NN_TILE - 1) / NN_TILE) * NN_TILE;
    int nchunks = (int)((c->nt + chunk - 1) / chunk);
    grid = dim3(qtiles, nchunks, n_inst);
}

// enqueue for instances [0, n_inst): keys := NN of pose(src)
static int enqueue_nn(fgoicp_ctx* c, int n_inst, int src_sel, int pose_sel, int rooted, int check_done)
{
    char* inst = (char*)c->d_icp;
    if (c->nn_mode != 1)
    {
        CellGrid g;
        g.start = c->d_cell_start; g.pts = c->d_cell_M;
        g.nx = c->cnx; g.ny = c->cny; g.nz = c->cnz; g.h = c->cell_h; g.inv_h = c->cell_inv_h;
        fg_set_coarse(c, g);
        const int qpb = NNG_WARPS * 32 / NN_LPQ;            // queries per block
        // inside the ICP loop the key buffer carries the previous pass's winners (all-ones before the first pass)
        const float4* warm = (check_done && !getenv("FGOICP_NN_NO_WARM")) ? c->d_model : nullptr;
        // winner memo (ICP loop only): scans reach `margin` beyond the winner so that the proven clearance covers the
        // rounding-level offset between the two searches of an iteration and the small moves of late iterations
        static const float margin_cfg = getenv("FGOICP_NN_MARGIN") ? (float)atof(getenv("FGOICP_NN_MARGIN")) : FG_NN_MARGIN;
        float4* memo = (warm && margin_cfg > 0.0f) ? c->d_nnmemo : nullptr;
        const float margin = memo ? margin_cfg : 0.0f;
        dim3 grid((unsigned)((c->ns + qpb - 1) / qpb), (unsigned)n_inst);
        if (rooted)
            k_nn_grid<1><<<grid, NNG_WARPS * 32, 0, c->stream>>>(g, c->lut, c->res, c->d_data, c->d_work, (int)c->ns, inst, src_sel, pose_sel, c->d_nnkey, check_done, warm, memo, margin);
        else
            k_nn_grid<0><<<grid, NNG_WARPS * 32, 0, c->stream>>>(g, c->lut, c->res, c->d_data, c->d_work, (int)c->ns, inst, src_sel, pose_sel, c->d_nnkey, check_done, warm, memo, margin);
        FG_CUDA(cudaGetLastError());
        return FGOICP_OK;
    }
    dim3 grid; int chunk;
    nn_geometry(c, grid, chunk, n_inst);
    FG_CUDA(cudaMemsetAsync(c->d_nnkey, 0xff, sizeof(unsigned long long) * c->ns * n_inst, c->stream));
    if (rooted)
        k_nn_brute<1><<<grid, NN_THREADS, 0, c->stream>>>(c->d_model, (int)c->nt, chunk, c->d_data, c->d_work, (int)c->ns, inst, src_sel, pose_sel, c->d_nnkey, check_done);
    else
        k_nn_brute<0><<<grid, NN_THREADS, 0, c->stream>>>(c->d_model, (int)c->nt, chunk, c->d_data, c->d_work, (int)c->ns, inst, src_sel, pose_sel, c->d_nnkey, check_done);
    FG_CUDA(cudaGetLastError());
    return FGOICP_OK;
}

// seed poses of n instances -> pose0 slots (through pinned staging)
static int upload_seeds(fgoicp_ctx* c, const float* R0s, const float* t0s, int n)
{
    int rc = fg_ensure_icp_capacity(c, n);
    if (rc) return rc;
    rc = fg::ensure_pinned(c, (size_t)n * (sizeof(IcpInst) + 64) + 4096);
    if (rc) return rc;
    float* hp = (float*)c->h_pinned;
    for (int k = 0; k < n; ++k)
    {
        memcpy(hp + 12 * k, R0s + 9 * k, 9 * sizeof(float));
        memcpy(hp + 12 * k + 9, t0s + 3 * k, 3 * sizeof(float));
    }
    FG_CUDA(cudaMemcpy2DAsync((char*)c->d_icp + offsetof(IcpInst, pose0), ICP_INST_BYTES, hp, 12 * sizeof(float),
                              12 * sizeof(float), n, cudaMemcpyHostToDevice, c->stream));
    return FGOICP_OK;
}

extern "C" int fgoicp_sse(fgoicp_ctx* c, const float R[9], const float t[3], float* sse)
{
    FG_ARG(c && R && t && sse, "NULL pointer");
    FG_CUDA(cudaSetDevice(c->device));
    int rc = upload_seeds(c, R, t, 1);
    if (rc) return rc;
    rc = enqueue_nn(c, 1, SRC_DATA, POSE_SEED, 0, 0);
    if (rc) return rc;
    float* d_out = (float*)((char*)c->d_icp + (size_t)ICP_INST_BYTES * c->icp_capacity);
    k_sse_reduce<<<1, 1024, 0, c->stream>>>(c->d_nnkey, (int)c->ns, (char*)c->d_icp, 0, d_out, (unsigned int)c->trim_k);
    FG_CUDA(cudaGetLastError());
    float* hp = (float*)c->h_pinned;
    FG_CUDA(cudaMemcpyAsync(hp + 32, d_out, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    *sse = hp[32];
    return FGOICP_OK;
}

extern "C" int fgoicp_nn(fgoicp_ctx* c, const float R[9], const float t[3], int rooted, int32_t* idx, float* d2)
{
    FG_ARG(c && R && t, "NULL pointer");
    FG_CUDA(cudaSetDevice(c->device));
    int rc = upload_seeds(c, R, t, 1);
    if (rc) return rc;
    rc = enqueue_nn(c, 1, SRC_DATA, POSE_SEED, rooted, 0);
    if (rc) return rc;
    size_t ns = c->ns;
    rc = fg::ensure_scratch(c, ns * 8);
    if (rc) return rc;
    int* d_idx = (int*)c->d_scratch;
    float* d_d2 = (float*)c->d_scratch + ns;
    const float* d_pose = (const float*)((char*)c->d_icp + offsetof(IcpInst, pose0));
    k_nn_finish<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>(c->d_nnkey, c->d_model, c->d_data, (int)ns, d_pose, d_idx, d_d2, c->d_data_orig);
    FG_CUDA(cudaGetLastError());
    if (idx) FG_CUDA(cudaMemcpyAsync(idx, d_idx, ns * 4, cudaMemcpyDeviceToHost, c->stream));
    if (d2) FG_CUDA(cudaMemcpyAsync(d2, d_d2, ns * 4, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FGOICP_OK;
}

// Number of instance slots a batch runs on (FGOICP_ICP_SLOTS overrides, 1..256).
static int icp_slots(const fgoicp_ctx* c)
{
    int s = ICP_MAX_BATCH;
    if (const char* e = getenv("FGOICP_ICP_SLOTS")) s = std::min(256, std::max(1, atoi(e)));
    while (s > 1 && (size_t)s * c->ns * 41 > ((size_t)1 << 30)) s /= 2;         // 41 bytes per slot and data point
    return s;
}

int fg_ensure_icp_jobs(fgoicp_ctx* c, int n)
{
    size_t need = sizeof(IcpQueue) + (size_t)n * (12 * sizeof(float) + sizeof(IcpResult));
    if (need <= c->icp_jobs_bytes) return FGOICP_OK;
    FG_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_icp_jobs); c->d_icp_jobs = nullptr; c->icp_jobs_bytes = 0;
    need = std::max(need * 2, (size_t)1 << 20);
    FG_CUDA(cudaMalloc(&c->d_icp_jobs, need));
    c->icp_jobs_bytes = need;
    return FGOICP_OK;
}

// Runs n independent ICPs to completion on a pool of instance slots (instance = blockIdx.y / blockIdx.x of every
// kernel); finished slots pull the next pending job on the device (k_icp_next).
// R0s[n][9], t0s[n][3] -> sse[n], R[n][9], t[n][3], iters[n].  Used by fgoicp_icp and by the level driver.
int fg_icp_run_batch(fgoicp_ctx* c, const float* R0s, const float* t0s, int n, int max_iter, float thr,
                     float* sse, float* R, float* t, int* iters)
{
    FG_RANGE("fgoicp icp batch");
    if (n <= 0) return FGOICP_OK;
    const int S = std::min(n, icp_slots(c));
    int rc = fg_ensure_icp_capacity(c, S);
    if (rc) return rc;
    rc = fg_ensure_icp_jobs(c, n);
    if (rc) return rc;
    rc = fg::ensure_pinned(c, (size_t)n * (12 * sizeof(float) + sizeof(IcpResult)) + 8192);
    if (rc) return rc;
    // device layout: queue | seeds[n][12] | results[n]
    IcpQueue* d_q = (IcpQueue*)c->d_icp_jobs;
    float* d_seeds = (float*)((char*)c->d_icp_jobs + sizeof(IcpQueue));
    IcpResult* d_res = (IcpResult*)(d_seeds + 12 * (size_t)n);
    IcpQueue* hq = (IcpQueue*)c->h_pinned;
    float* hseeds = (float*)((char*)c->h_pinned + 4096);
    for (int k = 0; k < n; ++k)
    {
        memcpy(hseeds + 12 * (size_t)k, R0s + 9 * (size_t)k, 9 * sizeof(float));
        memcpy(hseeds + 12 * (size_t)k + 9, t0s + 3 * (size_t)k, 3 * sizeof(float));
    }
    hq->n_jobs = n; hq->next = S; hq->finished = 0; hq->pad = 0;
    FG_CUDA(cudaMemcpyAsync(d_q, hq, sizeof(IcpQueue), cudaMemcpyHostToDevice, c->stream));
    FG_CUDA(cudaMemcpyAsync(d_seeds, hseeds, 12 * sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));

    char* inst = (char*)c->d_icp;
    const int ns = (int)c->ns;
    const long long guard_max = ((long long)(n + S - 1) / S + 1) * ((long long)max_iter + 2) + 8;
    bool all_done = false, results_on_host = false;
    const bool trimmed = c->trim_k > 0 && c->trim_k < c->ns;
    // Which driver: the persistent loop kernel (one launch, one synchronisation, nothing for the host to do) for
    // batches that fit the slot pool -- the latency-bound refinements of the coarse levels and of run()'s first and last
    // ICP -- and the launch chain for batches that refill their slots many times (hundreds to thousands of refinements on
    // partial-overlap scans): there the searches are pure throughput, the chain's one-warp-per-query launches run them
    // ~10 % faster (36 instead of 32 resident warps per SM, no barrier at the end of every stage) and its host polls are
    // amortised.  Same results either way, bit for bit (tests/test_gpu_parity.py).
    const bool use_loop = (c->icp_mode == 2 || (c->icp_mode == 0 && n <= S)) && !trimmed && c->nn_mode != 1;
    if (use_loop)
    {
        // ---- persistent loop kernel: one cooperative launch, one synchronisation for the whole batch
        // block shape of the loop kernel (FGOICP_ICP_SHAPE selects among the instantiations for experiments)
        struct Shape { const char* name; const void* fn; int threads, minb; };
        static const Shape shapes[] = {
            { "512x2", (const void*)k_icp_loop<512, 2, 0>, 512, 2 },
            { "512x2o", (const void*)k_icp_loop<512, 2, 1>, 512, 2 },
            { "256x5o", (const void*)k_icp_loop<256, 5, 1>, 256, 5 },
            { "256x4", (const void*)k_icp_loop<256, 4, 0>, 256, 4 },
            { "512x1", (const void*)k_icp_loop<512, 1, 0>, 512, 1 },
            { "384x3", (const void*)k_icp_loop<384, 3, 0>, 384, 3 },
            { "640x2", (const void*)k_icp_loop<640, 2, 0>, 640, 2 },
        };
        static const int shape_idx = []() {
            const char* e = getenv("FGOICP_ICP_SHAPE");
            if (e) for (int k = 0; k < (int)(sizeof(shapes) / sizeof(shapes[0])); ++k) if (!strcmp(e, shapes[k].name)) return k;
            return 0;
        }();
        const Shape& shape = shapes[shape_idx];
        const int ICPL_THREADS = shape.threads;
        if (c->icp_loop_grid == 0)
        {
            int occ = 0;
            FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, shape.fn, shape.threads, 0));
            if (occ < 1) { fg::set_error("k_icp_loop does not fit on this device"); return FGOICP_ERR_STATE; }
            c->icp_loop_grid = c->sm_count * std::min(occ, shape.minb);
            if (const char* e = getenv("FGOICP_ICP_GRID")) c->icp_loop_grid = std::max(1, std::min(c->icp_loop_grid, atoi(e)));
        }
        static const float margin_cfg = getenv("FGOICP_NN_MARGIN") ? (float)atof(getenv("FGOICP_NN_MARGIN")) : FG_NN_MARGIN;
        IcpLoopArgs a;
        a.g.start = c->d_cell_start; a.g.pts = c->d_cell_M;
        a.g.nx = c->cnx; a.g.ny = c->cny; a.g.nz = c->cnz; a.g.h = c->cell_h; a.g.inv_h = c->cell_inv_h;
        fg_set_coarse(c, a.g);
        a.L = c->lut; a.res = c->res;
        a.data = c->d_data; a.model = c->d_model; a.work = c->d_work; a.keys = c->d_nnkey; a.memo = c->d_nnmemo;
        a.inst = inst; a.ns = ns; a.S = S;
        a.jobs = d_seeds; a.results = d_res; a.q = d_q; a.max_iter = max_iter; a.thr = thr;
        a.ctl = (IcpLoopCtl*)c->d_icp_loop;
        a.arrive = (unsigned int*)((char*)c->d_icp_loop + 1024);
        a.part = (double*)((char*)c->d_icp_loop + ICPL_HEAD);
        a.miss = (unsigned int*)((char*)c->d_icp_loop + ICPL_HEAD + icp_loop_part_bytes(c, c->icp_capacity));
        a.margin = getenv("FGOICP_NN_NO_WARM") ? 0.0f : std::max(0.0f, margin_cfg);
        a.guard_max = guard_max;
        // cooperative scans of heavy queries are OFF by default: measured neutral on W5 / W3 and -10 % on W4 at 2,500 rows
        // (scans are bound by instruction issue, not by the latency of single heavy ones); FGOICP_NN_HEAVY_ROWS=<n> enables
        static const int heavy_cfg = getenv("FGOICP_NN_HEAVY_ROWS") ? atoi(getenv("FGOICP_NN_HEAVY_ROWS")) : 0;
        a.heavy_rows = heavy_cfg > 0 ? heavy_cfg : 0x7fffffff;
        FG_CUDA(cudaMemsetAsync(a.ctl, 0, ICPL_HEAD, c->stream));
        // no more blocks than there is work for: a barrier costs time per participating block
        // (enough warps that a first pass -- every query a full scan -- hands each warp about four of them)
        const long long want = ((long long)S * ns * 16 + 63) / 64 / (ICPL_THREADS / 32);
        const int grid = (int)std::max<long long>(1, std::min<long long>(c->icp_loop_grid, want));
        void* params[] = { &a };
        FG_CUDA(cudaLaunchCooperativeKernel(shape.fn, dim3((unsigned)grid), dim3((unsigned)ICPL_THREADS), params, 0, c->stream));
        IcpLoopCtl* hctl = (IcpLoopCtl*)((char*)c->h_pinned + 2048);
        FG_CUDA(cudaMemcpyAsync(hq, d_q, sizeof(IcpQueue), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaMemcpyAsync(hctl, a.ctl, sizeof(IcpLoopCtl), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaMemcpyAsync((char*)c->h_pinned + 4096, d_res, sizeof(IcpResult) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        all_done = hq->finished >= n && hctl->error == 0;
        results_on_host = all_done;
        if (getenv("FGOICP_ICP_LOG"))
        {
            fprintf(stderr, "[icp loop %s] jobs %d slots %d grid %d trips %u full scans: rooted %u squared %u (heavy %u) | stage us:", shape.name, n, S, grid,
                    hctl->iterations, hctl->scans_a, hctl->scans_b, hctl->heavy);
            for (int k = 0; k < 9; ++k) fprintf(stderr, " %.0f", hctl->stage_ns[k] * 1e-3);
            fprintf(stderr, "\n");
        }
    }
    else
    {
    dim3 pgrid((unsigned)((ns + 255) / 256), (unsigned)S);
    k_icp_assign<<<S, 1, 0, c->stream>>>(inst, d_seeds, max_iter, thr);
    FG_CUDA(cudaGetLastError());
    // iterations enqueued between polls of the queue (finished slots with no job left cost only empty launches)
    const int burst = 8;
    for (long long guard = 0; guard <= guard_max + burst && !all_done; guard += burst)
    {
        for (int b = 0; b < burst; ++b)
        {
            k_icp_prepare<<<pgrid, 256, 0, c->stream>>>(c->d_data, c->d_work, c->d_nnkey, ns, inst);             // icp3d.cu:85
            // one loop body per slot still running; every kernel returns at once for idle ones
            rc = enqueue_nn(c, S, SRC_WORK, POSE_NONE, 1, 1);                                        // icp3d.cu:146
            if (rc) return rc;
            const unsigned char* inl = nullptr;
            int n_in = ns;
            if (c->trim_k > 0 && c->trim_k < c->ns)
            {
                // trimmed registration: the Procrustes step sees the trim_k closest correspondences only
                k_icp_select<<<S, 1024, 0, c->stream>>>(c->d_nnkey, ns, inst, (unsigned int)c->trim_k, c->d_inl, c->d_data_orig);
                inl = c->d_inl; n_in = (int)c->trim_k;
            }
            double* part = (double*)c->d_icp_part;
            unsigned int* counters = (unsigned int*)((char*)c->d_icp_part + sizeof(double) * ICP_NB * ICP_PART * (size_t)c->icp_capacity);
            dim3 rgrid(ICP_NB, (unsigned)S);
            k_icp_centroids<<<rgrid, ICP_BT, 0, c->stream>>>(c->d_work, c->d_nnkey, c->d_model, ns, inst, inl, n_in, part, counters);
            k_icp_procrustes<<<rgrid, ICP_BT, 0, c->stream>>>(c->d_work, c->d_nnkey, c->d_model, ns, inst, inl, part, counters);
            k_icp_transform<<<pgrid, 256, 0, c->stream>>>(c->d_data, c->d_work, ns, inst, SRC_WORK, POSE_INC, 1);  // icp3d.cu:100
            rc = enqueue_nn(c, S, SRC_DATA, POSE_CUR, 0, 1);                                         // icp3d.cu:103
            if (rc) return rc;
            k_sse_reduce<<<S, 1024, 0, c->stream>>>(c->d_nnkey, ns, inst, 1, nullptr, (unsigned int)c->trim_k);
            k_icp_next<<<S, 1, 0, c->stream>>>(inst, d_seeds, d_res, d_q, max_iter, thr);   // loop head of the next iteration / next job
            FG_CUDA(cudaGetLastError());
        }
        FG_CUDA(cudaMemcpyAsync(hq, d_q, sizeof(IcpQueue), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        all_done = hq->finished >= n;
    }
    }
    if (!all_done) { fg::set_error("ICP loop did not terminate"); return FGOICP_ERR_STATE; }
    IcpResult* hres = (IcpResult*)((char*)c->h_pinned + 4096);
    if (!results_on_host)
    {
        FG_CUDA(cudaMemcpyAsync(hres, d_res, sizeof(IcpResult) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (int k = 0; k < n; ++k)
    {
        if (sse) sse[k] = hres[k].sse;
        if (R) memcpy(R + 9 * (size_t)k, hres[k].R, 9 * sizeof(float));
        if (t) memcpy(t + 3 * (size_t)k, hres[k].t, 3 * sizeof(float));
        if (iters) iters[k] = hres[k].iters;
    }
    return FGOICP_OK;
}

// Everything run() needs for its refinements, allocated once at context creation so that the timed search allocates
// nothing (first-ICP latency used to grow with the number of processes contending in the allocator).
int fg_icp_prealloc(fgoicp_ctx* c)
{
    int rc = fg_ensure_icp_capacity(c, icp_slots(c));
    if (rc) return rc;
    rc = fg_ensure_icp_jobs(c, 4096);
    if (rc) return rc;
    return fg::ensure_pinned(c, (size_t)4096 * (12 * sizeof(float) + sizeof(IcpResult)) + 8192);
}

extern "C" int fgoicp_set_icp_mode(fgoicp_ctx* c, int mode)
{
    FG_ARG(c, "NULL context");
    FG_ARG(mode >= 0 && mode <= 2, "icp mode must be 0 (automatic), 1 (launch chain) or 2 (persistent loop kernel)");
    c->icp_mode = mode;
    return FGOICP_OK;
}

int fg_icp_run(fgoicp_ctx* c, const float R0[9], const float t0[3], int max_iter, float thr,
               float* sse, float R[9], float t[3], int* iters)
{
    return fg_icp_run_batch(c, R0, t0, 1, max_iter, thr, sse, R, t, iters);
}

extern "C" int fgoicp_icp(fgoicp_ctx* c, const float R0[9], const float t0[3], int max_iter, float thr,
                          float* sse, float R[9], float t[3], int* iters)
{
    FG_ARG(c && R0 && t0, "NULL pointer");
    FG_ARG(max_iter >= 0, "max_iter must be non-negative");
    FG_CUDA(cudaSetDevice(c->device));
    return fg_icp_run_batch(c, R0, t0, 1, max_iter, thr, sse, R, t, iters);
}

extern "C" int fgoicp_icp_batch(fgoicp_ctx* c, const float* R0s, const float* t0s, int n, int max_iter, float thr,
                                float* sse, float* R, float* t, int* iters)
{
    FG_ARG(c, "NULL context");
    FG_ARG(n >= 0, "n must be non-negative");
    FG_ARG(max_iter >= 0, "max_iter must be non-negative");
    if (n == 0) return FGOICP_OK;
    FG_ARG(R0s && t0s, "NULL pointer");
    FG_CUDA(cudaSetDevice(c->device));
    return fg_icp_run_batch(c, R0s, t0s, n, max_iter, thr, sse, R, t, iters);
}
