// preprocess.cu -- FastGoICP constructor preprocessing on the device.
//
// Replaces the host passes of the reference constructor (fgoicp/fgoicp.hpp:13-19):
//   center_point_cloud(pcs), center_point_cloud(pct)    fgoicp/fgoicp.cpp:176-195
//   scale_point_clouds(pct, pcs) / get_scaling_factor   fgoicp/fgoicp.cpp:197-220, 271-287
//   get_point_cloud_ranges(pct)                          fgoicp/fgoicp.cpp:222-268
//
// Everything here is order-independent (max, min, element-wise subtract and multiply) EXCEPT the centroid, which
// the reference accumulates as a serial fp32 sum in index order.  The default mode reproduces that sum bit for bit:
// one block per cloud streams 1024-point tiles through shared memory (all threads load the next tile while three
// lanes -- one per coordinate -- run the dependent fp32 add chains over the current one).  The chain costs one FADD
// latency (~4 cycles) per point.  FGOICP_PRE_TREE_CENTROID swaps it for a deterministic
// parallel fp64 reduction (fixed partition, fixed combination order) for clouds of millions of points; the centroid
// then differs from the reference's in the last fp32 bits.
#include "common.cuh"

#include <math.h>
#include <string.h>
#include <vector>

namespace
{
    constexpr int TILE_PTS = 1024;              // points per shared-memory tile
    constexpr int TILE_F = TILE_PTS * 3;        // floats per tile
    constexpr int SEQ_THREADS = 256;
    constexpr int F_PER_THREAD = TILE_F / SEQ_THREADS;     // 12
    constexpr int TREE_BLOCKS = 296;            // 2 per SM; partition is fixed so the result does not depend on the GPU
    constexpr int TREE_THREADS = 256;

    struct PreDev                               // device-side result block (one per call)
    {
        float centroid[2][3];                   // [0] = source (data), [1] = target (model)
        unsigned absmax_bits[2];                // max |coord| after centring, as IEEE bits (non-negative => ordered)
        unsigned bbox_min_key[3], bbox_max_key[3];   // order-preserving keys of the target's scaled min / max
        float scale;
    };

    __host__ __device__ inline unsigned fkey(float v)
    {
        unsigned b;
#ifdef __CUDA_ARCH__
        b = __float_as_uint(v);
#else
        memcpy(&b, &v, 4);
#endif
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    inline float funkey(unsigned k)
    {
        unsigned b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
        float v; memcpy(&v, &b, 4);
        return v;
    }

    // Serial fp32 sum in index order (fgoicp.cpp:180-185), then centroid /= n (fgoicp.cpp:186).
    // blockIdx.x selects the cloud.  The tile is stored per coordinate (x[], y[], z[]) so that lane a < 3 reads four
    // consecutive values of coordinate a with one 128-bit shared-memory load; loads run 16 adds ahead of their use
    // in two alternating register sets, so the chain advances at one FADD latency per point.
    constexpr int SOA_STRIDE = TILE_PTS + 44;   // multiple of 4 (float4 alignment), 12 mod 32 (bank spread), >= 16 floats of slack for the read-ahead

    __device__ __forceinline__ float chain4(float s, const float4& v)
    {
        s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
        return s;
    }

    __global__ void __launch_bounds__(SEQ_THREADS) k_pre_centroid_seq(const float* __restrict__ pts0, size_t n0,
                                                                      const float* __restrict__ pts1, size_t n1, PreDev* out)
    {
        __shared__ __align__(16) float tile[2][3 * SOA_STRIDE];
        const float* pts = blockIdx.x == 0 ? pts0 : pts1;
        const size_t n = blockIdx.x == 0 ? n0 : n1;
        const size_t nf = n * 3;
        const int tid = threadIdx.x;
        float reg[F_PER_THREAD];
        float s = 0.0f;                                     // lanes 0..2 of warp 0: running sum of coordinate `tid`

        auto load = [&](size_t base)
        {
#pragma unroll
            for (int k = 0; k < F_PER_THREAD; ++k)
            {
                size_t i = base + (size_t)k * SEQ_THREADS + tid;
                reg[k] = i < nf ? pts[i] : 0.0f;
            }
        };
        auto stash = [&](int buf)
        {
#pragma unroll
            for (int k = 0; k < F_PER_THREAD; ++k)
            {
                int li = k * SEQ_THREADS + tid;             // flat index inside the tile: point li / 3, coordinate li % 3
                tile[buf][(li % 3) * SOA_STRIDE + li / 3] = reg[k];
            }
        };

        const size_t ntiles = (nf + TILE_F - 1) / TILE_F;
        if (ntiles > 0) { load(0); stash(0); }
        __syncthreads();
        for (size_t t = 0; t < ntiles; ++t)
        {
            const int cur = (int)(t & 1);
            if (t + 1 < ntiles) load((t + 1) * TILE_F);     // global loads in flight while the chain runs
            if (tid < 3)
            {
                size_t left = n - t * TILE_PTS;
                const float* p = &tile[cur][tid * SOA_STRIDE];
                if (left >= (size_t)TILE_PTS)
                {
                    const float4* q = reinterpret_cast<const float4*>(p);
                    float4 a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3];
#pragma unroll 1
                    for (int g = 0; g < TILE_PTS / 32; ++g)
                    {
                        const float4 b0 = q[8 * g + 4], b1 = q[8 * g + 5], b2 = q[8 * g + 6], b3 = q[8 * g + 7];
                        s = chain4(chain4(chain4(chain4(s, a0), a1), a2), a3);
                        a0 = q[8 * g + 8]; a1 = q[8 * g + 9]; a2 = q[8 * g + 10]; a3 = q[8 * g + 11];   // last pass reads the slack
                        s = chain4(chain4(chain4(chain4(s, b0), b1), b2), b3);
                    }
                }
                else
                {
                    const int cnt = (int)left;
                    int j = 0;
                    for (; j + 4 <= cnt; j += 4) s = chain4(s, *reinterpret_cast<const float4*>(p + j));
                    for (; j < cnt; ++j) s = __fadd_rn(s, p[j]);
                }
            }
            if (t + 1 < ntiles) stash(cur ^ 1);
            __syncthreads();
        }
        if (tid < 3) out->centroid[blockIdx.x][tid] = __fdiv_rn(s, (float)n);
    }

    // Deterministic parallel alternative: fp64 partial sums over a fixed partition.
    __global__ void __launch_bounds__(TREE_THREADS) k_pre_centroid_partial(const float* __restrict__ pts, size_t n, double* partial)
    {
        __shared__ double sh[3][TREE_THREADS];
        const size_t per = (n + TREE_BLOCKS - 1) / TREE_BLOCKS;
        const size_t lo = (size_t)blockIdx.x * per;
        const size_t hi = lo + per < n ? lo + per : n;
        double a[3] = { 0.0, 0.0, 0.0 };
        for (size_t i = lo + threadIdx.x; i < hi; i += TREE_THREADS)
        {
            a[0] += (double)pts[3 * i]; a[1] += (double)pts[3 * i + 1]; a[2] += (double)pts[3 * i + 2];
        }
        for (int c = 0; c < 3; ++c) sh[c][threadIdx.x] = a[c];
        __syncthreads();
        for (int w = TREE_THREADS / 2; w > 0; w >>= 1)
        {
            if (threadIdx.x < w)
                for (int c = 0; c < 3; ++c) sh[c][threadIdx.x] += sh[c][threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x < 3) partial[(size_t)blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
    }
    __global__ void k_pre_centroid_final(const double* partial, size_t n, int which, PreDev* out)
    {
        if (threadIdx.x < 3)
        {
            double s = 0.0;
            for (int b = 0; b < TREE_BLOCKS; ++b) s += partial[(size_t)b * 3 + threadIdx.x];
            out->centroid[which][threadIdx.x] = (float)(s / (double)n);
        }
    }

    // pc[i] -= centroid (fgoicp.cpp:188-192) and max |coord| of the centred cloud (fgoicp.cpp:200-217)
    __global__ void __launch_bounds__(256) k_pre_center(float* __restrict__ pts, size_t nf, int which, PreDev* out)
    {
        const float c0 = out->centroid[which][0], c1 = out->centroid[which][1], c2 = out->centroid[which][2];
        float m = 0.0f;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (size_t)gridDim.x * blockDim.x)
        {
            int a = (int)(i % 3);
            float v = __fsub_rn(pts[i], a == 0 ? c0 : (a == 1 ? c1 : c2));
            pts[i] = v;
            m = fmaxf(m, fabsf(v));
        }
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) atomicMax(&out->absmax_bits[which], __float_as_uint(m));
    }

    // s = 1 / max|coord| (fgoicp.cpp:218); the source's alone (reference) or over both clouds (FGOICP_PRE_SCALE_BOTH)
    __global__ void k_pre_scale_factor(PreDev* out, int both)
    {
        float m = __uint_as_float(out->absmax_bits[0]);
        if (both) m = fmaxf(m, __uint_as_float(out->absmax_bits[1]));
        out->scale = __fdiv_rn(1.0f, m);
    }

    // pc[i] *= s (fgoicp.cpp:271-287) and, for the target, its per-axis range (fgoicp.cpp:222-268)
    __global__ void __launch_bounds__(256) k_pre_scale(float* __restrict__ pts, size_t nf, int want_bbox, PreDev* out)
    {
        const float s = out->scale;
        unsigned lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u };
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (size_t)gridDim.x * blockDim.x)
        {
            float v = __fmul_rn(pts[i], s);
            pts[i] = v;
            if (want_bbox)
            {
                int a = (int)(i % 3);
                unsigned k = fkey(v);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (a == c) { lo[c] = min(lo[c], k); hi[c] = max(hi[c], k); }
            }
        }
        if (want_bbox)
        {
#pragma unroll
            for (int c = 0; c < 3; ++c)
            {
                unsigned l = __reduce_min_sync(0xffffffffu, lo[c]);
                unsigned h = __reduce_max_sync(0xffffffffu, hi[c]);
                if ((threadIdx.x & 31) == 0)
                {
                    atomicMin(&out->bbox_min_key[c], l);
                    atomicMax(&out->bbox_max_key[c], h);
                }
            }
        }
    }

    int select_device(int device)
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
        {
            fg::set_error("no CUDA device available: this library has no CPU fallback");
            return FGOICP_ERR_CUDA;
        }
        FG_ARG(device >= 0 && device < ndev, "device index out of range");
        FG_CUDA(cudaSetDevice(device));
        return FGOICP_OK;
    }

    // Enqueues the whole preprocessing on `st`; d_out must be zero/identity-initialised by the caller (init_out).
    int enqueue(float* d_model, size_t nt, float* d_data, size_t ns, unsigned flags, cudaStream_t st,
                PreDev* d_out, double* d_partial)
    {
        PreDev init;
        memset(&init, 0, sizeof(init));
        for (int c = 0; c < 3; ++c) { init.bbox_min_key[c] = 0xFFFFFFFFu; init.bbox_max_key[c] = 0u; }
        FG_CUDA(cudaMemcpyAsync(d_out, &init, sizeof(init), cudaMemcpyHostToDevice, st));
        if (flags & FGOICP_PRE_TREE_CENTROID)
        {
            k_pre_centroid_partial<<<TREE_BLOCKS, TREE_THREADS, 0, st>>>(d_data, ns, d_partial);
            k_pre_centroid_final<<<1, 32, 0, st>>>(d_partial, ns, 0, d_out);
            k_pre_centroid_partial<<<TREE_BLOCKS, TREE_THREADS, 0, st>>>(d_model, nt, d_partial);
            k_pre_centroid_final<<<1, 32, 0, st>>>(d_partial, nt, 1, d_out);
        }
        else
            k_pre_centroid_seq<<<2, SEQ_THREADS, 0, st>>>(d_data, ns, d_model, nt, d_out);
        auto blocks = [](size_t nf) { size_t b = (nf + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b)); };
        k_pre_center<<<blocks(ns * 3), 256, 0, st>>>(d_data, ns * 3, 0, d_out);
        k_pre_center<<<blocks(nt * 3), 256, 0, st>>>(d_model, nt * 3, 1, d_out);
        k_pre_scale_factor<<<1, 1, 0, st>>>(d_out, (flags & FGOICP_PRE_SCALE_BOTH) ? 1 : 0);
        k_pre_scale<<<blocks(ns * 3), 256, 0, st>>>(d_data, ns * 3, 0, d_out);
        k_pre_scale<<<blocks(nt * 3), 256, 0, st>>>(d_model, nt * 3, 1, d_out);
        FG_CUDA(cudaGetLastError());
        return FGOICP_OK;
    }

    void publish(const PreDev& r, float ms, fgoicp_normalisation* out)
    {
        for (int a = 0; a < 3; ++a)
        {
            out->offset_pcs[a] = -r.centroid[0][a];          // center_point_cloud returns -centroid (fgoicp.cpp:194)
            out->offset_pct[a] = -r.centroid[1][a];
            out->bbox_min[a] = funkey(r.bbox_min_key[a]);
            out->bbox_max[a] = funkey(r.bbox_max_key[a]);
        }
        out->scale = r.scale;
        out->device_ms = ms;
    }

    int check_args(const float* model, size_t nt, const float* data, size_t ns, unsigned flags, const fgoicp_normalisation* out)
    {
        FG_ARG(model && data && out, "NULL pointer");
        FG_ARG(nt > 0 && ns > 0, "empty point cloud");
        FG_ARG(nt < (size_t)1 << 31 && ns < (size_t)1 << 31, "point cloud too large");
        FG_ARG((flags & ~(FGOICP_PRE_TREE_CENTROID | FGOICP_PRE_SCALE_BOTH)) == 0, "unknown preprocessing flag");
        return FGOICP_OK;
    }
}

extern "C" int fgoicp_preprocess_dev(float* d_model_xyz, size_t nt, float* d_data_xyz, size_t ns,
                                     int device, unsigned flags, void* cuda_stream, fgoicp_normalisation* out)
{
    FG_RANGE("fgoicp_preprocess_dev");
    int rc = check_args(d_model_xyz, nt, d_data_xyz, ns, flags, out);
    if (rc) return rc;
    rc = select_device(device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    PreDev* d_out = nullptr;
    double* d_partial = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() { if (d_out) cudaFree(d_out); if (d_partial) cudaFree(d_partial); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
#define FG_PTRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fg::cuda_fail(e__, #call, __FILE__, __LINE__); } } while (0)
    FG_PTRY(cudaMalloc(&d_out, sizeof(PreDev)));
    FG_PTRY(cudaMalloc(&d_partial, sizeof(double) * 3 * TREE_BLOCKS));
    FG_PTRY(cudaEventCreate(&e0));
    FG_PTRY(cudaEventCreate(&e1));
    FG_PTRY(cudaEventRecord(e0, st));
    rc = enqueue(d_model_xyz, nt, d_data_xyz, ns, flags, st, d_out, d_partial);
    if (rc) { cleanup(); return rc; }
    FG_PTRY(cudaEventRecord(e1, st));
    PreDev r;
    FG_PTRY(cudaMemcpyAsync(&r, d_out, sizeof(r), cudaMemcpyDeviceToHost, st));
    FG_PTRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    FG_PTRY(cudaEventElapsedTime(&ms, e0, e1));
    publish(r, ms, out);
    cleanup();
    return FGOICP_OK;
}

extern "C" int fgoicp_preprocess(float* model_xyz, size_t nt, float* data_xyz, size_t ns,
                                 int device, unsigned flags, fgoicp_normalisation* out)
{
    FG_RANGE("fgoicp_preprocess");
    int rc = check_args(model_xyz, nt, data_xyz, ns, flags, out);
    if (rc) return rc;
    rc = select_device(device);
    if (rc) return rc;
    float* d_m = nullptr; float* d_d = nullptr;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() { if (d_m) cudaFree(d_m); if (d_d) cudaFree(d_d); if (st) cudaStreamDestroy(st); };
    FG_PTRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    FG_PTRY(cudaMalloc(&d_m, sizeof(float) * 3 * nt));
    FG_PTRY(cudaMalloc(&d_d, sizeof(float) * 3 * ns));
    FG_PTRY(cudaMemcpyAsync(d_m, model_xyz, sizeof(float) * 3 * nt, cudaMemcpyHostToDevice, st));
    FG_PTRY(cudaMemcpyAsync(d_d, data_xyz, sizeof(float) * 3 * ns, cudaMemcpyHostToDevice, st));
    rc = fgoicp_preprocess_dev(d_m, nt, d_d, ns, device, flags, st, out);
    if (rc) { cleanup(); return rc; }
    FG_PTRY(cudaMemcpyAsync(model_xyz, d_m, sizeof(float) * 3 * nt, cudaMemcpyDeviceToHost, st));
    FG_PTRY(cudaMemcpyAsync(data_xyz, d_d, sizeof(float) * 3 * ns, cudaMemcpyDeviceToHost, st));
    FG_PTRY(cudaStreamSynchronize(st));
    cleanup();
    return FGOICP_OK;
#undef FG_PTRY
}
