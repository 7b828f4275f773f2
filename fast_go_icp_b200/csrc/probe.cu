// probe.cu -- measurement hooks (not on the product path): raw scattered-gather bandwidth of this GPU.
//
// The bound kernels are gather-bound: each evaluation pulls one 32-byte corner-packed cell from a
// grid much larger than L2, at an address that is effectively random.  HBM cannot serve random
// 32-byte sectors at its streaming rate, so the honest ceiling for those kernels is the rate at which
// the memory system serves *independent random gathers* of the same width.  This kernel measures it:
// every thread issues `per_thread` gathers of `width` bytes at pseudo-random aligned offsets of a
// buffer of `bytes` bytes, with `ilp` independent loads in flight, and folds the data into a checksum.
#include "common.cuh"

__device__ __forceinline__ unsigned fg_hash(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int WIDTH_F4>   // gather width in 16-byte units: 1, 2 (one 256-bit load), 4, 8
__global__ void __launch_bounds__(256)
k_gather_probe(const float4* __restrict__ buf, unsigned long long n_slots, int per_thread, float* __restrict__ sink)
{
    unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < per_thread; it += 4)
    {
        float part[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            unsigned long long r = ((unsigned long long)fg_hash(tid * 977u + (unsigned)(it + k) * 0x9e3779b9u) << 16) ^ fg_hash(tid + 7919u * (unsigned)(it + k));
            unsigned long long slot = r % n_slots;
            const float4* p = buf + slot * WIDTH_F4;
            float s = 0.f;
            if (WIDTH_F4 == 2)
            {
                float v[8];
                fg_ld256((const float*)p, v);
                s = v[0] + v[7];
            }
            else
            {
#pragma unroll
                for (int j = 0; j < WIDTH_F4; ++j) { float4 v = __ldg(p + j); s += v.x + v.w; }
            }
            part[k] = s;
        }
        acc += part[0] + part[1] + part[2] + part[3];
    }
    if (acc == 123.456f) sink[0] = acc;     // keeps the loads alive
}

// Returns useful gathered GB/s for gathers of `width_bytes` (16, 32, 64, 128) over the first `bytes` of
// a scratch buffer (allocated here).  Timed with CUDA events on the context stream.
extern "C" int fgoicp_gather_probe(fgoicp_ctx* c, size_t bytes, int width_bytes, int blocks_per_sm, float* out_gbps)
{
    FG_ARG(c && out_gbps, "NULL pointer");
    FG_ARG(width_bytes == 16 || width_bytes == 32 || width_bytes == 64 || width_bytes == 128, "width must be 16/32/64/128");
    FG_CUDA(cudaSetDevice(c->device));
    float4* buf = nullptr;
    float* sink = nullptr;
    FG_CUDA(cudaMalloc(&buf, bytes));
    FG_CUDA(cudaMalloc(&sink, 16));
    FG_CUDA(cudaMemsetAsync(buf, 0, bytes, c->stream));
    unsigned long long n_slots = bytes / (size_t)width_bytes;
    int per_thread = 256;
    int blocks = c->sm_count * blocks_per_sm * 8;
    cudaEvent_t e0 = c->ev0, e1 = c->ev1;
    for (int rep = 0; rep < 2; ++rep)
    {
        if (rep == 1) FG_CUDA(cudaEventRecord(e0, c->stream));
        switch (width_bytes)
        {
        case 16:  k_gather_probe<1><<<blocks, 256, 0, c->stream>>>(buf, n_slots, per_thread, sink); break;
        case 32:  k_gather_probe<2><<<blocks, 256, 0, c->stream>>>(buf, n_slots, per_thread, sink); break;
        case 64:  k_gather_probe<4><<<blocks, 256, 0, c->stream>>>(buf, n_slots, per_thread, sink); break;
        default:  k_gather_probe<8><<<blocks, 256, 0, c->stream>>>(buf, n_slots, per_thread, sink); break;
        }
    }
    FG_CUDA(cudaEventRecord(e1, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    FG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double total = (double)blocks * 256.0 * per_thread * width_bytes;
    *out_gbps = (float)(total / (ms * 1e-3) / 1e9);
    cudaFree(buf); cudaFree(sink);
    return FGOICP_OK;
}
