// TEST INFRASTRUCTURE (oracle/_ref build only).  Not part of the shipped product path.
//
// extern "C" facade over the UNMODIFIED reference sources (fgoicp/*.cu, *.cpp under
// /root/reference), compiled by oracle/build_ref.py into oracle/_ref/libfgoicp_ref.so against the
// GLM / Eigen stand-ins (include/glm, oracle/ref_shim/Eigen).  It lets GPU tests put the real
// reference kernels (real tex3D, real per-cube launches, real thrust reductions, real host BnB)
// next to our kernels and next to the CPU oracle on the same inputs.
//
// `private` is redefined for the reference headers only, to reach Registration / LUT internals
// without touching the reference sources; class layout is unaffected.
#include <cuda.h>
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>
#include <array>
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <iostream>
#include <queue>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#define private public
#include "fgoicp.hpp"
#include "icp3d.hpp"
#undef private

using namespace icp;

namespace
{
    glm::mat3 to_mat3(const float* a) { return glm::mat3(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); }
    void from_mat3(const glm::mat3& R, float* a) { for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) a[c * 3 + r] = R[c][r]; }

    __global__ void k_ref_sample(cudaTextureObject_t tex, float scale, float3 offset, const float* q, int n, float* out)
    {
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        // body of NearestNeighborLUT::search (reference registration.cu:320-328)
        float x = (q[3 * i] + offset.x) * scale;
        float y = (q[3 * i + 1] + offset.y) * scale;
        float z = (q[3 * i + 2] + offset.z) * scale;
        out[i] = tex3D<float>(tex, x, y, z);
    }

    // sin(half_angle) exactly as kernComputeBounds forms it per thread (reference registration.cu:41-42), with the
    // reference's own M_SQRT3 / M_PI (fgoicp/common.hpp:17-19) and this build's flags (same libdevice sin, no fast math):
    // pins the four constants the oracle uses without going through the library under test
    __global__ void k_ref_rot_sin(const float* spans, int n, float* out)
    {
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        float half_angle = spans[i] * M_SQRT3 * M_PI / 2.0f;
        out[i] = sin(half_angle);
    }

    // raw texel fetch through the same texture object: unnormalised coordinate i + 0.5 hits texel i exactly
    __global__ void k_ref_texels(cudaTextureObject_t tex, int dx, int dy, int dz, float* out)
    {
        size_t n = (size_t)dx * dy * dz;
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        int x = (int)(i % dx), y = (int)((i / dx) % dy), z = (int)(i / ((size_t)dx * dy));
        out[i] = tex3D<float>(tex, x + 0.5f, y + 0.5f, z + 0.5f);
    }
}

extern "C"
{
    // FastGoICP ctor: target first (reference fgoicp.hpp:13)
    void* ref_create(const float* model, size_t nt, const float* data, size_t ns, float res, float mse_threshold)
    {
        std::vector<glm::vec3> pct(nt), pcs(ns);
        for (size_t i = 0; i < nt; ++i) pct[i] = glm::vec3(model[3 * i], model[3 * i + 1], model[3 * i + 2]);
        for (size_t i = 0; i < ns; ++i) pcs[i] = glm::vec3(data[3 * i], data[3 * i + 1], data[3 * i + 2]);
        Logger::set_verbose(false);
        return new FastGoICP(std::move(pct), std::move(pcs), res, mse_threshold);
    }

    void ref_destroy(void* h) { delete static_cast<FastGoICP*>(h); }

    // normalised clouds and preprocessing constants as the reference computed them
    void ref_get_preprocessed(void* h, float* model, float* data, float* offset_pcs, float* offset_pct,
                              float* scale, float* bbox_min, float* bbox_max)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        memcpy(model, f->pct.data(), sizeof(float) * 3 * f->nt);
        memcpy(data, f->pcs.data(), sizeof(float) * 3 * f->ns);
        for (int a = 0; a < 3; ++a)
        {
            offset_pcs[a] = f->offset_pcs[a]; offset_pct[a] = f->offset_pct[a];
            bbox_min[a] = f->target_bounds[a].first; bbox_max[a] = f->target_bounds[a].second;
        }
        *scale = f->scaling_factor;
    }

    void ref_lut_dims(void* h, int* dims)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        dims[0] = f->registration.nnlut.dims.x; dims[1] = f->registration.nnlut.dims.y; dims[2] = f->registration.nnlut.dims.z;
    }

    int ref_lut_download(void* h, float* out)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        NearestNeighborLUT& L = f->registration.nnlut;
        size_t n = (size_t)L.dims.x * L.dims.y * L.dims.z;
        float* d = nullptr;
        if (cudaMalloc(&d, n * sizeof(float)) != cudaSuccess) return -1;
        k_ref_texels<<<(unsigned)((n + 255) / 256), 256>>>(L.texObj, L.dims.x, L.dims.y, L.dims.z, d);
        cudaError_t e = cudaMemcpy(out, d, n * sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(d);
        return e == cudaSuccess ? 0 : -1;
    }

    int ref_lut_sample(void* h, const float* q, int n, float* out)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        NearestNeighborLUT& L = f->registration.nnlut;
        float *dq = nullptr, *dout = nullptr;
        cudaMalloc(&dq, sizeof(float) * 3 * n); cudaMalloc(&dout, sizeof(float) * n);
        cudaMemcpy(dq, q, sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
        k_ref_sample<<<(n + 255) / 256, 256>>>(L.texObj, L.scale, L.offset, dq, n, dout);
        cudaError_t e = cudaMemcpy(out, dout, sizeof(float) * n, cudaMemcpyDeviceToHost);
        cudaFree(dq); cudaFree(dout);
        return e == cudaSuccess ? 0 : -1;
    }

    // Registration::compute_sse_error(rnode, tnodes, fix_rot, pool)  (reference registration.cu:88-152)
    void ref_bounds(void* h, const float* rot_xyz_span, int fix_rot, const float* tcubes, int T, float* lb, float* ub)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        RotNode rnode(rot_xyz_span[0], rot_xyz_span[1], rot_xyz_span[2], rot_xyz_span[3], 0.0f, 0.0f);
        std::vector<TransNode> tnodes;
        for (int i = 0; i < T; ++i) tnodes.emplace_back(tcubes[4 * i], tcubes[4 * i + 1], tcubes[4 * i + 2], tcubes[4 * i + 3], 0.0f, 0.0f);
        auto [l, u] = f->registration.compute_sse_error(rnode, tnodes, fix_rot != 0, f->stream_pool);
        for (int i = 0; i < T; ++i) { lb[i] = l[i]; ub[i] = u[i]; }
    }

    float ref_sse(void* h, const float* R, const float* t)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        return f->registration.compute_sse_error(to_mat3(R), glm::vec3(t[0], t[1], t[2]));
    }

    float ref_icp(void* h, const float* R0, const float* t0, int max_iter, float thr, float* R, float* t)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        IterativeClosestPoint3D icp3d(f->registration, f->pct, f->pcs, max_iter, thr, to_mat3(R0), glm::vec3(t0[0], t0[1], t0[2]));
        auto [sse, Ro, to] = icp3d.run();
        from_mat3(Ro, R); t[0] = to.x; t[1] = to.y; t[2] = to.z;
        return sse;
    }

    // FastGoICP::branch_and_bound_R3 (reference fgoicp.cpp:102-174) with best_sse set first
    float ref_bnb_r3(void* h, const float* rot_xyz_span, int fix_rot, float best_sse, float* best_t)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        float saved = f->best_sse;
        f->best_sse = best_sse;
        RotNode rnode(rot_xyz_span[0], rot_xyz_span[1], rot_xyz_span[2], rot_xyz_span[3], 0.0f, best_sse);
        auto [ub, bt] = f->branch_and_bound_R3(rnode, fix_rot != 0);
        f->best_sse = saved;
        best_t[0] = bt.x; best_t[1] = bt.y; best_t[2] = bt.z;
        return ub;
    }

    // FastGoICP::run (reference fgoicp.cpp:10-30): R, t in ORIGINAL coordinates; sse normalised
    float ref_run(void* h, float* R, float* t, float* R_norm, float* t_norm)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        std::streambuf* old = std::cout.rdbuf();
        std::ostringstream sink;
        std::cout.rdbuf(sink.rdbuf());          // the reference logs through std::cout
        auto [Ro, to] = f->run();
        std::cout.rdbuf(old);
        from_mat3(Ro, R); t[0] = to.x; t[1] = to.y; t[2] = to.z;
        auto [Rn, tn] = f->get_best_transform();
        from_mat3(Rn, R_norm); t_norm[0] = tn.x; t_norm[1] = tn.y; t_norm[2] = tn.z;
        return f->get_best_error();
    }

    // run() with the reference's Debug log captured: the sequence of best errors it prints after every refinement
    // ("New best error: ...", reference fgoicp.cpp:85; 6 significant digits) -- lets a parity report name the first
    // refinement at which another implementation's search parts ways.  Slower than ref_run (every cudaCheckError logs).
    float ref_run_trace(void* h, float* R, float* t, float* trace, int cap, int* n_trace)
    {
        FastGoICP* f = static_cast<FastGoICP*>(h);
        std::streambuf* old = std::cout.rdbuf();
        std::ostringstream sink;
        std::cout.rdbuf(sink.rdbuf());
        Logger::set_verbose(true);
        auto [Ro, to] = f->run();
        Logger::set_verbose(false);
        std::cout.rdbuf(old);
        from_mat3(Ro, R); t[0] = to.x; t[1] = to.y; t[2] = to.z;
        const std::string log = sink.str();
        const std::string key = "New best error: ";
        int n = 0;
        for (size_t pos = log.find(key); pos != std::string::npos; pos = log.find(key, pos + 1))
        {
            if (n < cap) trace[n] = (float)atof(log.c_str() + pos + key.size());
            ++n;
        }
        *n_trace = n;
        return f->get_best_error();
    }

    int ref_rot_sin(const float* spans, int n, float* out)
    {
        float *ds = nullptr, *dout = nullptr;
        cudaMalloc(&ds, sizeof(float) * n); cudaMalloc(&dout, sizeof(float) * n);
        cudaMemcpy(ds, spans, sizeof(float) * n, cudaMemcpyHostToDevice);
        k_ref_rot_sin<<<(n + 63) / 64, 64>>>(ds, n, dout);
        cudaError_t e = cudaMemcpy(out, dout, sizeof(float) * n, cudaMemcpyDeviceToHost);
        cudaFree(ds); cudaFree(dout);
        return e == cudaSuccess ? 0 : -1;
    }

    float ref_sse_threshold(void* h) { return static_cast<FastGoICP*>(h)->sse_threshold; }
}
