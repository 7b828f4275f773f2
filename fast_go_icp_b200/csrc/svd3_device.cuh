// svd3_device.cuh -- double-precision 3x3 SVD and closest rotation, usable on host and device.
//
// Replaces closest_orthogonal_approximation() of the reference (fgoicp/icp3d.cu:110-138), which
// round-trips the 3x3 cross-covariance to the host and calls Eigen::JacobiSVD<Matrix3d>.  Here a
// one-sided Jacobi SVD runs in a single device thread so the ICP loop never leaves the GPU.
// Every operation is an explicit IEEE round-to-nearest op (no FMA contraction), so host and
// device agree bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#if defined(__CUDA_ARCH__)
#define FG_DMUL(a, b) __dmul_rn((a), (b))
#define FG_DADD(a, b) __dadd_rn((a), (b))
#define FG_DSUB(a, b) __dsub_rn((a), (b))
#define FG_DDIV(a, b) __ddiv_rn((a), (b))
#define FG_DSQRT(a)   __dsqrt_rn((a))
#else
#define FG_DMUL(a, b) ((a) * (b))
#define FG_DADD(a, b) ((a) + (b))
#define FG_DSUB(a, b) ((a) - (b))
#define FG_DDIV(a, b) ((a) / (b))
#define FG_DSQRT(a)   sqrt((a))
#endif

__host__ __device__ inline double fg_dot3d(const double* a, const double* b)
{
    return FG_DADD(FG_DADD(FG_DMUL(a[0], b[0]), FG_DMUL(a[1], b[1])), FG_DMUL(a[2], b[2]));
}

__host__ __device__ inline void fg_cross3d(const double* a, const double* b, double* c)
{
    c[0] = FG_DSUB(FG_DMUL(a[1], b[2]), FG_DMUL(a[2], b[1]));
    c[1] = FG_DSUB(FG_DMUL(a[2], b[0]), FG_DMUL(a[0], b[2]));
    c[2] = FG_DSUB(FG_DMUL(a[0], b[1]), FG_DMUL(a[1], b[0]));
}

// A (row-major, A[r*3+c]) = U * diag(S) * V^T, singular values in decreasing order.
__host__ __device__ inline void fg_svd3(const double* A, double* U, double* S, double* V)
{
    double u[3][3], v[3][3];    // [col][row]
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
        {
            u[j][i] = A[i * 3 + j];
            v[j][i] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        bool rotated = false;
        // unrolled over the three (p, q) pairs so that u and v stay in registers on the device (same operations,
        // same order)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int p = 0; p < 2; ++p)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int q = p + 1; q < 3; ++q)
            {
                double alpha = fg_dot3d(u[p], u[p]);
                double beta = fg_dot3d(u[q], u[q]);
                double gamma = fg_dot3d(u[p], u[q]);
                if (gamma == 0.0 || fabs(gamma) <= FG_DMUL(1e-17, FG_DSQRT(FG_DMUL(alpha, beta)))) continue;
                double zeta = FG_DDIV(FG_DSUB(beta, alpha), FG_DMUL(2.0, gamma));
                double t = FG_DDIV(zeta >= 0.0 ? 1.0 : -1.0,
                                   FG_DADD(fabs(zeta), FG_DSQRT(FG_DADD(1.0, FG_DMUL(zeta, zeta)))));
                double c = FG_DDIV(1.0, FG_DSQRT(FG_DADD(1.0, FG_DMUL(t, t))));
                double s = FG_DMUL(c, t);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < 3; ++i)
                {
                    double up = u[p][i], uq = u[q][i], vp = v[p][i], vq = v[q][i];
                    u[p][i] = FG_DSUB(FG_DMUL(c, up), FG_DMUL(s, uq));
                    u[q][i] = FG_DADD(FG_DMUL(s, up), FG_DMUL(c, uq));
                    v[p][i] = FG_DSUB(FG_DMUL(c, vp), FG_DMUL(s, vq));
                    v[q][i] = FG_DADD(FG_DMUL(s, vp), FG_DMUL(c, vq));
                }
                rotated = true;
            }
        if (!rotated) break;
    }
    double sig[3];
    int order[3] = { 0, 1, 2 };
    for (int j = 0; j < 3; ++j) sig[j] = FG_DSQRT(fg_dot3d(u[j], u[j]));
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (sig[order[j]] < sig[order[j + 1]]) { int tmp = order[j]; order[j] = order[j + 1]; order[j + 1] = tmp; }
    double uu[3][3], vv[3][3];
    double tiny = FG_DADD(FG_DMUL(sig[order[0]], 1e-300), 1e-300);
    for (int j = 0; j < 3; ++j)
    {
        int src = order[j];
        S[j] = sig[src];
        for (int i = 0; i < 3; ++i)
        {
            vv[j][i] = v[src][i];
            uu[j][i] = (sig[src] > tiny) ? FG_DDIV(u[src][i], sig[src]) : 0.0;
        }
    }
    if (!(S[0] > tiny))
    {
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) uu[j][i] = (i == j) ? 1.0 : 0.0;
    }
    else
    {
        if (!(S[1] > tiny))
        {
            double e[3] = { 0.0, 0.0, 0.0 };
            int k = 0;
            if (fabs(uu[0][1]) < fabs(uu[0][k])) k = 1;
            if (fabs(uu[0][2]) < fabs(uu[0][k])) k = 2;
            e[k] = 1.0;
            fg_cross3d(uu[0], e, uu[1]);
            double n = FG_DSQRT(fg_dot3d(uu[1], uu[1]));
            for (int i = 0; i < 3; ++i) uu[1][i] = FG_DDIV(uu[1][i], n);
        }
        if (!(S[2] > tiny)) fg_cross3d(uu[0], uu[1], uu[2]);
    }
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
        {
            U[i * 3 + j] = uu[j][i];
            V[i * 3 + j] = vv[j][i];
        }
}

__host__ __device__ inline double fg_det3d(const double* M)
{
    double a = FG_DMUL(M[0], FG_DSUB(FG_DMUL(M[4], M[8]), FG_DMUL(M[5], M[7])));
    double b = FG_DMUL(M[1], FG_DSUB(FG_DMUL(M[3], M[8]), FG_DMUL(M[5], M[6])));
    double c = FG_DMUL(M[2], FG_DSUB(FG_DMUL(M[3], M[7]), FG_DMUL(M[4], M[6])));
    return FG_DADD(FG_DSUB(a, b), c);
}

// ABt: glm column-major float 3x3 (sum of outer products a b^T, stored [col][row]).
// Rout: column-major float.  R = V diag(1, 1, det(V U^T)) U^T with U S V^T = (math) sum a b^T.
__host__ __device__ inline void fg_closest_rotation(const float* ABt, float* Rout)
{
    double H[9], U[9], S[3], V[9], VUt[9], Rd[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) H[r * 3 + c] = (double)ABt[c * 3 + r];
    fg_svd3(H, U, S, V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc = FG_DADD(acc, FG_DMUL(V[i * 3 + k], U[j * 3 + k]));
            VUt[i * 3 + j] = acc;
        }
    double d = fg_det3d(VUt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k)
                acc = FG_DADD(acc, FG_DMUL(FG_DMUL(V[i * 3 + k], (k == 2 ? d : 1.0)), U[j * 3 + k]));
            Rd[i * 3 + j] = acc;
        }
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Rout[c * 3 + r] = (float)Rd[r * 3 + c];
}
