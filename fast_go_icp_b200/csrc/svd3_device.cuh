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

// One Jacobi rotation of columns (p, q) of u and v; returns whether it rotated.  Columns are passed as three
// scalars each so that everything stays in registers on the device (no indexed local arrays).
__host__ __device__ inline bool fg_jacobi_pair(double& up0, double& up1, double& up2, double& uq0, double& uq1, double& uq2,
                                               double& vp0, double& vp1, double& vp2, double& vq0, double& vq1, double& vq2)
{
    double alpha = FG_DADD(FG_DADD(FG_DMUL(up0, up0), FG_DMUL(up1, up1)), FG_DMUL(up2, up2));
    double beta = FG_DADD(FG_DADD(FG_DMUL(uq0, uq0), FG_DMUL(uq1, uq1)), FG_DMUL(uq2, uq2));
    double gamma = FG_DADD(FG_DADD(FG_DMUL(up0, uq0), FG_DMUL(up1, uq1)), FG_DMUL(up2, uq2));
    // converged when the two columns are orthogonal to within one unit roundoff (2^-52).  Round 1 tested against 1e-17,
    // which is below what double arithmetic can reach: the sweeps then kept rotating rounding noise until the 60-sweep cap
    // (38.6 sweeps on average against 4.4; the same float rotation matrix in all but 1 of 1.8e6 elements) -- and this
    // single thread is what every other block of the persistent ICP kernel waits for.
    if (gamma == 0.0 || fabs(gamma) <= FG_DMUL(2.220446049250313e-16, FG_DSQRT(FG_DMUL(alpha, beta)))) return false;
    double zeta = FG_DDIV(FG_DSUB(beta, alpha), FG_DMUL(2.0, gamma));
    double t = FG_DDIV(zeta >= 0.0 ? 1.0 : -1.0,
                       FG_DADD(fabs(zeta), FG_DSQRT(FG_DADD(1.0, FG_DMUL(zeta, zeta)))));
    double c = FG_DDIV(1.0, FG_DSQRT(FG_DADD(1.0, FG_DMUL(t, t))));
    double s = FG_DMUL(c, t);
#define FG_ROT2(P, Q) { double a_ = P, b_ = Q; P = FG_DSUB(FG_DMUL(c, a_), FG_DMUL(s, b_)); Q = FG_DADD(FG_DMUL(s, a_), FG_DMUL(c, b_)); }
    FG_ROT2(up0, uq0) FG_ROT2(vp0, vq0)
    FG_ROT2(up1, uq1) FG_ROT2(vp1, vq1)
    FG_ROT2(up2, uq2) FG_ROT2(vp2, vq2)
#undef FG_ROT2
    return true;
}

// A (row-major, A[r*3+c]) = U * diag(S) * V^T, singular values in decreasing order.
// One-sided Jacobi on the columns of A (u) accumulating V (v); same operations in the same order as the indexed
// formulation (pairs (0,1), (0,2), (1,2) per sweep; bubble sort of the singular values), written without indexed
// local arrays so that the device keeps u, v in registers.
__host__ __device__ inline void fg_svd3(const double* A, double* U, double* S, double* V)
{
    double u00 = A[0], u01 = A[3], u02 = A[6];      // column 0 of A: rows 0, 1, 2
    double u10 = A[1], u11 = A[4], u12 = A[7];      // column 1
    double u20 = A[2], u21 = A[5], u22 = A[8];      // column 2
    double v00 = 1.0, v01 = 0.0, v02 = 0.0, v10 = 0.0, v11 = 1.0, v12 = 0.0, v20 = 0.0, v21 = 0.0, v22 = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        bool r01 = fg_jacobi_pair(u00, u01, u02, u10, u11, u12, v00, v01, v02, v10, v11, v12);
        bool r02 = fg_jacobi_pair(u00, u01, u02, u20, u21, u22, v00, v01, v02, v20, v21, v22);
        bool r12 = fg_jacobi_pair(u10, u11, u12, u20, u21, u22, v10, v11, v12, v20, v21, v22);
        if (!(r01 || r02 || r12)) break;
    }
    double s0 = FG_DSQRT(FG_DADD(FG_DADD(FG_DMUL(u00, u00), FG_DMUL(u01, u01)), FG_DMUL(u02, u02)));
    double s1 = FG_DSQRT(FG_DADD(FG_DADD(FG_DMUL(u10, u10), FG_DMUL(u11, u11)), FG_DMUL(u12, u12)));
    double s2 = FG_DSQRT(FG_DADD(FG_DADD(FG_DMUL(u20, u20), FG_DMUL(u21, u21)), FG_DMUL(u22, u22)));
    // bubble sort, decreasing: (0,1), (1,2), (0,1); a swap moves the whole column triple (sigma, u, v)
#define FG_SWAPD(a, b) { double t_ = a; a = b; b = t_; }
#define FG_CSWAP(sa, ua0, ua1, ua2, va0, va1, va2, sb, ub0, ub1, ub2, vb0, vb1, vb2) \
    if (sa < sb) { FG_SWAPD(sa, sb) FG_SWAPD(ua0, ub0) FG_SWAPD(ua1, ub1) FG_SWAPD(ua2, ub2) FG_SWAPD(va0, vb0) FG_SWAPD(va1, vb1) FG_SWAPD(va2, vb2) }
    FG_CSWAP(s0, u00, u01, u02, v00, v01, v02, s1, u10, u11, u12, v10, v11, v12)
    FG_CSWAP(s1, u10, u11, u12, v10, v11, v12, s2, u20, u21, u22, v20, v21, v22)
    FG_CSWAP(s0, u00, u01, u02, v00, v01, v02, s1, u10, u11, u12, v10, v11, v12)
#undef FG_CSWAP
#undef FG_SWAPD
    double tiny = FG_DADD(FG_DMUL(s0, 1e-300), 1e-300);
    S[0] = s0; S[1] = s1; S[2] = s2;
    double uu[3][3], vv[3][3];                     // [col][row]; static indices only below
    vv[0][0] = v00; vv[0][1] = v01; vv[0][2] = v02;
    vv[1][0] = v10; vv[1][1] = v11; vv[1][2] = v12;
    vv[2][0] = v20; vv[2][1] = v21; vv[2][2] = v22;
    uu[0][0] = (s0 > tiny) ? FG_DDIV(u00, s0) : 0.0; uu[0][1] = (s0 > tiny) ? FG_DDIV(u01, s0) : 0.0; uu[0][2] = (s0 > tiny) ? FG_DDIV(u02, s0) : 0.0;
    uu[1][0] = (s1 > tiny) ? FG_DDIV(u10, s1) : 0.0; uu[1][1] = (s1 > tiny) ? FG_DDIV(u11, s1) : 0.0; uu[1][2] = (s1 > tiny) ? FG_DDIV(u12, s1) : 0.0;
    uu[2][0] = (s2 > tiny) ? FG_DDIV(u20, s2) : 0.0; uu[2][1] = (s2 > tiny) ? FG_DDIV(u21, s2) : 0.0; uu[2][2] = (s2 > tiny) ? FG_DDIV(u22, s2) : 0.0;
    if (!(S[0] > tiny))
    {
        uu[0][0] = 1.0; uu[0][1] = 0.0; uu[0][2] = 0.0;
        uu[1][0] = 0.0; uu[1][1] = 1.0; uu[1][2] = 0.0;
        uu[2][0] = 0.0; uu[2][1] = 0.0; uu[2][2] = 1.0;
    }
    else
    {
        if (!(S[1] > tiny))
        {
            // unit vector along the smallest component of uu[0] (first minimum), crossed with uu[0]
            int k = 0;
            if (fabs(uu[0][1]) < fabs(uu[0][0])) k = 1;
            if (fabs(uu[0][2]) < fabs(k == 0 ? uu[0][0] : uu[0][1])) k = 2;
            double e[3] = { k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0 };
            fg_cross3d(uu[0], e, uu[1]);
            double n = FG_DSQRT(fg_dot3d(uu[1], uu[1]));
            uu[1][0] = FG_DDIV(uu[1][0], n); uu[1][1] = FG_DDIV(uu[1][1], n); uu[1][2] = FG_DDIV(uu[1][2], n);
        }
        if (!(S[2] > tiny)) fg_cross3d(uu[0], uu[1], uu[2]);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 3; ++j)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
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
