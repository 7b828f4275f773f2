/* TEST INFRASTRUCTURE (oracle).  Not part of the shipped product path.
 *
 * Double-precision SVD of a 3x3 matrix by one-sided (Hestenes) Jacobi rotations.
 *
 * The reference calls Eigen::JacobiSVD<Eigen::Matrix3d> (fgoicp/icp3d.cu:118-121).  Eigen 3
 * is an un-vendored external dependency (fgoicp/CMakeLists.txt:21, version ">= 3.3", not
 * pinned, not installed here), so its published algorithm is restated instead: Jacobi SVD,
 * full U and V, singular values sorted in decreasing order.  The quantity the reference
 * derives from it -- R = V diag(1,1,det(V U^T)) U^T, icp3d.cu:123-133 -- is unique whenever
 * sigma_2 != sigma_3, so two correct double-precision SVDs agree to ~1e-15 on it.
 *
 * Row-major 3x3 arrays: a[r*3+c].  A = U * diag(s) * V^T.
 */
#ifndef FGOICP_ORACLE_SVD3_H
#define FGOICP_ORACLE_SVD3_H

#include <math.h>

static inline void orc_svd3_cross(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static inline void orc_svd3(const double A[9], double U[9], double S[3], double V[9])
{
    double u[3][3]; /* u[col][row] : working columns of A*V */
    double v[3][3]; /* v[col][row] */
    int i, j, sweep;
    for (j = 0; j < 3; ++j)
        for (i = 0; i < 3; ++i)
        {
            u[j][i] = A[i * 3 + j];
            v[j][i] = (i == j) ? 1.0 : 0.0;
        }

    for (sweep = 0; sweep < 60; ++sweep)
    {
        int rotated = 0;
        int p, q;
        for (p = 0; p < 2; ++p)
            for (q = p + 1; q < 3; ++q)
            {
                double alpha = u[p][0] * u[p][0] + u[p][1] * u[p][1] + u[p][2] * u[p][2];
                double beta = u[q][0] * u[q][0] + u[q][1] * u[q][1] + u[q][2] * u[q][2];
                double gamma = u[p][0] * u[q][0] + u[p][1] * u[q][1] + u[p][2] * u[q][2];
                /* converged at one unit roundoff (2^-52); same constant as csrc/svd3_device.cuh, whose twin this is */
                if (gamma == 0.0 || fabs(gamma) <= 2.220446049250313e-16 * sqrt(alpha * beta)) continue;
                {
                    double zeta = (beta - alpha) / (2.0 * gamma);
                    double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    double c = 1.0 / sqrt(1.0 + t * t);
                    double s = c * t;
                    for (i = 0; i < 3; ++i)
                    {
                        double up = u[p][i], uq = u[q][i];
                        double vp = v[p][i], vq = v[q][i];
                        u[p][i] = c * up - s * uq;
                        u[q][i] = s * up + c * uq;
                        v[p][i] = c * vp - s * vq;
                        v[q][i] = s * vp + c * vq;
                    }
                    rotated = 1;
                }
            }
        if (!rotated) break;
    }

    {
        double sig[3];
        int order[3] = { 0, 1, 2 };
        for (j = 0; j < 3; ++j)
            sig[j] = sqrt(u[j][0] * u[j][0] + u[j][1] * u[j][1] + u[j][2] * u[j][2]);
        /* sort indices by decreasing sigma (stable) */
        for (i = 0; i < 2; ++i)
            for (j = 0; j < 2 - i; ++j)
                if (sig[order[j]] < sig[order[j + 1]])
                {
                    int tmp = order[j]; order[j] = order[j + 1]; order[j + 1] = tmp;
                }
        {
            double uu[3][3], vv[3][3];
            double smax = sig[order[0]];
            double tiny = smax * 1e-300 + 1e-300;
            for (j = 0; j < 3; ++j)
            {
                int src = order[j];
                S[j] = sig[src];
                for (i = 0; i < 3; ++i)
                {
                    vv[j][i] = v[src][i];
                    uu[j][i] = (sig[src] > tiny) ? u[src][i] / sig[src] : 0.0;
                }
            }
            /* complete U when trailing singular values vanish */
            if (!(S[0] > tiny))
            {
                for (j = 0; j < 3; ++j)
                    for (i = 0; i < 3; ++i)
                        uu[j][i] = (i == j) ? 1.0 : 0.0;
            }
            else
            {
                if (!(S[1] > tiny))
                {
                    /* any unit vector orthogonal to uu[0] */
                    double e[3] = { 0.0, 0.0, 0.0 };
                    double n;
                    int k = 0;
                    if (fabs(uu[0][1]) < fabs(uu[0][k])) k = 1;
                    if (fabs(uu[0][2]) < fabs(uu[0][k])) k = 2;
                    e[k] = 1.0;
                    orc_svd3_cross(uu[0], e, uu[1]);
                    n = sqrt(uu[1][0] * uu[1][0] + uu[1][1] * uu[1][1] + uu[1][2] * uu[1][2]);
                    for (i = 0; i < 3; ++i) uu[1][i] /= n;
                }
                if (!(S[2] > tiny))
                {
                    orc_svd3_cross(uu[0], uu[1], uu[2]);
                }
            }
            for (j = 0; j < 3; ++j)
                for (i = 0; i < 3; ++i)
                {
                    U[i * 3 + j] = uu[j][i];
                    V[i * 3 + j] = vv[j][i];
                }
        }
    }
}

static inline double orc_det3(const double M[9])
{
    return M[0] * (M[4] * M[8] - M[5] * M[7])
         - M[1] * (M[3] * M[8] - M[5] * M[6])
         + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* R = V * diag(1, 1, det(V U^T)) * U^T  for H = U S V^T  (fgoicp/icp3d.cu:123-133). */
static inline void orc_closest_rotation(const double H[9], double R[9])
{
    double U[9], S[3], V[9], VUt[9], d;
    int i, j, k;
    orc_svd3(H, U, S, V);
    for (i = 0; i < 3; ++i)
        for (j = 0; j < 3; ++j)
        {
            double acc = 0.0;
            for (k = 0; k < 3; ++k) acc += V[i * 3 + k] * U[j * 3 + k];
            VUt[i * 3 + j] = acc;
        }
    d = orc_det3(VUt);
    for (i = 0; i < 3; ++i)
        for (j = 0; j < 3; ++j)
        {
            double acc = 0.0;
            for (k = 0; k < 3; ++k)
                acc += V[i * 3 + k] * (k == 2 ? d : 1.0) * U[j * 3 + k];
            R[i * 3 + j] = acc;
        }
}

#endif /* FGOICP_ORACLE_SVD3_H */
