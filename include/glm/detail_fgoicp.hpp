// Minimal stand-in for the parts of GLM that the fast-go-icp caller code touches.
//
// GLM itself is not installed in this image, yet the reference's
// src/main.cpp:4 includes <glm/vec3.hpp> and its public API (fgoicp/fgoicp.hpp:13,29)
// traffics in glm::vec3 / glm::mat3.  This header is written from scratch and only
// promises GLM's *observable semantics* for the handful of operations used:
//   - vec3 / mat3 are plain aggregates of floats, mat3 is COLUMN-major: m[col][row];
//   - mat3(float d) builds d * identity; mat3(9 floats) fills column by column;
//   - mat3 * vec3, mat3 * mat3 evaluate their three-term sums left to right;
//   - vec3 / scalar divides component-wise (no reciprocal multiply).
// If a real GLM is on the include path first, it wins and this file is never seen.
#ifndef FGOICP_GLM_SHIM_DETAIL_HPP
#define FGOICP_GLM_SHIM_DETAIL_HPP

#include <cmath>
#include <cstddef>

#if defined(__CUDACC__)
#define FGOICP_GLM_HD __host__ __device__
#else
#define FGOICP_GLM_HD
#endif

namespace glm
{
    typedef int length_t;

    struct vec3
    {
        float x, y, z;

        vec3() = default;
        FGOICP_GLM_HD explicit vec3(float s) : x(s), y(s), z(s) {}
        FGOICP_GLM_HD vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}

        FGOICP_GLM_HD float& operator[](length_t i) { return (&x)[i]; }
        FGOICP_GLM_HD const float& operator[](length_t i) const { return (&x)[i]; }

        FGOICP_GLM_HD vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
        FGOICP_GLM_HD vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
        FGOICP_GLM_HD vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
        FGOICP_GLM_HD vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
    };

    FGOICP_GLM_HD inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    FGOICP_GLM_HD inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    FGOICP_GLM_HD inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
    FGOICP_GLM_HD inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
    FGOICP_GLM_HD inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
    FGOICP_GLM_HD inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
    FGOICP_GLM_HD inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
    FGOICP_GLM_HD inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
    FGOICP_GLM_HD inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }

    FGOICP_GLM_HD inline float dot(const vec3& a, const vec3& b)
    {
        vec3 tmp(a * b);
        return tmp.x + tmp.y + tmp.z;
    }
    FGOICP_GLM_HD inline float length(const vec3& v) { return sqrtf(dot(v, v)); }
    FGOICP_GLM_HD inline float distance(const vec3& p0, const vec3& p1) { return length(p1 - p0); }

    struct mat3
    {
        vec3 value[3];   // three columns

        mat3() = default;
        FGOICP_GLM_HD explicit mat3(float d)
        {
            value[0] = vec3(d, 0.0f, 0.0f);
            value[1] = vec3(0.0f, d, 0.0f);
            value[2] = vec3(0.0f, 0.0f, d);
        }
        FGOICP_GLM_HD mat3(float x0, float y0, float z0,
                           float x1, float y1, float z1,
                           float x2, float y2, float z2)
        {
            value[0] = vec3(x0, y0, z0);
            value[1] = vec3(x1, y1, z1);
            value[2] = vec3(x2, y2, z2);
        }
        // GLM's converting constructor: nine scalars of any arithmetic type, cast to float.
        template <typename X0, typename Y0, typename Z0,
                  typename X1, typename Y1, typename Z1,
                  typename X2, typename Y2, typename Z2>
        FGOICP_GLM_HD mat3(X0 x0, Y0 y0, Z0 z0, X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2)
        {
            value[0] = vec3(static_cast<float>(x0), static_cast<float>(y0), static_cast<float>(z0));
            value[1] = vec3(static_cast<float>(x1), static_cast<float>(y1), static_cast<float>(z1));
            value[2] = vec3(static_cast<float>(x2), static_cast<float>(y2), static_cast<float>(z2));
        }
        FGOICP_GLM_HD mat3(const vec3& c0, const vec3& c1, const vec3& c2)
        {
            value[0] = c0; value[1] = c1; value[2] = c2;
        }

        FGOICP_GLM_HD vec3& operator[](length_t c) { return value[c]; }
        FGOICP_GLM_HD const vec3& operator[](length_t c) const { return value[c]; }
    };

    typedef mat3 mat3x3;

    FGOICP_GLM_HD inline vec3 operator*(const mat3& m, const vec3& v)
    {
        return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                    m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                    m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
    }

    FGOICP_GLM_HD inline mat3 operator*(const mat3& a, const mat3& b)
    {
        mat3 r;
        for (length_t c = 0; c < 3; ++c)
        {
            r[c] = vec3(a[0][0] * b[c][0] + a[1][0] * b[c][1] + a[2][0] * b[c][2],
                        a[0][1] * b[c][0] + a[1][1] * b[c][1] + a[2][1] * b[c][2],
                        a[0][2] * b[c][0] + a[1][2] * b[c][1] + a[2][2] * b[c][2]);
        }
        return r;
    }

    FGOICP_GLM_HD inline mat3 operator+(const mat3& a, const mat3& b)
    {
        return mat3(a[0] + b[0], a[1] + b[1], a[2] + b[2]);
    }

    FGOICP_GLM_HD inline mat3 operator*(const mat3& a, float s)
    {
        return mat3(a[0] * s, a[1] * s, a[2] * s);
    }

    FGOICP_GLM_HD inline mat3 transpose(const mat3& m)
    {
        return mat3(m[0][0], m[1][0], m[2][0],
                    m[0][1], m[1][1], m[2][1],
                    m[0][2], m[1][2], m[2][2]);
    }

    // outerProduct(c, r)[i] = c * r[i]  (column i)
    FGOICP_GLM_HD inline mat3 outerProduct(const vec3& c, const vec3& r)
    {
        return mat3(c * r[0], c * r[1], c * r[2]);
    }
}

#endif // FGOICP_GLM_SHIM_DETAIL_HPP
