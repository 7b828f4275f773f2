// fgoicp/common.hpp -- shared types of the B200-native fast-go-icp drop-in.
//
// Source-compatible with the reference header of the same path for everything its callers
// (src/main.cpp, src/utilities.hpp) use: namespace icp, Logger / LogLevel / Logger::set_verbose,
// the Point3D / PointCloud aliases, `using std::string;` at global scope and the transitive
// standard includes those callers rely on (reference fgoicp/common.hpp:3-21, 171-269).
// Unlike the reference this header pulls in no CUDA headers: the device side sits behind the
// C ABI in <fgoicp_c.h>.
#ifndef FGOICP_B200_COMMON_HPP
#define FGOICP_B200_COMMON_HPP

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include <math.h>
#include <glm/vec3.hpp>
#include <glm/mat3x3.hpp>

// The reference redefines M_PI as a float literal and adds two more (fgoicp/common.hpp:17-19);
// callers that include this header see the same values.
#ifdef M_PI
#undef M_PI
#endif
#define M_PI    3.141592653589793f
#define M_INF   1E+10f
#define M_SQRT3 1.732050807568877f

using std::string;

namespace icp
{
    typedef glm::vec3 Point3D;
    using PointCloud = std::vector<Point3D>;

    // Quaternion-vector parametrised rotation (reference fgoicp/common.hpp:30-69): (x, y, z) is the
    // vector part of a unit quaternion, w = sqrt(1 - |q|^2).  Outside the unit ball R stays the
    // identity and r keeps |q|^2; inside, r = |q|.  R is what glm's column-major constructor makes
    // of the nine listed values, i.e. the transpose of the textbook quaternion matrix.
    struct Rotation
    {
        float x, y, z, r;
        glm::mat3 R;

        Rotation() : Rotation(0.0f, 0.0f, 0.0f) {}
        Rotation(float x_, float y_, float z_) : x(x_), y(y_), z(z_), r(x_ * x_ + y_ * y_ + z_ * z_), R(1.0f)
        {
            if (r > 1.0f) { return; }
            float ww = 1.0f - r;
            float w = std::sqrt(ww);
            float wx = w * x, xx = x * x;
            float wy = w * y, xy = x * y, yy = y * y;
            float wz = w * z, xz = x * z, yz = y * z, zz = z * z;
            R = glm::mat3(ww + xx - yy - zz, 2 * (xy - wz), 2 * (xz + wy),
                          2 * (xy + wz), ww - xx + yy - zz, 2 * (yz - wx),
                          2 * (xz - wy), 2 * (yz + wx), ww - xx - yy + zz);
            r = std::sqrt(r);
        }
        bool in_SO3() const { return r <= 1.0f; }
    };

    // SO(3) search cube.  Heap order: smaller lb first, ties -> larger span first
    // (reference fgoicp/common.hpp:75-104).
    struct RotNode
    {
        Rotation q;
        float span;
        float lb, ub;

        RotNode(float x, float y, float z, float span_, float lb_, float ub_) : q(x, y, z), span(span_), lb(lb_), ub(ub_) {}

        friend bool operator<(const RotNode& a, const RotNode& b)
        {
            if (a.lb == b.lb) { return a.span < b.span; }
            return a.lb > b.lb;
        }
        bool overlaps_SO3() const
        {
            return q.r - 2 * span * (std::abs(q.x) + std::abs(q.y) + std::abs(q.z)) + 3 * span * span <= 1;
        }
    };

    // R^3 search cube, same ordering (reference fgoicp/common.hpp:110-128).
    struct TransNode
    {
        glm::vec3 t;
        float span;
        float lb, ub;

        TransNode(float x, float y, float z, float span_, float lb_, float ub_) : t(x, y, z), span(span_), lb(lb_), ub(ub_) {}

        friend bool operator<(const TransNode& a, const TransNode& b)
        {
            if (a.lb == b.lb) { return a.span < b.span; }
            return a.lb > b.lb;
        }
    };

    enum class LogLevel { Debug, Info, Warning, Error };

    // RAII line logger: streams into a buffer, prints on destruction with a level tag, a wall-clock
    // stamp and an ANSI colour; Debug lines are dropped unless set_verbose(true)
    // (behaviour of reference fgoicp/common.hpp:171-269).
    class Logger
    {
    public:
        explicit Logger(LogLevel level) : level_(level) {}
        Logger() : Logger(LogLevel::Debug) {}

        template <typename T>
        Logger& operator<<(const T& msg) { buffer_ << msg; return *this; }

        Logger& operator<<(const glm::vec3& v)
        {
            buffer_ << std::fixed << std::setprecision(6) << v.x << "\t" << v.y << "\t" << v.z;
            return *this;
        }

        Logger& operator<<(const glm::mat3& m)
        {
            buffer_ << std::fixed << std::setprecision(4);
            for (int row = 0; row < 3; ++row)
            {
                buffer_ << "\t" << m[0][row] << "\t" << m[1][row] << "\t" << m[2][row];
                if (row < 2) buffer_ << "\n";
            }
            return *this;
        }

        ~Logger()
        {
            if (level_ == LogLevel::Debug && !verbose_) return;
            static const char* const colour[] = { "\033[34m", "\033[32m", "\033[33m", "\033[31m" };
            static const char* const name[] = { "Debug", "Info", "Warning", "Error" };
            int k = static_cast<int>(level_);
            std::cout << colour[k] << "[" << name[k] << " " << stamp() << "] " << buffer_.str() << "\033[0m" << "\n";
        }

        static void set_verbose(bool verbose) { verbose_ = verbose; }
        static bool verbose() { return verbose_; }

    private:
        LogLevel level_;
        std::ostringstream buffer_;
        inline static bool verbose_ = false;

        static std::string stamp()
        {
            std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
            std::tm buf{};
#if defined(_WIN32) || defined(_WIN64)
            localtime_s(&buf, &now);
#else
            localtime_r(&now, &buf);
#endif
            std::ostringstream ss;
            ss << std::put_time(&buf, "%H:%M:%S");
            return ss.str();
        }
    };
}

#endif // FGOICP_B200_COMMON_HPP
