// fgoicp/fgoicp.hpp -- icp::FastGoICP, the registration API of fast-go-icp, backed by the
// B200-native CUDA hot path behind <fgoicp_c.h>.
//
// Drop-in for the reference class of the same name (reference fgoicp/fgoicp.hpp:10-108): same
// constructor (target first, by-value clouds, lut_resolution, mse_threshold), same run() returning
// (R, t) that maps source -> target in the caller's original coordinates, same accessors.  The
// reference's src/main.cpp compiles against this header unchanged.
#ifndef FGOICP_B200_FGOICP_HPP
#define FGOICP_B200_FGOICP_HPP

#include "common.hpp"
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>

struct fgoicp_ctx;
struct fgoicp_level_stats;
namespace icp { struct DevicePool; }

namespace icp
{
    class FastGoICP
    {
    public:
        // How the outer SO(3) search is scheduled.
        enum class Schedule
        {
            Level,      // level-synchronous frontier: every surviving cube of a level is searched concurrently
            BestFirst   // the reference's serial best-first order (fgoicp.cpp:32-100), one cube at a time
        };

        struct Options
        {
            Schedule schedule = Schedule::Level;
            int device = 0;
            std::vector<int> devices;  // more than one entry: one context per device, every wave of a level sharded over them (env FGOICP_DEVICES=0,1,...)
            int sampler = -1;        // FGOICP_SAMPLER_*, -1 = library default
            int wave1 = 32;          // first-wave size of a level (then x4, x16, rest): early ICPs tighten best_sse; 0: no split
            bool skip_dead_lb = true; // skip the leaf level's rotation-uncertainty searches (they cannot change any output)
            bool verbose_levels = false;
            bool device_preprocess = false; // centre / scale / range the clouds on the GPU (fgoicp_preprocess; env FGOICP_DEVICE_PREPROCESS)
            unsigned preprocess_flags = 0;  // FGOICP_PRE_* (0 = bit-identical to the reference's host code)
            float trim_fraction = 0.0f; // > 0: trimmed registration over the (1 - trim_fraction) * ns best points (extension; env FGOICP_TRIM_FRACTION)
            // called (under the snapshot mutex, from the thread inside run()) every time a new incumbent (SSE, R, t) is
            // published -- the moments the reference's viewer would see best_* change (fgoicp.cpp:79-84, 22-23)
            std::function<void(float, const glm::mat3&, const glm::vec3&)> on_best;
        };

        struct Stats
        {
            std::uint64_t bound_evals = 0;   // (rotation cube x translation cube x data point) evaluations
            std::uint32_t rot_cubes = 0;     // rotation cubes whose inner searches were run
            std::uint32_t icp_runs = 0;
            std::uint32_t icp_iters = 0;
            std::uint32_t levels = 0;
            float ctor_ms = 0.f;             // preprocessing + NN-grid build (the reference builds its LUT here too)
            float lut_build_ms = 0.f;        // device time of the grid build alone
            float run_ms = 0.f;              // wall time of run()
            float ms_bnb_ub = 0.f, ms_icp = 0.f, ms_bnb_lb = 0.f;
        };

        FastGoICP(std::vector<glm::vec3> _pct, std::vector<glm::vec3> _pcs, float _lut_resolution, float _mse_threshold);
        FastGoICP(std::vector<glm::vec3> _pct, std::vector<glm::vec3> _pcs, float _lut_resolution, float _mse_threshold,
                  const Options& options);
        ~FastGoICP();
        FastGoICP(const FastGoICP&) = delete;
        FastGoICP& operator=(const FastGoICP&) = delete;

        using Result_t = std::tuple<glm::mat3, glm::vec3>;
        Result_t run();

        // Interfaces for visualization (reference fgoicp.hpp:32-43).  The reference's companion viewer polls these
        // from another thread without any lock (fgoicp.hpp:40-43, fgoicp.cpp:71-72); here every read returns a
        // consistent snapshot: writers publish (error, R, t) together under the same mutex.
        float get_best_error() const { std::lock_guard<std::mutex> g(snap_mutex_); return best_sse; }   // SSE, normalised frame
        Result_t get_best_transform() const { std::lock_guard<std::mutex> g(snap_mutex_); return { best_rotation, best_translation }; }
        Result_t get_last_transform() const { std::lock_guard<std::mutex> g(snap_mutex_); return { last_rotation, last_translation }; }
        // one consistent (SSE, R, t) triple of the incumbent, in the normalised frame
        std::tuple<float, glm::mat3, glm::vec3> get_best_snapshot() const
        {
            std::lock_guard<std::mutex> g(snap_mutex_);
            return { best_sse, best_rotation, best_translation };
        }

        // Extras (not in the reference)
        float get_best_mse() const { return get_best_error() / static_cast<float>(n_inliers ? n_inliers : ns); }
        const Stats& stats() const { return stats_; }
        float scaling() const { return scaling_factor; }
        fgoicp_ctx* context() const { return ctx_; }
        const PointCloud& normalised_source() const { return pcs; }
        const PointCloud& normalised_target() const { return pct; }

    private:
        PointCloud pcs;   // source (data) cloud, centred and scaled
        PointCloud pct;   // target (model) cloud, centred and scaled by the same factor
        size_t ns, nt;
        size_t n_inliers = 0;      // points entering every sum (ns unless trimming is on)

        glm::vec3 offset_pcs;
        glm::vec3 offset_pct;
        float scaling_factor;
        std::array<std::pair<float, float>, 3> target_bounds;

        float best_sse;
        glm::mat3 best_rotation;
        glm::vec3 best_translation;
        float mse_threshold;
        float sse_threshold;

        glm::mat3 last_rotation{ 1.0f };
        glm::vec3 last_translation{ 0.0f };

        mutable std::mutex snap_mutex_;
        void publish_best(float e, const glm::mat3& R, const glm::vec3& t)
        {
            std::lock_guard<std::mutex> g(snap_mutex_);
            best_sse = e; best_rotation = R; best_translation = t;
            if (options_.on_best) options_.on_best(e, R, t);
        }
        void publish_last(const glm::mat3& R, const glm::vec3& t)
        {
            std::lock_guard<std::mutex> g(snap_mutex_);
            last_rotation = R; last_translation = t;
        }

        Options options_;
        Stats stats_;
        fgoicp_ctx* ctx_ = nullptr;            // context of the first device (ICPs of run(), best-first schedule)
        std::vector<fgoicp_ctx*> ctxs_;        // one per device; ctxs_[0] == ctx_
        std::unique_ptr<DevicePool> pool_;     // one long-lived host thread per additional device (csrc/fgoicp_host.cpp)

        void init(float lut_resolution);
        void preprocess_clouds();
        glm::vec3 center_point_cloud(PointCloud& pc);
        float scale_point_clouds(PointCloud& pct, PointCloud& pcs);
        std::array<std::pair<float, float>, 3> get_point_cloud_ranges(PointCloud& pc);
        glm::vec3 restore_translation(glm::mat3 _R, glm::vec3 _t)
        {
            return _t / scaling_factor + _R * offset_pcs - offset_pct;     // reference fgoicp.hpp:87-90
        }

        float icp(int max_iter, float thr, const glm::mat3& R0, const glm::vec3& t0, glm::mat3& R, glm::vec3& t);
        void search_level_synchronous();
        void level_ub(const float* cubes, int m, float* ub, float* bt, float& level_best, float* bR, float* bT, fgoicp_level_stats& st);
        void level_lb(const float* cubes, int n, float* lb, fgoicp_level_stats& st);
        void search_best_first();
    };
}

#endif // FGOICP_B200_FGOICP_HPP
