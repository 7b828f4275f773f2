// fgoicp_host.cpp -- icp::FastGoICP: host driver of the Go-ICP search over the C ABI.
//
// Mirrors the behaviour of the reference's fgoicp/fgoicp.cpp (run(), branch_and_bound_SO3(),
// preprocessing) but owns no CUDA code: every device operation goes through <fgoicp_c.h>.
// Two schedules of the outer SO(3) search are provided:
//   Level     -- level-synchronous frontier; all surviving cubes of a level run their inner searches
//                concurrently on the GPU (one thread block each).  This is the production path and
//                the one that shards across GPUs (see fast_go_icp_b200/driver.py).
//   BestFirst -- the reference's serial order (fgoicp.cpp:32-100), for side-by-side parity runs.
#include <fgoicp/fgoicp.hpp>
#include <fgoicp_c.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <thread>

namespace icp
{
    // One long-lived host thread per additional device.  A thread that touches the CUDA runtime for the first time pays
    // milliseconds for its thread-local state and for binding the device's context, so spawning fresh threads for every
    // wave of every level (round 1) made two GPUs slower than one: 186 ms against 132 ms for run() on W5.
    struct DevicePool
    {
        struct Worker
        {
            std::thread th;
            std::mutex m;
            std::condition_variable cv;
            std::function<void()> task;
            bool busy = false, stop = false;
        };
        std::vector<std::unique_ptr<Worker>> workers;

        explicit DevicePool(size_t n)
        {
            for (size_t k = 0; k < n; ++k)
            {
                workers.emplace_back(new Worker());
                Worker* w = workers.back().get();
                w->th = std::thread([w]()
                {
                    std::unique_lock<std::mutex> lk(w->m);
                    while (true)
                    {
                        w->cv.wait(lk, [w]() { return w->stop || (w->busy && w->task); });
                        if (w->stop) return;
                        std::function<void()> f;
                        f.swap(w->task);
                        lk.unlock();
                        f();
                        lk.lock();
                        w->busy = false;
                        w->cv.notify_all();
                    }
                });
            }
        }
        ~DevicePool()
        {
            for (auto& w : workers)
            {
                { std::lock_guard<std::mutex> lk(w->m); w->stop = true; }
                w->cv.notify_all();
                if (w->th.joinable()) w->th.join();
            }
        }
        void post(size_t k, std::function<void()> f)
        {
            Worker* w = workers[k].get();
            { std::lock_guard<std::mutex> lk(w->m); w->task = std::move(f); w->busy = true; }
            w->cv.notify_all();
        }
        void wait(size_t k)
        {
            Worker* w = workers[k].get();
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [w]() { return !w->busy; });
        }
        void wait_all() { for (size_t k = 0; k < workers.size(); ++k) wait(k); }
    };

    namespace
    {
        struct ApiError : std::runtime_error
        {
            explicit ApiError(const std::string& what) : std::runtime_error(what) {}
        };

        void check(int rc, const char* what)
        {
            if (rc != FGOICP_OK)
                throw ApiError(std::string(what) + ": " + fgoicp_last_error());
        }

        void to_array(const glm::mat3& R, float* out)
        {
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) out[c * 3 + r] = R[c][r];
        }

        glm::mat3 to_mat3(const float* a)
        {
            return glm::mat3(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
        }

        // joins whatever threads were started, also when starting one of them (or anything after) throws
        struct JoinAll
        {
            std::vector<std::thread>& th;
            ~JoinAll() { for (auto& t : th) if (t.joinable()) t.join(); }
        };

        double now_ms()
        {
            using clk = std::chrono::steady_clock;
            return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
        }
    }

    FastGoICP::FastGoICP(std::vector<glm::vec3> _pct, std::vector<glm::vec3> _pcs, float _lut_resolution, float _mse_threshold)
        : FastGoICP(std::move(_pct), std::move(_pcs), _lut_resolution, _mse_threshold, Options())
    {}

    // Member order and preprocessing order follow the reference constructor (fgoicp.hpp:13-25):
    // centre source, centre target, scale both by the SOURCE's factor, bounding box of the target.
    FastGoICP::FastGoICP(std::vector<glm::vec3> _pct, std::vector<glm::vec3> _pcs, float _lut_resolution, float _mse_threshold,
                         const Options& options)
        : pcs(std::move(_pcs)), pct(std::move(_pct)), ns{ pcs.size() }, nt{ pct.size() },
          offset_pcs(0.0f), offset_pct(0.0f), scaling_factor(1.0f),
          best_sse(M_INF), best_rotation(1.0f), best_translation(0.0f),
          mse_threshold(_mse_threshold),
          sse_threshold(ns * mse_threshold),
          options_(options)
    {
        if (const char* s = std::getenv("FGOICP_SCHEDULE"))
        {
            if (std::strcmp(s, "bestfirst") == 0) options_.schedule = Schedule::BestFirst;
            if (std::strcmp(s, "level") == 0) options_.schedule = Schedule::Level;
        }
        if (const char* s = std::getenv("FGOICP_SAMPLER")) options_.sampler = std::atoi(s);
        if (const char* s = std::getenv("FGOICP_DEVICE")) options_.device = std::atoi(s);
        if (const char* s = std::getenv("FGOICP_WAVE1")) options_.wave1 = std::atoi(s);
        if (const char* s = std::getenv("FGOICP_SKIP_DEAD_LB")) options_.skip_dead_lb = std::atoi(s) != 0;
        if (const char* s = std::getenv("FGOICP_TRIM_FRACTION")) options_.trim_fraction = static_cast<float>(std::atof(s));
        if (const char* s = std::getenv("FGOICP_DEVICE_PREPROCESS")) options_.device_preprocess = std::atoi(s) != 0;
        if (const char* s = std::getenv("FGOICP_DEVICES"))
        {
            options_.devices.clear();
            for (const char* p = s; *p;)
            {
                char* end = nullptr;
                long d = std::strtol(p, &end, 10);
                if (end == p) break;
                options_.devices.push_back(static_cast<int>(d));
                p = (*end == ',') ? end + 1 : end;
            }
        }
        if (options_.devices.empty()) options_.devices.push_back(options_.device);
        options_.device = options_.devices.front();
        preprocess_clouds();
        init(_lut_resolution);
    }

    // Reference constructor order (fgoicp.hpp:16-19): centre source, centre target, scale both, range of the target --
    // on the host in the reference's own fp32 order, or on the GPU through the C ABI (same bits with flags = 0).
    void FastGoICP::preprocess_clouds()
    {
        static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "glm::vec3 must be three packed floats");
        if (!options_.device_preprocess)
        {
            offset_pcs = center_point_cloud(pcs);
            offset_pct = center_point_cloud(pct);
            scaling_factor = scale_point_clouds(pct, pcs);
            target_bounds = get_point_cloud_ranges(pct);
            return;
        }
        fgoicp_normalisation n;
        check(fgoicp_preprocess(reinterpret_cast<float*>(pct.data()), nt, reinterpret_cast<float*>(pcs.data()), ns,
                                options_.device, options_.preprocess_flags, &n), "fgoicp_preprocess");
        offset_pcs = glm::vec3(n.offset_pcs[0], n.offset_pcs[1], n.offset_pcs[2]);
        offset_pct = glm::vec3(n.offset_pct[0], n.offset_pct[1], n.offset_pct[2]);
        scaling_factor = n.scale;
        for (int a = 0; a < 3; ++a) target_bounds[a] = std::make_pair(n.bbox_min[a], n.bbox_max[a]);
    }

    void FastGoICP::init(float lut_resolution)
    {
        double t0 = now_ms();
        float bmin[3] = { target_bounds[0].first, target_bounds[1].first, target_bounds[2].first };
        float bmax[3] = { target_bounds[0].second, target_bounds[1].second, target_bounds[2].second };
        static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "glm::vec3 must be three packed floats");
        unsigned flags = FGOICP_BUILD_PACKED;
        if (options_.sampler == FGOICP_SAMPLER_TEX) flags |= FGOICP_BUILD_TEX;
        // one context per device (clouds and grids are replicated); built concurrently, one host thread each
        const size_t K = options_.devices.size();
        ctxs_.assign(K, nullptr);
        {
            std::vector<std::string> errors(K);
            auto make = [&](size_t k)
            {
                int rc = fgoicp_ctx_create(reinterpret_cast<const float*>(pct.data()), nt,
                                           reinterpret_cast<const float*>(pcs.data()), ns,
                                           bmin, bmax, lut_resolution, options_.devices[k], flags, &ctxs_[k]);
                if (rc == FGOICP_OK && options_.sampler >= 0) rc = fgoicp_set_sampler(ctxs_[k], options_.sampler);
                if (rc != FGOICP_OK) errors[k] = std::string("fgoicp_ctx_create: ") + fgoicp_last_error();   // message is per thread
            };
            // a constructor that throws never runs its destructor: release the contexts that were created (tens of MB to
            // GBs of HBM each) before passing the error on
            auto release = [&]() { for (fgoicp_ctx*& c : ctxs_) { fgoicp_ctx_destroy(c); c = nullptr; } ctx_ = nullptr; };
            try
            {
                if (K == 1) make(0);
                else
                {
                    std::vector<std::thread> th;
                    JoinAll guard{ th };
                    for (size_t k = 0; k < K; ++k) th.emplace_back(make, k);
                }
            }
            catch (...) { release(); throw; }
            ctx_ = ctxs_[0];
            for (size_t k = 0; k < K; ++k)
                if (!errors[k].empty()) { release(); throw ApiError(errors[k]); }
        }
        try
        {
        if (K > 1) pool_.reset(new DevicePool(K - 1));
        fgoicp_info info;
        check(fgoicp_ctx_info(ctx_, &info), "fgoicp_ctx_info");
        if (info.dims[0] >= 1024 || info.dims[1] >= 1024 || info.dims[2] >= 1024)
            Logger(LogLevel::Warning) << "Dims " << info.dims[0] << ", " << info.dims[1] << ", " << info.dims[2]
                                      << " is large, consider a lower LUT resolution";
        stats_.lut_build_ms = info.build_ms;
        if (options_.trim_fraction > 0.0f)
        {
            // trimmed registration (extension; the reference only parses `trim`): sums over the n_inliers smallest
            // residuals, threshold scaled accordingly (Go-ICP: SSEThresh = MSEThresh * inlierNum)
            std::uint64_t k = ns;
            for (fgoicp_ctx* c : ctxs_) check(fgoicp_set_trim(c, options_.trim_fraction, &k), "fgoicp_set_trim");
            n_inliers = static_cast<size_t>(k);
            sse_threshold = static_cast<float>(n_inliers) * mse_threshold;
        }
        }
        catch (...)
        {
            for (fgoicp_ctx*& c : ctxs_) { fgoicp_ctx_destroy(c); c = nullptr; }
            ctx_ = nullptr;
            throw;
        }
        stats_.ctor_ms = static_cast<float>(now_ms() - t0);
    }

    FastGoICP::~FastGoICP()
    {
        pool_.reset();                                   // workers first: none may still hold a context
        for (fgoicp_ctx* c : ctxs_) fgoicp_ctx_destroy(c);
    }

    // One wave of the fixed-rotation phase (inner searches + ICP on promising cubes) over all devices.  The cubes are
    // dealt round-robin; every device works against the same wave-start best_sse, so which cube is refined and what
    // every search returns do not depend on the number of devices.  The new incumbent is the MIN over
    // (sse, index of the cube within the wave) -- the cube a single device's ascending scan would have picked.
    void FastGoICP::level_ub(const float* cubes, int m, float* ub, float* bt, float& level_best, float* bR, float* bT,
                             fgoicp_level_stats& st)
    {
        const int K = static_cast<int>(ctxs_.size());
        if (K == 1 || m < 2)
        {
            check(fgoicp_so3_level_ub(ctx_, cubes, m, best_sse, sse_threshold, ub, bt, &level_best, bR, bT, &st), "fgoicp_so3_level_ub");
            return;
        }
        struct Shard
        {
            std::vector<float> cubes, ub, bt;
            float best, R[9], t[3];
            fgoicp_level_stats st;
            std::string error;
        };
        std::vector<Shard> sh(K);
        const float start_best = best_sse;
        for (int k = 0; k < K; ++k)
        {
            const int mk = (m - k + K - 1) / K;
            sh[k].cubes.resize(4 * static_cast<size_t>(mk)); sh[k].ub.resize(mk); sh[k].bt.resize(3 * static_cast<size_t>(mk));
            for (int j = 0; j < mk; ++j) std::memcpy(&sh[k].cubes[4 * j], cubes + 4 * (k + j * K), 4 * sizeof(float));
            sh[k].best = start_best;
            std::memcpy(sh[k].R, bR, sizeof(sh[k].R)); std::memcpy(sh[k].t, bT, sizeof(sh[k].t));
            std::memset(&sh[k].st, 0, sizeof(sh[k].st)); sh[k].st.best_icp_index = -1;
        }
        static const bool log_waves = std::getenv("FGOICP_WAVE_LOG") != nullptr;
        std::vector<double> t_shard(K, 0.0);
        auto work = [&](int k)
        {
            Shard& s = sh[k];
            const int mk = static_cast<int>(s.ub.size());
            if (mk == 0) return;
            const double ts = now_ms();
            int rc = fgoicp_so3_level_ub(ctxs_[k], s.cubes.data(), mk, start_best, sse_threshold, s.ub.data(), s.bt.data(),
                                         &s.best, s.R, s.t, &s.st);
            t_shard[k] = now_ms() - ts;
            if (rc != FGOICP_OK) s.error = std::string("fgoicp_so3_level_ub: ") + fgoicp_last_error();
        };
        // devices 1.. on their own long-lived threads, device 0 on this one; the shards outlive the tasks (wait_all below)
        const double tw = now_ms();
        for (int k = 1; k < K; ++k) pool_->post(static_cast<size_t>(k - 1), [&work, k]() { work(k); });
        work(0);
        pool_->wait_all();
        if (log_waves)
        {
            std::fprintf(stderr, "[wave ub] m %d wall %.3f ms | per device: call ms (device ub + icp ms):", m, now_ms() - tw);
            for (int k = 0; k < K; ++k) std::fprintf(stderr, " %.3f (%.3f + %.3f)", t_shard[k], sh[k].st.ms_bnb_ub, sh[k].st.ms_icp);
            std::fprintf(stderr, "\n");
        }
        std::memset(&st, 0, sizeof(st)); st.best_icp_index = -1;
        int winner = -1, winner_index = 0;
        for (int k = 0; k < K; ++k)
        {
            Shard& s = sh[k];
            if (!s.error.empty()) throw ApiError(s.error);
            const int mk = static_cast<int>(s.ub.size());
            for (int j = 0; j < mk; ++j)
            {
                ub[k + j * K] = s.ub[j];
                std::memcpy(bt + 3 * (k + j * K), &s.bt[3 * j], 3 * sizeof(float));
            }
            st.evals += s.st.evals; st.n_icp += s.st.n_icp; st.icp_iters += s.st.icp_iters;
            st.ms_bnb_ub = std::max(st.ms_bnb_ub, s.st.ms_bnb_ub); st.ms_icp = std::max(st.ms_icp, s.st.ms_icp);
            if (s.st.best_icp_index >= 0 && s.best < start_best)
            {
                const int gi = k + s.st.best_icp_index * K;
                if (winner < 0 || s.best < sh[winner].best || (s.best == sh[winner].best && gi < winner_index)) { winner = k; winner_index = gi; }
            }
        }
        if (winner >= 0)
        {
            level_best = sh[winner].best;
            std::memcpy(bR, sh[winner].R, sizeof(sh[winner].R)); std::memcpy(bT, sh[winner].t, sizeof(sh[winner].t));
            st.best_icp_index = winner_index;
        }
    }

    // Rotation-uncertainty searches of a level over all devices (independent per cube).
    void FastGoICP::level_lb(const float* cubes, int n, float* lb, fgoicp_level_stats& st)
    {
        const int K = static_cast<int>(ctxs_.size());
        if (K == 1 || n < 2)
        {
            check(fgoicp_so3_level_lb(ctx_, cubes, n, best_sse, sse_threshold, lb, &st), "fgoicp_so3_level_lb");
            return;
        }
        std::vector<std::vector<float>> sc(K), sl(K);
        std::vector<fgoicp_level_stats> ss(K);
        std::vector<std::string> errors(K);
        for (int k = 0; k < K; ++k)
        {
            const int nk = (n - k + K - 1) / K;
            sc[k].resize(4 * static_cast<size_t>(nk)); sl[k].resize(nk);
            for (int j = 0; j < nk; ++j) std::memcpy(&sc[k][4 * j], cubes + 4 * (k + j * K), 4 * sizeof(float));
            std::memset(&ss[k], 0, sizeof(ss[k]));
        }
        auto work = [&](int k)
        {
            const int nk = static_cast<int>(sl[k].size());
            if (nk == 0) return;
            int rc = fgoicp_so3_level_lb(ctxs_[k], sc[k].data(), nk, best_sse, sse_threshold, sl[k].data(), &ss[k]);
            if (rc != FGOICP_OK) errors[k] = std::string("fgoicp_so3_level_lb: ") + fgoicp_last_error();
        };
        for (int k = 1; k < K; ++k) pool_->post(static_cast<size_t>(k - 1), [&work, k]() { work(k); });
        work(0);
        pool_->wait_all();
        std::memset(&st, 0, sizeof(st));
        for (int k = 0; k < K; ++k)
        {
            if (!errors[k].empty()) throw ApiError(errors[k]);
            for (size_t j = 0; j < sl[k].size(); ++j) lb[k + static_cast<int>(j) * K] = sl[k][j];
            st.evals += ss[k].evals;
            st.ms_bnb_lb = std::max(st.ms_bnb_lb, ss[k].ms_bnb_lb);
        }
    }

    float FastGoICP::icp(int max_iter, float thr, const glm::mat3& R0, const glm::vec3& t0, glm::mat3& R, glm::vec3& t)
    {
        float r0[9], tt0[3] = { t0.x, t0.y, t0.z }, r[9], tt[3], sse = 0.f;
        int iters = 0;
        to_array(R0, r0);
        check(fgoicp_icp(ctx_, r0, tt0, max_iter, thr, &sse, r, tt, &iters), "fgoicp_icp");
        R = to_mat3(r);
        t = glm::vec3(tt[0], tt[1], tt[2]);
        stats_.icp_runs += 1;
        stats_.icp_iters += static_cast<std::uint32_t>(iters);
        return sse;
    }

    // reference fgoicp.cpp:10-30
    FastGoICP::Result_t FastGoICP::run()
    {
        double t0 = now_ms();
        glm::mat3 icp_R; glm::vec3 icp_t;
        // Initial ICP from the identity; only its error is kept (fgoicp.cpp:12-14)
        {
            float e0 = icp(100, 0.05, glm::mat3(1.0f), glm::vec3(0.0f), icp_R, icp_t);
            publish_best(e0, best_rotation, best_translation);
        }
        Logger(LogLevel::Info) << "Initial ICP best error: " << best_sse
                               << "\n\tRotation:\n" << icp_R
                               << "\n\tTranslation: " << icp_t;

        if (options_.schedule == Schedule::BestFirst) search_best_first();
        else search_level_synchronous();

        // Refine the best transform (fgoicp.cpp:22-23)
        glm::mat3 R; glm::vec3 t;
        {
            float e1 = icp(100, 0.0005, best_rotation, best_translation, R, t);
            publish_best(e1, R, t);
        }

        Logger(LogLevel::Info) << "Searching over! Best Error: " << best_sse
                               << "\n\tRotation:\n" << best_rotation
                               << "\n\tTranslation: " << restore_translation(best_rotation, best_translation);
        stats_.run_ms = static_cast<float>(now_ms() - t0);
        return { best_rotation, restore_translation(best_rotation, best_translation) };
    }

    // Level-synchronous form of branch_and_bound_SO3 (fgoicp.cpp:32-100).  Per level:
    //   1. nodes with best_sse - lb <= sse_threshold are finished (the reference's stop rule :44 applies to
    //      the smallest lb first, hence to every node it would still hold);
    //   2. spawn the 8 octants (:49-59), drop cubes missing the unit ball (:61), pass cubes whose centre is
    //      outside the ball on unevaluated (:62-66);
    //   3. inner search with the rotation fixed for ALL remaining children at once -> ub, t* (:69),
    //      ICP from (R, t*) where ub < 1.8 best_sse (:74-88);
    //   4. inner search with rotation uncertainty for all children -> lb (:90), drop lb >= best_sse (:92).
    void FastGoICP::search_level_synchronous()
    {
        std::vector<RotNode> frontier;
        frontier.emplace_back(0.0f, 0.0f, 0.0f, 1.0f, 0.0f, best_sse);
        while (!frontier.empty())
        {
            std::vector<RotNode> open;
            for (const RotNode& n : frontier)
                if (!(best_sse - n.lb <= sse_threshold)) open.push_back(n);
            if (open.empty()) break;
            float span = open.front().span / 2.0f;
            if (span < 0.05f) break;

            std::vector<RotNode> next, eval;
            for (const RotNode& n : open)
                for (char j = 0; j < 8; ++j)
                {
                    RotNode child(n.q.x - span + (j >> 0 & 1) * n.span,
                                  n.q.y - span + (j >> 1 & 1) * n.span,
                                  n.q.z - span + (j >> 2 & 1) * n.span,
                                  span, n.lb, n.ub);
                    if (!child.overlaps_SO3()) continue;
                    if (!child.q.in_SO3()) { next.push_back(child); continue; }
                    eval.push_back(child);
                }
            const int n = static_cast<int>(eval.size());
            std::vector<float> cubes(4 * static_cast<size_t>(n)), ub(n), bt(3 * static_cast<size_t>(n)), lb(n);
            for (int i = 0; i < n; ++i)
            {
                cubes[4 * i] = eval[i].q.x; cubes[4 * i + 1] = eval[i].q.y; cubes[4 * i + 2] = eval[i].q.z;
                cubes[4 * i + 3] = eval[i].span;
            }
            // Fixed-rotation phase in (up to) two waves: the children of the parents with the smallest
            // fixed-rotation error first, so that what their ICPs find tightens best_sse for the rest.
            std::vector<int> order(n);
            for (int i = 0; i < n; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return eval[a].ub < eval[b].ub; });
            std::vector<std::vector<int>> waves;               // sizes wave1, 4*wave1, 16*wave1, then the rest
            for (int lo = 0, size = options_.wave1; lo < n; size *= 4)
            {
                int hi = (size <= 0 || waves.size() >= 3) ? n : std::min(n, lo + size);
                waves.emplace_back(order.begin() + lo, order.begin() + hi);
                lo = hi;
            }
            fgoicp_level_stats st_ub{}, st_lb{};
            for (auto& wave : waves)
            {
                if (wave.empty()) continue;
                std::sort(wave.begin(), wave.end());
                const int m = static_cast<int>(wave.size());
                std::vector<float> wc(4 * static_cast<size_t>(m)), wub(m), wbt(3 * static_cast<size_t>(m));
                for (int k = 0; k < m; ++k) std::memcpy(&wc[4 * k], &cubes[4 * wave[k]], 4 * sizeof(float));
                float bR[9], bT[3] = { best_translation.x, best_translation.y, best_translation.z };
                to_array(best_rotation, bR);
                float level_best = best_sse;
                fgoicp_level_stats st{};
                level_ub(wc.data(), m, wub.data(), wbt.data(), level_best, bR, bT, st);
                for (int k = 0; k < m; ++k)
                {
                    ub[wave[k]] = wub[k];
                    std::memcpy(&bt[3 * wave[k]], &wbt[3 * k], 3 * sizeof(float));
                }
                if (level_best < best_sse)
                {
                    publish_best(level_best, to_mat3(bR), glm::vec3(bT[0], bT[1], bT[2]));
                    Logger(LogLevel::Debug) << "New best error: " << best_sse
                                            << "\n\tRotation:\n" << best_rotation
                                            << "\n\tTranslation: " << restore_translation(best_rotation, best_translation);
                }
                st_ub.evals += st.evals; st_ub.n_icp += st.n_icp; st_ub.icp_iters += st.icp_iters;
                st_ub.ms_bnb_ub += st.ms_bnb_ub; st_ub.ms_icp += st.ms_icp;
            }
            if (options_.skip_dead_lb && span / 2.0f < 0.05f)
            {
                // Leaf level: the children of these cubes are never evaluated (reference fgoicp.cpp:53), so their
                // lower bounds can only feed the loop's exit test; the reference computes them anyway (:90).
                // Skipping them changes no output.
                std::fill(lb.begin(), lb.end(), 0.0f);
            }
            else
                level_lb(cubes.data(), n, lb.data(), st_lb);
            for (int i = 0; i < n; ++i)
            {
                if (lb[i] >= best_sse) continue;
                eval[i].lb = lb[i];
                eval[i].ub = ub[i];
                next.push_back(eval[i]);
            }
            if (n > 0)
            {
                publish_last(eval[n - 1].q.R, glm::vec3(bt[3 * (n - 1)], bt[3 * (n - 1) + 1], bt[3 * (n - 1) + 2]));
            }
            stats_.levels += 1;
            stats_.rot_cubes += static_cast<std::uint32_t>(n);
            stats_.bound_evals += st_ub.evals + st_lb.evals;
            stats_.icp_runs += st_ub.n_icp;
            stats_.icp_iters += st_ub.icp_iters;
            stats_.ms_bnb_ub += st_ub.ms_bnb_ub; stats_.ms_icp += st_ub.ms_icp; stats_.ms_bnb_lb += st_lb.ms_bnb_lb;
            if (options_.verbose_levels || Logger::verbose())
                Logger(LogLevel::Debug) << "level span " << span << ": " << n << " cubes, " << st_ub.n_icp << " ICPs, "
                                        << (st_ub.evals + st_lb.evals) << " bound evals, best " << best_sse
                                        << ", survivors " << next.size();
            frontier.swap(next);
        }
    }

    // The reference's own serial order (fgoicp.cpp:32-100); std::priority_queue<RotNode> as there.
    void FastGoICP::search_best_first()
    {
        std::priority_queue<RotNode> rcandidates;
        rcandidates.push(RotNode(0.0f, 0.0f, 0.0f, 1.0f, 0.0f, best_sse));
        while (!rcandidates.empty())
        {
            RotNode rnode = rcandidates.top();
            rcandidates.pop();
            if (best_sse - rnode.lb <= sse_threshold) break;
            float span = rnode.span / 2.0f;
            for (char j = 0; j < 8; ++j)
            {
                if (span < 0.05f) continue;
                RotNode child(rnode.q.x - span + (j >> 0 & 1) * rnode.span,
                              rnode.q.y - span + (j >> 1 & 1) * rnode.span,
                              rnode.q.z - span + (j >> 2 & 1) * rnode.span,
                              span, rnode.lb, rnode.ub);
                if (!child.overlaps_SO3()) continue;
                if (!child.q.in_SO3()) { rcandidates.push(child); continue; }

                float cube[4] = { child.q.x, child.q.y, child.q.z, child.span };
                float ub = 0.f, bt[3] = { 0, 0, 0 }, lb = 0.f, dummy[3];
                std::uint64_t ev = 0;
                check(fgoicp_bnb_r3(ctx_, cube, 1, best_sse, sse_threshold, &ub, bt, &ev), "fgoicp_bnb_r3");
                stats_.bound_evals += ev; stats_.rot_cubes += 1;
                publish_last(child.q.R, glm::vec3(bt[0], bt[1], bt[2]));
                if (ub < best_sse * 1.8)
                {
                    glm::mat3 R; glm::vec3 t;
                    float e = icp(100, 0.005, child.q.R, last_translation, R, t);
                    if (e < best_sse) publish_best(e, R, t);
                }
                check(fgoicp_bnb_r3(ctx_, cube, 0, best_sse, sse_threshold, &lb, dummy, &ev), "fgoicp_bnb_r3");
                stats_.bound_evals += ev;
                if (lb >= best_sse) continue;
                child.lb = lb; child.ub = ub;
                rcandidates.push(child);
            }
        }
    }

    // reference fgoicp.cpp:176-195: serial fp32 sum in index order, returns -centroid
    glm::vec3 FastGoICP::center_point_cloud(PointCloud& pc)
    {
        glm::vec3 centroid(0.0f);
        for (size_t i = 0; i < pc.size(); ++i) centroid += pc[i];
        centroid /= static_cast<float>(pc.size());
        for (size_t i = 0; i < pc.size(); ++i) pc[i] -= centroid;
        return -centroid;
    }

    // reference fgoicp.cpp:197-220, 271-287: 1 / max |coordinate| of the SOURCE, applied to both clouds
    float FastGoICP::scale_point_clouds(PointCloud& target, PointCloud& source)
    {
        float max_abs = std::numeric_limits<float>::lowest();
        for (const auto& p : source)
            max_abs = std::max(max_abs, std::max(std::abs(p.x), std::max(std::abs(p.y), std::abs(p.z))));
        float s = 1.0f / max_abs;
        for (auto& p : source) p *= s;
        for (auto& p : target) p *= s;
        return s;
    }

    // reference fgoicp.cpp:222-268
    std::array<std::pair<float, float>, 3> FastGoICP::get_point_cloud_ranges(PointCloud& pc)
    {
        std::array<std::pair<float, float>, 3> ranges;
        for (auto& r : ranges) r = std::make_pair(std::numeric_limits<float>::max(), std::numeric_limits<float>::lowest());
        for (const auto& p : pc)
            for (int a = 0; a < 3; ++a)
            {
                ranges[a].first = std::min(ranges[a].first, p[a]);
                ranges[a].second = std::max(ranges[a].second, p[a]);
            }
        return ranges;
    }
}
