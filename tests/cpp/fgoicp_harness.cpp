// Test harness for the drop-in C++ class icp::FastGoICP (reference fgoicp/fgoicp.hpp:10-108), used the way the
// reference's src/main.cpp:46-53 uses it: construct with (target, source, lut_resolution, mse_threshold), run(),
// read the error.  Reads two raw float32 xyz files, prints one line of hex floats so the Python test can compare
// bit patterns:   R[9 column-major] t[3] sse scale ctor_ms run_ms bound_evals rot_cubes icp_runs
//
//   fgoicp_harness <model.f32> <data.f32> <lut_resolution> <mse_threshold>
//
// Options come from the environment (FGOICP_SCHEDULE, FGOICP_DEVICE_PREPROCESS, FGOICP_TRIM_FRACTION ...), exactly
// as they would for the unchanged reference CLI.
#include <fgoicp/fgoicp.hpp>

#include <cstdio>
#include <cstdlib>
#include <fstream>

static std::vector<glm::vec3> read_cloud(const char* path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    std::vector<glm::vec3> pts(static_cast<size_t>(bytes) / sizeof(glm::vec3));
    f.read(reinterpret_cast<char*>(pts.data()), static_cast<std::streamsize>(pts.size() * sizeof(glm::vec3)));
    return pts;
}

int main(int argc, char** argv)
{
    if (argc != 5) { std::fprintf(stderr, "usage: %s model.f32 data.f32 lut_resolution mse_threshold\n", argv[0]); return 2; }
    try
    {
        icp::Logger::set_verbose(false);
        std::vector<glm::vec3> pct = read_cloud(argv[1]), pcs = read_cloud(argv[2]);
        icp::FastGoICP fgoicp(std::move(pct), std::move(pcs), static_cast<float>(std::atof(argv[3])),
                              static_cast<float>(std::atof(argv[4])));
        auto [R, t] = fgoicp.run();
        std::printf("RESULT");
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) std::printf(" %a", static_cast<double>(R[c][r]));
        for (int a = 0; a < 3; ++a) std::printf(" %a", static_cast<double>(t[a]));
        std::printf(" %a %a %a %a", static_cast<double>(fgoicp.get_best_error()), static_cast<double>(fgoicp.scaling()),
                    static_cast<double>(fgoicp.stats().ctor_ms), static_cast<double>(fgoicp.stats().run_ms));
        std::printf(" %a %a %a\n", static_cast<double>(fgoicp.stats().bound_evals), static_cast<double>(fgoicp.stats().rot_cubes),
                    static_cast<double>(fgoicp.stats().icp_runs));
        return 0;
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "ERROR %s\n", e.what());
        return 1;
    }
}
