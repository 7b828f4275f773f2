"""python -m fast_go_icp_b200.cli -c <config.toml> [-v]

The reference's src/main.cpp (lines 8-58) over the B200 library: same flags (-c/--config required, -v/--verbose),
same TOML keys, same flow (load target, load source, construct, time run()), plus what main.cpp drops: the result is
written to `[io] output` and the transformed source cloud to `[io] visualization` when those keys are set.
"""
import argparse
import sys
import time

import numpy as np

from . import cloudio, driver


def main(argv=None):
    ap = argparse.ArgumentParser(prog="fast-go-icp")
    ap.add_argument("-c", "--config", required=True, help="Path to the configuration file")
    ap.add_argument("-v", "--verbose", action="store_true", help="Enable verbose output")
    ap.add_argument("--schedule", default="level", choices=["level", "bestfirst"])
    a = ap.parse_args(argv)
    cfg = cloudio.Config(a.config)
    p = cfg.params
    target = cloudio.load_cloud(cfg.resolve(cfg.target), p.target_subsample, p.seed)
    print("[Info] Target point cloud: %d points" % len(target), file=sys.stderr)
    source = cloudio.load_cloud(cfg.resolve(cfg.source), p.source_subsample, p.seed + 1)
    print("[Info] Source point cloud: %d points" % len(source), file=sys.stderr)
    # `trim_fraction` > 0 selects the trimmed registration (extension; `trim` alone is parsed and ignored like the
    # reference does, src/utilities.hpp:94)
    g = driver.FastGoICP(target, source, p.lut_resolution, p.mse_threshold, schedule=a.schedule,
                         trim_fraction=p.trim_fraction)
    t0 = time.perf_counter()
    R, t = g.run()
    dt = time.perf_counter() - t0
    print("[Info] Fast Go-ICP finished, time elapsed: %.6f seconds" % dt, file=sys.stderr)
    mse = float(g.best_sse) / g.n_inliers           # the sums run over the inliers only when trimming is on
    if a.verbose:
        print("[Debug] R =\n%s\nt = %s\nMSE = %.6g" % (R, t, mse), file=sys.stderr)
    if cfg.output:
        cloudio.write_result_toml(cfg.output, R, t, mse, float(g.best_sse),
                                  {"seconds": dt, "bound_evals": int(g.stats["bound_evals"]), "icp_runs": int(g.stats["icp_runs"]),
                                   "inliers": int(g.n_inliers), "trim_fraction": float(p.trim_fraction)})
    if cfg.visualization:
        cloudio.write_ply(cfg.visualization, source.astype(np.float64) @ R.T + t)
    g.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
