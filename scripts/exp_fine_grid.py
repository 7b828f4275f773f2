"""Sanity at the reference's own resolution (test/bunny.toml: lut_resolution 0.002, ~18k / 3k points): grid of
~900^3 nodes (2.9 GB dense, 23 GB corner-packed)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=18000, ns=3000, sigma=0.002, seed=77)
t0 = time.perf_counter()
g = driver.FastGoICP(w["model"], w["data"], 0.002, 1e-3, flags=capi.BUILD_PACKED)
info = g.ctx.info()
print("ctor %.1f ms (grid build %.1f ms) dims %s dense %.2f GB packed %.2f GB" % ((time.perf_counter() - t0) * 1e3, info.build_ms, list(info.dims), info.grid_bytes / 1e9, info.packed_bytes / 1e9), flush=True)
R, t = g.run(); s = g.stats
err = float(np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1))))
print("run %.1f ms | evals %.3e | icps %d | mse %.3e | rot err %.3f deg t err %.4f" % (s["run_ms"], s["bound_evals"], s["icp_runs"], float(g.best_sse) / 3000, err, float(np.linalg.norm(t - w["t_true"]))), flush=True)
# bounds: phased vs plain on the big grid
rot, tc = workloads.bound_microbench(512, 32, seed=5)
g.ctx.set_phased(True); a = g.ctx.bounds_multi(rot, False, tc)
g.ctx.set_phased(False); b = g.ctx.bounds_multi(rot, False, tc)
print("phased == plain on the fine grid:", np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))
g.close()
g2 = driver.FastGoICP(w["model"], w["data"], 0.002, 1e-5, flags=capi.BUILD_PACKED)
R, t = g2.run(); s = g2.stats
print("mse_threshold 1e-5: run %.1f ms | evals %.3e | rot cubes %d | mse %.3e" % (s["run_ms"], s["bound_evals"], s["rot_cubes"], float(g2.best_sse) / 3000), flush=True)
g2.close()
