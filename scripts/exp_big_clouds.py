"""Sanity with large clouds: 300k-point model, 80k-point data (phase-ordered == plain, NN grid == brute force, run())."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=300_000, ns=80_000, sigma=0.005, seed=31)
t0 = time.perf_counter()
g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED)
print("ctor %.1f ms, grid build %.1f ms" % ((time.perf_counter() - t0) * 1e3, g.ctx.info().build_ms), flush=True)
rot, tc = workloads.bound_microbench(256, 32, seed=5)
g.ctx.set_phased(True); t1 = time.perf_counter(); a = g.ctx.bounds_multi(rot, False, tc); tp = time.perf_counter() - t1
g.ctx.set_phased(False); t1 = time.perf_counter(); b = g.ctx.bounds_multi(rot, False, tc); tq = time.perf_counter() - t1
g.ctx.set_phased(True)
print("phased == plain:", np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "| %.2e vs %.2e evals/s (host-buffer calls)" % (256 * 32 * 80000 / tp, 256 * 32 * 80000 / tq), flush=True)
I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
i0, d0 = g.ctx.nn(I, np.zeros(3, np.float32), False)
g.ctx.set_nn_mode(1); i1, d1 = g.ctx.nn(I, np.zeros(3, np.float32), False); g.ctx.set_nn_mode(0)
print("cell-grid NN == brute force:", np.array_equal(i0, i1) and np.array_equal(d0, d1), flush=True)
R, t = g.run(); s = g.stats
err = float(np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1))))
print("run %.1f ms | evals %.3e | icps %d | mse %.3e | rot err %.3f deg t err %.4f" % (s["run_ms"], s["bound_evals"], s["icp_runs"], float(g.best_sse) / 80000, err, float(np.linalg.norm(t - w["t_true"]))), flush=True)
g.close()
