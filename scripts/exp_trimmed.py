"""Throughput of the trimmed operators on W5: trimmed bound kernel at bench geometry, trimmed run()."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n_rot = 1024
rot, tc = workloads.bound_microbench(n_rot, 32, seed=7)
d_rot, d_tc = torch.from_numpy(rot).to(dev), torch.from_numpy(tc).to(dev)
d_lb, d_ub = torch.empty(n_rot, 32, device=dev), torch.empty(n_rot, 32, device=dev)
for rho in (0.0, 0.1, 0.3):
    ctx.set_trim(rho)
    for _ in range(2): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("trim %.1f: %.2f ms per %d pairs x 10000 points = %.3e evals/s" % (rho, ms, n_rot * 32, n_rot * 32 * 10000 / ms * 1e3), flush=True)
ctx.set_stream(0); ctx.close()
# trimmed run() with 20 % gross outliers
rng = np.random.default_rng(5)
data = w["data"].copy()
bad = rng.choice(len(data), 2000, replace=False)
lo, hi = data.min(0) - 0.2, data.max(0) + 0.2
data[bad] = (lo + rng.random((2000, 3)) * (hi - lo)).astype(np.float32)
for rho in (0.0, 0.25):
    g = driver.FastGoICP(w["model"], data, 0.005, 1e-4, flags=capi.BUILD_PACKED, trim_fraction=rho)
    R, t = g.run(); s = g.stats
    err = float(np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1))))
    print("run() with 20%% outliers, trim %.2f: %.1f ms | evals %.3e | icps %d | mse over inliers %.3e | rot err %.3f deg, t err %.4f"
          % (rho, s["run_ms"], s["bound_evals"], s["icp_runs"], float(g.best_sse) / g.n_inliers, err, float(np.linalg.norm(t - w["t_true"]))), flush=True)
    g.close()
