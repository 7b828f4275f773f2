import sys, os
sys.path.insert(0, "/root/repo")
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair()
ws = workloads.synthetic_pair(nt=3000, ns=400, seed=3)
g = driver.FastGoICP(ws["model"], ws["data"], 0.03, 1e-4); g.run(); g.close()
for wave1 in (4, 8, 16, 32, 64, 0):
    for rep in range(2):
        g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED, wave1=wave1)
        g.run(); s = g.stats
    print("wave1 %3d: run %.1f ms | ub %.1f icp %.1f lb %.1f | icps %d iters %d evals %.3e sse %.7g" % (wave1, s["run_ms"], s["ms_bnb_ub"], s["ms_icp"], s["ms_bnb_lb"], s["icp_runs"], s["icp_iters"], s["bound_evals"], float(g.best_sse)), flush=True)
    g.close()
