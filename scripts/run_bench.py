"""Times FastGoICP.run() on W5 (after a small warm-up run that loads every kernel): 3 repetitions, per-phase ms."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
ws = workloads.synthetic_pair(nt=3000, ns=400, seed=3)
gw = driver.FastGoICP(ws["model"], ws["data"], 0.03, 1e-4, flags=capi.BUILD_PACKED); gw.run(); gw.close()
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
reps = int(os.environ.get("REPS", "3"))
for rep in range(reps):
    g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED)
    R, t = g.run()
    s = g.stats
    print("run %.1f ms | ub %.1f icp %.1f lb %.1f | evals %.3e icps %d iters %d | sse %.6f" %
          (s["run_ms"], s["ms_bnb_ub"], s["ms_icp"], s["ms_bnb_lb"], s["bound_evals"], s["icp_runs"], s["icp_iters"], g.best_sse), flush=True)
    if rep == reps - 1 and os.environ.get("LEVELS_LOG"):
        for l in s["level_log"]:
            print("  ", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in l.items()})
    g.close()
