"""How much work does the reference's best-first order do on W5, compared with the level-synchronous schedule?"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
for sched in ("pool", "level", "pool"):
    g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED, schedule=sched)
    R, t = g.run(); s = g.stats
    print("   ", " ".join("[%d cubes span %.4f evals %.2e best %.2f]" % (l["cubes"], l["span"], l["evals"], l["best_sse"]) for l in s["level_log"]))
    err = float(np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1))))
    print("%-9s run %.1f ms | rot cubes %d | evals %.3e | icps %d | sse %.6f rot err %.3f" % (sched, s["run_ms"], s["rot_cubes"], s["bound_evals"], s["icp_runs"], g.best_sse, err), flush=True)
    g.close()
