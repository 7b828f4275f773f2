"""Experiment: run() on W5 with each sampler (packed = default, grid = dense 8-gather, tex = hardware tex3D)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
ws = workloads.synthetic_pair(nt=3000, ns=400, seed=3)
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
for name, smp in (("packed", capi.SAMPLER_PACKED), ("tex", capi.SAMPLER_TEX), ("grid", capi.SAMPLER_GRID)):
    gw = driver.FastGoICP(ws["model"], ws["data"], 0.03, 1e-4, flags=capi.BUILD_PACKED | capi.BUILD_TEX, sampler=smp); gw.run(); gw.close()
    for rep in range(2):
        g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED | capi.BUILD_TEX, sampler=smp)
        R, t = g.run(); s = g.stats
        err = float(np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1))))
        print("%-7s run %.1f ms | ub %.1f icp %.1f lb %.1f | evals %.3e | sse %.6f rot err %.3f" % (name, s["run_ms"], s["ms_bnb_ub"], s["ms_icp"], s["ms_bnb_lb"], s["bound_evals"], g.best_sse, err), flush=True)
        g.close()
