"""Experiment: does a spatially coherent data-point order (Morton) speed up the unordered-gather kernels?"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
def morton(p, bits=10):
    q = ((p - p.min(0)) / (p.max(0) - p.min(0) + 1e-9) * (2 ** bits - 1)).astype(np.uint64)
    code = np.zeros(len(p), np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(code, kind="stable")
ws = workloads.synthetic_pair(nt=3000, ns=400, seed=3)
gw = driver.FastGoICP(ws["model"], ws["data"], 0.03, 1e-4, flags=capi.BUILD_PACKED); gw.run(); gw.close()
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
for label, data in (("original order", w["data"]), ("morton order", w["data"][morton(w["data"])])):
    for rep in range(2):
        g = driver.FastGoICP(w["model"], data, 0.005, 1e-4, flags=capi.BUILD_PACKED)
        g.run(); s = g.stats
        print("%-15s run %.1f ms | ub %.1f icp %.1f lb %.1f | evals %.3e | sse %.6f" % (label, s["run_ms"], s["ms_bnb_ub"], s["ms_icp"], s["ms_bnb_lb"], s["bound_evals"], g.best_sse), flush=True)
        g.close()
