"""run() on one of the committed full-size pairs (tests/golden/<pair>_full.npz, or w5) for environment-variable sweeps:
    FGOICP_ICP_SLOTS=128 python scripts/run_repo_case.py dragon 0.005 1e-4 [reps] [trim_fraction] [wave1]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402
pair, res, mse = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
trim = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
wave1 = int(sys.argv[6]) if len(sys.argv) > 6 else 32
Rt = tt = None
if pair == "w5":
    w = workloads.synthetic_pair()
    model, data, Rt, tt = w["model"], w["data"], w["R_true"], w["t_true"]
else:
    z = np.load(os.path.join(ROOT, "tests", "golden", pair + "_full.npz"))
    model, data = z["model"], z["data"]
    if "R_move" in z.files:
        Rt, tt = z["R_move"].T, -z["R_move"].T @ z["t_move"]
ws = workloads.synthetic_pair(nt=3000, ns=400, seed=3)
g = driver.FastGoICP(ws["model"], ws["data"], 0.03, 1e-4, trim_fraction=trim); g.run(); g.close()      # loads the kernels
for _ in range(reps):
    g = driver.FastGoICP(model, data, res, mse, flags=capi.BUILD_PACKED, trim_fraction=trim, wave1=wave1)
    R, t = g.run(); s = g.stats
    err = ""
    if Rt is not None:
        err = " | rot err %.3f deg, t err %.4g" % (float(np.degrees(np.arccos(np.clip((np.trace(R @ Rt.T) - 1) / 2, -1, 1)))), float(np.linalg.norm(t - tt)))
    print("%s res %g mse %g trim %g wave1 %d: run %.1f ms | ub %.1f icp %.1f lb %.1f | sse %.8g (bits %08x) icps %d iters %d evals %.3e%s"
          % (pair, res, mse, trim, wave1, s["run_ms"], s["ms_bnb_ub"], s["ms_icp"], s["ms_bnb_lb"], float(g.best_sse),
             int(np.float32(g.best_sse).view(np.uint32)), s["icp_runs"], s["icp_iters"], s["bound_evals"], err), flush=True)
    g.close()
