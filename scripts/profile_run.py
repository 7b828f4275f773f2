"""Short driver for ncu: the full W5 run() (level-synchronous search); use -k regex:k_bnb_r3 -s 6 -c 1 to
capture the leaf-level fixed-rotation search launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402

w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED)
R, t = g.run()
print("run ms", g.stats["run_ms"], "evals", g.stats["bound_evals"])
g.close()
