"""One run() on the dragon pair (W3): the workload for an ncu capture of k_nn_grid (the ICP-bound case).
ncu --set full --clock-control none --import-source on -k regex:k_nn_grid -s 1500 -c 2 -o gpurun_out/nn_grid_w3 python scripts/profile_nn_w3.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from fast_go_icp_b200 import capi, driver  # noqa: E402

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dragon_full.npz"))
model, data = z["model"], z["data"]
g = driver.FastGoICP(model, data, 0.005, 1e-4, flags=capi.BUILD_PACKED)
g.run()
print("run_ms", g.stats["run_ms"], "icp ms", g.stats["ms_icp"], "iters", g.stats["icp_iters"])
g.close()
