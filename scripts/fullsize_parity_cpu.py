"""Full-size parity on the CPU: the oracle driven through the SAME level-synchronous driver as the CUDA path (Python
driver with tests/oracle_context.OracleContext in place of the CUDA context) against the result the CUDA path returned
on the B200 for the same clouds (profiles/repo_clouds_all_cases_r01.json).  Needs no GPU; minutes per case.

    python scripts/fullsize_parity_cpu.py bunny 0.005 1e-3 "W1 bunny res 0.005"
    python scripts/fullsize_parity_cpu.py bunny 0.005 1e-5 "W1 bunny res 0.005 mse 1e-5"
    python scripts/fullsize_parity_cpu.py w5 0.005 1e-4 "W5 synthetic 100k/10k"
    python scripts/fullsize_parity_cpu.py dragon 0.005 1e-3 "W3 dragon"
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_repo_clouds import load_pair  # noqa: E402
from fast_go_icp_b200 import driver  # noqa: E402
from oracle_context import OracleContext  # noqa: E402

pair, res, mse, case = sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
model, data, _, _ = load_pair(pair)
t0 = time.perf_counter()
g = driver.FastGoICP(model, data, res, mse, ctx_factory=OracleContext)
R, t = g.run()
print("case:", case, "(%d / %d points)" % (len(model), len(data)))
print("oracle through the driver: sse", repr(float(g.best_sse)), "evals", g.stats["bound_evals"], "icp runs / iterations",
      g.stats["icp_runs"], g.stats["icp_iters"], "rotation cubes", g.stats["rot_cubes"], "%.1f s" % (time.perf_counter() - t0))
o = [r for r in json.load(open(os.path.join(ROOT, "profiles", "repo_clouds_all_cases_r01.json"))) if r["case"] == case][0]["ours"]
print("CUDA path on the B200:     sse", repr(o["sse"]), "evals", o["bound_evals"], "icp runs", o["icp_runs"], "rotation cubes", o["rot_cubes"])
same = (np.float32(g.best_sse) == np.float32(o["sse"]) and np.array_equal(np.asarray(R, np.float32), np.asarray(o["R"], np.float32))
        and np.array_equal(np.asarray(t, np.float32), np.asarray(o["t"], np.float32)) and g.stats["bound_evals"] == o["bound_evals"])
print("SSE, R, t bit-identical and evaluation counts equal:", bool(same))
