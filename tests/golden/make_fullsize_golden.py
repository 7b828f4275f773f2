"""Builds tests/golden/bunny_full.npz: the reference repository's bunny pair at the size test/bunny.toml asks for
(17,973 model / 3,037 data points; clouds from scripts/make_full_clouds.py, i.e. the seeded loader on
data/bunny/{model,data}_bunny.txt) together with the result of the CUDA path's run() on a B200 as recorded by
scripts/bench_repo_clouds.py (profiles/repo_clouds_all_cases_r01.json).  Run HERE after those two scripts:

    python scripts/make_full_clouds.py && python tests/golden/make_fullsize_golden.py

tests/test_fullsize_parity.py then checks, bit for bit, that (CPU) the oracle driven through the same level-synchronous
driver and (GPU) the CUDA path both reproduce the recorded SSE, pose and evaluation counts."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
z = np.load(os.path.join(ROOT, "build", "workloads", "repo_clouds_full.npz"))
rows = json.load(open(os.path.join(ROOT, "profiles", "repo_clouds_all_cases_r01.json")))
out = dict(model=z["bunny_model"], data=z["bunny_data"])
# W5 (synthetic, regenerated from its seed by workloads.synthetic_pair()): result only
for tag, case in (("mse1e-3", "W1 bunny res 0.005"), ("mse1e-5", "W1 bunny res 0.005 mse 1e-5"), ("w5", "W5 synthetic 100k/10k")):
    o = [r for r in rows if r["case"] == case][0]["ours"]
    out["gpu_R_" + tag] = np.asarray(o["R"], np.float32)
    out["gpu_t_" + tag] = np.asarray(o["t"], np.float32)
    out["gpu_sse_" + tag] = np.float32(o["sse"])
    out["gpu_counts_" + tag] = np.array([o["bound_evals"], o["rot_cubes"], o["icp_runs"]], np.int64)
path = os.path.join(ROOT, "tests", "golden", "bunny_full.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})

# W3: the two dragon range scans at full size (75,305 / 10,000 points) -- the ICP-heavy case (2,498 refinements, 46,588
# ICP iterations; the CPU oracle reproduces the mse 1e-3 result bit for bit in 343 s, profiles/fullsize_parity_r01.log)
out = dict(model=z["dragon_model"], data=z["dragon_data"])
for tag, case in (("mse1e-3", "W3 dragon"), ("mse1e-4", "W3 dragon mse 1e-4")):
    o = [r for r in rows if r["case"] == case][0]["ours"]
    out["gpu_R_" + tag] = np.asarray(o["R"], np.float32)
    out["gpu_t_" + tag] = np.asarray(o["t"], np.float32)
    out["gpu_sse_" + tag] = np.float32(o["sse"])
    out["gpu_counts_" + tag] = np.array([o["bound_evals"], o["rot_cubes"], o["icp_runs"]], np.int64)
path = os.path.join(ROOT, "tests", "golden", "dragon_full.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
