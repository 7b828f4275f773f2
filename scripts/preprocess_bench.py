"""Device-side constructor preprocessing: device ms of fgoicp_preprocess against the host pass.
Writes gpurun_out/preprocess_r01.json.  Run on the GPU box: python scripts/preprocess_bench.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver  # noqa: E402

rows = []
for nt, ns in ((100_000, 10_000), (1_000_000, 100_000), (8_000_000, 1_000_000)):
    rng = np.random.default_rng(nt)
    model = rng.normal(size=(nt, 3)).astype(np.float32) * 30 + 5
    data = rng.normal(size=(ns, 3)).astype(np.float32) * 25 - 2
    row = dict(nt=nt, ns=ns)
    for name, flags in (("reference_order", capi.PRE_REFERENCE), ("tree_centroid", capi.PRE_TREE_CENTROID)):
        capi.preprocess(model, data, flags=flags)
        ms, wall = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            r = capi.preprocess(model, data, flags=flags)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(r["device_ms"])
        row[name] = dict(device_ms=float(np.median(ms)), wall_ms_with_copies=float(np.median(wall)))
    t0 = time.perf_counter()
    driver.preprocess(model, data)
    row["host_numpy_ms"] = (time.perf_counter() - t0) * 1e3
    rows.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/preprocess_r01.json", "w"), indent=1)
