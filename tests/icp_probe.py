"""Helper of test_gpu_parity.py::test_winner_memo_changes_no_result: runs a fixed set of refinements and prints every
output as hex floats.  FGOICP_NN_MARGIN is read once per process, so each setting needs its own process."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402


def main(which):
    if which == "synthetic":
        w = workloads.synthetic_pair(nt=30000, ns=4000, sigma=0.004, seed=17)
        model, data = w["model"], w["data"]
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "clouds_small.npz"))
        model, data = z[which + "_model"], z[which + "_data"]
    pp = driver.preprocess(model, data)
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.01, flags=capi.BUILD_PACKED)
    rng = np.random.default_rng(5)
    Rs, ts = [], []
    for k in range(20):
        v = (rng.uniform(-0.4, 0.4, 3) * (0.05 if k % 2 else 1.0)).astype(np.float32)
        Rs.append(driver.rotation_matrix(*v)[0])
        ts.append(rng.uniform(-0.1, 0.1, 3).astype(np.float32))
    out = {}
    for thr in (0.005, 0.00001):                       # the second one runs long: many late iterations with tiny moves
        e, R, t, it = ctx.icp_batch(np.array(Rs), np.array(ts), 100, thr)
        out[str(thr)] = dict(e=[float(x).hex() for x in e], R=[float(x).hex() for x in R.ravel()],
                             t=[float(x).hex() for x in t.ravel()], it=[int(x) for x in it])
    # the per-point searches behind fgoicp_nn / fgoicp_sse never use the memo; they must be unaffected
    idx, d2 = ctx.nn(Rs[0], ts[0], rooted=True)
    out["nn"] = dict(idx=[int(x) for x in idx[:200]], sse=float(ctx.sse(Rs[0], ts[0])).hex())
    ctx.close()
    print("PROBE " + json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1])
