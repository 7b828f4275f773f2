"""Times the exact NN search (fgoicp_sse = squared rule, one pass over the data cloud) on W5 for near and far poses."""
import os, sys, time, numpy as np
sys.path.insert(0, '.')
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
I = np.eye(3, dtype=np.float32).ravel()
poses = {"identity (far)": (I, np.zeros(3, np.float32))}
Rt = w["R_true"].astype(np.float32)
# true pose in the normalised frame: y_n = R x_n + t_n
s = pp["scale"]; t_n = (w["t_true"] + Rt @ (-pp["offset_pcs"]) + pp["offset_pct"]) * s
poses["true pose (near)"] = (Rt.T.ravel().copy(), t_n.astype(np.float32))
for k, c in enumerate([(0.5, 0.5, 0.5), (-0.5, 0.5, -0.5), (0.25, -0.75, 0.25)]):
    poses["cube %s" % (c,)] = (driver.rotation_matrix(*c)[0], np.zeros(3, np.float32))
for lv in os.environ.get("LEVELS", "1,2,3").split(","):
    os.environ["FGOICP_NN_LEVELS"] = lv
    for name, (R, t) in poses.items():
        ctx.sse(R, t)
        t0 = time.perf_counter()
        for _ in range(20):
            e = ctx.sse(R, t)
        ms = (time.perf_counter() - t0) / 20 * 1e3
        for rooted in (False, True):
            ctx.nn(R, t, rooted)
            t1 = time.perf_counter()
            for _ in range(10):
                idx, d2 = ctx.nn(R, t, rooted)
            print("      nn rooted=%s %.3f ms" % (rooted, (time.perf_counter() - t1) / 10 * 1e3))
        print("levels %s  %-28s sse %.4f  %.3f ms per pass  mean d %.4f" % (lv, name, e, ms, float(np.sqrt(d2).mean())), flush=True)
ctx.close()
