"""Times the z-phase-ordered bound kernel on the bench workload under the environment's FGOICP_PHASED_*
settings and checks it against the plain kernel (bit-identical sums expected)."""
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n_rot = int(os.environ.get("N_ROT", "4096"))
rot, tc = workloads.bound_microbench(n_rot, 32, seed=7)
d_rot, d_tc = torch.from_numpy(rot).to(dev), torch.from_numpy(tc).to(dev)
d_lb, d_ub = torch.empty(n_rot, 32, device=dev), torch.empty(n_rot, 32, device=dev)
def timed(label, reps=5):
    for _ in range(3): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(label, "ms %.3f" % ms, "evals/s %.3e" % (n_rot * 32 * 10000 / ms * 1e3), flush=True)
    return d_lb.cpu().numpy().copy(), d_ub.cpu().numpy().copy()
ref = None
if os.environ.get("CHECK", "0") == "1":
    ctx.set_phased(False); ref = timed("plain")
ctx.set_phased(True)
for pf in os.environ.get("PFS", "0,1,2,3").split(","):
    for lag in os.environ.get("LAGS", "1,2,3").split(","):
        os.environ["FGOICP_PHASED_PF"] = pf; os.environ["FGOICP_PHASED_LAG"] = lag
        got = timed("pf %s lag %s" % (pf, lag))
        if ref is not None:
            print("   equal to plain:", np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]))
ctx.close()
