import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FGOICP_PHASED"] = "1"
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402
n_rot = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
rot, tc = workloads.bound_microbench(n_rot, 32, seed=7)
dev = torch.device("cuda", 0)
d_rot, d_tc = torch.from_numpy(rot).to(dev), torch.from_numpy(tc).to(dev)
d_lb, d_ub = torch.empty(n_rot, 32, device=dev), torch.empty(n_rot, 32, device=dev)
for _ in range(3):
    ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
torch.cuda.synchronize()
print("ok", float(d_ub.min()))
ctx.close()
