import sys, os, numpy as np, torch
sys.path.insert(0,'.')
os.environ["FGOICP_PHASED"]="1"; os.environ.setdefault("FGOICP_PHASED_LAG","2")
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
dev=torch.device("cuda",0)
stream=torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n_rot=4096
rot, tc = workloads.bound_microbench(n_rot, 32, seed=7)
def run(tc, label):
    d_rot, d_tc = torch.from_numpy(rot).to(dev), torch.from_numpy(np.ascontiguousarray(tc)).to(dev)
    d_lb, d_ub = torch.empty(n_rot,32,device=dev), torch.empty(n_rot,32,device=dev)
    for _ in range(3): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, False, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(label,"ms %.3f"%ms,"evals/s %.3e"%(n_rot*32*10000/ms*1e3), flush=True)
run(tc, "microbench (216 distinct leaf t)")
tc1 = tc.copy(); tc1[..., :3] = tc[0,0,:3]
run(tc1, "all cubes the same t")
tc2 = tc.copy(); tc2[..., 2] = tc[0,0,2]
run(tc2, "same t.z, x/y varied")
tc3 = tc.copy(); tc3[..., :2] = tc[0,0,:2]
run(tc3, "same t.x,t.y, z varied")
ctx.close()
