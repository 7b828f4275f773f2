#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import sys, os, json, numpy as np, torch
sys.path.insert(0,'.')
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
pp = driver.preprocess(w["model"], w["data"])
res={}
for phased in (0,1):
    os.environ["FGOICP_PHASED"]=str(phased)
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
    dev=torch.device("cuda",0)
    stream=torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    for n_rot in (4096,):
        rot, tc = workloads.bound_microbench(n_rot, 32, seed=7)
        d_rot, d_tc = torch.from_numpy(rot).to(dev), torch.from_numpy(tc).to(dev)
        d_lb, d_ub = torch.empty(n_rot,32,device=dev), torch.empty(n_rot,32,device=dev)
        for fix_rot in (False, True):
            for _ in range(3): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, fix_rot, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(5): ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, fix_rot, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr())
            e1.record(); torch.cuda.synchronize()
            ms=e0.elapsed_time(e1)/5
            print("phased",phased,"n_rot",n_rot,"fix_rot",fix_rot,"ms %.3f"%ms,"evals/s %.3e"%(n_rot*32*10000/ms*1e3), flush=True)
            res[(phased,fix_rot)]=(d_lb.cpu().numpy().copy(), d_ub.cpu().numpy().copy())
    ctx.close()
for f in (False,True):
    a,b=res[(0,f)],res[(1,f)]
    print("fix_rot",f,"lb equal",np.array_equal(a[0],b[0]),"ub equal",np.array_equal(a[1],b[1]),"max rel", float(np.max(np.abs(a[1]-b[1])/np.maximum(np.abs(a[1]),1e-9))))
PY
