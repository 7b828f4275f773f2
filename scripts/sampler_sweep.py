"""Times the fused bound kernel with each grid sampler (dense 8-gather, corner-packed 256-bit gather,
hardware tex3D) on the W5 workload; prints evals/s and algorithmic GB/s.  Device-resident inputs,
CUDA events on the launching stream."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402


def main(out_path, nt=100_000, ns=10_000, res=0.005, n_rot=2048, T=32, steps=5):
    w = workloads.synthetic_pair(nt=nt, ns=ns, seed=1234)
    pp = driver.preprocess(w["model"], w["data"])
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], res, flags=capi.BUILD_PACKED | capi.BUILD_TEX)
    info = ctx.info()
    rot, tc = workloads.bound_microbench(n_rot, T, seed=7)
    dev = torch.device("cuda", 0)
    d_rot, d_tc = torch.from_numpy(rot).to(dev), torch.from_numpy(tc).to(dev)
    d_lb, d_ub = torch.empty(n_rot, T, device=dev), torch.empty(n_rot, T, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    out = {"dims": list(info.dims), "grid_MB": info.grid_bytes / 1e6, "packed_MB": info.packed_bytes / 1e6,
           "build_ms": info.build_ms, "n_rot": n_rot, "T": T, "ns": ns, "results": []}
    ref = None
    for name, s in (("grid", capi.SAMPLER_GRID), ("packed", capi.SAMPLER_PACKED), ("tex", capi.SAMPLER_TEX)):
        ctx.set_sampler(s)
        for fix_rot in (False, True):
            for _ in range(3):
                ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, fix_rot, d_tc.data_ptr(), T, d_lb.data_ptr(), d_ub.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                ctx.bounds_multi_dev(d_rot.data_ptr(), n_rot, fix_rot, d_tc.data_ptr(), T, d_lb.data_ptr(), d_ub.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            evals = n_rot * T * ns
            r = {"sampler": name, "fix_rot": fix_rot, "ms": ms, "evals_per_s": evals / ms * 1e3, "algo_GBps": 32 * evals / ms / 1e6}
            if name == "grid" and not fix_rot:
                ref = d_ub.clone()
            if not fix_rot and ref is not None and name != "grid":
                r["max_rel_vs_grid"] = float(((d_ub - ref).abs() / ref.abs().clamp_min(1e-9)).max().item())
            out["results"].append(r)
            print(json.dumps(r), flush=True)
    json.dump(out, open(out_path, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sampler_sweep.json")
