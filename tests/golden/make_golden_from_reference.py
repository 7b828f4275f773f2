"""Generates tests/golden/reference_small.npz by running the UNMODIFIED reference (oracle/_ref/libfgoicp_ref.so,
built by oracle/build_ref.py) on a B200: real buildLUTKernel, real tex3D, real kernComputeBounds + thrust
reductions, real ICP and real host branch-and-bound.  Run on the GPU box:

    python tests/golden/make_golden_from_reference.py gpurun_out/reference_small.npz

and copy the result to tests/golden/.  The CPU test-suite then checks the oracle against these vectors
(tests/test_golden.py), which pins the oracle to the reference itself rather than to our reading of it."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_problem  # noqa: E402
from fast_go_icp_b200 import workloads  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as REF  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_small.npz")
pp = make_problem()           # the same problem as the `small_problem` fixture
raw = pp["raw"]
MSE = 1e-4
ref = REF.Reference(raw["model"], raw["data"], float(pp["res"]), MSE)
g = {}
rp = ref.preprocessed()
for k in ("offset_pcs", "offset_pct", "bbox_min", "bbox_max"):
    g["pre_" + k] = rp[k]
g["pre_scale"] = np.float32(rp["scale"])
g["pre_model_sha"] = np.frombuffer(hashlib.sha256(rp["model"].tobytes()).digest(), np.uint8)
g["pre_data_sha"] = np.frombuffer(hashlib.sha256(rp["data"].tobytes()).digest(), np.uint8)
lut, dims = ref.lut()
g["lut_dims"] = dims
g["lut_sha"] = np.frombuffer(hashlib.sha256(lut.tobytes()).digest(), np.uint8)
g["lut_stride97"] = lut[::97].copy()
rng = np.random.default_rng(77)
q = rng.uniform(-1.2, 1.2, (3000, 3)).astype(np.float32)
g["tex_q"] = q
g["tex_val"] = ref.lut_sample(q)
rots = np.float32([[0.25, -0.25, 0.25, 0.25], [0.0625, 0.1875, -0.0625, 0.0625], [-0.375, 0.125, 0.375, 0.125]])
tcs = np.stack([workloads.translation_cube_list(32, level=2 + k, seed=40 + k) for k in range(3)])
g["bounds_rot"], g["bounds_tc"] = rots, tcs
lbs, ubs = np.zeros((3, 2, 32), np.float32), np.zeros((3, 2, 32), np.float32)
for r in range(3):
    for f in (0, 1):
        lbs[r, f], ubs[r, f] = ref.bounds(rots[r], bool(f), tcs[r])
g["bounds_lb"], g["bounds_ub"] = lbs, ubs
poses_R = np.stack([O.rotation(*v)[0] for v in np.float32([[0.2, 0.1, -0.1], [0, 0, 0], [-0.4, 0.3, 0.2]])])
poses_t = np.float32([[0.05, -0.02, 0.01], [0, 0, 0], [0.4, 0.3, -0.5]])
g["sse_R"], g["sse_t"] = poses_R, poses_t
g["sse_val"] = np.float32([ref.sse(poses_R[k], poses_t[k]) for k in range(3)])
icp_thr = np.float32([0.05, 0.005, 0.0005])
icp_out = np.zeros((3, 13), np.float32)
seeds_R = np.stack([np.eye(3, dtype=np.float32).ravel(), O.rotation(0.3, 0.1, -0.2)[0], O.rotation(-0.1, 0.05, 0.1)[0]])
seeds_t = np.float32([[0, 0, 0], [0.1, 0, -0.1], [0.02, 0.03, 0.0]])
for k in range(3):
    e, R, t = ref.icp(seeds_R[k], seeds_t[k], 100, float(icp_thr[k]))
    icp_out[k] = np.concatenate([[e], R, t])
g["icp_seed_R"], g["icp_seed_t"], g["icp_thr"], g["icp_out"] = seeds_R, seeds_t, icp_thr, icp_out
cubes = np.float32([[0.25, -0.25, 0.25, 0.25], [0.0625, 0.1875, -0.0625, 0.0625], [-0.5, 0.5, 0.5, 0.5],
                    [0.125, 0.125, 0.125, 0.125], [-0.1875, 0.0625, 0.3125, 0.0625]])
bnb = np.zeros((5, 2, 2, 4), np.float32)
for i, c in enumerate(cubes):
    for f in (0, 1):
        for j, bs in enumerate((1e10, 5.0)):
            ub, bt = ref.bnb_r3(c, bool(f), bs)
            bnb[i, f, j] = [ub, *bt]
g["bnb_cubes"], g["bnb_out"], g["bnb_best_sse"] = cubes, bnb, np.float32([1e10, 5.0])
g["sse_threshold"] = np.float32(ref.sse_threshold())
sse, R, t, Rn, tn = ref.run()
g["run_sse"], g["run_R"], g["run_t"], g["run_Rn"], g["run_tn"] = np.float32(sse), R, t, Rn, tn
g["mse_threshold"] = np.float32(MSE)
np.savez_compressed(out, **g)
print("wrote", out, {k: (v.shape if hasattr(v, "shape") else v) for k, v in g.items()})
print("run: sse", sse, "R", R, "t", t)
ref.close()
