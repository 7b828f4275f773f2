"""Generates tests/golden/reference_clouds.npz: outputs of the UNMODIFIED reference (oracle/_ref/libfgoicp_ref.so)
on the four cloud pairs of tests/golden/clouds_small.npz (subsamples of the reference repository's own bunny,
skull and dragon clouds, tests/golden/make_cloud_fixtures.py).  Run on the GPU box:

    python tests/golden/make_golden_clouds.py gpurun_out/reference_clouds.npz

and copy the result to tests/golden/.  tests/test_golden_clouds.py checks the oracle (CPU) and the CUDA path (GPU)
against these vectors."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import workloads  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as REF  # noqa: E402

PAIRS = ("bunny", "skull", "dragon", "overlap")
MSE = {"bunny": 1e-3, "skull": 1e-3, "dragon": 1e-3, "overlap": 1e-4}      # test/*.toml; W4: "tight"


def resolution_for(model, data):
    """Grid resolution that keeps the largest grid dimension near 80 nodes (the CPU oracle builds the grid by
    brute force); 3 significant digits so that it survives the round trip through the .npz exactly."""
    pp = O.preprocess(model, data)
    ext = float((pp["bbox_max"] - pp["bbox_min"]).max())
    return float("%.3g" % (ext / 80.0))


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_clouds.npz")
    C = np.load(os.path.join(ROOT, "tests", "golden", "clouds_small.npz"))
    g = {}
    for name in PAIRS:
        model, data = C[name + "_model"], C[name + "_data"]
        res = resolution_for(model, data)
        ref = REF.Reference(model, data, res, MSE[name])
        P = name + "_"
        g[P + "res"], g[P + "mse"] = np.float32(res), np.float32(MSE[name])
        rp = ref.preprocessed()
        for k in ("offset_pcs", "offset_pct", "bbox_min", "bbox_max"):
            g[P + "pre_" + k] = rp[k]
        g[P + "pre_scale"] = np.float32(rp["scale"])
        g[P + "pre_model_sha"], g[P + "pre_data_sha"] = sha(rp["model"]), sha(rp["data"])
        lut, dims = ref.lut()
        g[P + "lut_dims"], g[P + "lut_sha"], g[P + "lut_stride101"] = dims, sha(lut), lut[::101].copy()
        rng = np.random.default_rng(5)
        lo, hi = rp["bbox_min"] - 0.3, rp["bbox_max"] + 0.3
        q = (lo + rng.random((2000, 3)) * (hi - lo)).astype(np.float32)
        g[P + "tex_q"], g[P + "tex_val"] = q, ref.lut_sample(q)
        rots = np.float32([[0.25, -0.25, 0.25, 0.25], [0.0625, 0.1875, -0.0625, 0.0625], [-0.375, 0.125, 0.375, 0.125]])
        tcs = np.stack([workloads.translation_cube_list(32, level=1 + k, seed=60 + k) for k in range(3)])
        lbs, ubs = np.zeros((3, 2, 32), np.float32), np.zeros((3, 2, 32), np.float32)
        for r in range(3):
            for f in (0, 1):
                lbs[r, f], ubs[r, f] = ref.bounds(rots[r], bool(f), tcs[r])
        g[P + "bounds_rot"], g[P + "bounds_tc"], g[P + "bounds_lb"], g[P + "bounds_ub"] = rots, tcs, lbs, ubs
        I = np.eye(3, dtype=np.float32).ravel()
        z = np.zeros(3, np.float32)
        R1 = O.rotation(0.2, -0.1, 0.15)[0]
        t1 = np.float32([0.05, 0.02, -0.03])
        g[P + "sse_R"], g[P + "sse_t"] = np.stack([I, R1]), np.stack([z, t1])
        g[P + "sse_val"] = np.float32([ref.sse(I, z), ref.sse(R1, t1)])
        icp = np.zeros((2, 13), np.float32)
        for k, (R0, t0, thr) in enumerate([(I, z, 0.05), (R1, t1, 0.005)]):
            e, R, t = ref.icp(R0, t0, 100, thr)
            icp[k] = np.concatenate([[e], R, t])
        g[P + "icp_out"] = icp
        g[P + "sse_threshold"] = np.float32(ref.sse_threshold())
        sse, R, t, Rn, tn = ref.run()
        g[P + "run_sse"], g[P + "run_R"], g[P + "run_t"], g[P + "run_Rn"], g[P + "run_tn"] = np.float32(sse), R, t, Rn, tn
        print(name, "res", res, "dims", dims, "run sse", sse, "mse", sse / len(data), flush=True)
        ref.close()
    np.savez_compressed(out, **g)
    print("wrote", out)


if __name__ == "__main__":
    main()
