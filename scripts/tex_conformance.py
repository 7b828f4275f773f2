"""Pins the manual grid filter against the B200 texture unit: for each (weight rounding, interpolation
formula) candidate the oracle's filter is compared with tex3D<float> (cudaFilterModeLinear, clamp,
unnormalised) on random queries.  Writes profiles/tex_conformance_rNN.json."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402
from oracle import oracle as O  # noqa: E402


def ulps(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


def main(out_path):
    w = workloads.synthetic_pair(nt=20000, ns=1000, seed=3)
    pp = driver.preprocess(w["model"], w["data"])
    res = 0.02
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], res, flags=capi.BUILD_PACKED | capi.BUILD_TEX)
    lut, dims = ctx.lut_download()
    rng = np.random.default_rng(0)
    q = rng.uniform(-1.2, 1.2, (400000, 3)).astype(np.float32)
    tex = ctx.lut_sample(q, capi.SAMPLER_TEX)
    man = ctx.lut_sample(q, capi.SAMPLER_GRID)
    report = {"n": len(q), "dims": [int(d) for d in dims], "candidates": {}}
    for wm in (0, 1):
        for im in (0, 1):
            O.set_modes(wm, im)
            o = O.lut_sample(lut, dims, pp["bbox_min"], res, q)
            u = ulps(o, tex)
            rel = np.abs(o - tex) / np.maximum(np.abs(tex), 1e-12)
            report["candidates"]["weights=%s,interp=%s" % ("nearest" if wm == 0 else "trunc", "lerp" if im == 0 else "wsum")] = {
                "bit_exact_frac": float(np.mean(u == 0)), "within_1ulp_frac": float(np.mean(u <= 1)),
                "within_4ulp_frac": float(np.mean(u <= 4)), "max_rel": float(rel.max()), "p999_rel": float(np.quantile(rel, 0.999)),
                "median_rel": float(np.median(rel))}
    O.set_modes(0, 0)
    o = O.lut_sample(lut, dims, pp["bbox_min"], res, q)
    report["manual_kernel_equals_oracle_default"] = bool(np.array_equal(o, man))
    # texel-centre probes: u = i + 0.5 must return T[i] exactly under any weight rule
    json.dump(report, open(out_path, "w"), indent=1)
    print(json.dumps(report, indent=1))
    ctx.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tex_conformance.json")
