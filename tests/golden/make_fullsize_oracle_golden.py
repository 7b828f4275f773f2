"""Oracle-derived golden results of whole searches at full size (VERDICT r01, next #1b/c).

For each case the CPU oracle (oracle/fgoicp_oracle.c, pinned on the unmodified reference) is driven through the SAME
level-synchronous driver as the CUDA path (fast_go_icp_b200.driver with tests/oracle_context.OracleContext answering
every C-ABI call) and the result -- SSE bits, pose, evaluation / cube / refinement / iteration counts -- is written to
tests/golden/fullsize_oracle/<case>.json.  tests/test_fullsize_parity.py then requires the CUDA path to reproduce these
values bit for bit.  No GPU is involved in producing them.  Minutes per case on 8 cores:

    python tests/golden/make_fullsize_oracle_golden.py bunny_mse1e-3 bunny_mse1e-5 skull_mse1e-3 w5_mse1e-4 ...

Clouds: tests/golden/{bunny,dragon,skull,overlap}_full.npz (`model`, `data`; written by scripts/make_full_clouds.py +
this script's --clouds step from the reference repository's own files with the seeded loader); W5 is regenerated from
its seed (fast_go_icp_b200.workloads.synthetic_pair)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "bunny_mse1e-3": ("bunny", 0.005, 1e-3), "bunny_mse1e-5": ("bunny", 0.005, 1e-5),
    "skull_mse1e-3": ("skull", 0.005, 1e-3),
    "w5_mse1e-4": ("w5", 0.005, 1e-4),
    "dragon_mse1e-3": ("dragon", 0.005, 1e-3), "dragon_mse1e-4": ("dragon", 0.005, 1e-4),
    "overlap_mse1e-3": ("overlap", 0.005, 1e-3), "overlap_mse1e-4": ("overlap", 0.005, 1e-4),
    # trimmed registration (extension; fourth field = trim_fraction): the configurations bench.py times on W2 and W4
    "skull_trim0.1_mse1e-3": ("skull", 0.005, 1e-3, 0.1), "overlap_trim0.45_mse1e-4": ("overlap", 0.005, 1e-4, 0.45),
}


def load(pair):
    if pair == "w5":
        from fast_go_icp_b200 import workloads
        w = workloads.synthetic_pair()
        return w["model"], w["data"]
    z = np.load(os.path.join(HERE, pair + "_full.npz"))
    return z["model"], z["data"]


def write_clouds():
    """skull_full.npz / overlap_full.npz from build/workloads/repo_clouds_full.npz (scripts/make_full_clouds.py)."""
    z = np.load(os.path.join(ROOT, "build", "workloads", "repo_clouds_full.npz"))
    for pair in ("skull", "overlap"):
        out = dict(model=z[pair + "_model"], data=z[pair + "_data"], R_move=z[pair + "_R_move"], t_move=z[pair + "_t_move"])
        np.savez_compressed(os.path.join(HERE, pair + "_full.npz"), **out)
        print("wrote", pair + "_full.npz", out["model"].shape, out["data"].shape)


def run_case(name):
    from fast_go_icp_b200 import driver
    from oracle_context import OracleContext
    pair, res, mse = CASES[name][:3]
    trim = CASES[name][3] if len(CASES[name]) > 3 else 0.0
    model, data = load(pair)
    t0 = time.perf_counter()
    g = driver.FastGoICP(model, data, res, mse, ctx_factory=OracleContext, trim_fraction=trim)
    R, t = g.run()
    st = g.stats
    out = dict(case=name, pair=pair, nt=len(model), ns=len(data), lut_resolution=res, mse_threshold=mse, trim_fraction=trim,
               schedule="level", produced_by="CPU oracle through fast_go_icp_b200.driver (tests/oracle_context.OracleContext)",
               sse=float(g.best_sse), sse_bits=int(np.float32(g.best_sse).view(np.uint32)),
               R_bits=[int(x) for x in np.asarray(R, np.float32).ravel().view(np.uint32)],
               t_bits=[int(x) for x in np.asarray(t, np.float32).ravel().view(np.uint32)],
               R=np.asarray(R, np.float32).tolist(), t=np.asarray(t, np.float32).tolist(),
               bound_evals=int(st["bound_evals"]), rot_cubes=int(st["rot_cubes"]), icp_runs=int(st["icp_runs"]),
               icp_iters=int(st["icp_iters"]), cpu_seconds=time.perf_counter() - t0)
    g.close()
    with open(os.path.join(HERE, "fullsize_oracle", name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(name, "sse", repr(out["sse"]), "evals", out["bound_evals"], "cubes", out["rot_cubes"], "icp", out["icp_runs"], out["icp_iters"],
          "%.1f s" % out["cpu_seconds"], flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    if "--clouds" in args:
        write_clouds()
        args.remove("--clouds")
    for name in args:
        run_case(name)
