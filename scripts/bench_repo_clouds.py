"""End-to-end run() on the reference repository's own clouds at the sizes SURVEY.md section 8d names (W1-W4) and on
W5, with two baselines timed beside it on the same box:
  * "reference": the UNMODIFIED reference's icp::FastGoICP::run() (oracle/_ref, its kernels on GPU 0, one host thread);
  * "cpu_port":  the CPU oracle (oracle/fgoicp_oracle.c, OpenMP on all host cores; the reference has no CPU path of
                 its own): grid build and ICP searches through its exact k-d tree, bounds and searches as restated;
                 ctor (preprocessing + grid build) and run() are timed separately, like main.cpp:46-53.
Baselines run in subprocesses under a wall-clock cap; a capped run is reported as a lower bound.

    python scripts/make_full_clouds.py                  # in the container that mounts /root/reference
    python scripts/bench_repo_clouds.py [--cap 40]      # on the GPU box -> gpurun_out/repo_clouds_r01.json
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLOUDS = os.path.join(ROOT, "build", "workloads", "repo_clouds_full.npz")

# name, pair, lut_resolution, mse_threshold, baselines?
CASES = [
    ("W1 bunny (test/bunny.toml res)", "bunny", 0.002, 1e-3, False),
    ("W1 bunny res 0.005", "bunny", 0.005, 1e-3, True),
    ("W1 bunny res 0.005 mse 1e-5", "bunny", 0.005, 1e-5, True),
    ("W2 skull", "skull", 0.005, 1e-3, False),
    ("W3 dragon", "dragon", 0.005, 1e-3, False),
    ("W3 dragon mse 1e-4", "dragon", 0.005, 1e-4, False),
    ("W4 partial overlap", "overlap", 0.005, 1e-4, False),
    ("W5 synthetic 100k/10k", "w5", 0.005, 1e-4, False),
]


def load_pair(pair):
    if pair == "w5":
        from fast_go_icp_b200 import workloads
        w = workloads.synthetic_pair()
        return w["model"], w["data"], w["R_true"], w["t_true"]
    z = np.load(CLOUDS)
    model, data = z[pair + "_model"], z[pair + "_data"]
    if pair + "_R_move" in z:            # data = R_move * x + t_move  =>  the registration should return its inverse
        Rm, tm = z[pair + "_R_move"], z[pair + "_t_move"]
        return model, data, Rm.T, -Rm.T @ tm
    return model, data, None, None


def pose_err(R, t, Rt, tt):
    if Rt is None:
        return None, None
    ang = float(np.degrees(np.arccos(np.clip((np.trace(np.asarray(R, np.float64) @ Rt.T) - 1) / 2, -1, 1))))
    return ang, float(np.linalg.norm(np.asarray(t, np.float64) - tt))


def ours(pair, res, mse, reps=3):
    from fast_go_icp_b200 import capi, driver
    model, data, Rt, tt = load_pair(pair)
    runs = []
    for _ in range(reps):
        g = driver.FastGoICP(model, data, res, mse, flags=capi.BUILD_PACKED)
        R, t = g.run()
        st = g.stats
        info = g.ctx.info()
        ang, terr = pose_err(R, t, Rt, tt)
        runs.append(dict(run_ms=st["run_ms"], ctor_ms=st["ctor_ms"], lut_build_ms=st["lut_build_ms"], sse=float(g.best_sse),
                         mse=float(g.best_sse) / len(data), bound_evals=int(st["bound_evals"]), rot_cubes=int(st["rot_cubes"]),
                         icp_runs=int(st["icp_runs"]), icp_iters=int(st["icp_iters"]), ms_bnb_ub=st["ms_bnb_ub"], ms_icp=st["ms_icp"], ms_bnb_lb=st["ms_bnb_lb"],
                         rot_err_deg=ang, t_err=terr, R=np.asarray(R).tolist(), t=np.asarray(t).tolist(),
                         grid_dims=list(info.dims), packed_bytes=int(info.packed_bytes)))
        g.close()
    runs.sort(key=lambda r: r["run_ms"])
    med = runs[len(runs) // 2]
    med["run_ms_all"] = [r["run_ms"] for r in runs]
    search_ms = med["ms_bnb_ub"] + med["ms_bnb_lb"]
    med["search_evals_per_s"] = med["bound_evals"] / (search_ms * 1e-3) if search_ms > 0 else None
    return med


def baseline_child(kind, pair, res, mse):
    """Runs in a subprocess (so that it can be capped): prints one JSON line."""
    model, data, Rt, tt = load_pair(pair)
    if kind == "reference":
        from oracle import ref as REF
        t0 = time.perf_counter()
        r = REF.Reference(model, data, res, mse)
        ctor_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps(dict(stage="ctor", ctor_ms=ctor_ms)), flush=True)
        t0 = time.perf_counter()
        sse, R, t, _, _ = r.run()
        run_ms = (time.perf_counter() - t0) * 1e3
        R = np.asarray(R).reshape(3, 3).T
        ang, terr = pose_err(R, t, Rt, tt)
        print(json.dumps(dict(stage="run", run_ms=run_ms, ctor_ms=ctor_ms, sse=float(sse), rot_err_deg=ang, t_err=terr,
                              host_threads=1)), flush=True)
    else:
        from fast_go_icp_b200 import capi, driver
        from oracle import oracle as O
        t0 = time.perf_counter()
        pp = driver.preprocess(model, data)
        lut, dims = O.lut_build(pp["model"], pp["bbox_min"], pp["bbox_max"], res)      # k-d tree build, all host threads
        ctor_ms = (time.perf_counter() - t0) * 1e3
        try:    # the reference build's sin(half-angle) constants (tests/golden/reference_sin.json)
            j = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_sin.json")))
            O.set_sin_table(np.array(j["spans"], np.float32), np.array(j["reference_build_bits"], np.uint32).view(np.float32))
        except Exception:
            pass
        print(json.dumps(dict(stage="ctor", threads=O.num_threads(), ctor_ms=ctor_ms)), flush=True)
        t0 = time.perf_counter()
        sse, R, t, st = O.run(pp["model"], pp["data"], lut, dims, pp["bbox_min"], res, mse)
        run_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps(dict(stage="run", run_ms=run_ms, sse=float(sse), bound_evals=st["evals"],
                              rot_cubes=st["cubes"], icp_runs=st["icps"], host_threads=O.num_threads())), flush=True)


def baseline(kind, pair, res, mse, cap):
    cmd = [sys.executable, os.path.abspath(__file__), "--child", kind, pair, repr(res), repr(mse)]
    t0 = time.perf_counter()
    out, timed_out = "", False
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=cap)
        out = p.stdout
        if p.returncode != 0:
            return dict(error=p.stderr[-400:])
    except subprocess.TimeoutExpired as e:
        timed_out = True
        out = (e.stdout or b"").decode() if isinstance(e.stdout, (bytes, bytearray)) else (e.stdout or "")
    res_d = {}
    for line in out.splitlines():
        try:
            res_d.update(json.loads(line))
        except ValueError:
            pass
    if timed_out:
        res_d["capped"] = True
        res_d["run_ms_at_least"] = (time.perf_counter() - t0) * 1e3 - res_d.get("ctor_ms", 0.0)
    return res_d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", nargs=4)
    ap.add_argument("--cap", type=float, default=35.0, help="wall-clock cap per baseline run, seconds")
    ap.add_argument("--no-baselines", action="store_true")
    ap.add_argument("--skip", default=None, help="comma-separated substrings of case names to leave out")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="repo_clouds_r01.json")
    ap.add_argument("--only", default=None, help="substring filter on the case name")
    args = ap.parse_args()
    if args.child:
        kind, pair, res, mse = args.child
        baseline_child(kind, pair, float(res), float(mse))
        return
    from fast_go_icp_b200 import capi, driver, workloads
    ws = workloads.synthetic_pair(nt=3000, ns=400, seed=3)          # untimed warm-up: loads every kernel of the search
    gw = driver.FastGoICP(ws["model"], ws["data"], 0.03, 1e-4, flags=capi.BUILD_PACKED)
    gw.run()
    gw.close()
    rows = []
    for name, pair, res, mse, with_base in CASES:
        if args.only and not any(tok in name for tok in args.only.split(",")):
            continue
        if args.skip and any(tok in name for tok in args.skip.split(",")):
            continue
        model, data, _, _ = load_pair(pair)
        row = dict(case=name, nt=len(model), ns=len(data), lut_resolution=res, mse_threshold=mse, ours=ours(pair, res, mse, args.reps))
        if with_base and not args.no_baselines:
            row["reference"] = baseline("reference", pair, res, mse, args.cap)
            row["cpu_port"] = baseline("cpu_port", pair, res, mse, args.cap)
        rows.append(row)
        o = row["ours"]
        print("%-32s nt %6d ns %5d | ours run %8.1f ms (icp %7.1f ms, %d runs, %d iters) ctor %7.1f ms mse %.3e evals %.2e | ref %s | cpu %s"
              % (name, row["nt"], row["ns"], o["run_ms"], o["ms_icp"], o["icp_runs"], o["icp_iters"], o["ctor_ms"], o["mse"], o["bound_evals"],
                 json.dumps({k: v for k, v in row.get("reference", {}).items() if k in ("run_ms", "run_ms_at_least", "sse", "error")}),
                 json.dumps({k: v for k, v in row.get("cpu_port", {}).items() if k in ("run_ms", "run_ms_at_least", "sse", "error")})),
              flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", args.out), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
