"""The drop-in C++ class icp::FastGoICP on the GPU (reference fgoicp/fgoicp.hpp:10-108, used as src/main.cpp:46-53
uses it) through build/fgoicp_harness: against the Python mirror of the same driver, with device-side preprocessing,
and under the reference's own best-first schedule."""
import os
import subprocess

import numpy as np
import pytest

from fast_go_icp_b200 import build_harness, driver, workloads

pytestmark = pytest.mark.gpu


def _run_harness(tmp_path, w, res, mse, **env):
    exe = build_harness.build()
    w["model"].astype(np.float32).tofile(tmp_path / "model.f32")
    w["data"].astype(np.float32).tofile(tmp_path / "data.f32")
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    out = subprocess.run([exe, str(tmp_path / "model.f32"), str(tmp_path / "data.f32"), repr(res), repr(mse)],
                         capture_output=True, text=True, env=e, timeout=300)
    assert out.returncode == 0, out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()[1:]
    v = np.array([float.fromhex(x) for x in line])
    return dict(R=v[:9].astype(np.float32).reshape(3, 3).T, t=v[9:12].astype(np.float32), sse=np.float32(v[12]),
                scale=np.float32(v[13]))


@pytest.fixture(scope="module")
def problem():
    return workloads.synthetic_pair(nt=5000, ns=800, sigma=0.005, seed=12, max_angle=1.2)


def test_cpp_class_matches_python_driver(tmp_path, problem):
    w = problem
    cpp = _run_harness(tmp_path, w, 0.02, 1e-4)
    g = driver.FastGoICP(w["model"], w["data"], 0.02, 1e-4)
    R, t = g.run()
    # same C-ABI calls in the same order from both hosts: the same bits are expected; the stated tolerance is
    # BASELINE.json's (MSE 1e-6 relative, pose far inside the BnB leaf size)
    assert abs(float(cpp["sse"]) - float(g.best_sse)) <= 1e-6 * float(g.best_sse)
    assert np.allclose(cpp["R"], R, atol=1e-6) and np.allclose(cpp["t"], t, atol=1e-5)
    assert cpp["scale"] == np.float32(g.pp["scale"])
    g.close()
    ang = np.degrees(np.arccos(np.clip((np.trace(cpp["R"] @ w["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 2.0 and np.linalg.norm(cpp["t"] - w["t_true"]) < 0.03


def test_cpp_class_device_preprocess_is_bit_identical(tmp_path, problem):
    a = _run_harness(tmp_path, problem, 0.02, 1e-4)
    b = _run_harness(tmp_path, problem, 0.02, 1e-4, FGOICP_DEVICE_PREPROCESS=1)
    assert np.array_equal(a["R"], b["R"]) and np.array_equal(a["t"], b["t"]) and a["sse"] == b["sse"] and a["scale"] == b["scale"]


def test_cpp_class_frontier_sharded_over_two_gpus(tmp_path, problem):
    """FGOICP_DEVICES=0,1: one context per GPU inside the C++ class, waves dealt over them by host threads (the sharding
    logic itself is tested on the CPU in tests/test_cpp_host_cpu.py).  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    a = _run_harness(tmp_path, problem, 0.02, 1e-4)
    b = _run_harness(tmp_path, problem, 0.02, 1e-4, FGOICP_DEVICES="0,1")
    assert np.array_equal(a["R"], b["R"]) and np.array_equal(a["t"], b["t"]) and a["sse"] == b["sse"]


def test_cpp_class_reference_schedule(tmp_path, problem):
    """FGOICP_SCHEDULE=bestfirst: the reference's own visiting order (fgoicp.cpp:32-100), one cube at a time.  It stops
    as soon as best_sse - lb <= sse_threshold, so it agrees with the level schedule within that threshold."""
    a = _run_harness(tmp_path, problem, 0.02, 1e-4)
    b = _run_harness(tmp_path, problem, 0.02, 1e-4, FGOICP_SCHEDULE="bestfirst")
    thr = len(problem["data"]) * 1e-4
    assert abs(float(a["sse"]) - float(b["sse"])) <= thr
    ang = np.degrees(np.arccos(np.clip((np.trace(b["R"] @ problem["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 2.0


def test_cpp_class_reports_errors_as_exceptions(tmp_path, problem):
    exe = build_harness.build()
    np.zeros((0, 3), np.float32).tofile(tmp_path / "empty.f32")
    problem["data"].astype(np.float32).tofile(tmp_path / "data.f32")
    out = subprocess.run([exe, str(tmp_path / "empty.f32"), str(tmp_path / "data.f32"), "0.02", "1e-4"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "ERROR" in out.stderr


def test_progress_accessors_polled_during_run_on_the_gpu(tmp_path, problem):
    """SURVEY.md 8f N4 (reference fgoicp.hpp:32-43): a second host thread polls the visualisation accessors while run()
    drives the GPU; every (SSE, R, t) it sees is a published triple, never a torn mix, and the last one is final."""
    exe = build_harness.build()
    problem["model"].astype(np.float32).tofile(tmp_path / "model.f32")
    problem["data"].astype(np.float32).tofile(tmp_path / "data.f32")
    out = subprocess.run([exe, str(tmp_path / "model.f32"), str(tmp_path / "data.f32"), "0.02", "1e-4"], capture_output=True,
                         text=True, env=dict(os.environ, HARNESS_POLL="1"), timeout=300)
    assert out.returncode == 0, out.stderr
    f = [l for l in out.stdout.splitlines() if l.startswith("POLL")][-1].split()
    polls, distinct, published, torn, final_ok = int(f[2]), int(f[4]), int(f[6]), int(f[8]), int(f[10])
    assert polls >= 1 and published >= 3 and 1 <= distinct <= published      # how often the poller gets to run is up to the host scheduler
    assert torn == 0 and final_ok == 1


def test_unchanged_reference_cli_binary_runs_the_search(tmp_path, problem):
    """The reference's own src/main.cpp + src/utilities.hpp, compiled UNCHANGED against include/fgoicp/*.hpp and this
    library (fast_go_icp_b200/build_cli.py -> build/fast-go-icp; built where /root/reference is mounted, travels to the GPU
    box): config TOML -> loaders -> icp::FastGoICP -> run() (src/main.cpp:38-55).  Its log carries the final error."""
    import re
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "fast-go-icp")
    if not os.path.exists(exe):
        pytest.skip("build/fast-go-icp not built (needs the reference tree at build time)")
    for name, pts in (("model.txt", problem["model"]), ("data.txt", problem["data"])):
        with open(tmp_path / name, "w") as f:
            f.write("%d\n" % len(pts))
            np.savetxt(f, pts, fmt="%.7f")
    (tmp_path / "cfg.toml").write_text(
        '[info]\nversion = "1.0.0"\n[io]\ntarget = "%s"\nsource = "%s"\n'
        '[params]\ntrim = false\ntarget_subsample = 1.0\nsource_subsample = 0.5\nlut_resolution = 0.02\nmse_threshold = 1e-4\n'
        % (tmp_path / "model.txt", tmp_path / "data.txt"))
    out = subprocess.run([exe, "-c", str(tmp_path / "cfg.toml")], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    log = out.stdout + out.stderr
    m = re.search(r"Searching over! Best Error: ([0-9.eE+-]+)", log)
    assert m, log[-1500:]
    sse = float(m.group(1))
    # source_subsample = 0.5 of 800 points with the reference's unseeded sampler: about 400 points at sigma 0.005
    assert 0.0 < sse < 400 * 3 * (0.005 ** 2) * 4
    assert "Fast Go-ICP finished" in log or "finished" in log.lower()
