"""The drop-in C++ class icp::FastGoICP (include/fgoicp/fgoicp.hpp, csrc/fgoicp_host.cpp) WITHOUT a GPU: the same host
driver source, linked against an oracle-backed stand-in of the C ABI (tests/cpp/oracle_abi.c, test infrastructure), must
make the same decisions as the Python mirror of the driver over the oracle -- and, on the reference repository's bunny
pair at full size, land on the oracle-derived golden bits the CUDA path is held to (tests/golden/fullsize_oracle/)."""
import os
import subprocess

import numpy as np
import pytest

import cpu_harness
from fast_go_icp_b200 import driver, workloads
from oracle import oracle as O
from oracle_context import OracleContext

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bunny_full.npz")


def _run_cpp(tmp_path, model, data, res, mse, **env):
    exe = cpu_harness.build_cpu()
    np.ascontiguousarray(model, np.float32).tofile(tmp_path / "model.f32")
    np.ascontiguousarray(data, np.float32).tofile(tmp_path / "data.f32")
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    out = subprocess.run([exe, str(tmp_path / "model.f32"), str(tmp_path / "data.f32"), repr(res), repr(mse)],
                         capture_output=True, text=True, env=e, timeout=900)
    assert out.returncode == 0, out.stderr
    v = np.array([float.fromhex(x) for x in [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()[1:]])
    return dict(R=v[:9].astype(np.float32).reshape(3, 3).T, t=v[9:12].astype(np.float32), sse=np.float32(v[12]),
                scale=np.float32(v[13]), evals=int(v[16]), cubes=int(v[17]), icps=int(v[18]))


@pytest.fixture(scope="module")
def problem():
    return workloads.synthetic_pair(nt=1500, ns=200, sigma=0.005, seed=12, max_angle=1.2)


def test_cpp_level_schedule_equals_python_driver(tmp_path, problem):
    w = problem
    cpp = _run_cpp(tmp_path, w["model"], w["data"], 0.03, 1e-4)
    g = driver.FastGoICP(w["model"], w["data"], 0.03, 1e-4, ctx_factory=OracleContext)
    R, t = g.run()
    assert cpp["sse"] == np.float32(g.best_sse) and cpp["scale"] == np.float32(g.pp["scale"])
    assert np.array_equal(cpp["R"], np.asarray(R, np.float32)) and np.array_equal(cpp["t"], np.asarray(t, np.float32))
    assert [cpp["evals"], cpp["cubes"], cpp["icps"]] == [g.stats["bound_evals"], g.stats["rot_cubes"], g.stats["icp_runs"]]
    g.close()
    ang = np.degrees(np.arccos(np.clip((np.trace(cpp["R"] @ w["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 2.0 and np.linalg.norm(cpp["t"] - w["t_true"]) < 0.03


@pytest.mark.parametrize("devices", ["0,1", "0,1,2,3,4"])
def test_cpp_frontier_sharded_over_devices_changes_no_bit(tmp_path, problem, devices):
    """FGOICP_DEVICES: one context per device, every wave of a level dealt round-robin over them by host threads, the new
    incumbent taken as the MIN over (sse, cube index).  Same pose, SSE and total counts as one device (the stand-in's
    "devices" are all the CPU oracle: this checks the sharding and merging logic of the host driver)."""
    w = problem
    one = _run_cpp(tmp_path, w["model"], w["data"], 0.03, 1e-4)
    exe = cpu_harness.build_cpu()
    e = dict(os.environ, FGOICP_DEVICES=devices, ORACLE_ABI_TRACE="1")
    out = subprocess.run([exe, str(tmp_path / "model.f32"), str(tmp_path / "data.f32"), "0.03", "0.0001"], capture_output=True, text=True,
                         env=e, timeout=900)
    assert out.returncode == 0, out.stderr
    assert out.stderr.count("oracle_abi: context") == len(devices.split(","))
    many = _run_cpp(tmp_path, w["model"], w["data"], 0.03, 1e-4, FGOICP_DEVICES=devices)
    assert many["sse"] == one["sse"] and np.array_equal(many["R"], one["R"]) and np.array_equal(many["t"], one["t"])
    assert [many["evals"], many["cubes"], many["icps"]] == [one["evals"], one["cubes"], one["icps"]]


def test_cpp_reference_schedule_agrees_with_the_oracle_run(tmp_path, problem):
    """FGOICP_SCHEDULE=bestfirst is the reference's own order (fgoicp.cpp:32-100), which the oracle's orc_run restates:
    same registration (the heaps may order exact ties differently, so the comparison is the stated tolerance, not bits)."""
    w = problem
    cpp = _run_cpp(tmp_path, w["model"], w["data"], 0.03, 1e-4, FGOICP_SCHEDULE="bestfirst")
    pp = O.preprocess(w["model"], w["data"])
    lut, dims = O.lut_build(pp["model"], pp["bbox_min"], pp["bbox_max"], 0.03)
    e, R, t, st = O.run(pp["model"], pp["data"], lut, dims, pp["bbox_min"], 0.03, 1e-4)
    assert abs(float(cpp["sse"]) - float(e)) <= 1e-6 * float(e)
    t_orig = O.restore_translation(R, t, pp["scale"], pp["offset_pcs"], pp["offset_pct"])
    assert np.allclose(cpp["R"], np.asarray(R, np.float32).reshape(3, 3).T, atol=1e-6) and np.allclose(cpp["t"], t_orig, atol=1e-5)


def test_cpp_class_on_the_full_bunny_pair_lands_on_the_golden_bits(tmp_path):
    import json
    z = np.load(GOLD)
    with open(os.path.join(os.path.dirname(GOLD), "fullsize_oracle", "bunny_mse1e-3.json")) as f:
        want = json.load(f)
    cpp = _run_cpp(tmp_path, z["model"], z["data"], 0.005, 1e-3)
    bits = lambda a: [int(x) for x in np.asarray(a, np.float32).ravel().view(np.uint32)]
    assert int(cpp["sse"].view(np.uint32)) == want["sse_bits"]
    assert bits(cpp["R"]) == want["R_bits"] and bits(cpp["t"]) == want["t_bits"]
    assert [cpp["evals"], cpp["cubes"], cpp["icps"]] == [want["bound_evals"], want["rot_cubes"], want["icp_runs"]]


def test_cpp_class_reports_a_missing_gpu_for_device_preprocessing(tmp_path, problem):
    exe = cpu_harness.build_cpu()
    problem["model"].astype(np.float32).tofile(tmp_path / "m.f32")
    problem["data"].astype(np.float32).tofile(tmp_path / "d.f32")
    out = subprocess.run([exe, str(tmp_path / "m.f32"), str(tmp_path / "d.f32"), "0.03", "1e-4"], capture_output=True, text=True,
                         env=dict(os.environ, FGOICP_DEVICE_PREPROCESS="1"), timeout=120)
    assert out.returncode == 1 and "no CUDA device" in out.stderr


def test_progress_accessors_polled_from_a_second_thread_return_published_triples(tmp_path, problem):
    """SURVEY.md 8f N4 (reference fgoicp.hpp:32-43, the viewer's accessors): a second thread polls get_best_snapshot(),
    get_best_error(), get_best_transform(), get_last_transform() while run() works.  Every (SSE, R, t) it sees must be
    one of the triples the search published -- never a mix of two -- and the last one is what the accessors return."""
    w = problem
    exe = cpu_harness.build_cpu()
    np.ascontiguousarray(w["model"], np.float32).tofile(tmp_path / "model.f32")
    np.ascontiguousarray(w["data"], np.float32).tofile(tmp_path / "data.f32")
    out = subprocess.run([exe, str(tmp_path / "model.f32"), str(tmp_path / "data.f32"), "0.03", "0.0001"], capture_output=True,
                         text=True, env=dict(os.environ, HARNESS_POLL="1"), timeout=900)
    assert out.returncode == 0, out.stderr
    f = [l for l in out.stdout.splitlines() if l.startswith("POLL")][-1].split()
    polls, distinct, published, torn, final_ok = int(f[2]), int(f[4]), int(f[6]), int(f[8]), int(f[10])
    assert polls >= 1 and published >= 3 and 1 <= distinct <= published      # how often the poller gets to run is up to the host scheduler
    assert torn == 0 and final_ok == 1
