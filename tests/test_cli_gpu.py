"""N1 end to end on the GPU: config -> seeded loaders -> run() -> `output` TOML and `visualization` PLY."""
import tomllib

import numpy as np
import pytest

from fast_go_icp_b200 import cli, cloudio, workloads

pytestmark = pytest.mark.gpu


def test_cli_writes_result_and_visualisation(tmp_path):
    w = workloads.synthetic_pair(nt=6000, ns=1600, sigma=0.005, seed=9)
    for name, pts in (("model.txt", w["model"]), ("data.txt", w["data"])):
        with open(tmp_path / name, "w") as f:
            f.write("%d\n" % len(pts))
            np.savetxt(f, pts, fmt="%.7f")
    (tmp_path / "cfg.toml").write_text(
        '[io]\ntarget = "%s"\nsource = "%s"\noutput = "%s"\nvisualization = "%s"\n'
        '[params]\ntrim = true\ntarget_subsample = 0.5\nsource_subsample = 0.25\nlut_resolution = 0.03\n'
        'mse_threshold = 1e-4\nseed = 3\n' % (tmp_path / "model.txt", tmp_path / "data.txt", tmp_path / "out.toml", tmp_path / "viz.ply"))
    assert cli.main(["-c", str(tmp_path / "cfg.toml")]) == 0
    res = tomllib.loads((tmp_path / "out.toml").read_text())["result"]
    R, t = np.array(res["R"]), np.array(res["t"])
    ang = np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 3.0 and np.linalg.norm(t - w["t_true"]) < 0.05 and res["mse"] < 2e-3
    viz = cloudio.read_ply(str(tmp_path / "viz.ply"))
    src = cloudio.load_cloud(str(tmp_path / "data.txt"), 0.25, 4)                 # the CLI seeds the source with seed + 1
    assert len(viz) == len(src) and np.allclose(viz, src @ R.T + t, atol=1e-4)
    # same config, same seed -> the same clouds -> the same result, bit for bit
    assert cli.main(["-c", str(tmp_path / "cfg.toml")]) == 0
    assert tomllib.loads((tmp_path / "out.toml").read_text())["result"]["sse"] == res["sse"]


def test_cli_passes_trim_fraction_to_the_search(tmp_path):
    """`[params] trim_fraction` reaches the driver: with 25 % gross outliers in the source cloud the trimmed run
    recovers the pose and reports the MSE over the inliers (the CLI used to drop the key)."""
    w = workloads.synthetic_pair(nt=6000, ns=1200, sigma=0.004, seed=21)
    rng = np.random.default_rng(5)
    data = w["data"].copy()
    bad = rng.choice(len(data), len(data) // 4, replace=False)
    data[bad] = rng.uniform(-1.0, 1.0, (len(bad), 3)).astype(np.float32)
    for name, pts in (("model.txt", w["model"]), ("data.txt", data)):
        with open(tmp_path / name, "w") as f:
            f.write("%d\n" % len(pts))
            np.savetxt(f, pts, fmt="%.7f")
    (tmp_path / "cfg.toml").write_text(
        '[io]\ntarget = "%s"\nsource = "%s"\noutput = "%s"\n'
        '[params]\ntrim = true\ntrim_fraction = 0.3\nsource_subsample = 0.5\nlut_resolution = 0.03\nmse_threshold = 1e-4\n'
        % (tmp_path / "model.txt", tmp_path / "data.txt", tmp_path / "out.toml"))
    assert cli.main(["-c", str(tmp_path / "cfg.toml")]) == 0
    res = tomllib.loads((tmp_path / "out.toml").read_text())["result"]
    src = cloudio.load_cloud(str(tmp_path / "data.txt"), 0.5, 1)
    assert res["trim_fraction"] == 0.3 and res["inliers"] == len(src) - int(np.float32(len(src)) * np.float32(0.3))
    assert abs(res["mse"] - res["sse"] / res["inliers"]) <= 1e-6 * res["mse"]
    R, t = np.array(res["R"]), np.array(res["t"])
    ang = np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 3.0 and np.linalg.norm(t - w["t_true"]) < 0.05
