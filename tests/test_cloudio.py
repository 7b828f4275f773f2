"""N1: loaders, config and result artefacts (fast_go_icp_b200/cloudio.py) -- CPU only."""
import tomllib

import numpy as np
import pytest

from fast_go_icp_b200 import cloudio


def _write_txt(path, pts):
    with open(path, "w") as f:
        f.write("%d\n" % len(pts))
        np.savetxt(f, pts, fmt="%.6f")


def test_txt_and_ply_loaders_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
    _write_txt(tmp_path / "a.txt", pts)
    got = cloudio.read_txt(str(tmp_path / "a.txt"))
    assert got.shape == (500, 3) and np.allclose(got, pts, atol=1e-6)
    cloudio.write_ply(str(tmp_path / "a.ply"), pts)
    assert np.allclose(cloudio.read_ply(str(tmp_path / "a.ply")), pts, atol=1e-6)
    # binary little-endian with extra vertex properties (the layout of data/artec3d/data_skull.ply)
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    a = np.zeros(500, dt)
    a["x"], a["y"], a["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    with open(tmp_path / "b.ply", "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 500\nproperty float x\nproperty float y\n"
                b"property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
        f.write(a.tobytes())
    assert np.array_equal(cloudio.read_ply(str(tmp_path / "b.ply")), pts)
    with pytest.raises(ValueError):
        cloudio.load_cloud(str(tmp_path / "a.xyz"))


def test_subsampling_is_seeded_and_follows_the_reference_rule():
    pts = np.arange(3000, dtype=np.float32).reshape(1000, 3)
    a = cloudio.subsample(pts, 0.3, seed=7)
    b = cloudio.subsample(pts, 0.3, seed=7)
    c = cloudio.subsample(pts, 0.3, seed=8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert len(a) <= int(np.float32(1000) * np.float32(0.3))          # never more than floor(n p) points
    assert np.all(np.diff(a[:, 0]) > 0)                                 # file order kept
    assert len(cloudio.subsample(pts, 1.0, seed=0)) == 1000             # u <= 1 always accepts


def test_config_keys_defaults_and_clamps(tmp_path):
    (tmp_path / "c.toml").write_text('[info]\nversion = "0.2"\n[io]\ntarget = "t.txt"\nsource = "s.txt"\n'
                                     'output = "out.toml"\nvisualization = "viz.ply"\n[params]\nmode = 4\ntrim = true\n'
                                     'target_subsample = 0.5\nsource_subsample = 0.9\nlut_resolution = 0.002\nmse_threshold = 0.0\n')
    cfg = cloudio.Config(str(tmp_path / "c.toml"))
    p = cfg.params
    assert (cfg.target, cfg.source, cfg.output, cfg.visualization) == ("t.txt", "s.txt", "out.toml", "viz.ply")
    assert p.trim is True and p.trim_fraction == 0.0 and p.mode == 4
    assert p.target_subsample == 0.5 and p.source_subsample == 0.5      # source clamped to <= 0.5 (utilities.hpp:103)
    assert abs(p.lut_resolution - 0.002) < 1e-9 and p.mse_threshold == float(np.float32(1e-12))
    (tmp_path / "d.toml").write_text('[io]\ntarget = "t.txt"\nsource = "s.txt"\n')
    p = cloudio.Config(str(tmp_path / "d.toml")).params
    assert (p.target_subsample, p.source_subsample) == (1.0, 0.5) and abs(p.mse_threshold - 1e-3) < 1e-9
    assert abs(p.lut_resolution - 0.005) < 1e-9 and p.trim is False


def test_result_artefacts(tmp_path):
    R = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    t = np.array([0.5, -0.25, 2.0])
    cloudio.write_result_toml(str(tmp_path / "out.toml"), R, t, 1.5e-4, 1.5, {"seconds": 0.2, "bound_evals": 7})
    got = tomllib.loads((tmp_path / "out.toml").read_text())["result"]
    assert np.allclose(got["R"], R) and np.allclose(got["t"], t) and got["mse"] == 1.5e-4 and got["bound_evals"] == 7
