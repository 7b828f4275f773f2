"""CPU tests of the N>1 path: world_size-2 gloo process group, frontier sharded round-robin, MIN all-reduce
of the best SSE key, pose broadcast, all-gather of per-cube results.  The CUDA context is replaced by the
oracle-backed stand-in (tests/oracle_context.py); the sharded run must reproduce the single-process run
bit for bit, and both must recover the synthetic pose."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _problem():
    from fast_go_icp_b200 import workloads
    return workloads.synthetic_pair(nt=700, ns=120, sigma=0.005, seed=21, max_angle=0.9)


def _run(group_ready, trim=0.0):
    from fast_go_icp_b200 import driver
    from oracle import oracle as O
    from oracle_context import OracleContext
    O.set_num_threads(2)
    w = _problem()
    data = w["data"]
    if trim > 0:
        # gross outliers that only a trimmed registration ignores
        rng = np.random.default_rng(5)
        data = data.copy()
        bad = rng.choice(len(data), 24, replace=False)
        data[bad] = (data.min(0) + rng.random((24, 3)) * (data.max(0) - data.min(0))).astype(np.float32)
    g = driver.FastGoICP(w["model"], data, 0.04, 1e-4, ctx_factory=OracleContext, trim_fraction=trim)
    R, t = g.run()
    g.close()
    return dict(R=R, t=t, sse=float(g.best_sse), evals=g.stats["bound_evals"], cubes=g.stats["rot_cubes"],
                icps=g.stats["icp_runs"], levels=[(l["cubes"], l["survivors"], l["best_sse"]) for l in g.stats["level_log"]])


def _worker(rank, world, port, out_path, trim=0.0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _run(True, trim)
    if rank == 0:
        torch.save(res, out_path)
    # every rank must hold the same answer
    key = torch.tensor([np.float64(res["sse"]), float(np.sum(res["R"])), float(np.sum(res["t"]))], dtype=torch.float64)
    gathered = [torch.zeros_like(key) for _ in range(world)]
    dist.all_gather(gathered, key)
    assert all(torch.equal(gathered[0], k) for k in gathered)
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def single():
    return _run(False)


def test_single_process_recovers_pose(single):
    w = _problem()
    ang = np.degrees(np.arccos(np.clip((np.trace(single["R"] @ w["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 2.0 and np.linalg.norm(single["t"] - w["t_true"]) < 0.03
    assert single["cubes"] > 8 and single["icps"] >= 2


def test_comm_primitives_single_process():
    from fast_go_icp_b200.driver import _Comm
    c = _Comm()
    assert (c.rank, c.world) == (0, 1)
    assert c.min_key(123) == 123
    a = np.arange(12, dtype=np.float32).reshape(4, 3)
    assert np.array_equal(c.gather_rows(a, 4), a)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_is_bit_identical(tmp_path, single, world):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 400) + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    assert res["sse"] == single["sse"]
    assert np.array_equal(res["R"], single["R"]) and np.array_equal(res["t"], single["t"])
    assert res["levels"] == single["levels"]                 # same frontier, level by level
    assert res["cubes"] < single["cubes"]                    # rank 0 only searched its shard


def test_trimmed_sharded_run_is_bit_identical(tmp_path):
    """Trimmed registration through the sharded driver: same answer on 1 and 2 ranks."""
    one = _run(False, trim=0.2)
    out = str(tmp_path / "res_trim.pt")
    port = 29900 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out, 0.2), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert res["sse"] == one["sse"]
    assert np.array_equal(res["R"], one["R"]) and np.array_equal(res["t"], one["t"])
    assert res["levels"] == one["levels"]
    w = _problem()
    ang = np.degrees(np.arccos(np.clip((np.trace(one["R"] @ w["R_true"].T) - 1) / 2, -1, 1)))
    assert ang < 2.0 and np.linalg.norm(one["t"] - w["t_true"]) < 0.03
