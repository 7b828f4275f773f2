"""Full-size pairs from the reference repository's own clouds, per SURVEY.md section 8d (W1-W4), for the GPU
benchmark scripts/bench_repo_clouds.py.  Run HERE (the container mounting /root/reference):

    python scripts/make_full_clouds.py        ->  build/workloads/repo_clouds_full.npz   (git-ignored, travels to the GPU box)

Only point coordinates are read (data files).  Subsampling uses this repository's seeded loader, which applies the
reference's acceptance rule (src/utilities.hpp:149-163, 204-222) on a seeded generator.

  W1 bunny   : data/bunny/model_bunny.txt p=0.5 (seed 0) vs data/bunny/data_bunny.txt p=0.1 (seed 1)   (test/bunny.toml:15-19)
  W2 skull   : data/artec3d/data_skull.ply p=0.3 vs the recipe of scripts/transform_point_cloud.py:15-54, 79-84
               (Gaussian index window, 10 % of the points, Euler rotation U(0,2pi)^3, translation U(-5,5)^3), rng(1)
  W3 dragon  : all vertices of dragonClearSpace2_0.ply vs 10,000 points of dragonToes3_0.ply (rng(2), sorted indices)
  W4 overlap : data_skull.ply split by x-rank into two 70 % slabs (40 % common band); the upper one subsampled to
               10,000 points and moved by a seeded SE(3) (rng(3)); the lower one seeded-subsampled p=0.5
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from fast_go_icp_b200 import cloudio  # noqa: E402
from make_cloud_fixtures import euler_rotation  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(ROOT, "build", "workloads", "repo_clouds_full.npz")


def main():
    out = {}
    out["bunny_model"] = cloudio.load_cloud(os.path.join(REF, "bunny", "model_bunny.txt"), 0.5, 0)
    out["bunny_data"] = cloudio.load_cloud(os.path.join(REF, "bunny", "data_bunny.txt"), 0.1, 1)

    skull = cloudio.load_cloud(os.path.join(REF, "artec3d", "data_skull.ply"), 1.0, 0).astype(np.float64)
    n = len(skull)
    rng = np.random.default_rng(1)
    out["skull_model"] = cloudio.subsample(skull.astype(np.float32), 0.3, 1)
    idx = np.arange(n)
    prob = np.exp(-0.5 * ((idx - n // 2) / (n / 100.0)) ** 2)
    prob /= prob.sum()
    sel = rng.choice(idx, size=int(0.1 * n), replace=False, p=prob)
    R = euler_rotation(rng.uniform(0, 2 * np.pi, 3))
    t = rng.uniform(-5, 5, 3)
    out["skull_data"] = skull[sel] @ R.T + t
    out["skull_R_move"], out["skull_t_move"] = R, t

    rng = np.random.default_rng(2)
    out["dragon_model"] = cloudio.load_cloud(os.path.join(REF, "dragon", "dragonClearSpace2_0.ply"), 1.0, 0)
    dd = cloudio.load_cloud(os.path.join(REF, "dragon", "dragonToes3_0.ply"), 1.0, 0)
    out["dragon_data"] = dd[np.sort(rng.choice(len(dd), 10_000, replace=False))]

    rng = np.random.default_rng(3)
    order = np.argsort(skull[:, 0], kind="stable")
    lo, hi = skull[order[: int(0.7 * n)]], skull[order[int(0.3 * n):]]
    out["overlap_model"] = cloudio.subsample(lo.astype(np.float32), 0.5, 3)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    t = rng.uniform(-20, 20, 3)
    hi_s = hi[np.sort(rng.choice(len(hi), 10_000, replace=False))]
    out["overlap_data"] = hi_s @ R.T + t
    out["overlap_R_move"], out["overlap_t_move"] = R, t

    out = {k: np.ascontiguousarray(v, dtype=np.float32 if (v.ndim == 2 and v.shape[0] > 3) else np.float64) for k, v in out.items()}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print("%-16s %s %s" % (k, v.shape, v.dtype))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
