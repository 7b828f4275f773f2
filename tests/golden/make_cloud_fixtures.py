"""Builds tests/golden/clouds_small.npz: small, seeded subsamples of the reference repository's own clouds, one
pair per BASELINE.json config that names a repo cloud.  Run HERE (the container
that mounts /root/reference); the GPU box has no /root/reference, so tests only ever read the committed .npz.

    python tests/golden/make_cloud_fixtures.py

Only point coordinates are read (data files, not source code).  The reference's loaders subsample with an
unseeded std::random_device (src/utilities.hpp:149-163, 204-222); here the same Bernoulli rule is applied with
numpy's seeded generator so the fixtures are reproducible.

  bunny   : data/bunny/model_bunny.txt vs data/bunny/data_bunny.txt          (test/bunny.toml, W1)
  skull   : data/artec3d/data_skull.ply vs a windowed, moved 10 % sample of itself, following the recipe of
            scripts/transform_point_cloud.py:15-54, 79-84 (model_skull.ply is missing upstream; W2)
  dragon  : data/dragon/dragonClearSpace2_0.ply vs data/dragon/dragonToes3_0.ply (two range scans; W3)
  overlap : data_skull.ply split along x into two 70 % slabs sharing a 40 % band, the second one moved (W4)
"""
import os

import numpy as np

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clouds_small.npz")


def load_txt(path):
    """Go-ICP demo format (src/utilities.hpp:181-235): first line = count, then x y z per line."""
    with open(path) as f:
        n = int(f.readline())
        pts = np.loadtxt(f, dtype=np.float64).reshape(-1, 3)
    assert len(pts) == n
    return pts


def load_ply(path):
    """ASCII or binary little-endian PLY, vertex x y z first (what src/utilities.hpp:113-179 reads via tinyply)."""
    with open(path, "rb") as f:
        fmt, nv, props = None, 0, []
        in_vertex = False
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element"):
                in_vertex = line.split()[1] == "vertex"
                if in_vertex:
                    nv = int(line.split()[2])
            elif line.startswith("property") and in_vertex:
                props.append(line.split()[1:])
            elif line == "end_header":
                break
        if fmt == "ascii":
            rows = [f.readline().split()[:3] for _ in range(nv)]
            return np.array(rows, dtype=np.float64)
        assert fmt == "binary_little_endian"
        types = {"float": "<f4", "uchar": "u1", "double": "<f8", "int": "<i4"}
        dt = np.dtype([(p[1], types[p[0]]) for p in props])
        a = np.frombuffer(f.read(nv * dt.itemsize), dtype=dt, count=nv)
        return np.stack([a["x"], a["y"], a["z"]], axis=1).astype(np.float64)


def bernoulli(pts, p, rng):
    return pts[rng.random(len(pts)) < p]


def euler_rotation(angles):
    """scripts/transform_point_cloud.py:40-54: Rz @ Ry @ Rx."""
    a, b, c = angles
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return rz @ ry @ rx


def main():
    out = {}
    # W1 bunny ---------------------------------------------------------------------------------------
    rng = np.random.default_rng(0)
    model = load_txt(os.path.join(REF, "bunny", "model_bunny.txt"))
    data = load_txt(os.path.join(REF, "bunny", "data_bunny.txt"))
    out["bunny_model"] = bernoulli(model, 0.07, rng)
    out["bunny_data"] = bernoulli(data, 0.012, rng)
    # W2 skull ---------------------------------------------------------------------------------------
    rng = np.random.default_rng(1)
    skull = load_ply(os.path.join(REF, "artec3d", "data_skull.ply"))
    out["skull_model"] = bernoulli(skull, 0.03, rng)
    n = len(skull)
    idx = np.arange(n)
    prob = np.exp(-0.5 * ((idx - n // 2) / (n / 100.0)) ** 2)
    prob /= prob.sum()
    sel = rng.choice(idx, size=int(0.1 * n), replace=False, p=prob)
    R = euler_rotation(rng.uniform(0, 2 * np.pi, 3))
    t = rng.uniform(-5, 5, 3)
    moved = skull[sel] @ R.T + t
    out["skull_data"] = moved[rng.random(len(moved)) < 0.035]
    out["skull_R_move"], out["skull_t_move"] = R, t
    # W3 dragon --------------------------------------------------------------------------------------
    rng = np.random.default_rng(2)
    dm = load_ply(os.path.join(REF, "dragon", "dragonClearSpace2_0.ply"))
    dd = load_ply(os.path.join(REF, "dragon", "dragonToes3_0.ply"))
    out["dragon_model"] = bernoulli(dm, 0.04, rng)
    out["dragon_data"] = dd[np.sort(rng.choice(len(dd), 350, replace=False))]
    # W4 partial overlap -----------------------------------------------------------------------------
    rng = np.random.default_rng(3)
    order = np.argsort(skull[:, 0], kind="stable")
    lo, hi = skull[order[: int(0.7 * n)]], skull[order[int(0.3 * n):]]
    out["overlap_model"] = bernoulli(lo, 0.04, rng)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    t = rng.uniform(-20, 20, 3)
    hi_s = hi[np.sort(rng.choice(len(hi), 360, replace=False))]
    out["overlap_data"] = hi_s @ R.T + t
    out["overlap_R_move"], out["overlap_t_move"] = R, t
    out = {k: np.ascontiguousarray(v, dtype=np.float32 if v.ndim == 2 and v.shape[1] == 3 and v.shape[0] > 3 else np.float64)
           for k, v in out.items()}
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print("%-16s %s %s" % (k, v.shape, v.dtype))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
