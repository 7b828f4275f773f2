"""CPU tests of the oracle itself: independent cross-checks (numpy / scipy) of every stage."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from fast_go_icp_b200 import workloads
from oracle import oracle as O


def test_rotation_is_transposed_quaternion_matrix():
    rng = np.random.default_rng(0)
    for _ in range(50):
        v = rng.uniform(-0.6, 0.6, 3)
        if v @ v > 1:
            continue
        R, r = O.rotation(*v.astype(np.float32))
        w = np.sqrt(1 - v @ v)
        Rq = workloads.quat_to_matrix([w, *v])
        # column-major storage of the glm constructor => math matrix is the TRANSPOSE
        assert np.allclose(R.reshape(3, 3), Rq, atol=2e-6)         # storage [c][r] == Rq[c][r]
        assert np.allclose(R.reshape(3, 3).T, Rq.T, atol=2e-6)
        assert abs(r - np.linalg.norm(v)) < 1e-6
    R, r = O.rotation(0.9, 0.9, 0.9)                               # outside the ball: R = I, r = |q|^2 (Q3)
    assert np.array_equal(R, np.eye(3, dtype=np.float32).ravel()) and abs(r - 2.43) < 1e-5


def test_overlap_rule():
    assert O.overlaps_so3(0.5, 0.5, 0.5, 0.5)
    assert not O.overlaps_so3(0.9375, 0.9375, 0.9375, 0.0625)
    assert O.in_so3(0.5, 0.5, 0.5) and not O.in_so3(0.75, 0.75, 0.75)


def test_preprocess_matches_numpy(small_problem):
    pp = small_problem
    raw = pp["raw"]
    cs = raw["data"].astype(np.float64).mean(0)
    ct = raw["model"].astype(np.float64).mean(0)
    assert np.allclose(-pp["offset_pcs"], cs, atol=1e-5) and np.allclose(-pp["offset_pct"], ct, atol=1e-5)
    s = 1.0 / np.abs(raw["data"] - cs).max()
    assert abs(pp["scale"] - s) < 1e-4 * s
    assert abs(np.abs(pp["data"]).max() - 1.0) < 1e-6                  # source fills [-1, 1]
    assert np.allclose(pp["bbox_min"], pp["model"].min(0)) and np.allclose(pp["bbox_max"], pp["model"].max(0))


def test_lut_matches_kdtree(small_problem):
    pp = small_problem
    lut, dims = pp["lut"], pp["dims"]
    assert tuple(dims) == tuple(np.ceil((pp["bbox_max"] - pp["bbox_min"]) / pp["res"]).astype(int))
    P = pp["model"] + (-pp["bbox_min"])
    tree = cKDTree(P.astype(np.float64))
    zz, yy, xx = np.meshgrid(*[np.arange(d) for d in dims[::-1]], indexing="ij")
    nodes = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1) * np.float64(pp["res"])
    d, _ = tree.query(nodes)
    assert np.allclose(lut, d ** 2, rtol=2e-5, atol=1e-9)


def _numpy_trilinear(lut, dims, bbox_min, res, q):
    """float64 restatement of unnormalised linear texture filtering with clamp, 8-bit weights."""
    T = lut.reshape(dims[2], dims[1], dims[0]).astype(np.float64)
    u = (q.astype(np.float32) + (-bbox_min).astype(np.float32)) * np.float32(1.0 / np.float32(res))
    xf = np.floor(u.astype(np.float64) * 256.0 + 0.5).astype(np.int64) - 128     # round half up (B200 texture unit)
    i = xf >> 8
    a = (xf & 255) / 256.0
    out = np.zeros(len(q))
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ix = np.clip(i[:, 0] + dx, 0, dims[0] - 1)
                iy = np.clip(i[:, 1] + dy, 0, dims[1] - 1)
                iz = np.clip(i[:, 2] + dz, 0, dims[2] - 1)
                w = (a[:, 0] if dx else 1 - a[:, 0]) * (a[:, 1] if dy else 1 - a[:, 1]) * (a[:, 2] if dz else 1 - a[:, 2])
                out += w * T[iz, iy, ix]
    return out


def test_lut_sample_matches_numpy(small_problem):
    pp = small_problem
    rng = np.random.default_rng(1)
    q = rng.uniform(-1.3, 1.3, (5000, 3)).astype(np.float32)         # includes points outside the bbox (clamp)
    got = O.lut_sample(pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]), q)
    want = _numpy_trilinear(pp["lut"], pp["dims"], pp["bbox_min"], pp["res"], q)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7)
    # both interpolation formulas agree to fp32 rounding
    O.set_modes(0, 1)
    got2 = O.lut_sample(pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]), q)
    O.set_modes(0, 0)
    assert np.allclose(got, got2, rtol=1e-5, atol=1e-7)


def test_lut_sample_half_texel_shift(small_problem):
    """Q6: sampling at lattice node i*res + res/2 returns exactly T[i] (no +0.5 correction in the reference)."""
    pp = small_problem
    dims, res = pp["dims"], np.float64(pp["res"])
    i = np.array([[3, 4, 5], [10, 2, 7]])
    q = (i + 0.5) * res + pp["bbox_min"].astype(np.float64)
    got = O.lut_sample(pp["lut"], dims, pp["bbox_min"], float(pp["res"]), q.astype(np.float32))
    T = pp["lut"].reshape(dims[2], dims[1], dims[0])
    assert np.allclose(got, [T[5, 4, 3], T[7, 2, 10]], rtol=2e-3)


def test_bounds_invariants(small_problem):
    pp = small_problem
    rng = np.random.default_rng(2)
    R, _ = O.rotation(0.1, -0.2, 0.05)
    tc = workloads.translation_cube_list(24, level=3, seed=3)
    lb1, ub1 = O.bounds(pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]), pp["data"], R, 0.125, True, tc)
    lb0, ub0 = O.bounds(pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]), pp["data"], R, 0.125, False, tc)
    assert np.all(lb1 <= ub1) and np.all(lb0 <= ub0)
    assert np.all(ub0 <= ub1 + 1e-6) and np.all(lb0 <= lb1 + 1e-6)   # rotation slack only loosens
    # fixed-rotation "ub" is the LUT-interpolated SSE at (R, t): close to the exact SSE as long as the
    # transformed cloud stays inside the model's bounding box (outside, the texture clamps -- Q6)
    tin = np.array([[0.02, -0.01, 0.03, 0.0625], [-0.03, 0.02, 0.0, 0.0625]], np.float32)
    _, ubin = O.bounds(pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]), pp["data"], R, 0.125, True, tin)
    for c in range(2):
        exact = O.sse(pp["model"], pp["data"], R, tin[c, :3])
        assert abs(ubin[c] - exact) < 0.35 * exact + 1e-3
    # a zero-span translation cube has lb == ub
    tc0 = tc.copy(); tc0[:, 3] = 0
    l, u = O.bounds(pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]), pp["data"], R, 0.125, True, tc0)
    assert np.array_equal(l, u)
    del rng


def test_nn_matches_kdtree_and_tie_rules(small_problem):
    pp = small_problem
    R, _ = O.rotation(0.2, 0.1, -0.1)
    t = np.array([0.05, -0.02, 0.01], np.float32)
    idx, d2 = O.nn(pp["model"], pp["data"], R, t, rooted=False)
    q = pp["data"] @ R.reshape(3, 3) + t        # R stored column-major: (R_math @ p) = p @ R_storage
    d, j = cKDTree(pp["model"].astype(np.float64)).query(q.astype(np.float64))
    assert np.mean(idx == j) > 0.995
    assert np.allclose(d2, d ** 2, rtol=1e-4, atol=1e-9)
    # exact duplicates: the lowest index must win in both modes
    model = np.concatenate([pp["model"][:50], pp["model"][:50]])
    for rooted in (False, True):
        idx, _ = O.nn(model, pp["model"][:50], rooted=rooted)
        assert np.array_equal(idx, np.arange(50))


def test_closest_orthogonal_is_kabsch():
    rng = np.random.default_rng(3)
    for k in range(20):
        A = rng.normal(size=(40, 3))
        Rt = workloads.quat_to_matrix(rng.normal(size=4) / 1.0)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        Rt = workloads.quat_to_matrix(q)
        B = A @ Rt.T
        H = A.T @ B                                   # math: sum a b^T
        ABt = H.T.astype(np.float32).ravel()          # glm storage [c][r] = H[r][c]
        R = O.closest_orthogonal(ABt).reshape(3, 3).T # back to math layout
        assert np.allclose(R, Rt, atol=1e-5), k
        assert abs(np.linalg.det(R) - 1) < 1e-5
    # reflection case: determinant fix keeps a proper rotation
    H = np.diag([1.0, 1.0, -1.0])
    R = O.closest_orthogonal(H.T.astype(np.float32).ravel()).reshape(3, 3).T
    assert abs(np.linalg.det(R) - 1) < 1e-6


def test_icp_recovers_small_motion(small_problem):
    pp = small_problem
    # start from the true pose perturbed slightly: ICP must return to (near) the noise floor
    raw = pp["raw"]
    s = float(pp["scale"])
    R_true = raw["R_true"]
    # pose in the normalised frame: y_n = R x_n + t_n, t_n = s (R c_s + t - c_t)
    cs, ct = -pp["offset_pcs"].astype(np.float64), -pp["offset_pct"].astype(np.float64)
    t_n = s * (R_true @ cs + raw["t_true"] - ct)
    R0 = R_true.T.astype(np.float32).ravel()          # column-major storage of R_true
    e_true = O.sse(pp["model"], pp["data"], R0, t_n.astype(np.float32))
    e, R, t, it = O.icp(pp["model"], pp["data"], 100, 0.0005, R0, (t_n + 0.03).astype(np.float32))
    assert it >= 2 and e < 1.5 * e_true
    assert np.allclose(R.reshape(3, 3).T, R_true, atol=0.03)


def test_bnb_r3_finds_translation(small_problem):
    pp = small_problem
    raw = pp["raw"]
    s = float(pp["scale"])
    cs, ct = -pp["offset_pcs"].astype(np.float64), -pp["offset_pct"].astype(np.float64)
    t_n = s * (raw["R_true"] @ cs + raw["t_true"] - ct)
    # quaternion vector of R_true^T-storage convention: find (x, y, z) whose orc_rotation matches R_true
    Rm = raw["R_true"]
    w = np.sqrt(max(0.0, 1 + np.trace(Rm))) / 2
    # orc_rotation(x,y,z) storage == textbook matrix of (w, x, y, z) read row-major => math matrix is its transpose,
    # i.e. the rotation of quaternion (w, -x, -y, -z)
    v = -np.array([Rm[2, 1] - Rm[1, 2], Rm[0, 2] - Rm[2, 0], Rm[1, 0] - Rm[0, 1]]) / (4 * w)
    R, _ = O.rotation(*v.astype(np.float32))
    assert np.allclose(R.reshape(3, 3).T, Rm, atol=1e-5)
    thr = len(pp["data"]) * 1e-4
    ub, bt, evals, nb = O.bnb_r3(pp["model"], pp["data"], pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"]),
                                 np.array([*v, 0.0625], np.float32), True, 1e10, thr)
    assert np.all(np.abs(bt - t_n) <= 0.0625 + 1e-6)      # within the leaf cube
    assert evals == nb * len(pp["data"]) or evals <= nb * 32 * len(pp["data"])
    assert ub < 40 * thr


def test_kdtree_search_equals_the_brute_force_scans():
    """The k-d tree behind the oracle's NN (the CPU baseline's stand-in for a nanoflann ICP) returns the SAME index and
    the same distance bits as the literal restatement of the reference's scans (registration.cu:160-172,
    icp3d.cu:11-28), under both tie rules: random clouds, every point triplicated, queries on model points, clustered
    points, degenerate (flat / collinear / identical) clouds."""
    rng = np.random.default_rng(77)
    base = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    clouds = {
        "random": base,
        "triplicated": np.concatenate([base[:900]] * 3)[rng.permutation(2700)],
        "clustered": (base[:40, None, :] + rng.normal(scale=1e-4, size=(40, 60, 3))).reshape(-1, 3).astype(np.float32),
        "flat": np.concatenate([base[:1500, :2], np.zeros((1500, 1), np.float32)], axis=1),
        "collinear": np.outer(np.linspace(-1, 1, 500), [1, 2, -1]).astype(np.float32),
        "identical": np.tile(base[:1], (200, 1)),
        "quantised": np.round(base * 8) / 8,                       # many exactly equal distances
    }
    try:
        for name, model in clouds.items():
            model = np.ascontiguousarray(model, np.float32)
            q = np.concatenate([rng.uniform(-1.3, 1.3, (600, 3)), model[::7][:200], np.round(rng.uniform(-1, 1, (200, 3)) * 8) / 8]).astype(np.float32)
            R, _ = O.rotation(0.2, -0.1, 0.3)
            t = np.array([0.05, -0.02, 0.1], np.float32)
            for rooted in (False, True):
                for pose in ((None, None), (R, t)):
                    O.set_nn_mode(0)
                    i0, d0 = O.nn(model, q, pose[0], pose[1], rooted=rooted)
                    O.set_nn_mode(1)
                    i1, d1 = O.nn(model, q, pose[0], pose[1], rooted=rooted)
                    assert np.array_equal(i0, i1), (name, rooted)
                    assert np.array_equal(d0, d1), (name, rooted)
    finally:
        O.set_nn_mode(1)


def test_kdtree_grid_build_equals_the_brute_force_build():
    """The oracle's fast grid build (k-d tree over the shifted points) writes the same bits as the literal restatement of
    buildLUTKernel (registration.cu:258-278: every node scans every point), including off-centre boxes, coarse and fine
    resolutions, and degenerate clouds."""
    rng = np.random.default_rng(12)
    base = rng.normal(size=(2500, 3)).astype(np.float32) * np.array([0.4, 0.25, 0.1], np.float32) + np.float32(0.05)
    cases = [(base, 0.05), (base, 0.011),
             (np.concatenate([base[:600, :2], np.full((600, 1), 0.3, np.float32)], axis=1), 0.02),
             ((np.round(base[:800] * 16) / 16).astype(np.float32), 0.03)]
    try:
        for model, res in cases:
            model = np.ascontiguousarray(model, np.float32)
            mn, mx = model.min(0) - np.float32(0.013), model.max(0) + np.float32(0.021)
            O.set_lut_mode(0)
            a, da = O.lut_build(model, mn, mx, res)
            O.set_lut_mode(1)
            b, db = O.lut_build(model, mn, mx, res)
            assert np.array_equal(da, db) and np.array_equal(a, b)
    finally:
        O.set_lut_mode(1)
