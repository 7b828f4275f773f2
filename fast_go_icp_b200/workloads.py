"""Deterministic synthetic workloads (BASELINE.json configs / SURVEY.md section 8d).

W5 "synthetic": 100,000-point model on a bumpy closed surface, 10,000 data points = model points +
N(0, 0.01^2) noise under a random SE(3) pose.  Everything is seeded; nothing is read from disk, so
the same clouds exist in this container and on the GPU box.
"""
import numpy as np


def bumpy_surface(n, rng):
    """n points on r(theta, phi) = 0.6 + 0.15 sin(3 theta) cos(2 phi) + 0.1 cos(5 phi) + 0.08 sin(theta) sin(phi),
    directions uniform on the sphere, scaled so that max |coordinate| = 0.9.

    The first three terms are SURVEY.md section 8d's formula; on its own it is invariant under a half-turn
    about the x axis (theta -> pi - theta, phi -> -phi), i.e. it has TWO globally optimal registrations 180
    degrees apart.  The last term breaks that symmetry so that the optimum is unique and pose recovery
    can be asserted."""
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    theta = np.arccos(np.clip(v[:, 2], -1.0, 1.0))
    phi = np.arctan2(v[:, 1], v[:, 0])
    r = 0.6 + 0.15 * np.sin(3 * theta) * np.cos(2 * phi) + 0.1 * np.cos(5 * phi) + 0.08 * np.sin(theta) * np.sin(phi)
    p = v * r[:, None]
    return (p * (0.9 / np.abs(p).max())).astype(np.float32)


def quat_to_matrix(q):
    """Textbook rotation matrix of unit quaternion (w, x, y, z): y = R @ x."""
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def synthetic_pair(nt=100_000, ns=10_000, sigma=0.01, seed=1234, max_angle=None):
    """Returns dict(model, data, R_true, t_true) with model ~= R_true @ data + t_true (up to noise).

    data_i = R_move @ (model_sel_i + noise) + t_move, so the registration that maps data onto the
    model is R_true = R_move^T, t_true = -R_move^T t_move."""
    rng = np.random.default_rng(seed)
    model = bumpy_surface(nt, rng)
    sel = rng.choice(nt, ns, replace=False)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    if max_angle is not None:
        # limit the rotation angle (keeps small CPU tests fast)
        ang = 2 * np.arccos(abs(q[0]))
        if ang > max_angle:
            axis = q[1:] / np.linalg.norm(q[1:])
            q = np.concatenate([[np.cos(max_angle / 2)], np.sin(max_angle / 2) * axis])
    R_move = quat_to_matrix(q)
    t_move = rng.uniform(-0.3, 0.3, 3)
    noisy = model[sel].astype(np.float64) + rng.normal(scale=sigma, size=(ns, 3))
    data = (noisy @ R_move.T + t_move).astype(np.float32)
    return dict(model=model, data=data, R_true=R_move.T.copy(), t_true=-R_move.T @ t_move)


def rotation_cube_list(n, seed=7):
    """n rotation cubes (x, y, z, half_span = 0.0625) with centres inside the unit ball: the leaf level
    of the reference's SO(3) search (spans 0.5 .. 0.0625, fgoicp/fgoicp.cpp:53)."""
    rng = np.random.default_rng(seed)
    span = 0.0625
    out = []
    while len(out) < n:
        # leaf centres are odd multiples of the half-span
        c = (2 * rng.integers(-8, 8, size=3) + 1) * span
        if float(c @ c) <= 1.0:
            out.append([c[0], c[1], c[2], span])
    return np.array(out, np.float32)


def translation_cube_list(T=32, level=4, seed=11):
    """T translation cubes of one depth of the inner search: centres -1 + (2i+1) 2^-level, half-span 2^-level."""
    rng = np.random.default_rng(seed)
    span = 2.0 ** -level
    idx = rng.integers(0, 2 ** level, size=(T, 3))
    c = -1.0 + (2 * idx + 1) * span
    return np.concatenate([c, np.full((T, 1), span)], axis=1).astype(np.float32)


def bound_microbench(n_rot=4096, T=32, seed=7):
    """The pure-throughput bound workload: n_rot leaf rotation cubes x T leaf translation cubes each, no
    pruning.  Translation cubes are the leaf cubes (half-span 0.0625) whose centres
    lie in [-0.3125, 0.3125]^3 -- where a real search spends its deep levels, and where nearly every
    transformed point falls INSIDE the model's bounding box, so every evaluation is a genuine scattered
    gather (nothing is served by the clamped border cells)."""
    rot = rotation_cube_list(n_rot, seed)
    rng = np.random.default_rng(seed + 1)
    span = 0.0625
    idx = rng.integers(5, 11, size=(n_rot, T, 3))                 # of 16 leaf cells per axis
    tc = np.empty((n_rot, T, 4), np.float32)
    tc[..., :3] = -1.0 + (2 * idx + 1) * span
    tc[..., 3] = span
    return rot, tc
