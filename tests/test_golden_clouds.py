"""The reference repository's own clouds (BASELINE.json configs 1-4 / SURVEY.md W1-W4) at test size.

tests/golden/clouds_small.npz   : seeded subsamples of data/bunny, data/artec3d and data/dragon
                                   (tests/golden/make_cloud_fixtures.py, run where /root/reference is mounted)
tests/golden/reference_clouds.npz: what the UNMODIFIED reference (oracle/_ref on a B200) computes on them
                                   (tests/golden/make_golden_clouds.py)

CPU part pins the oracle to the reference on these clouds; GPU part pins the CUDA path to the reference AND
to the oracle.  Tolerances as in tests/test_golden.py: bit-exact preprocessing and grid cells; per-cube bounds
within 1e-4 relative + 1e-6 x ns absolute of the reference (texture-unit blend + CUB summation order,
DESIGN.md 3.3/3.5) and within 1 ulp of the oracle; exact SSE 2e-6; ICP 2e-5 (5 % / 0.5 % stop rules);
run(): final SSE within 1e-3 relative (the final ICP stops at a 0.05 % improvement)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CLOUDS = np.load(os.path.join(HERE, "golden", "clouds_small.npz"))
_gpath = os.path.join(HERE, "golden", "reference_clouds.npz")
G = np.load(_gpath) if os.path.exists(_gpath) else None
PAIRS = ("bunny", "skull", "dragon", "overlap")
ULP = 2.4e-7

pytestmark = pytest.mark.skipif(G is None, reason="tests/golden/reference_clouds.npz not generated yet")


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


_cache = {}


def problem(name):
    """Preprocessed pair + oracle grid, cached per session."""
    if name not in _cache:
        pp = O.preprocess(CLOUDS[name + "_model"], CLOUDS[name + "_data"])
        pp["res"] = float(G[name + "_res"])
        pp["lut"], pp["dims"] = O.lut_build(pp["model"], pp["bbox_min"], pp["bbox_max"], pp["res"])
        _cache[name] = pp
    return _cache[name]


def _lut(pp):
    return pp["lut"], pp["dims"], pp["bbox_min"], pp["res"]


@pytest.mark.parametrize("name", PAIRS)
def test_oracle_preprocessing_and_grid_bit_exact(name):
    pp, P = problem(name), name + "_"
    for k in ("offset_pcs", "offset_pct", "bbox_min", "bbox_max"):
        assert np.array_equal(pp[k], G[P + "pre_" + k]), k
    assert np.float32(pp["scale"]) == G[P + "pre_scale"]
    assert np.array_equal(_sha(pp["model"]), G[P + "pre_model_sha"]) and np.array_equal(_sha(pp["data"]), G[P + "pre_data_sha"])
    assert np.array_equal(pp["dims"], G[P + "lut_dims"])
    assert np.array_equal(pp["lut"][::101], G[P + "lut_stride101"])
    assert np.array_equal(_sha(pp["lut"]), G[P + "lut_sha"])            # every cell


@pytest.mark.parametrize("name", PAIRS)
def test_oracle_sampling_bounds_sse_icp_vs_reference(name):
    pp, P = problem(name), name + "_"
    ns = len(pp["data"])
    got = O.lut_sample(*_lut(pp), G[P + "tex_q"])
    rel = np.abs(got - G[P + "tex_val"]) / np.maximum(np.abs(G[P + "tex_val"]), 1e-6)
    assert np.median(rel) < 1e-5 and np.quantile(rel, 0.99) < 5e-3 and rel.max() < 0.1
    for r in range(3):
        rot = G[P + "bounds_rot"][r]
        R, _ = O.rotation(*rot[:3])
        for f in (0, 1):
            lb, ub = O.bounds(*_lut(pp), pp["data"], R, float(rot[3]), bool(f), G[P + "bounds_tc"][r])
            assert np.allclose(ub, G[P + "bounds_ub"][r, f], rtol=1e-4, atol=1e-6 * ns)
            assert np.allclose(lb, G[P + "bounds_lb"][r, f], rtol=1e-4, atol=1e-6 * ns)
    for k in range(2):
        e = O.sse(pp["model"], pp["data"], G[P + "sse_R"][k], G[P + "sse_t"][k])
        assert abs(e - G[P + "sse_val"][k]) <= 2e-6 * G[P + "sse_val"][k]
    for k, thr in enumerate((0.05, 0.005)):
        e, R, t, _ = O.icp(pp["model"], pp["data"], 100, thr, G[P + "sse_R"][k], G[P + "sse_t"][k])
        want = G[P + "icp_out"][k]
        assert abs(e - want[0]) <= 2e-5 * want[0]
        assert np.allclose(R, want[1:10], atol=1e-3) and np.allclose(t, want[10:13], atol=1e-3)


# bunny and skull take seconds on the CPU; dragon (170 s) and overlap (70 s) run when FGOICP_SLOW_TESTS=1 -- verified
# once (all four agree with the reference to 1e-7 in the pose) -- and always through the CUDA path on the GPU below
_RUN_PAIRS = PAIRS if os.environ.get("FGOICP_SLOW_TESTS") == "1" else ("bunny", "skull")


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("name", _RUN_PAIRS)
def test_oracle_full_run_vs_reference(name):
    """End to end on the CPU: the oracle's best-first run() against the reference's own run()."""
    pp, P = problem(name), name + "_"
    e, R, t, _ = O.run(pp["model"], pp["data"], *_lut(pp), float(G[P + "mse"]))
    # BASELINE.json north_star: MSE within 1e-6 relative (achieved: bunny 2.3e-7, skull 9e-8; pose 3e-7 / 3e-8)
    assert abs(e - G[P + "run_sse"]) <= 1e-6 * G[P + "run_sse"]
    assert np.allclose(R, G[P + "run_Rn"], atol=2e-6, rtol=0) and np.allclose(t, G[P + "run_tn"], atol=2e-6, rtol=0)


# ---- GPU ------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def ctxs():
    from fast_go_icp_b200 import capi
    made = {}

    def get(name):
        if name not in made:
            pp = problem(name)
            c = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], pp["res"],
                             flags=capi.BUILD_PACKED | capi.BUILD_TEX)
            made[name] = c          # (the oracle's sin constants come from tests/golden/reference_sin.json: conftest)
        return made[name]
    yield get
    for c in made.values():
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", PAIRS)
def test_cuda_vs_reference_and_oracle(name, ctxs):
    from fast_go_icp_b200 import capi
    pp, P, ctx = problem(name), name + "_", ctxs(name)
    ns = len(pp["data"])
    lut, dims = ctx.lut_download()
    assert np.array_equal(dims, G[P + "lut_dims"]) and np.array_equal(_sha(lut), G[P + "lut_sha"])
    assert np.array_equal(lut, pp["lut"])
    assert np.array_equal(ctx.lut_sample(G[P + "tex_q"], capi.SAMPLER_TEX), G[P + "tex_val"])     # same hardware path
    assert np.array_equal(ctx.lut_sample(G[P + "tex_q"], capi.SAMPLER_PACKED), O.lut_sample(*_lut(pp), G[P + "tex_q"]))
    ctx.set_sampler(capi.SAMPLER_PACKED)
    for r in range(3):
        rot = G[P + "bounds_rot"][r]
        R, _ = O.rotation(*rot[:3])
        for f in (0, 1):
            lb, ub = ctx.bounds_batch(R, float(rot[3]), bool(f), G[P + "bounds_tc"][r])
            assert np.allclose(ub, G[P + "bounds_ub"][r, f], rtol=1e-4, atol=1e-6 * ns)
            assert np.allclose(lb, G[P + "bounds_lb"][r, f], rtol=1e-4, atol=1e-6 * ns)
            wl, wu = O.bounds(*_lut(pp), pp["data"], R, float(rot[3]), bool(f), G[P + "bounds_tc"][r])
            assert np.allclose(ub, wu, rtol=ULP, atol=0) and np.allclose(lb, wl, rtol=ULP, atol=0)
    for k in range(2):
        R0, t0 = G[P + "sse_R"][k], G[P + "sse_t"][k]
        e = ctx.sse(R0, t0)
        assert abs(e - G[P + "sse_val"][k]) <= 2e-6 * G[P + "sse_val"][k]
        assert e == O.sse(pp["model"], pp["data"], R0, t0)
        for rooted in (False, True):
            idx, d2 = ctx.nn(R0, t0, rooted)
            widx, wd2 = O.nn(pp["model"], pp["data"], R0, t0, rooted)
            assert np.array_equal(idx, widx) and np.array_equal(d2, wd2)
    for k, thr in enumerate((0.05, 0.005)):
        e, R, t, it = ctx.icp(G[P + "sse_R"][k], G[P + "sse_t"][k], 100, thr)
        want = G[P + "icp_out"][k]
        assert abs(e - want[0]) <= 2e-5 * want[0]
        assert np.allclose(R, want[1:10], atol=1e-3) and np.allclose(t, want[10:13], atol=1e-3)
        we, wR, wt, wit = O.icp(pp["model"], pp["data"], 100, thr, G[P + "sse_R"][k], G[P + "sse_t"][k])
        assert it == wit and abs(e - we) <= 1e-6 * we
    thr = float(G[P + "sse_threshold"])
    cubes = np.float32([[0.25, -0.25, 0.25, 0.25], [-0.0625, 0.1875, 0.0625, 0.0625]])
    for fix_rot in (True, False):
        ub, bt, ev = ctx.bnb_r3_batch(cubes, fix_rot, 1e10, thr)
        for i, c in enumerate(cubes):
            wub, wbt, wev, _ = O.bnb_r3(pp["model"], pp["data"], *_lut(pp), c, fix_rot, 1e10, thr)
            assert ev[i] == wev and np.isclose(ub[i], wub, rtol=ULP, atol=0) and np.array_equal(bt[i], wbt)


# (relative SSE, absolute pose) tolerance of run() against the reference's run() where the north_star figures
# (1e-6, far inside the BnB leaf) cannot be asserted; every other pair is held to them.  Filled from the achieved
# differences recorded on the B200 (profiles/run_parity_r02.md).
RUN_TOL = {}


@pytest.mark.gpu
@pytest.mark.parametrize("name", PAIRS)
def test_cuda_full_run_vs_reference(name):
    """run() through the Python driver against the reference's own run() on the same pair.
    * reference visiting order ("bestfirst"): final SSE within BASELINE.json's 1e-6 relative of the reference's, pose
      within 2e-6 in the normalised frame (RUN_TOL below lists the pairs where that cannot hold, and why);
    * level-synchronous schedule (default; whole levels in flight): the search stops as soon as
      best_sse - lb <= sse_threshold, so a different visiting order may stop at a different incumbent; both are
      optimal to within sse_threshold, which is what is asserted."""
    from fast_go_icp_b200 import driver
    P = name + "_"
    g = driver.FastGoICP(CLOUDS[name + "_model"], CLOUDS[name + "_data"], float(G[P + "res"]), float(G[P + "mse"]),
                         schedule="bestfirst")
    R, t = g.run()
    sse_tol, pose_tol = RUN_TOL.get(name, (1e-6, 2e-6))
    assert abs(g.best_sse - G[P + "run_sse"]) <= sse_tol * G[P + "run_sse"]
    assert np.allclose(g.best_R, G[P + "run_Rn"], atol=pose_tol, rtol=0) and np.allclose(g.best_t, G[P + "run_tn"], atol=pose_tol, rtol=0)
    assert np.allclose(t, G[P + "run_t"], atol=10 * pose_tol / float(g.pp["scale"]), rtol=0)
    g.close()
    g = driver.FastGoICP(CLOUDS[name + "_model"], CLOUDS[name + "_data"], float(G[P + "res"]), float(G[P + "mse"]))
    g.run()
    assert g.best_sse <= G[P + "run_sse"] + float(G[P + "sse_threshold"])
    g.close()
