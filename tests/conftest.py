import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def make_problem(nt=2000, ns=300, res=0.03, seed=5, sigma=0.01, max_angle=None):
    """Small seeded registration problem, preprocessed the way FastGoICP's constructor does it."""
    from fast_go_icp_b200 import workloads
    from oracle import oracle as O
    w = workloads.synthetic_pair(nt=nt, ns=ns, sigma=sigma, seed=seed, max_angle=max_angle)
    pp = O.preprocess(w["model"], w["data"])
    pp["res"] = np.float32(res)
    pp["raw"] = w
    return pp


@pytest.fixture(scope="session")
def small_problem():
    from oracle import oracle as O
    pp = make_problem()
    lut, dims = O.lut_build(pp["model"], pp["bbox_min"], pp["bbox_max"], float(pp["res"]))
    pp["lut"], pp["dims"] = lut, dims
    return pp


@pytest.fixture(scope="session")
def gpu_ctx(small_problem):
    """CUDA context over the same small problem (all three samplers built)."""
    from fast_go_icp_b200 import capi
    pp = small_problem
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], float(pp["res"]),
                       flags=capi.BUILD_PACKED | capi.BUILD_TEX)
    # both sides must use the same sin(half-angle) constants: take the device's
    from oracle import oracle as O
    spans = np.array([1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125], np.float32)
    O.set_sin_table(spans, ctx.rot_sin(spans))
    yield ctx
    ctx.close()
