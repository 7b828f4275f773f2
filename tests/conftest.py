import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def install_reference_sin():
    """The oracle's sin(half-angle) constants (registration.cu:41-42) come from tests/golden/reference_sin.json: values
    produced by a kernel compiled inside the reference build (tests/golden/make_reference_sin.py), not by the library
    under test (the device under test used to supply them)."""
    import json
    from oracle import oracle as O
    with open(os.path.join(ROOT, "tests", "golden", "reference_sin.json")) as f:
        j = json.load(f)
    spans = np.array(j["spans"], np.float32)
    vals = np.array(j["reference_build_bits"], np.uint32).view(np.float32)
    O.set_sin_table(spans, vals)
    return spans, vals


def make_problem(nt=2000, ns=300, res=0.03, seed=5, sigma=0.01, max_angle=None):
    """Small seeded registration problem, preprocessed the way FastGoICP's constructor does it."""
    from fast_go_icp_b200 import workloads
    from oracle import oracle as O
    w = workloads.synthetic_pair(nt=nt, ns=ns, sigma=sigma, seed=seed, max_angle=max_angle)
    pp = O.preprocess(w["model"], w["data"])
    pp["res"] = np.float32(res)
    pp["raw"] = w
    return pp


@pytest.fixture(scope="session", autouse=True)
def _reference_sin_table():
    install_reference_sin()


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
    # the oracle uses the reference build's sin(half-angle) constants (golden file), never the device's
    install_reference_sin()
    yield ctx
    ctx.close()
