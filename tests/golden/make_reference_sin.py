"""Generates tests/golden/reference_sin.json on the GPU box: sin(span * sqrt3 * pi / 2) for the rotation half-spans the
search uses, evaluated by a kernel compiled INSIDE the reference build (oracle/ref_shim/ref_capi.cu: the expression of
registration.cu:41-42 with the reference's own M_SQRT3 / M_PI and the flags oracle/build_ref.py compiles the reference
with) -- i.e. the constants the reference's bound kernel uses, obtained without the library under test.  The file also
records what the library's device code and numpy's float32 sin return for the same arguments (all three agree bit for bit
on B200 / CUDA 12.9).  The oracle's tests install the reference-build values (tests/conftest.py); a GPU test checks the
library against them.

    python tests/golden/make_reference_sin.py gpurun_out/reference_sin.json     # then copy to tests/golden/
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import capi, driver, workloads  # noqa: E402
from oracle import ref as REF  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_sin.json")
spans = np.array([1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125], np.float32)
r = REF.rot_sin(spans)
w = workloads.synthetic_pair(nt=500, ns=50, seed=1)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.1, flags=0)
d = ctx.rot_sin(spans)
ctx.close()
h = np.array([np.sin(np.float32(np.float32(np.float32(s * np.float32(1.732050807568877)) * np.float32(3.141592653589793)) * np.float32(0.5)),
                     dtype=np.float32) for s in spans], np.float32)
json.dump(dict(spans=spans.tolist(), reference_build_bits=[int(x) for x in r.view(np.uint32)],
               library_bits=[int(x) for x in d.view(np.uint32)], numpy_sinf_bits=[int(x) for x in h.view(np.uint32)],
               reference_build=r.tolist()), open(out, "w"))
print(open(out).read())
