"""Drop-in demonstration: writes a synthetic cloud pair in the reference's .txt format plus a TOML config
with the reference's keys, then runs build/fast-go-icp -- the reference's UNCHANGED src/main.cpp linked
against this repository's library."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import workloads  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "cli_demo")
os.makedirs(out, exist_ok=True)
w = workloads.synthetic_pair(nt=20000, ns=4000, seed=5)


def write_txt(path, pts):
    with open(path, "w") as f:
        f.write("%d\n" % len(pts))
        np.savetxt(f, pts, fmt="%.6f")


write_txt(os.path.join(out, "model.txt"), w["model"])
write_txt(os.path.join(out, "data.txt"), w["data"])
with open(os.path.join(out, "demo.toml"), "w") as f:
    f.write('[info]\nversion = "0.2"\n[io]\ntarget = "%s"\nsource = "%s"\n[params]\ntrim = false\n'
            'target_subsample = 1.0\nsource_subsample = 0.5\nlut_resolution = 0.01\nmse_threshold = 1e-4\n'
            % (os.path.join(out, "model.txt"), os.path.join(out, "data.txt")))
exe = os.path.join(ROOT, "build", "fast-go-icp")
r = subprocess.run([exe, "-c", os.path.join(out, "demo.toml"), "-v"], capture_output=True, text=True)
print(r.stdout[-3000:])
print(r.stderr[-2000:], file=sys.stderr)
print("expected R^T rows (ground truth):\n", w["R_true"], "\nexpected t:", w["t_true"])
sys.exit(r.returncode)
