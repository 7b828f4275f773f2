"""Cloud loaders, TOML config and result artefacts: the callers' side of the hot path.

Mirrors the reference's src/utilities.hpp (Config: lines 18-106, loaders: 113-260) with the two things it lacks:
  * subsampling is SEEDED.  The reference draws from an unseeded std::random_device (utilities.hpp:149-151,
    204-206), so no two runs see the same clouds; here the same acceptance rule runs on numpy's PCG64 with a
    seed taken from the config (`[params] seed`, default 0).  The rule itself is kept, quirks included:
    a point is accepted when u <= p and fewer than floor(n * p) points were accepted so far (Q19); the PLY loader
    stops reading at that count (head-biased), the TXT loader keeps reading -- same accepted set either way.
  * the `output` and `visualization` keys of the shipped TOML files (test/bunny.toml:10-11), which the reference
    parses into comments only (utilities.hpp:86-87), are honoured: `output` receives R, t, MSE as TOML and
    `visualization` the transformed source cloud as an ASCII PLY.
`trim` (bool) is parsed like the reference does; `trim_fraction` (new, default 0 = the reference's behaviour)
selects the trimmed registration.  `mode` is accepted and ignored, as in the reference (Q18).
"""
import os
import tomllib

import numpy as np


class Params:
    def __init__(self):
        self.trim = False
        self.trim_fraction = 0.0
        self.target_subsample = 1.0
        self.source_subsample = 1.0
        self.lut_resolution = 0.005
        self.mse_threshold = 1e-3
        self.seed = 0
        self.mode = None


class Config:
    """[io] target, source, output, visualization; [params] as in the reference, clamped the same way
    (utilities.hpp:99-104: subsample in [1e-5, 1], source_subsample additionally <= 0.5, mse_threshold >= 1e-12)."""

    def __init__(self, path):
        with open(path, "rb") as f:
            tbl = tomllib.load(f)
        self.path = path
        io = tbl.get("io", {})
        self.target = io.get("target", "")
        self.source = io.get("source", "")
        self.output = io.get("output", "")
        self.visualization = io.get("visualization", "")
        p, q = Params(), tbl.get("params", {})
        p.trim = bool(q.get("trim", False))
        p.trim_fraction = float(np.clip(q.get("trim_fraction", 0.0), 0.0, 0.9))
        p.target_subsample = float(np.clip(np.float32(q.get("target_subsample", 1.0)), np.float32(1e-5), np.float32(1.0)))
        p.source_subsample = float(np.clip(np.float32(q.get("source_subsample", 1.0)), np.float32(1e-5), np.float32(0.5)))
        p.lut_resolution = float(np.float32(q.get("lut_resolution", 0.005)))
        p.mse_threshold = float(max(np.float32(q.get("mse_threshold", 1e-3)), np.float32(1e-12)))
        p.seed = int(q.get("seed", 0))
        p.mode = q.get("mode")
        self.params = p

    def resolve(self, rel):
        """Paths in the shipped configs are relative to the directory the CLI runs in (../data/...); fall back to
        the config file's own directory."""
        if os.path.isabs(rel) or os.path.exists(rel):
            return rel
        return os.path.join(os.path.dirname(os.path.abspath(self.path)), rel)


def read_txt(path):
    """Go-ICP demo format: first token = point count, then x y z triples (utilities.hpp:181-235)."""
    with open(path) as f:
        tok = f.read().split()
    n = int(tok[0])
    if n <= 0:
        raise ValueError("Invalid number of points in the TXT file: %s" % path)
    pts = np.array(tok[1:1 + 3 * n], dtype=np.float32)
    if pts.size != 3 * n:
        raise ValueError("Error reading point data from TXT file: %s" % path)
    return pts.reshape(n, 3)


def read_ply(path):
    """Vertex x, y, z of an ASCII or binary little-endian PLY (what tinyply hands utilities.hpp:113-179)."""
    types = {"char": "i1", "uchar": "u1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4",
             "float": "<f4", "double": "<f8", "float32": "<f4", "float64": "<f8", "uint8": "u1", "int32": "<i4"}
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("not a PLY file: %s" % path)
        fmt, nv, props, in_vertex, seen_vertex = None, 0, [], False, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY header without end_header: %s" % path)
            w = line.decode("ascii", "replace").split()
            if not w:
                continue
            if w[0] == "format":
                fmt = w[1]
            elif w[0] == "element":
                if seen_vertex and not in_vertex:
                    pass
                in_vertex = w[1] == "vertex"
                if in_vertex:
                    nv, seen_vertex = int(w[2]), True
            elif w[0] == "property" and in_vertex:
                if w[1] == "list":
                    raise ValueError("list property in the vertex element: %s" % path)
                props.append((w[2], w[1]))
            elif w[0] == "end_header":
                break
        names = [p[0] for p in props]
        if not all(k in names for k in ("x", "y", "z")):
            raise ValueError("PLY file missing 'x', 'y', or 'z' vertex properties.")
        if nv <= 0:
            raise ValueError("No vertices found in the PLY file.")
        if fmt == "ascii":
            ix = [names.index(k) for k in ("x", "y", "z")]
            rows = np.array([f.readline().split() for _ in range(nv)])
            return rows[:, ix].astype(np.float32)
        if fmt != "binary_little_endian":
            raise ValueError("unsupported PLY format %s" % fmt)
        dt = np.dtype([(n, types[t]) for n, t in props])
        a = np.frombuffer(f.read(nv * dt.itemsize), dtype=dt, count=nv)
        return np.stack([a["x"], a["y"], a["z"]], axis=1).astype(np.float32)


def subsample(points, p, seed):
    """The reference's acceptance rule (utilities.hpp:149-163, 204-222) with a seeded generator: walk the points in
    file order, accept when u <= p, never more than floor(n * p) points."""
    n = len(points)
    cap = int(np.float32(n) * np.float32(p))                 # static_cast<size_t>(total_points * subsample), fp32
    u = np.random.default_rng(seed).random(n, dtype=np.float32)
    take = np.nonzero(u <= np.float32(p))[0][:cap]
    return points[take]


def load_cloud(path, p=1.0, seed=0):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".ply":
        pts = read_ply(path)
    elif ext == ".txt":
        pts = read_txt(path)
    else:
        raise ValueError("Unsupported file extension: %s" % ext.lstrip("."))
    return subsample(pts, p, seed)


def write_result_toml(path, R, t, mse, sse, extra=None):
    """`output` artefact: the registration y = R x + t (original coordinates), MSE and SSE (normalised frame)."""
    R = np.asarray(R, np.float64).reshape(3, 3)
    t = np.asarray(t, np.float64).reshape(3)
    lines = ["# fast-go-icp result: target ~= R * source + t", "[result]",
             "R = [%s]" % ", ".join("[%s]" % ", ".join("%.9g" % v for v in row) for row in R),
             "t = [%s]" % ", ".join("%.9g" % v for v in t),
             "mse = %.9g" % mse, "sse = %.9g" % sse]
    for k, v in (extra or {}).items():
        lines.append("%s = %s" % (k, ("%.9g" % v) if isinstance(v, float) else ('"%s"' % v if isinstance(v, str) else v)))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def write_ply(path, points):
    """`visualization` artefact: ASCII PLY of the transformed source cloud."""
    pts = np.asarray(points, np.float32).reshape(-1, 3)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n" % len(pts))
        np.savetxt(f, pts, fmt="%.7g")
