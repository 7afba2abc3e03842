import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# parity bars from BASELINE.json north_star: rel = max|x - ref| / max|ref| per tensor
TOL_F64 = 1e-10
TOL_F32 = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def rel_err(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, "shape %s vs reference %s" % (x.shape, ref.shape)
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(x))) if x.size else 0.0
    return float(np.max(np.abs(x - ref))) / scale


def row_scaled_err(x, ref):
    """max over rows (last axis) of max|x - ref| / max|ref| of that row.  `rel_err` scales by the largest entry of the
    whole tensor; rows of Minv span 1e-3 .. 1e4 (light distal links) and small rows would never be checked by it.
    Rows whose reference is identically zero (structural zeros) must be exactly zero."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape
    scale = np.max(np.abs(ref), axis=-1)
    err = np.max(np.abs(x - ref), axis=-1)
    zero = scale == 0.0
    assert np.all(err[zero] == 0.0), "structural-zero row is not zero"
    return float(np.max(err[~zero] / scale[~zero])) if np.any(~zero) else 0.0


def make_robot(name):
    from rbdreference_b200 import robots
    if name == "tree9":
        return robots.random_tree(9, seed=1)
    if name == "tree13":
        return robots.random_tree(13, seed=2, branching=0.5, prismatic=0.3)
    return robots.by_name(name)


GOLDEN_CASES = ["iiwa14", "hyq", "atlas", "tree9", "tree13"]


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    name = request.param
    data = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return name, make_robot(name), data


FB_CASES = ["hyq", "atlas", "iiwa14", "tree9"]


def make_fb_robot(name):
    """Floating-base robots of tests/golden/fb_<name>.npz (oracle/make_golden.py --fb)."""
    from rbdreference_b200 import robots
    if name == "tree9":
        return robots.FloatingBaseRobot(robots.random_tree(9, seed=1), name="tree9_fb")
    return robots.by_name(name + "_fb")


def load_fb_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, "fb_" + name + ".npz")))


def load_fbpass_golden(name):
    """tests/golden/fbpass_<name>.npz (oracle/make_golden.py --fbpass): every array the unmodified reference's eight
    per-pass helpers take and return for floating-base robots."""
    return dict(np.load(os.path.join(GOLDEN_DIR, "fbpass_" + name + ".npz")))


def load_ee_golden(name):
    """tests/golden/ee_<name>.npz (oracle/make_golden.py --ee): the unmodified reference's
    end_effector_pose / end_effector_pose_gradient.  -> (q, [(names, offset, pose, grad), ...])"""
    import json
    d = np.load(os.path.join(GOLDEN_DIR, "ee_" + name + ".npz"))
    specs = json.loads(str(d["specs"]))
    cases = []
    for si, names in enumerate(specs):
        for oi, off in enumerate(d["offsets"]):
            cases.append((names, [list(off)], d["pose_s%d_o%d" % (si, oi)], d["grad_s%d_o%d" % (si, oi)]))
    return d["q"], cases


def random_states(n, B, seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-np.pi, np.pi, (B, n)), rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, n)))
