"""Shared helpers for the test-suite: seeded cases (same ones the golden files were made from), oracle runs and
comparison metrics."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from eggfusion_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))


def _golden_mod():
    # imported lazily: make_golden imports torch
    mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(mod)
    return mod


def case_names():
    return ["c1_identity", "c1_posed_bg", "small_deg0_ragged", "small_deg1", "small_deg2"]


def case_inputs(name):
    return _golden_mod().case_inputs(name)


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


def oracle_run(name, backward=True):
    cam, sc, g, bg, mask, deg = case_inputs(name)
    M = sc["shs"].shape[1]
    oc = orc.cam_from_synthetic(cam, deg, M, bg=bg)
    f = orc.forward(oc, sc["xyz"], sc["scales"], sc["rotations"], sc["opacity"], sc["shs"], tile_mask=mask)
    b = None
    if backward:
        b = orc.backward(oc, f, sc["xyz"], sc["scales"], sc["rotations"], sc["shs"], g["color"], g["normal"],
                         g["depth"], g["opacity"])
    return cam, sc, g, bg, mask, deg, f, b


def rel_err(a, b):
    """max |a-b| / max |b|  -- the tolerance of BASELINE.json ("within 1e-4 rel fp32") applied tensor-wise."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    s = np.abs(b).max() if b.size else 0.0
    return d / (s + 1e-30)


def frac_close(a, b, rtol, atol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 1.0
    return float(np.mean(np.abs(a - b) <= atol + rtol * np.abs(b)))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def screen_block_from_oracle(b, P):
    """[P,16] screen-gradient block in the layout of include/eggsplat.h from the oracle's backward dict."""
    sg = np.zeros((P, 16), np.float32)
    sg[:, 0:2] = b["dL_dmean2D"][:, :2]
    sg[:, 2] = b["dL_dconic"][:, 0]
    sg[:, 3] = b["dL_dconic"][:, 1]
    sg[:, 4] = b["dL_dconic"][:, 3]
    sg[:, 5] = b["dL_dopacity"][:, 0]
    sg[:, 6:9] = b["dL_dcolors"]
    sg[:, 9:12] = b["dL_dnormal"]
    sg[:, 12] = b["dL_ddepth"][:, 0]
    return sg


def fusion_inputs(name="fusion_320x240"):
    return _golden_mod().fusion_inputs(name)


def tracking_inputs():
    return _golden_mod().tracking_inputs()
