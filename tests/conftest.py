import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_files(kind):
    return sorted(glob.glob(os.path.join(GOLDEN, f"{kind}_*.npz")))


def load_golden(path, dtype=torch.float32):
    """-> dict with 'sd0' (state before), 'sd1' (running stats after one train forward),
    'grad' (reference grads) and the remaining top-level arrays, all as torch tensors."""
    z = np.load(path)
    g = {"sd0": {}, "sd1": {}, "grad": {}, "name": os.path.basename(path)}
    for k in z.files:
        a = z[k]
        if "/" in k:
            grp, key = k.split("/", 1)
            t = torch.from_numpy(a)
            if t.is_floating_point():
                t = t.to(dtype)
            g[grp][key] = t
        else:
            t = torch.from_numpy(np.asarray(a))
            g[k] = t.to(dtype) if t.is_floating_point() else t
    return g


# ---- parity tolerances (written here once; BASELINE.json north_star: 1e-4 on outputs) ---------------------
OUT_TOL = 2e-5            # outputs: every forward product is fp32-accurate (3-term TF32 split on the tcgen05 path)
GRAD_TOL_FP32 = 3e-4      # gradients, fp32 SIMT paths: relative to the tensor's own largest entry (summation-order noise
                          # of float atomics over ~1e5-row reductions: measured up to 1.7e-4 on FD003, B=70)
GRAD_TOL_TC = 5e-3        # gradients through the single-pass TF32 backward products of the tcgen05 path: operand
                          # rounding 2^-12 per factor, three products deep, and reductions over ~1e5 rows that cancel
                          # (BatchNorm shifts, biases).  Measured per tensor: 1e-4 .. 8e-4 typical, 2.7e-3 worst (a bias gradient that
                          # is the sum of 9e4 rows cancelling to ~1 % of their size)
                          # (scripts/tc_parity.py, scripts/tc_parity_model.py; profiles/r02_tc_parity.md)


def tc_shape(C, H, N, w=2):
    """True when plan_blocks_tc (csrc/stg_block_tc.cu) takes the shape: tcgen05 path, TF32 backward products."""
    return (not os.environ.get("STG_NO_TC")) and C <= 16 and H <= 8 and w == 2 and 2 * N <= 64


def rel_err(a, b):
    """max|a-b| / max|b| -- a TRUE relative error (no floor at 1)."""
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def assert_close_rel(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e < tol, f"{what}: relative error {e:.3e} >= {tol:.1e} (max|ref| {float(b.abs().max()):.3e})"


def assert_grads_close(got, ref, tol, what=""):
    """Every gradient tensor within `tol` of the reference RELATIVE TO ITS OWN largest entry.  Tensors whose
    reference is more than four orders of magnitude below the largest gradient of the case are numerically zero
    (e.g. the bias of a Linear followed by a training-mode BatchNorm: its gradient is a sum that cancels exactly,
    SURVEY 9.3) and are held to an absolute 1e-6 of that largest gradient instead."""
    gmax = max(float(v.abs().max()) for v in ref.values())
    bad = []
    for k, r in ref.items():
        g = got[k]
        scale, err = float(r.abs().max()), float((g - r).abs().max())
        if scale < 1e-4 * gmax:
            ok, shown = err <= 1e-6 * gmax, f"abs {err:.2e} (zero-level tensor, limit {1e-6 * gmax:.2e})"
        else:
            ok, shown = err <= tol * scale, f"rel {err / scale:.2e} (limit {tol:.1e})"
        if not ok:
            bad.append(f"{k}: {shown}")
    assert not bad, f"{what} gradient mismatch: " + "; ".join(bad)
