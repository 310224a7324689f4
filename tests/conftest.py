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
