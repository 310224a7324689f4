"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference tree.

The reference (Frank-Wang-oss/GNN_RUL_Benchmarking @ 9325667, mounted read-only at
/root/reference) does not import as shipped in this image:
  * models/FC_STGNN/Model_Base.py:5 imports matplotlib (absent, unused on the path)
  * utils.py:17 imports thop (absent, unused on the path)
  * Model_Base.py:58,119,151 call .cuda() unconditionally
  * trainer.py:92,94 use np.Inf (removed in NumPy 2)
  * dataloader/dataloader.py:62-63 torch.load without weights_only=False
`install()` applies the five shims *in this process only* and puts the reference on
sys.path.  Nothing is copied out of the reference and nothing is written into it.

This file exists only so that tests/golden/make_golden.py can run the real reference
in the build container.  /root/reference does not exist on the GPU box, so nothing in
`-m gpu` tests, smoke() or bench.py imports this module.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GNN_RUL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "FC_STGNN"))


def install() -> None:
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True            # never write __pycache__ into the read-only tree
    import numpy as np
    import torch

    for name in ("matplotlib", "matplotlib.pyplot", "thop"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["thop"].profile = lambda *a, **k: (0, 0)

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self          # CPU oracle runs
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if not getattr(torch.load, "_stg_shim", False):
        _orig_load = torch.load

        def _load(*a, **k):
            k.setdefault("weights_only", False)
            return _orig_load(*a, **k)

        _load._stg_shim = True
        torch.load = _load
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def fc_stgnn_classes():
    """Returns (FC_STGNN_RUL, GraphConvpoolMPNN_block_v6) from the real reference."""
    install()
    from models.FC_STGNN.Model import FC_STGNN_RUL            # noqa: E402
    from models.FC_STGNN.Model_Base import GraphConvpoolMPNN_block_v6
    return FC_STGNN_RUL, GraphConvpoolMPNN_block_v6
