"""Debug (library built with -DSTG_HEAD_TIMING): per-CTA timeline of the tensor-core head kernels."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from gnn_rul_benchmarking_b200 import _lib
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS

dev = torch.device("cuda:0")
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "FD004"]
alg = get_algorithm_class("FC_STGNN")(cfg, TRAIN_PARAMS, dev).to(dev)
alg.train()
X = torch.rand(256, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], device=dev)
y = torch.rand(256, 1, device=dev)
for _ in range(3):
    alg.step(X, y)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * (2 * 1024 * 8))()
assert lib.stg_debug_head_cta_times(buf) == 0
for kid, name in enumerate(("fc1", "bwd1")):
    rows = [[buf[(kid * 1024 + i) * 8 + q] for q in range(8)] for i in range(1024)]
    rows = [r for r in rows if r[0]]
    if not rows:
        continue
    n = len(rows)
    q = lambda k: sorted(r[k] for r in rows)
    print(f"{name}: {n} CTAs; cycles after entry p50/p100 per stamp: " +
          ", ".join(f"[{k}] {q(k)[n // 2]}/{q(k)[-1]}" for k in range(1, 6)))
