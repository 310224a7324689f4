"""Debug (library built with -DSTG_ENC_TIMING): per-CTA timeline of the fast patch-encoder phases."""
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
buf = (C.c_ulonglong * (9 * 1024 * 8))()
assert lib.stg_debug_enc_cta_times(buf) == 0
for ph in range(9):
    rows = [[buf[(ph * 1024 + i) * 8 + q] for q in range(8)] for i in range(1024)]
    rows = [r for r in rows if r[0]]
    if not rows:
        continue
    n = len(rows)
    t0 = min(r[0] for r in rows)
    st = sorted(r[0] - t0 for r in rows)
    q = lambda k: sorted(r[k] for r in rows)
    pro, loop, end = q(1), q(2), q(3)
    print(f"PH{ph}: {n} CTAs; start ns p50/p100 = {st[n // 2]}/{st[-1]}; cycles after entry (p50/p100): "
          f"prologue {pro[n // 2]}/{pro[-1]}, tiles done {loop[n // 2]}/{loop[-1]}, end {end[n // 2]}/{end[-1]}; "
          f"inner stamps 7/4/5/6 p50 = {q(7)[n // 2]}/{q(4)[n // 2]}/{q(5)[n // 2]}/{q(6)[n // 2]}")
