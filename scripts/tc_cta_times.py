"""Debug (library built with -DSTG_TC_TIMING): start / end time of every CTA of the tcgen05 block kernels."""
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
for name, fn in (("fwd", lib.stg_debug_tc_cta_times_fwd), ("bwd", lib.stg_debug_tc_cta_times_bwd)):
    buf = (C.c_ulonglong * 3072)()
    assert fn(buf) == 0
    rows = [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(1024) if buf[3 * i]]
    t0 = min(r[0] for r in rows)
    starts = sorted(r[0] - t0 for r in rows)
    ends = sorted(r[1] - t0 for r in rows)
    dur = sorted(r[1] - r[0] for r in rows)
    n = len(rows)
    per_sm = {}
    for r in rows:
        per_sm.setdefault(r[2], []).append(r)
    print(f"{name}: {n} CTAs on {len(per_sm)} SMs; start ns p0/p50/p100 = {starts[0]}/{starts[n // 2]}/{starts[-1]}; "
          f"end ns p0/p50/p100 = {ends[0]}/{ends[n // 2]}/{ends[-1]}; duration ns p0/p50/p100 = {dur[0]}/{dur[n // 2]}/{dur[-1]}; "
          f"CTAs per SM min/max = {min(len(v) for v in per_sm.values())}/{max(len(v) for v in per_sm.values())}")
