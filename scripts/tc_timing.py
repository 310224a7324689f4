"""Debug: phase stamps of one tile of k_block_fwd_tc (library built with -DSTG_TC_TIMING into /tmp)."""
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
buf = (C.c_longlong * 64)()
half = (C.c_longlong * 32)()
for kern, fn in enumerate((lib.stg_debug_tc_stamps_fwd, lib.stg_debug_tc_stamps_bwd)):
    assert fn(half) == 0
    for i in range(32):
        buf[kern * 32 + i] = half[i]
names = {0: ["start", "x stored", "bar1", "issued FV", "FV done", "F/V stored", "bar2", "issued S", "S done", "softmax+agg", "-", "-",
             "-", "end"],
         1: ["start", "loads+stores", "bar1", "issued dA", "dA done", "softmax bwd", "bar2", "issued dF/dV", "dF/dV done",
             "dFV stored", "bar3", "issued dx/G", "dx/G done", "end"]}
for kern in (0, 1):
    print("forward" if kern == 0 else "backward", "tile, cycles (thread 0 | thread 64)")
    base = kern * 32
    prev = [buf[base], buf[base + 16]]
    for i in range(1, 14):
        cur = [buf[base + i], buf[base + 16 + i]]
        if cur[0] <= 0 or (kern == 0 and i in (10, 11)):
            continue
        print(f"  {names[kern][i]:14s} +{cur[0] - prev[0]:6d} | +{cur[1] - prev[1]:6d}")
        prev = cur
    print(f"  total {prev[0] - buf[base]}")
    if kern == 0:
        print(f"  CTA 0 forward: kernel start -> tile-loop end {buf[base + 10] - buf[base + 14]}, reductions + atomics "
              f"{buf[base + 11] - buf[base + 10]}, -> kernel end {buf[base + 15] - buf[base + 11]}")
    print(f"  CTA 0: kernel start -> this (2nd) tile start {buf[base] - buf[base + 14]}, tile end -> kernel end {buf[base + 15] - buf[base + 13]}, "
          f"whole CTA {buf[base + 15] - buf[base + 14]}")
