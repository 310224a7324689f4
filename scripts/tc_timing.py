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
buf = (C.c_longlong * 32)()
assert lib.stg_debug_tc_stamps(buf) == 0
for t in range(2):
    st = [buf[t * 16 + i] for i in range(14)]
    print("thread", 0 if t == 0 else 64, [st[i] - st[0] for i in range(14)])
names = ["start", "x stored", "bar1", "issued FV", "FV done", "F/V stored", "bar2", "issued S", "S done", "softmax+st", "bar3",
         "issued Z", "Z done", "end"]
st = [buf[i] for i in range(14)]
for i in range(1, 14):
    print(f"  {names[i]:12s} +{st[i] - st[i - 1]:6d}  (thread 64: +{buf[16 + i] - buf[16 + i - 1]:6d})")
