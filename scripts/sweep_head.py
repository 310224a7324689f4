"""S1 step time under the head kernels' split knobs (STG_FC1_KSPLIT, STG_BWD1_SLICES are read at every launch plan)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS

dev = torch.device("cuda:0")
cfg, B = CONFIGS["FD004"], 256
X = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], device=dev)
y = torch.rand(B, 1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(env):
    for k in ("STG_FC1_KSPLIT", "STG_BWD1_SLICES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    torch.manual_seed(0)
    alg = get_algorithm_class("FC_STGNN")(cfg, TRAIN_PARAMS, dev).to(dev)
    alg.train()
    alg.enable_cuda_graph(B)
    for _ in range(10):
        alg.step(X, y)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
    for a, b in ev:
        flush.zero_()
        a.record(); alg.step(X, y); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(t[5:-5]) / len(t[5:-5]) * 1e3


base = run({})
print(f"default                      {base:7.1f} us")
for ks in (1, 2, 4, 6, 8):
    print(f"STG_FC1_KSPLIT={ks:<2d}            {run({'STG_FC1_KSPLIT': str(ks)}):7.1f} us", flush=True)
for sl in (2, 4, 6, 8, 12, 16):
    print(f"STG_BWD1_SLICES={sl:<2d}           {run({'STG_BWD1_SLICES': str(sl)}):7.1f} us", flush=True)
print(f"default again                {run({}):7.1f} us")
