"""Per-kernel CUDA-event times of one eager training step on a named FC_STGNN hyper-parameter set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rul_benchmarking_b200._lib import kernel_profile
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["FD001", "NCMAPSS"]:
    B = 256
    cfg = CONFIGS[name]
    alg = get_algorithm_class("FC_STGNN")(cfg, TRAIN_PARAMS, dev).to(dev)
    alg.train()
    X = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], device=dev)
    y = torch.rand(B, 1, device=dev)
    for _ in range(3):
        alg.step(X, y)
    torch.cuda.synchronize()
    with kernel_profile() as kp:
        for _ in range(5):
            alg.step(X, y)
        torch.cuda.synchronize()
    print(name)
    for k, (ms, n) in kp.result().items():
        print(f"   {k:24s} {ms / n * 1e3:9.1f} us x {n}")
