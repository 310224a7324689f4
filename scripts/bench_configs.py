"""Step time of the fused update on every FC_STGNN hyper-parameter set (not the headline bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS

dev = torch.device("cuda:0")
for name, B in (("FD001", 256), ("FD002", 256), ("FD003", 256), ("FD004", 256), ("NCMAPSS", 256), ("S2", 256),
                ("FD004", 1024), ("FD004", 4096)):
    cfg = CONFIGS[name]
    torch.manual_seed(0)
    alg = get_algorithm_class("FC_STGNN")(cfg, TRAIN_PARAMS, dev).to(dev)
    alg.train()
    X = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], device=dev)
    y = torch.rand(B, 1, device=dev)
    alg.enable_cuda_graph(B)
    for _ in range(5):
        alg.step(X, y)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n = 30
    for _ in range(n):
        loss = alg.step(X, y)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / n
    print(f"{name:8s} B={B:5d}  {ms*1e3:8.1f} us/step  {B/ms*1e3:10.0f} windows/s  loss {float(loss):.4f}", flush=True)

# ASTGCNN (BASELINE configs[2], N-CMAPSS shape, batch 512): native TCN / adjacency / Chebyshev aggregation
import warnings
from gnn_rul_benchmarking_b200.configs import ASTGCNN_CONFIGS
for name, B in (("CMAPSS", 100), ("NCMAPSS", 512)):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alg = get_algorithm_class("ASTGCNN")(ASTGCNN_CONFIGS[name], TRAIN_PARAMS, dev).to(dev)
    alg.train()
    N = ASTGCNN_CONFIGS[name]["num_nodes"]
    X, y = torch.rand(B, N, 50, device=dev), torch.rand(B, 1, device=dev)
    for _ in range(5):
        alg.update(X, y, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        alg.update(X, y, 1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"ASTGCNN {name:8s} B={B:5d}  {ms*1e3:8.1f} us/step  {B/ms*1e3:10.0f} windows/s", flush=True)
