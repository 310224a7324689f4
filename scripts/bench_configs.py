"""Step time of the fused update on every FC_STGNN hyper-parameter set (not the headline bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS

dev = torch.device("cuda:0")
ONLY = os.environ.get("BENCH_ONLY", "")        # "siblings": skip the FC_STGNN / ASTGCNN tables
for name, B in () if ONLY == "siblings" else (("FD001", 256), ("FD002", 256), ("FD003", 256), ("FD004", 256), ("NCMAPSS", 256), ("S2", 256),
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
for name, B in () if ONLY == "siblings" else (("CMAPSS", 100), ("NCMAPSS", 512)):
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

# the other drop-in models at the shapes BASELINE.json / configs/hparams.py name (eager update(), wall clock)
SIBLINGS = [
    ("ASTGCNN", ASTGCNN_CONFIGS["NCMAPSS"], (512, 20, 50)),
    ("ST_GCN", dict(num_patch=40, patch_size=64, dropout=0.2), (128, 2560)),
    ("STGNN", dict(patch_size=5, num_patch=10, num_nodes=20, hidden_dim=64, K=3, top_k=10), (512, 20, 50)),
    ("STMSGCN", dict(num_patch=160, patch_size=16, interval=6, band_width=5, gcn_dims=[16, 64, 16, 1], gru_hidden_dim=8), (128, 2560)),
    ("GAT_LSTM", dict(num_patch=40, patch_size=64, hidden_dim=[300, 200, 100], lstm_hidden_dim=[30, 20], dropout=0.2), (128, 2560)),
    ("HAGCN", dict(patch_size=10, num_patch=5, encoder_hidden_dim=60, hidden_dim=64, output_dim=32), (256, 14, 50)),
    ("SAGCN", dict(num_patch=160, patch_size=16, gcn_hidden_dim=100, attention_hidden_dim=100), (128, 2560)),
]
for name, cfg, shape in SIBLINGS:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hp = dict(TRAIN_PARAMS, alpha=100)
        alg = get_algorithm_class(name)(cfg, hp, dev).to(dev)
    alg.train()
    X, y = torch.rand(*shape, device=dev), torch.rand(shape[0], 1, device=dev)
    for _ in range(5):
        alg.update(X, y, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        alg.update(X, y, 1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    try:
        alg.enable_cuda_graph(X, y)
    except Exception as e:          # report and keep going: the eager number stands
        print(f"{name:8s} X{list(shape)}  eager {ms*1e3:8.1f} us/step {shape[0]/ms*1e3:9.0f} windows/s | graph capture failed: {str(e)[:150]}", flush=True)
        torch.cuda.synchronize()
        continue
    for _ in range(3):
        alg.update(X, y, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        alg.update(X, y, 1)
    torch.cuda.synchronize()
    msg = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:8s} X{list(shape)}  eager {ms*1e3:8.1f} us/step {shape[0]/ms*1e3:9.0f} windows/s | "
          f"graph {msg*1e3:8.1f} us/step {shape[0]/msg*1e3:9.0f} windows/s", flush=True)
