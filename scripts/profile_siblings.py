"""One eager update of every sibling drop-in model at its BASELINE.json shape (for `ncu -k regex:"k_(adj|agg|gat|tcn|patch|rnn)"`:
the native kernels behind SURVEY.md 8 row a12)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import ASTGCNN_CONFIGS, TRAIN_PARAMS

dev = torch.device("cuda:0")
CASES = [
    ("ASTGCNN", ASTGCNN_CONFIGS["NCMAPSS"], (512, 20, 50)),
    ("ST_GCN", dict(num_patch=40, patch_size=64, dropout=0.2), (128, 2560)),
    ("STGNN", dict(patch_size=5, num_patch=10, num_nodes=20, hidden_dim=64, K=3, top_k=10), (512, 20, 50)),
    ("STMSGCN", dict(num_patch=160, patch_size=16, interval=6, band_width=5, gcn_dims=[16, 64, 16, 1], gru_hidden_dim=8), (128, 2560)),
    ("GAT_LSTM", dict(num_patch=40, patch_size=64, hidden_dim=[300, 200, 100], lstm_hidden_dim=[30, 20], dropout=0.2), (128, 2560)),
    ("HAGCN", dict(patch_size=10, num_patch=5, encoder_hidden_dim=60, hidden_dim=64, output_dim=32), (256, 14, 50)),
    ("SAGCN", dict(num_patch=160, patch_size=16, gcn_hidden_dim=100, attention_hidden_dim=100), (128, 2560)),
]
only = sys.argv[1:]
for name, cfg, shape in CASES:
    if only and name not in only:
        continue
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alg = get_algorithm_class(name)(cfg, dict(TRAIN_PARAMS, alpha=100), dev).to(dev)
    alg.train()
    X, y = torch.rand(*shape, device=dev), torch.rand(shape[0], 1, device=dev)
    torch.cuda.nvtx.range_push(name)
    alg.update(X, y, 1)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(name, "done", flush=True)
