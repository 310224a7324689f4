"""Debug aid: tensor-core backward vs SIMT backward of the graph-conv block, per gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rul_benchmarking_b200.fc_stgnn import GraphConvpoolMPNN_block_v6

dev = torch.device("cuda:0")
shapes = [(7, 25, 14, 16, 8, 1), (7, 25, 14, 16, 8, 2), (2, 12, 14, 48, 24, 1), (4, 50, 21, 14, 7, 1), (3, 9, 5, 6, 3, 1),
          (2, 2, 2, 4, 2, 1), (2, 13, 20, 32, 16, 1), (2, 3, 26, 16, 8, 2)]
for (B, T, N, C, H, s) in shapes:
    torch.manual_seed(1)
    blk = GraphConvpoolMPNN_block_v6(C, H, N, 10, time_window_size=2, stride=s, decay=0.7, pool_choice="mean").to(dev)
    blk.train()
    x = torch.randn(B, T, N, C, device=dev)
    L = (T - 2) // s + 1
    dout = torch.randn(B, L, N, H, device=dev)
    res = {}
    for mode in ("simt", "mma"):
        if mode == "simt":
            os.environ.pop("STG_MMA_BWD", None)
        else:
            os.environ["STG_MMA_BWD"] = "1"
        blk.zero_grad()
        xg = x.clone().requires_grad_(True)
        out = blk(xg)
        (out * dout).sum().backward()
        res[mode] = {"x": xg.grad.clone(), **{k: p.grad.clone() for k, p in blk.named_parameters()}}
    print(f"shape B{B} T{T} N{N} C{C} H{H} s{s}")
    for k in res["simt"]:
        a, b = res["simt"][k], res["mma"][k]
        d = (a - b).abs()
        print(f"   {k:40s} max|simt| {float(a.abs().max()):.3e}  maxdiff {float(d.max()):.3e}")
    if "x" in res["simt"]:
        d = (res["simt"]["x"] - res["mma"]["x"]).abs().amax(dim=(0, 2, 3))
        print("   dx diff per t:", [f"{float(v):.1e}" for v in d])
