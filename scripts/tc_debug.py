"""Debug aid: block backward through the tcgen05 path vs the SIMT path (STG_NO_TC=1) on one golden."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import load_golden  # noqa: E402


def run(g, dev):
    from gnn_rul_benchmarking_b200.fc_stgnn import GraphConvpoolMPNN_block_v6
    sd = g["sd0"]
    C = sd["graph_construction.mapping.weight"].shape[0]
    H = sd["MPNN.theta.0.weight"].shape[0]
    N = g["x"].shape[2]
    blk = GraphConvpoolMPNN_block_v6(C, H, N, 10, time_window_size=2, stride=int(g["stride"]), decay=0.7,
                                     pool_choice="mean")
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(dev).train()
    xg = g["x"].to(dev).clone().requires_grad_(True)
    out = blk(xg)
    (out * g["dout"].to(dev)).sum().backward()
    return out.detach().cpu(), xg.grad.cpu(), {k: p.grad.cpu() for k, p in blk.named_parameters()}


def main():
    name = sys.argv[1]
    g = load_golden(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name))
    dev = torch.device("cuda:0")
    os.environ.pop("STG_NO_TC", None)
    o1, dx1, g1 = run(g, dev)
    os.environ["STG_NO_TC"] = "1"
    o0, dx0, g0 = run(g, dev)
    ref = g["grad"]["x"]
    print("dx shape", tuple(ref.shape), "max|ref|", float(ref.abs().max()))
    e = (dx1 - ref).abs()
    print("tc   err by t:", [f"{float(v):.1e}" for v in e.amax(dim=(0, 2, 3))])
    print("simt err by t:", [f"{float(v):.1e}" for v in (dx0 - ref).abs().amax(dim=(0, 2, 3))])
    print("tc   err by n:", [f"{float(v):.1e}" for v in e.amax(dim=(0, 1, 3))])
    print("tc   err by c:", [f"{float(v):.1e}" for v in e.amax(dim=(0, 1, 2))])
    print("tc   err by b:", [f"{float(v):.1e}" for v in e.amax(dim=(1, 2, 3))])
    print("ref  max by t:", [f"{float(v):.1e}" for v in ref.abs().amax(dim=(0, 2, 3))])
    for k in g1:
        print(k, "tc-simt", float((g1[k] - g0[k]).abs().max()), "ref max", float(g["grad"][k].abs().max()))


if __name__ == "__main__":
    main()
