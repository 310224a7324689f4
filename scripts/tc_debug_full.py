"""Debug aid: full-size block backward, tcgen05 path vs SIMT path: where does dx differ?"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def main():
    from gnn_rul_benchmarking_b200.fc_stgnn import GraphConvpoolMPNN_block_v6
    N, C, H, stride, B = (int(v) for v in sys.argv[1:6])
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    blk = GraphConvpoolMPNN_block_v6(C, H, N, 10, time_window_size=2, stride=stride, decay=0.7, pool_choice="mean").to(dev)
    blk.train()
    x = torch.randn(B, 50, N, C, device=dev)
    L = (50 - 2) // stride + 1
    dout = torch.randn(B, L, N, H, device=dev)
    res = []
    for env in ({}, {"STG_NO_TC": "1"}):
        os.environ.pop("STG_NO_TC", None)
        os.environ.update(env)
        blk.zero_grad()
        xg = x.clone().requires_grad_(True)
        (blk(xg) * dout).sum().backward()
        res.append(xg.grad.clone())
    e = (res[0] - res[1]).abs()
    print("max ref", float(res[1].abs().max()), "max err", float(e.max()))
    print("err by t:", [f"{float(v):.0e}" for v in e.amax(dim=(0, 2, 3))])
    eb = e.amax(dim=(1, 2, 3))
    bad = (eb > 1e-2 * float(res[1].abs().max())).nonzero().flatten().tolist()
    print("bad samples:", bad[:40], "of", B)
    print("err by n:", [f"{float(v):.0e}" for v in e.amax(dim=(0, 1, 3))])
    print("err by c:", [f"{float(v):.0e}" for v in e.amax(dim=(0, 1, 2))])
    if bad:
        b = bad[0]
        print("sample", b, "err by t:", [f"{float(v):.0e}" for v in e[b].amax(dim=(1, 2))])
        t = int(e[b].amax(dim=(1, 2)).argmax())
        print(" t", t, "err by n:", [f"{float(v):.0e}" for v in e[b, t].amax(dim=1)])
        n = int(e[b, t].amax(dim=1).argmax())
        print(" tc  :", [f"{float(v):.3f}" for v in res[0][b, t, n]])
        print(" simt:", [f"{float(v):.3f}" for v in res[1][b, t, n]])


if __name__ == "__main__":
    main()
