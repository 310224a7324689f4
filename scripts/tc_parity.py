"""Per-tensor error of the tcgen05 block path vs the reference goldens (and vs the SIMT path, STG_NO_TC=1).
Diagnostic: prints max|a-b| / max|ref| for every output / gradient of every block golden."""
import glob
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import load_golden  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max()) / max(1e-30, float(b.abs().max()))


def run(path, dev):
    from gnn_rul_benchmarking_b200.fc_stgnn import GraphConvpoolMPNN_block_v6
    g = load_golden(path)
    sd = g["sd0"]
    C = sd["graph_construction.mapping.weight"].shape[0]
    H = sd["MPNN.theta.0.weight"].shape[0]
    N = g["x"].shape[2]
    blk = GraphConvpoolMPNN_block_v6(C, H, N, 10, time_window_size=2, stride=int(g["stride"]), decay=0.7,
                                     pool_choice="mean")
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(dev)
    x = g["x"].to(dev)
    res = {}
    blk.eval()
    with torch.no_grad():
        res["out_eval"] = rel(blk(x).cpu(), g["out_eval"])
    blk.train()
    xg = x.clone().requires_grad_(True)
    out = blk(xg)
    res["out_train"] = rel(out.detach().cpu(), g["out_train"])
    (out * g["dout"].to(dev)).sum().backward()
    res["dx"] = rel(xg.grad.cpu(), g["grad"]["x"])
    for k, p in blk.named_parameters():
        res["d" + k] = rel(p.grad.cpu(), g["grad"][k])
    return res


def main():
    dev = torch.device("cuda:0")
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "block_*.npz")))
    if len(sys.argv) > 1:
        files = [f for f in files if any(a in os.path.basename(f) for a in sys.argv[1:])]
    elif os.environ.get("TC_PARITY_ISOLATE", "1") == "1":
        import subprocess
        for path in files:      # one process per golden: a faulting kernel poisons its CUDA context
            subprocess.run([sys.executable, __file__, os.path.basename(path)])
        return
    for path in files:
        name = os.path.basename(path)
        for mode in ("tc", "simt"):
            if mode == "simt":
                os.environ["STG_NO_TC"] = "1"
            else:
                os.environ.pop("STG_NO_TC", None)
            try:
                r = run(path, dev)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(name, mode, "ERROR", repr(e)[:300])
                continue
            print(name, mode, " ".join(f"{k.replace('graph_construction.', 'gc.')}={v:.1e}" for k, v in r.items()))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
