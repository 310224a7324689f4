"""Per-tensor gradient error of the whole model (tcgen05 block path) vs the reference model goldens."""
import glob
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import load_golden  # noqa: E402
from oracle import fc_stgnn_oracle as orc  # noqa: E402


class PinnedDropout(torch.nn.Module):
    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        return x * self.keep / (1.0 - self.p) if self.training else x


def main():
    from gnn_rul_benchmarking_b200.fc_stgnn import FC_STGNN_RUL
    dev = torch.device("cuda:0")
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_*.npz"))):
        g = load_golden(path)
        cfg = orc.CONFIGS[g["name"].split("_")[1]]
        model = FC_STGNN_RUL(**cfg)
        model.load_state_dict(g["sd0"], strict=False)
        model = model.to(dev)
        X, y = g["X"].to(dev), g["y"].to(dev)
        model.eval()
        with torch.no_grad():
            e_eval = float((model(X).cpu() - g["y_eval"]).abs().max())
        model.positional_encoding.dropout = PinnedDropout(g["keep"].float().to(dev), 0.1)
        model.train()
        pred = model(X)
        e_tr = float((pred.detach().cpu() - g["y_train"]).abs().max())
        torch.nn.functional.mse_loss(pred, y).backward()
        gmax = max(float(v.abs().max()) for v in g["grad"].values())
        print(f"== {g['name']} out err eval {e_eval:.1e} train {e_tr:.1e}  gmax {gmax:.2e}")
        for k, p in model.named_parameters():
            r = g["grad"][k]
            sc, e = float(r.abs().max()), float((p.grad.cpu() - r).abs().max())
            flag = " <<<" if e > 1e-3 * sc and sc >= 1e-4 * gmax else ""
            print(f"   {k:45s} scale {sc:.2e} ({sc / gmax:.1e} of gmax) rel {e / max(sc, 1e-30):.1e}{flag}")


if __name__ == "__main__":
    main()
