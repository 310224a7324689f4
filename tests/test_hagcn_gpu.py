"""HAGCN (BASELINE.json configs[4]) drop-in: native cosine adjacency + dense aggregations (GIN, SAGPool) around the
native Bi-LSTM recurrence (stg_rnn_*) vs the UNMODIFIED reference model (tests/golden/aux_metrics_data.npz; encoder dropouts pinned)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
CFG = dict(patch_size=10, num_patch=5, encoder_hidden_dim=60, hidden_dim=64, output_dim=32)
TAG = "hagcn_p10"


class PinnedDropout(torch.nn.Module):
    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        return x * self.keep / (1.0 - self.p) if self.training else x


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)      # relative to the tensor's own largest entry


def _sub(grp):
    pre = f"{TAG}/{grp}/"
    return {k[len(pre):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre)}


def test_state_dict_layout_matches_reference():
    from gnn_rul_benchmarking_b200.hagcn import HAGCN_model
    sd, ref = HAGCN_model(**CFG).state_dict(), _sub("sd0")
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


@pytest.mark.gpu
def test_model_matches_reference():
    from gnn_rul_benchmarking_b200.hagcn import HAGCN_model
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False       # the goldens are CPU fp32; cuDNN's LSTM defaults to TF32
    model = HAGCN_model(**CFG)
    model.load_state_dict(_sub("sd0"), strict=True)
    model = model.to(dev)
    X, y = torch.from_numpy(Z[f"{TAG}/X"]).to(dev), torch.from_numpy(Z[f"{TAG}/y"]).to(dev)
    model.eval()
    with torch.no_grad():
        assert _rel(model(X).cpu(), torch.from_numpy(Z[f"{TAG}/y_eval"])) < 5e-5
    model.train()
    model.TD.drop2 = PinnedDropout(torch.from_numpy(Z[f"{TAG}/keep0"]).float().to(dev), 0.2)
    model.TD.drop3 = PinnedDropout(torch.from_numpy(Z[f"{TAG}/keep1"]).float().to(dev), 0.2)
    pred, kl = model(X, train=True)
    assert _rel(pred.detach().cpu(), torch.from_numpy(Z[f"{TAG}/y_train"])) < 5e-5
    assert abs(float(kl.detach()) - float(Z[f"{TAG}/kl"])) < 1e-6 + 1e-4 * abs(float(Z[f"{TAG}/kl"]))
    (torch.nn.functional.mse_loss(pred, y) + 100.0 * kl).backward()
    named = dict(model.named_parameters())
    grads = _sub("grad")
    assert len(grads) > 30
    # per tensor, relative to its own largest entry -- floored at 1e-4 of the model's largest gradient entry: several
    # HAGCN gradients are analytically zero (biases in front of a node softmax, everything behind the 1-node pool's
    # softmax) and hold only rounding noise (1e-8 .. 1e-6) in the reference as well
    floor = 1e-4 * max(float(v.abs().max()) for v in grads.values())
    for k, ref in grads.items():
        err = float((named[k].grad.cpu() - ref).abs().max()) / max(floor, float(ref.abs().max()))
        assert err < 1e-3, (k, err)


@pytest.mark.gpu
def test_algorithm_update_runs():
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    dev = torch.device("cuda:0")
    hp = {"learning_rate": 1e-3, "weight_decay": 1e-4, "alpha": 100}
    alg = get_algorithm_class("HAGCN")(CFG, hp, dev).to(dev)
    X, y = torch.from_numpy(Z[f"{TAG}/X"]).to(dev), torch.from_numpy(Z[f"{TAG}/y"]).to(dev)
    l0 = alg.update(X, y, 1)["loss"]
    for _ in range(30):
        l1 = alg.update(X, y, 1)["loss"]
    assert np.isfinite(l1) and l1 < l0
