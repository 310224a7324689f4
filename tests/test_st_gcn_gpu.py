"""ST_GCN (BASELINE.json configs[2]) drop-in: native patch statistics + Pearson adjacency + A.X aggregation +
TemporalConvNet vs the UNMODIFIED reference model (tests/golden/aux_metrics_data.npz; dropout masks pinned)."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
CFG = {"stgcn_40": dict(num_patch=40, patch_size=64, dropout=0.2), "stgcn_160": dict(num_patch=160, patch_size=16, dropout=0.2)}


class PinnedDropout(torch.nn.Module):
    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        return x * self.keep / (1.0 - self.p) if self.training else x


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)      # relative to the tensor's own largest entry


def _sub(tag, grp):
    pre = f"{tag}/{grp}/"
    return {k[len(pre):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre)}


def _build(tag):
    from gnn_rul_benchmarking_b200.st_gcn import ST_GCN_model
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ST_GCN_model(**CFG[tag])


def test_state_dict_layout_matches_reference():
    sd, ref = _build("stgcn_40").state_dict(), _sub("stgcn_40", "sd0")
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


@pytest.mark.gpu
def test_patch_statistics_match_reference():
    from gnn_rul_benchmarking_b200.primitives import segment_and_compute_features
    x = torch.from_numpy(Z["stats/x"]).cuda()
    got = segment_and_compute_features(x).cpu()
    ref = torch.from_numpy(Z["stats/f"])
    assert float(((got - ref).abs() / (1.0 + ref.abs())).max()) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["stgcn_40", "stgcn_160"])
def test_model_matches_reference(tag):
    dev = torch.device("cuda:0")
    model = _build(tag)
    model.load_state_dict(_sub(tag, "sd0"), strict=True)
    model = model.to(dev)
    for li, layer in enumerate(model.sg_tcn.layers):
        layer[2] = PinnedDropout(torch.from_numpy(Z[f"{tag}/keep{li}"]).to(dev), 0.2)
    X, y = torch.from_numpy(Z[f"{tag}/X"]).to(dev), torch.from_numpy(Z[f"{tag}/y"]).to(dev)
    model.eval()
    with torch.no_grad():
        assert _rel(model(X).cpu(), torch.from_numpy(Z[f"{tag}/y_eval"])) < 5e-5
    model.train()
    pred = model(X)
    assert _rel(pred.detach().cpu(), torch.from_numpy(Z[f"{tag}/y_train"])) < 5e-5
    torch.nn.functional.mse_loss(pred, y).backward()
    named = dict(model.named_parameters())
    for k, ref in _sub(tag, "grad").items():
        assert _rel(named[k].grad.cpu(), ref) < 2e-4, k
    sd = model.state_dict()
    for k, ref in _sub(tag, "sd1").items():
        assert torch.allclose(sd[k].cpu().to(ref.dtype), ref, atol=1e-5, rtol=1e-4), k
