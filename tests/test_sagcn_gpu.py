"""SAGCN drop-in: native temporal patch statistics + cosine adjacency + sym-norm GCN aggregation, cuFFT spectral
statistics and the closed-form cumulative features vs the UNMODIFIED reference (tests/golden/aux_metrics_data.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
CFG = dict(num_patch=160, patch_size=16, gcn_hidden_dim=100, attention_hidden_dim=100)


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)      # relative to the tensor's own largest entry


def _sub(grp):
    pre = f"sagcn/{grp}/"
    return {k[len(pre):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre)}


def test_state_dict_layout_matches_reference():
    from gnn_rul_benchmarking_b200.sagcn import SAGCN_model
    sd, ref = SAGCN_model(**CFG).state_dict(), _sub("sd0")
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


def test_cumulative_features_closed_form_equals_the_loop():
    from gnn_rul_benchmarking_b200.sagcn import generate_cumulative_features
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 17, 5, generator=g)
    ref = torch.zeros_like(x)                       # models/SAGCN/Model.py:7-19, restated literally
    for p in range(1, 18):
        cs = torch.cumsum(x[:, :p, :], dim=1)
        ref[:, p - 1, :] = cs[:, p - 1, :] / torch.sqrt(cs[:, p - 1, :].abs().clamp_min(1e-12))
    assert torch.allclose(generate_cumulative_features(x), ref, atol=1e-6)


@pytest.mark.gpu
def test_feature_extraction_matches_reference():
    from gnn_rul_benchmarking_b200.primitives import extract_temporal_features
    from gnn_rul_benchmarking_b200.sagcn import extract_features, extract_frequency_features
    x = torch.from_numpy(Z["sagcn/stats_x"]).cuda()
    for got, ref in ((extract_temporal_features(x), Z["sagcn/stats_t"]), (extract_frequency_features(x), Z["sagcn/stats_f"])):
        ref = torch.from_numpy(ref)
        assert float(((got.cpu() - ref).abs() / (1.0 + ref.abs())).max()) < 3e-5
    X = torch.from_numpy(Z["sagcn/X"]).cuda()
    feat = extract_features(X.reshape(3, 160, 16)).cpu()
    assert _rel(feat, torch.from_numpy(Z["sagcn/feat"])) < 2e-5


@pytest.mark.gpu
def test_model_matches_reference():
    from gnn_rul_benchmarking_b200.sagcn import SAGCN_model
    dev = torch.device("cuda:0")
    model = SAGCN_model(**CFG)
    model.load_state_dict(_sub("sd0"), strict=True)
    model = model.to(dev).train()
    X, y = torch.from_numpy(Z["sagcn/X"]).to(dev), torch.from_numpy(Z["sagcn/y"]).to(dev)
    pred = model(X)
    assert _rel(pred.detach().cpu(), torch.from_numpy(Z["sagcn/y_train"])) < 5e-5
    torch.nn.functional.mse_loss(pred, y).backward()
    named = dict(model.named_parameters())
    for k, ref in _sub("grad").items():
        assert _rel(named[k].grad.cpu(), ref) < 5e-4, k


@pytest.mark.gpu
def test_algorithm_update_runs():
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    dev = torch.device("cuda:0")
    alg = get_algorithm_class("SAGCN")(CFG, {"learning_rate": 1e-4, "weight_decay": 1e-4}, dev).to(dev)
    X, y = torch.from_numpy(Z["sagcn/X"]).to(dev), torch.from_numpy(Z["sagcn/y"]).to(dev)
    l0 = alg.update(X, y, 1)["loss"]
    for _ in range(30):
        l1 = alg.update(X, y, 1)["loss"]
    assert np.isfinite(l1) and l1 < l0
