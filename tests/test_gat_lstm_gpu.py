"""GAT_LSTM (BASELINE.json configs[3]) drop-in: native 11 patch statistics + dense graph attention (forward and
backward) vs the UNMODIFIED reference (tests/golden/aux_metrics_data.npz; attention-dropout masks pinned)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
CFG = dict(num_patch=40, patch_size=64, hidden_dim=[300, 200, 100], lstm_hidden_dim=[30, 20], dropout=0.2)


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)      # relative to the tensor's own largest entry


def _sub(tag, grp):
    pre = f"{tag}/{grp}/"
    return {k[len(pre):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre)}


def test_state_dict_layout_matches_reference():
    from gnn_rul_benchmarking_b200.gat_lstm import GAT_LSTM_model
    sd, ref = GAT_LSTM_model(**CFG).state_dict(), _sub("gatlstm", "sd0")
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


@pytest.mark.gpu
def test_patch_statistics_match_reference():
    from gnn_rul_benchmarking_b200.primitives import extract_features
    got = extract_features(torch.from_numpy(Z["gat/stats_x"]).cuda()).cpu()
    ref = torch.from_numpy(Z["gat/stats_f"])
    assert float(((got - ref).abs() / (1.0 + ref.abs())).max()) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag,fin,fout,training", [("gat_layer_eval", 11, 20, False), ("gat_layer_train", 30, 50, True)])
def test_attention_layer_matches_reference(tag, fin, fout, training):
    from gnn_rul_benchmarking_b200.gat_lstm import GraphAttentionLayer
    dev = torch.device("cuda:0")
    layer = GraphAttentionLayer(fin, fout, 0.2, 0.1)
    layer.load_state_dict(_sub(tag, "sd"), strict=True)
    layer = layer.to(dev).train(training)
    if training:
        layer.keep_mask = torch.from_numpy(Z[f"{tag}/keep"]).to(dev)
    h = torch.from_numpy(Z[f"{tag}/h"]).to(dev).requires_grad_()
    y = layer(h, torch.from_numpy(Z[f"{tag}/adj"]).to(dev))
    assert _rel(y.detach().cpu(), torch.from_numpy(Z[f"{tag}/y"])) < 2e-5
    (y * torch.from_numpy(Z[f"{tag}/w"]).to(dev)).sum().backward()
    assert _rel(h.grad.cpu(), torch.from_numpy(Z[f"{tag}/dh"])) < 1e-4
    named = dict(layer.named_parameters())
    for k, ref in _sub(tag, "grad").items():
        assert _rel(named[k].grad.cpu(), ref) < 1e-4, k


@pytest.mark.gpu
def test_per_graph_adjacency_and_errors():
    from gnn_rul_benchmarking_b200.primitives import gat_attention
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    Wh = torch.randn(4, 9, 6, generator=g).to(dev).requires_grad_()
    aw, ab = torch.randn(1, 12, generator=g).to(dev).requires_grad_(), torch.randn(1, generator=g).to(dev).requires_grad_()
    adj = (torch.rand(4, 9, 9, generator=g) > 0.4).float().to(dev)
    keep = (torch.rand(4, 9, 9, generator=g) > 0.3).float().to(dev)
    out = gat_attention(Wh, aw, ab, adj, keep, 0.3, 0.1, 0.01)
    out.square().sum().backward()
    # float64 restatement of models/GAT_LSTM/Model.py:87-109
    W64, a64, b64 = Wh.detach().double().requires_grad_(), aw.detach().double().requires_grad_(), ab.detach().double().requires_grad_()
    s, t = W64 @ a64[0, :6], W64 @ a64[0, 6:]
    e = torch.nn.functional.leaky_relu(s[:, :, None] + t[:, None, :] + b64, 0.1)
    att = torch.softmax(e, dim=2) * keep.double() / 0.7 * adj.double()
    ref = torch.nn.functional.leaky_relu(att @ W64)
    ref.square().sum().backward()
    assert _rel(out.detach().double(), ref.detach()) < 1e-5
    assert _rel(Wh.grad.double(), W64.grad) < 1e-4 and _rel(aw.grad.double(), a64.grad) < 1e-4
    assert _rel(ab.grad.double(), b64.grad) < 1e-4
    with pytest.raises(ValueError):
        gat_attention(Wh, aw[:, :5], ab, adj)
    with pytest.raises(RuntimeError):
        gat_attention(Wh.cpu(), aw.cpu(), ab.cpu(), adj.cpu())


@pytest.mark.gpu
def test_model_matches_reference():
    from gnn_rul_benchmarking_b200.gat_lstm import GAT_LSTM_model
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False       # the goldens are CPU fp32; cuDNN's LSTM defaults to TF32
    model = GAT_LSTM_model(**CFG)
    model.load_state_dict(_sub("gatlstm", "sd0"), strict=True)
    model = model.to(dev)
    X, y = torch.from_numpy(Z["gatlstm/X"]).to(dev), torch.from_numpy(Z["gatlstm/y"]).to(dev)
    model.eval()
    with torch.no_grad():
        assert _rel(model(X).cpu(), torch.from_numpy(Z["gatlstm/y_eval"])) < 5e-5
    model.train()
    for li, layer in enumerate(model.gat_layers):
        layer.keep_mask = torch.from_numpy(Z[f"gatlstm/keep{li}"]).to(dev)
    pred = model(X)
    assert _rel(pred.detach().cpu(), torch.from_numpy(Z["gatlstm/y_train"])) < 5e-5
    torch.nn.functional.mse_loss(pred, y).backward()
    named = dict(model.named_parameters())
    for k, ref in _sub("gatlstm", "grad").items():
        assert _rel(named[k].grad.cpu(), ref) < 5e-4, k


@pytest.mark.gpu
def test_algorithm_update_runs():
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    dev = torch.device("cuda:0")
    alg = get_algorithm_class("GAT_LSTM")(CFG, {"learning_rate": 1e-3, "weight_decay": 1e-4}, dev).to(dev)
    X, y = torch.from_numpy(Z["gatlstm/X"]).to(dev), torch.from_numpy(Z["gatlstm/y"]).to(dev)
    l0 = alg.update(X, y, 1)["loss"]
    for _ in range(30):
        l1 = alg.update(X, y, 1)["loss"]
    assert np.isfinite(l1) and l1 < l0
