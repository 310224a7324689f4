"""STGNN and STMSGCN drop-ins: native adjacency (Gaussian top-k / outer product) + Chebyshev / sym-norm GCN
aggregation vs the UNMODIFIED reference models (tests/golden/aux_metrics_data.npz, made by make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
CFG = {"stgnn_fd4": ("STGNN", dict(patch_size=50, num_patch=1, num_nodes=14, hidden_dim=64, K=3, top_k=10)),
       "stgnn_nc": ("STGNN", dict(patch_size=5, num_patch=10, num_nodes=20, hidden_dim=64, K=3, top_k=10)),
       "stmsgcn": ("STMSGCN", dict(num_patch=40, patch_size=64, interval=4, band_width=10,
                                   gcn_dims=[16, 64, 16, 1], gru_hidden_dim=8))}


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)      # relative to the tensor's own largest entry


def _sub(tag, grp):
    pre = f"{tag}/{grp}/"
    return {k[len(pre):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre)}


def _build(tag):
    from gnn_rul_benchmarking_b200 import stgnn
    kind, cfg = CFG[tag]
    return (stgnn.STGNN_model if kind == "STGNN" else stgnn.STMSGCN_model)(**cfg)


@pytest.mark.parametrize("tag", list(CFG))
def test_state_dict_layout_matches_reference(tag):
    sd, ref = _build(tag).state_dict(), _sub(tag, "sd0")
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


@pytest.mark.gpu
def test_gram_adjacency_matches_bmm():
    from gnn_rul_benchmarking_b200.primitives import gram_adjacency
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 13, 5, generator=g).cuda().requires_grad_()
    w = torch.randn(7, 13, 13, generator=g).cuda()
    a = gram_adjacency(x)
    (a * w).sum().backward()
    xr = x.detach().double().requires_grad_()
    ar = torch.bmm(xr, xr.transpose(1, 2))
    (ar * w.double()).sum().backward()
    assert _rel(a.detach().double(), ar.detach()) < 1e-6
    assert _rel(x.grad.double(), xr.grad) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CFG))
def test_model_matches_reference(tag):
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False       # the goldens are CPU fp32; cuDNN's GRU defaults to TF32
    model = _build(tag)
    model.load_state_dict(_sub(tag, "sd0"), strict=True)
    model = model.to(dev).train()
    X, y = torch.from_numpy(Z[f"{tag}/X"]).to(dev), torch.from_numpy(Z[f"{tag}/y"]).to(dev)
    pred = model(X)
    assert _rel(pred.detach().cpu(), torch.from_numpy(Z[f"{tag}/y_train"])) < 1e-4
    torch.nn.functional.mse_loss(pred, y).backward()
    named = dict(model.named_parameters())
    grads = _sub(tag, "grad")
    assert grads
    for k, ref in grads.items():
        assert _rel(named[k].grad.cpu(), ref) < 5e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("name,tag", [("STGNN", "stgnn_fd4"), ("STMSGCN", "stmsgcn")])
def test_algorithm_update_runs(name, tag):
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    dev = torch.device("cuda:0")
    alg = get_algorithm_class(name)(CFG[tag][1], {"learning_rate": 1e-3, "weight_decay": 0.0}, dev).to(dev)
    X, y = torch.from_numpy(Z[f"{tag}/X"]).to(dev), torch.from_numpy(Z[f"{tag}/y"]).to(dev)
    l0 = alg.update(X, y, 1)["loss"]
    for _ in range(20):
        l1 = alg.update(X, y, 1)["loss"]
    assert np.isfinite(l1) and l1 < l0


@pytest.mark.gpu
@pytest.mark.parametrize("name,tag", [("STGNN", "stgnn_nc"), ("STMSGCN", "stmsgcn")])
def test_cuda_graph_update_matches_eager(name, tag):
    """_ModelAlgorithm.enable_cuda_graph: the captured update (forward + backward + Adam) replays to the same
    parameters as the eager update, and enabling it leaves the training state untouched."""
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    hp = {"learning_rate": 1e-3, "weight_decay": 1e-4}
    X, y = torch.from_numpy(Z[f"{tag}/X"]).to(dev), torch.from_numpy(Z[f"{tag}/y"]).to(dev)
    algs = []
    for _ in range(2):
        a = get_algorithm_class(name)(CFG[tag][1], hp, dev)
        a.model.load_state_dict(_sub(tag, "sd0"), strict=True)
        algs.append(a.to(dev).train())
    eager, graphed = algs
    eager.update(X, y, 1)                         # one eager step on both first: the optimizer state carries over
    graphed.update(X, y, 1)
    before = {k: v.clone() for k, v in graphed.model.state_dict().items()}
    graphed.enable_cuda_graph(X, y)
    for k, v in graphed.model.state_dict().items():
        assert torch.equal(v, before[k]), k
    for _ in range(4):
        le = eager.update(X * 0.9, y, 1)["loss"]
        lg = graphed.update(X * 0.9, y, 1)["loss"]
    assert abs(le - lg) < 1e-5 * max(1.0, abs(le))
    for (k, p), q in zip(eager.model.named_parameters(), graphed.model.parameters()):
        assert _rel(q.detach().cpu(), p.detach().cpu()) < 2e-5, k
