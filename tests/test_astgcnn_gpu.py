"""ASTGCNN (BASELINE.json configs[2]) drop-in: native TCN + Gaussian adjacency + Chebyshev aggregation vs the
UNMODIFIED reference model's outputs, gradients and running statistics (tests/golden/aux_metrics_data.npz,
make_golden.aux_golden), plus a few optimisation steps of the mirrored Algorithm wrapper."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
CFG = {"astgcnn_c": dict(num_nodes=14, time_length=50, encoder_out_dim=50, output_dim=64, K=3),
       "astgcnn_n": dict(num_nodes=20, time_length=50, encoder_out_dim=50, output_dim=64, K=3)}


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)      # relative to the tensor's own largest entry


def _sub(tag, grp):
    pre = f"{tag}/{grp}/"
    return {k[len(pre):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre)}


def test_state_dict_layout_matches_reference():
    """CPU: key set and shapes equal the reference's (checkpoint interchange), incl. the dead weight-norm branch."""
    from gnn_rul_benchmarking_b200.astgcnn import ASTGCNN_model
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ASTGCNN_model(**CFG["astgcnn_c"])
    sd, ref = model.state_dict(), _sub("astgcnn_c", "sd0")
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    with pytest.raises(RuntimeError):
        model(torch.rand(2, 14, 50))                     # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["astgcnn_c", "astgcnn_n"])
def test_model_matches_reference(tag):
    from gnn_rul_benchmarking_b200.astgcnn import ASTGCNN_model
    dev = torch.device("cuda:0")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ASTGCNN_model(**CFG[tag])
    model.load_state_dict(_sub(tag, "sd0"), strict=True)
    model = model.to(dev)
    X, y = torch.from_numpy(Z[f"{tag}/X"]).to(dev), torch.from_numpy(Z[f"{tag}/y"]).to(dev)
    model.eval()
    with torch.no_grad():
        assert _rel(model(X).cpu(), torch.from_numpy(Z[f"{tag}/y_eval"])) < 2e-5
    model.train()
    pred = model(X)
    assert _rel(pred.detach().cpu(), torch.from_numpy(Z[f"{tag}/y_train"])) < 2e-5
    torch.nn.functional.mse_loss(pred, y).backward()
    grads = _sub(tag, "grad")
    named = dict(model.named_parameters())
    for k, ref in grads.items():
        assert _rel(named[k].grad.cpu(), ref) < 1e-4, k
    sd = model.state_dict()
    for k, ref in _sub(tag, "sd1").items():
        assert torch.allclose(sd[k].cpu().to(ref.dtype), ref, atol=1e-5, rtol=1e-4), k


@pytest.mark.gpu
def test_algorithm_wrapper_trains():
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    from gnn_rul_benchmarking_b200.configs import ASTGCNN_CONFIGS, TRAIN_PARAMS
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alg = get_algorithm_class("ASTGCNN")(ASTGCNN_CONFIGS["NCMAPSS"], TRAIN_PARAMS, dev).to(dev)
    alg.train()
    g = torch.Generator().manual_seed(1)
    X, y = torch.rand(64, 20, 50, generator=g).to(dev), torch.rand(64, 1, generator=g).to(dev)
    l0 = alg.update(X, y, 1)["loss"]
    for _ in range(20):
        l1 = alg.update(X, y, 1)["loss"]
    assert l1 < l0
    assert int(alg.model.tcn.conv_block1[2].num_batches_tracked) == 21
