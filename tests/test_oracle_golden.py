"""CPU: pins oracle/fc_stgnn_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import copy
import os

import pytest
import torch

from conftest import golden_files, load_golden
from oracle import fc_stgnn_oracle as orc

TOL = 2e-5          # fp32 re-association noise; the contract bound is 1e-4


def _cfg_of(name):
    return orc.CONFIGS[name.split("_")[1]]


def _clone(sd):
    return {k: v.clone() for k, v in sd.items()}


@pytest.mark.parametrize("path", golden_files("block"), ids=os.path.basename)
def test_block_forward_and_grads(path):
    g = load_golden(path)
    stride = int(g["stride"])
    x = g["x"]
    N = x.shape[2]
    assert torch.allclose(orc.decay_mask(N, 2, 0.7), g["mask"], atol=1e-7)
    sd = _clone(g["sd0"])
    out = orc.block_forward(x, sd, "", stride, training=False)
    assert (out - g["out_eval"]).abs().max() < TOL
    sd = {k: (v.clone().requires_grad_(True) if orc.is_param(k) else v.clone()) for k, v in g["sd0"].items()}
    xg = x.clone().requires_grad_(True)
    out = orc.block_forward(xg, sd, "", stride, training=True)
    assert (out - g["out_train"]).abs().max() < TOL
    (out * g["dout"]).sum().backward()
    scale = lambda t: max(1.0, float(t.abs().max()))
    assert (xg.grad - g["grad"]["x"]).abs().max() < 1e-4 * scale(g["grad"]["x"])
    for k, ref in g["grad"].items():
        if k == "x":
            continue
        assert (sd[k].grad - ref).abs().max() < 2e-4 * scale(ref), k
    for k, ref in g["sd1"].items():
        assert torch.allclose(sd[k].to(ref.dtype), ref, atol=1e-5, rtol=1e-5), k


@pytest.mark.parametrize("path", golden_files("block"), ids=os.path.basename)
def test_block_manual_backward_float64(path):
    """The hand-derived, re-associated backward (the kernel spec) == autograd, float64."""
    g = load_golden(path, torch.float64)
    stride = int(g["stride"])
    sd = {k: (v.clone().requires_grad_(True) if orc.is_param(k) else v.clone()) for k, v in g["sd0"].items()}
    xg = g["x"].clone().requires_grad_(True)
    out = orc.block_forward(xg, sd, "", stride, training=True, update_stats=False)
    (out * g["dout"]).sum().backward()
    with torch.no_grad():
        man = orc.block_backward_manual(g["x"], {k: v.detach() for k, v in sd.items()}, "", stride, g["dout"])
    assert (man["x"] - xg.grad).abs().max() < 1e-10
    for k, v in man.items():
        if k != "x":
            assert (v - sd[k].grad).abs().max() < 1e-9, k


@pytest.mark.parametrize("path", golden_files("model"), ids=os.path.basename)
def test_model_forward_and_grads(path):
    g = load_golden(path)
    cfg = _cfg_of(g["name"])
    C = 2 * cfg["hidden_dim"]
    assert torch.allclose(orc.positional_table(64, C), g["pe_head"], atol=1e-6)
    base = _clone(g["sd0"])
    base["positional_encoding.pe"] = orc.positional_table(5000, C).unsqueeze(0)
    with torch.no_grad():
        y = orc.model_forward(g["X"], _clone(base), cfg, training=False)
    assert (y - g["y_eval"]).abs().max() < TOL

    sd = {k: (v.clone().requires_grad_(True) if orc.is_param(k) else v.clone()) for k, v in base.items()}
    Xg = g["X"].clone().requires_grad_(True)
    pred = orc.model_forward(Xg, sd, cfg, training=True, dropout_keep=g["keep"].float())
    assert (pred - g["y_train"]).abs().max() < TOL
    loss = torch.nn.functional.mse_loss(pred, g["y"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for k, ref in g["grad"].items():
        got = Xg.grad if k == "X" else sd[k].grad
        assert (got - ref).abs().max() < 1e-4 * max(1.0, float(ref.abs().max())) + 1e-6, k
    for k, ref in g["sd1"].items():
        assert torch.allclose(sd[k].to(ref.dtype), ref, atol=1e-5, rtol=1e-4), k


def test_update_rule_matches_torch_adam_semantics():
    """OracleAlgorithm.update == forward/mse/backward/Adam(weight_decay) (algorithms.py:60-76)."""
    cfg = orc.CONFIGS["FD004"]
    alg = orc.OracleAlgorithm(cfg, orc.TRAIN_HPARAMS, seed=3)
    torch.manual_seed(0)
    X, y = torch.rand(8, 14, 50), torch.rand(8, 1)
    l0 = alg.update(X, y)["loss"]
    for _ in range(5):
        l1 = alg.update(X, y)["loss"]
    assert l1 < l0
    assert int(alg.sd["MPNN1.BN.num_batches_tracked"]) == 6
    assert orc.rmse([0.5, 0.5], [0.0, 1.0], 125.0) == pytest.approx(62.5)
