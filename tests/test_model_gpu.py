"""GPU parity of the drop-in FC_STGNN_RUL module against the reference-generated model goldens
(eval forward, train forward with the dropout mask pinned, every parameter gradient, running
statistics) -- the same quantities tests/test_oracle_golden.py pins for the CPU oracle."""
import os

import pytest
import torch

from conftest import (GRAD_TOL_FP32, GRAD_TOL_TC, OUT_TOL, assert_close_rel, assert_grads_close, golden_files,
                      load_golden, tc_shape)
from oracle import fc_stgnn_oracle as orc

pytestmark = pytest.mark.gpu


class PinnedDropout(torch.nn.Module):
    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        return x * self.keep / (1.0 - self.p) if self.training else x


@pytest.mark.parametrize("path", golden_files("model"), ids=os.path.basename)
def test_model_matches_reference_golden(path):
    from gnn_rul_benchmarking_b200.fc_stgnn import FC_STGNN_RUL
    g = load_golden(path)
    cfg = orc.CONFIGS[g["name"].split("_")[1]]
    dev = torch.device("cuda:0")
    model = FC_STGNN_RUL(**cfg)
    assert torch.allclose(model.positional_encoding.pe[0, :64], g["pe_head"], atol=1e-6)
    missing, unexpected = model.load_state_dict(g["sd0"], strict=False)
    assert missing == ["positional_encoding.pe"] and not unexpected
    model = model.to(dev)
    X, y = g["X"].to(dev), g["y"].to(dev)
    model.eval()
    with torch.no_grad():
        assert_close_rel(model(X).cpu(), g["y_eval"], OUT_TOL, "eval output")
    model.positional_encoding.dropout = PinnedDropout(g["keep"].float().to(dev), 0.1)
    model.train()
    pred = model(X)
    assert_close_rel(pred.detach().cpu(), g["y_train"], OUT_TOL, "train output")
    loss = torch.nn.functional.mse_loss(pred, y)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    tol = GRAD_TOL_TC if tc_shape(2 * cfg["hidden_dim"], cfg["hidden_dim"], cfg["num_node"]) else GRAD_TOL_FP32
    got = {k: p.grad.cpu() for k, p in model.named_parameters()}
    assert_grads_close(got, {k: g["grad"][k] for k in got}, tol, g["name"])
    sd = model.state_dict()
    for k, ref in g["sd1"].items():
        assert torch.allclose(sd[k].cpu().to(ref.dtype), ref, atol=1e-5, rtol=1e-4), k


def test_state_dict_keys_match_reference_layout():
    """checkpoint interchange (utils.py:111-120): key set and shapes equal the reference's."""
    from gnn_rul_benchmarking_b200.fc_stgnn import FC_STGNN_RUL
    g = load_golden(golden_files("model")[0])
    cfg = orc.CONFIGS[g["name"].split("_")[1]]
    sd = FC_STGNN_RUL(**cfg).state_dict()
    assert set(sd) == set(g["sd0"]) | {"positional_encoding.pe"}
    for k, v in g["sd0"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
