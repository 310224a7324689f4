"""flat_optim: flat parameter / gradient buffers + the one-kernel Adam for the sibling models' update rule
(algorithms/algorithms.py:139-163: torch.optim.Adam(lr, weight_decay) over model.parameters()), including the
reference's never-used modules (TemporalConvNet.net0 / net1, models/ST_GCN/Model.py:110-132), which torch's Adam skips
because their .grad stays None."""
import warnings

import pytest
import torch
import torch.nn as nn


class _WithDeadModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.live = nn.Linear(3, 2)
        self.dead = nn.Linear(3, 5)            # never used in forward
        self.bn = nn.BatchNorm1d(2)

    def forward(self, x):
        return self.bn(self.live(x))


def test_used_parameters_and_flat_views_cpu():
    from gnn_rul_benchmarking_b200.flat_optim import FlatParams, find_used_parameters
    torch.manual_seed(0)
    m = _WithDeadModule().train()
    x = torch.randn(4, 3)
    rm = m.bn.running_mean.clone()
    used = find_used_parameters(m, lambda: m(x).sum())
    assert {id(p) for p in used} == {id(m.live.weight), id(m.live.bias), id(m.bn.weight), id(m.bn.bias)}
    assert all(p.grad is None for p in m.parameters())            # the probe leaves no gradients behind
    assert torch.equal(m.bn.running_mean, rm)                     # ... and does not move BatchNorm statistics
    before = [p.detach().clone() for p in used]
    fl = FlatParams(used)
    assert fl.n % 4 == 0 and all(o % 4 == 0 for o in fl.offsets)
    for p, b, off in zip(used, before, fl.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == fl.param.data_ptr() + 4 * off
        assert p.grad.data_ptr() == fl.grad.data_ptr() + 4 * off
    m(x).sum().backward()                                         # autograd accumulates into the flat views
    assert float(fl.grad.abs().sum()) > 0
    for p, off in zip(used, fl.offsets):
        assert torch.equal(fl.grad[off:off + p.numel()].view_as(p), p.grad)
    assert m.dead.weight.grad is None
    used[0].grad = torch.ones_like(used[0])                       # a replaced .grad is gathered back
    fl.gather_stray_grads()
    assert torch.equal(fl.grad[:used[0].numel()], torch.ones(used[0].numel()))
    fl.check_bound()
    m.double()                                                    # re-allocates the parameters: the flat views are gone
    with pytest.raises(RuntimeError):
        fl.check_bound()


def _st_gcn(dev):
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    from gnn_rul_benchmarking_b200.configs import TRAIN_PARAMS
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alg = get_algorithm_class("ST_GCN")(dict(num_patch=20, patch_size=50, dropout=0.0), TRAIN_PARAMS, dev).to(dev)
    return alg.train()


@pytest.mark.gpu
def test_flat_adam_matches_torch_adam_on_st_gcn():
    """Three updates of ST_GCN (sensor-as-patch shape of BASELINE configs[2]) with torch.optim.Adam and with the flat
    one-kernel Adam: same losses and parameters; the never-used TemporalConvNet parameters do not move in either."""
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    a = _st_gcn(dev)
    torch.manual_seed(1)
    b = _st_gcn(dev)
    b.load_state_dict(a.state_dict())
    g = torch.Generator().manual_seed(2)
    X, y = torch.rand(32, 20, 50, generator=g).to(dev), torch.rand(32, 1, generator=g).to(dev)
    opt = b.use_flat_optimizer(X, y)
    named = dict(b.model.named_parameters())
    unused = [k for k, p in named.items() if p not in opt.state]
    assert unused and all(("net" in k or "downsample" in k) for k in unused), unused
    init = {k: named[k].detach().clone() for k in unused}
    for it in range(3):
        la, lb = a.update(X, y, it)["loss"], b.update(X, y, it)["loss"]
        assert abs(la - lb) < 1e-5 * (abs(la) + 1e-6), (it, la, lb)
    pa = dict(a.model.named_parameters())
    for k, p in named.items():
        d = float((p.detach() - pa[k].detach()).abs().max())
        assert d < 3 * 1e-3 * 0.05, (k, d)           # Adam moves an entry by <= lr per step; agreement to 5 % of that
    for k in unused:
        assert torch.equal(named[k].detach(), init[k])
        assert torch.equal(pa[k].detach(), init[k])


@pytest.mark.gpu
def test_flat_adam_update_captured_in_cuda_graph_matches_eager():
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    a = _st_gcn(dev)
    torch.manual_seed(3)
    b = _st_gcn(dev)
    b.load_state_dict(a.state_dict())
    g = torch.Generator().manual_seed(4)
    X, y = torch.rand(16, 20, 50, generator=g).to(dev), torch.rand(16, 1, generator=g).to(dev)
    a.use_flat_optimizer(X, y)
    b.use_flat_optimizer(X, y)
    b.enable_cuda_graph(X, y)
    for it in range(3):
        la, lb = a.update(X, y, it)["loss"], b.update(X, y, it)["loss"]
        assert abs(la - lb) < 1e-5 * (abs(la) + 1e-6), (it, la, lb)
    for (k, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert float((p.detach() - q.detach()).abs().max()) < 3 * 1e-3 * 0.05, k
    assert int(b.optimizer.step_dev) == 3
