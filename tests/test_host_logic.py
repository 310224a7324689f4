"""CPU: host-side logic -- config tables, algorithm registry, no-CPU-fallback behaviour, window
sharding, and the data-parallel gradient exchange on world_size-2 gloo."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs_match_oracle_tables():
    from gnn_rul_benchmarking_b200 import configs
    from oracle import fc_stgnn_oracle as orc
    assert configs.CONFIGS == orc.CONFIGS
    assert configs.TRAIN_PARAMS == orc.TRAIN_HPARAMS
    for name, cfg in configs.CONFIGS.items():
        T, h, N = cfg["num_patch"], cfg["hidden_dim"], cfg["num_node"]
        L = (T - 2) + 1 + (T - 2) // 2 + 1
        assert cfg["num_windows"] == L, name                              # fc1 input = h * num_windows * N
        k, P = cfg["encoder_conv_kernel"], cfg["patch_size"]
        l1 = P + 2 * (k // 2) - k + 1
        assert cfg["encoder_time_out"] == l1 + 2 - k + 1, name


def test_algorithm_registry_and_error_behaviour():
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS
    with pytest.raises(NotImplementedError):          # algorithms.py:31-32
        get_algorithm_class("NOPE")
    alg = get_algorithm_class("FC_STGNN")(CONFIGS["FD001"], TRAIN_PARAMS, "cpu")
    keys = set(alg.state_dict())
    assert "model.MPNN1.graph_construction.mapping.weight" in keys and "model.fc.fc4.bias" in keys
    assert "model.MPNN1.pre_relation" not in keys     # plain attribute in the reference (Model_Base.py:187)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        alg.update(torch.rand(2, 14, 50), torch.rand(2, 1), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        alg.model(torch.rand(2, 14, 50))


def test_mask_matrix_closed_form():
    from gnn_rul_benchmarking_b200.fc_stgnn import Mask_Matrix
    m = Mask_Matrix(3, 2, 0.7)
    assert m.shape == (6, 6)
    assert torch.allclose(m[:3, :3], torch.ones(3, 3)) and torch.allclose(m[:3, 3:], torch.full((3, 3), 0.7))


def test_shard_indices_partition():
    from gnn_rul_benchmarking_b200.dp import shard_indices
    n, world = 1003, 4
    parts = [list(shard_indices(n, r, world)) for r in range(world)]
    assert len({len(p) for p in parts}) == 1                      # equal counts: no rank waits in a collective
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(1000))                               # tail dropped, no overlap
    parts = [list(shard_indices(n, r, world, drop_tail=False)) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(n))


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gnn_rul_benchmarking_b200.dp import FlatGradAllReduce
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                                  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    # a registered module that forward never uses (the reference's TemporalConvNet keeps net0 / net1 like this,
    # models/ST_GCN/Model.py:110-132): its parameters get no gradient and must not stall or corrupt the exchange
    net.add_module("dead", torch.nn.Linear(3, 3))
    net.forward = lambda x: net[3](net[2](net[1](net[0](x))))
    hook = FlatGradAllReduce(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, weight_decay=1e-4)
    g = torch.Generator().manual_seed(7)
    X, y = torch.rand(8, 6, generator=g), torch.rand(8, 1, generator=g)
    Xr, yr = X[rank::world], y[rank::world]                        # window sharding
    for _ in range(3):                                             # the reference's update order (algorithms.py:72-74)
        loss = torch.nn.functional.mse_loss(net(Xr), yr)
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert net.dead.weight.grad is None
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((hook.n_allreduce, [t.tolist() for t in gathered]))
    dist.destroy_process_group()


def test_flat_grad_allreduce_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_ar, params = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_ar == 3                                               # exactly one collective per backward
    assert torch.allclose(torch.tensor(params[0]), torch.tensor(params[1]), atol=1e-7)   # replicas stay identical


def test_bench_reads_measured_peaks_in_any_documented_shape(tmp_path, monkeypatch):
    """bench.py's roofline denominator: MEASURED_PEAKS.json `hbm_gbs` (flat or nested, sustained preferred),
    else the profiling recipe's 6.65 TB/s fallback, and it says which."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_peak_gbs() == (6650.0, "fallback")
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6457.4, "bf16_tflops": 1590.0}))
    assert bench.measured_peak_gbs() == (6457.4, "measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm": {"burst_gbs": 6500.0, "sustained_gbs": 6300.0}}))
    assert bench.measured_peak_gbs() == (6300.0, "measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.measured_peak_gbs() == (6650.0, "fallback")


def test_sibling_algorithms_are_registered_with_reference_names():
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    for name in ("FC_STGNN", "ASTGCNN", "ST_GCN", "STGNN", "STMSGCN", "GAT_LSTM", "HAGCN", "SAGCN"):
        assert get_algorithm_class(name).__name__ == name


def test_tall_weight_gradient_helpers_match_autograd():
    """primitives.tall_gemm_t / col_sum / tall_linear / _ChebProject (split-K weight gradients of the sibling models'
    projections, pure torch: same numbers as the plain autograd formulas, any row count incl. ragged tails)."""
    import torch
    from gnn_rul_benchmarking_b200 import primitives as P
    g = torch.Generator().manual_seed(0)
    for R in (7, 300, 70003):
        a, b = torch.randn(R, 24, generator=g)[:, 3:19], torch.randn(R, 8, generator=g)      # strided columns
        ref = a.double().t() @ b.double()
        assert float((P.tall_gemm_t(a, b).double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max()) + 1e-6
        assert float((P.col_sum(a).double() - a.double().sum(0)).abs().max()) <= 2e-6 * R ** 0.5 + 1e-6
    x = torch.randn(40, 80, 5, generator=g, requires_grad=True)
    w = torch.randn(7, 5, generator=g, requires_grad=True)
    bias = torch.randn(7, generator=g, requires_grad=True)
    y = P.tall_linear(x, w, bias)
    assert y.shape == (40, 80, 7)
    y.square().sum().backward()
    got = [t.grad.clone() for t in (x, w, bias)]
    for t in (x, w, bias):
        t.grad = None
    torch.nn.functional.linear(x, w, bias).square().sum().backward()
    for a_, t in zip(got, (x, w, bias)):
        assert float((a_ - t.grad).abs().max()) <= 1e-5 * float(t.grad.abs().max())
    T = torch.randn(300, 3, 20, 6, generator=g, requires_grad=True)
    F = torch.randn(3, 6, 4, generator=g, requires_grad=True)
    P._ChebProject.apply(T, F).square().sum().backward()
    got = (T.grad.clone(), F.grad.clone())
    T.grad = F.grad = None
    torch.einsum("bknf,kfo->bno", T, F).square().sum().backward()
    assert float((got[0] - T.grad).abs().max()) <= 1e-5 * float(T.grad.abs().max())
    assert float((got[1] - F.grad).abs().max()) <= 1e-5 * float(F.grad.abs().max())


def test_rnn_weight_packing_matches_torch_layout():
    """rnn._layer_weights: what the recurrence kernel receives -- stacked W_hh per direction, b_ih + b_hh folded into the
    input projection (GRU: b_hn kept apart because it sits inside r * (W_hn h + b_hn))."""
    import torch
    from gnn_rul_benchmarking_b200 import rnn
    torch.manual_seed(1)
    m = rnn.LSTM(5, 6, batch_first=True, bidirectional=True)
    wih, whh, bias, bhn = rnn._layer_weights(m, 4)
    assert wih.shape == (48, 5) and whh.shape == (2, 24, 6) and bias.shape == (48,) and bhn is None
    assert torch.equal(whh[1], m.weight_hh_l0_reverse) and torch.equal(wih[24:], m.weight_ih_l0_reverse)
    assert torch.allclose(bias[:24], m.bias_ih_l0 + m.bias_hh_l0)
    g = rnn.GRU(4, 3, batch_first=True)
    wih, whh, bias, bhn = rnn._layer_weights(g, 3)
    assert whh.shape == (1, 9, 3) and bhn.shape == (1, 3)
    assert torch.allclose(bias[:6], g.bias_ih_l0[:6] + g.bias_hh_l0[:6])
    assert torch.allclose(bias[6:], g.bias_ih_l0[6:]) and torch.allclose(bhn[0], g.bias_hh_l0[6:])
    nb = rnn.GRU(4, 3, batch_first=True, bias=False)
    _, _, bias, bhn = rnn._layer_weights(nb, 3)
    assert bias is None and float(bhn.abs().sum()) == 0.0
