"""Rows 8f-2 / 8f-3: the device-resident data path and the evaluation metrics against outputs of the
UNMODIFIED reference (tests/golden/aux_metrics_data.npz, produced by make_golden.aux_golden from
utils._calc_metrics* and dataloader.Load_Dataset + DataLoader)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fc_stgnn_oracle as orc

Z = np.load(os.path.join(GOLDEN, "aux_metrics_data.npz"))
# the reference evaluates its metrics in float32 numpy; the oracle / kernel accumulate in float64 -> 5e-6 relative


@pytest.mark.parametrize("i", range(4))
def test_oracle_metrics_match_reference(i):
    pred, real = Z[f"m{i}/pred"], Z[f"m{i}/real"]
    got = orc.calc_metrics(pred, real, 125.0)
    assert np.allclose(got, Z[f"m{i}/all"], rtol=5e-6, atol=1e-7)


def test_device_loader_reproduces_reference_batches():
    """Same on-disk format, same channel-first fix-up, same shuffled batch order as the reference's
    Load_Dataset + DataLoader(shuffle=True) after torch.manual_seed(3) -- on CPU tensors here."""
    from gnn_rul_benchmarking_b200.data import DeviceLoader, DeviceWindowDataset
    samples = [a for a in Z["data/samples"]]                       # list of [50, 14] windows
    ds = DeviceWindowDataset(samples, Z["data/labels"], "cpu")
    assert tuple(ds.x_data.shape) == (23, 14, 50)
    torch.manual_seed(3)
    dl = DeviceLoader(ds, 5, shuffle=True, drop_last=False)
    assert len(dl) == 5
    xs, ys = zip(*list(dl))
    assert [x.shape[0] for x in xs] == [5, 5, 5, 5, 3]
    assert np.array_equal(torch.cat(xs).numpy(), Z["data/x_batches"])
    assert np.array_equal(torch.cat(ys).numpy(), Z["data/y_batches"])
    # data-parallel sharding: every rank sees a disjoint, equal share of every batch
    torch.manual_seed(3)
    r0 = torch.cat([x for x, _ in DeviceLoader(ds, 6, shuffle=True, drop_last=True, rank=0, world=2)])
    torch.manual_seed(3)
    r1 = torch.cat([x for x, _ in DeviceLoader(ds, 6, shuffle=True, drop_last=True, rank=1, world=2)])
    assert r0.shape == r1.shape == (9, 14, 50)
    both = torch.cat([r0, r1]).reshape(18, -1)
    assert torch.unique(both, dim=0).shape[0] == 18


def test_data_generator_reads_reference_format(tmp_path):
    from gnn_rul_benchmarking_b200.data import data_generator
    rng = np.random.default_rng(1)
    for name, n in (("train", 12), ("test", 5)):
        torch.save({"samples": [rng.uniform(0, 1, (50, 14)).astype(np.float32) for _ in range(n)],
                    "labels": rng.uniform(0, 1, (n, 1)).astype(np.float32), "max_ruls": 125}, tmp_path / f"{name}.pt")

    class Cfg:                      # configs/data_model_configs.py:8-16
        normalize, shuffle, drop_last = False, True, False

    tr, te, max_rul = data_generator(str(tmp_path), Cfg, {"batch_size": 4}, "cpu")
    assert max_rul == 125 and len(tr) == 3 and len(te) == 2
    X, y = next(iter(te))
    assert X.shape == (4, 14, 50) and y.shape == (4, 1) and X.dtype == torch.float32


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(4))
def test_device_metrics_match_reference(i):
    from gnn_rul_benchmarking_b200 import metrics
    dev = torch.device("cuda:0")
    pred, real = torch.from_numpy(Z[f"m{i}/pred"]).to(dev), torch.from_numpy(Z[f"m{i}/real"]).to(dev)
    assert np.allclose(metrics.calc_metrics(pred, real, 125.0), Z[f"m{i}/all"], rtol=5e-6, atol=1e-7)
    assert np.allclose(metrics.calc_metrics_aeroengine(pred, real, 125.0), Z[f"m{i}/aero"], rtol=5e-6, atol=1e-7)
    assert np.allclose(metrics.calc_metrics_bearing(pred, real, 125.0), Z[f"m{i}/bearing"], rtol=5e-6, atol=1e-7)
    with pytest.raises(RuntimeError):
        metrics.calc_metrics(pred.cpu(), real.cpu(), 125.0)


ADJ_CASES = ["pcc", "cosine", "cosine_big", "gauss", "gauss2_top10", "gauss2_top3"]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ADJ_CASES)
def test_adjacency_builders_match_reference(tag):
    """SURVEY 2.2 primitives A2-A4 (sibling models): device forward / backward vs the reference functions'
    own outputs and autograd gradients."""
    from gnn_rul_benchmarking_b200 import primitives as P
    dev = torch.device("cuda:0")
    fn = {"pcc": P.pcc_graph_construction, "cosine": P.cosine_distance, "cosine_big": P.cosine_distance,
          "gauss": P.gaussian_adjacency, "gauss2_top10": lambda t: P.compute_adjacency_matrix(t, 10),
          "gauss2_top3": lambda t: P.compute_adjacency_matrix(0.3 * t, 3)}[tag]
    x = torch.from_numpy(Z[f"adj/{tag}/x"]).to(dev).requires_grad_(True)
    a = fn(x)
    ref_a = torch.from_numpy(Z[f"adj/{tag}/a"])
    assert a.shape == ref_a.shape
    assert float((a.detach().cpu() - ref_a).abs().max()) < 2e-5
    (a * torch.from_numpy(Z[f"adj/{tag}/da"]).to(dev)).sum().backward()
    ref_dx = torch.from_numpy(Z[f"adj/{tag}/dx"])
    assert float((x.grad.cpu() - ref_dx).abs().max()) < 1e-4 * max(1.0, float(ref_dx.abs().max()))
    with pytest.raises(RuntimeError):
        P.cosine_distance(torch.randn(2, 4, 8))           # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["gcn", "gcn_wide", "cheb", "cheb_small"])
def test_aggregation_layers_match_reference(tag):
    """SURVEY 2.2 primitives M2 / M3: GCNLayer (STMSGCN) and ChebNet (ASTGCNN) mirrors -- native aggregation
    + library GEMM -- vs the reference modules' outputs and every gradient (dX, dA, parameters)."""
    from gnn_rul_benchmarking_b200 import primitives as P
    dev = torch.device("cuda:0")
    pre = f"layer/{tag}/"
    params = {k[len(pre) + 2:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(pre + "p/")}
    if tag.startswith("gcn"):
        layer = P.GCNLayer(params["linear.weight"].shape[1], params["linear.weight"].shape[0])
    else:
        layer = P.ChebNet(params["filters"].shape[1], params["filters"].shape[2], 3)
    layer.load_state_dict(params, strict=True)
    layer = layer.to(dev)
    x = torch.from_numpy(Z[pre + "x"]).to(dev).requires_grad_(True)
    a = torch.from_numpy(Z[pre + "a"]).to(dev).requires_grad_(True)
    y = layer(x, a)
    rel = lambda got, ref: float((got.cpu() - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    assert rel(y.detach(), torch.from_numpy(Z[pre + "y"])) < 2e-5
    (y * torch.from_numpy(Z[pre + "dy"]).to(dev)).sum().backward()
    assert rel(x.grad, torch.from_numpy(Z[pre + "dx"])) < 1e-4
    assert rel(a.grad, torch.from_numpy(Z[pre + "da"])) < 1e-4
    for k, p in layer.named_parameters():
        assert rel(p.grad, torch.from_numpy(Z[pre + "g/" + k])) < 1e-4, k
