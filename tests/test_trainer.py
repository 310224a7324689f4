"""Flat on-disk window format and the device-resident epoch loop (8f-2 / 8f-3):
data.save_flat / load_flat / data_generator against the reference's pickled .pt layout (Data_read_CMAPSS.py:323-324,
dataloader.py:60-94), and trainer.DeviceTrainer against the oracle driven the way trainer.py:101-177 drives the
reference (same DataLoader shuffling stream, AverageMeter that is never reset, _calc_metrics per epoch)."""
import numpy as np
import pytest
import torch


class Cfg:                      # configs/data_model_configs.py:8-16
    normalize, shuffle, drop_last = False, True, False


def _split(rng, n, L=50, C=14):
    return {"samples": [rng.uniform(0, 1, (L, C)).astype(np.float32) for _ in range(n)],
            "labels": rng.uniform(0, 1, (n, 1)).astype(np.float32), "max_ruls": 125}


def test_flat_format_round_trip_equals_reference_pickle(tmp_path):
    from gnn_rul_benchmarking_b200.data import convert_pt, data_generator, load_flat
    rng = np.random.default_rng(5)
    a, b = tmp_path / "pt", tmp_path / "flat"
    a.mkdir(), b.mkdir()
    for name, n in (("train", 13), ("test", 6)):
        torch.save(_split(rng, n), a / f"{name}.pt")
        convert_pt(str(a / f"{name}.pt"), str(b / f"{name}.stgw"))
    fl = load_flat(str(b / "train.stgw"))
    assert fl["samples"].dtype == np.float32 and fl["samples"].shape == (13, 50, 14) and fl["max_ruls"] == 125.0
    out = []
    for d in (a, b):
        torch.manual_seed(11)
        tr, te, mr = data_generator(str(d), Cfg, {"batch_size": 4}, "cpu")
        out.append(([x for x, _ in tr], [y for _, y in te], mr))
    assert out[0][2] == out[1][2] == 125
    for u, v in zip(out[0][0] + out[0][1], out[1][0] + out[1][1]):
        assert torch.equal(u, v)


def test_flat_format_dict_test_sets_and_bad_files(tmp_path):
    """Bearing test sets are dicts keyed by bearing id / float (dataloader.py:82-90)."""
    from gnn_rul_benchmarking_b200.data import data_generator, load_flat, save_flat
    rng = np.random.default_rng(6)
    save_flat(str(tmp_path / "train.stgw"), {"samples": rng.uniform(0, 1, (9, 2560)), "labels": rng.uniform(0, 1, 9),
                                            "max_ruls": {"b1": 100.0, 2.0: 50}})
    save_flat(str(tmp_path / "test.stgw"), {"samples": {"b1": rng.uniform(0, 1, (3, 2560)), 2.0: rng.uniform(0, 1, (4, 2560))},
                                           "labels": {"b1": rng.uniform(0, 1, 3), 2.0: rng.uniform(0, 1, 4)}, "max_ruls": None})
    tr, te, mr = data_generator(str(tmp_path), Cfg, {"batch_size": 2}, "cpu")
    assert set(te) == {"b1", 2.0} and mr == {"b1": 100.0, 2.0: 50.0}
    X, y = next(iter(te[2.0]))
    assert X.shape == (2, 1, 2560) and y.shape == (2, 1)
    assert len(tr.dataset) == 9 and tr.dataset.x_data.shape == (9, 1, 2560)
    bad = tmp_path / "bad.stgw"
    bad.write_bytes(b"not a window file at all")
    with pytest.raises(ValueError):
        load_flat(str(bad))
    whole = (tmp_path / "train.stgw").read_bytes()
    (tmp_path / "cut.stgw").write_bytes(whole[:len(whole) // 2])
    with pytest.raises(ValueError):
        load_flat(str(tmp_path / "cut.stgw"))


def _synthetic(n, gen):
    y = torch.rand(n, 1, generator=gen)
    t = torch.linspace(0, 1, 50).view(1, 1, 50)
    slope = torch.randn(1, 14, 1, generator=gen) * 0.5
    X = 0.5 + slope * (1.0 - y.view(n, 1, 1)) * t + 0.05 * torch.randn(n, 14, 50, generator=gen)
    return X.clamp(0, 1), y


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_device_trainer_reproduces_the_reference_epoch_loop(use_graph, tmp_path):
    """Three epochs on a synthetic FD004-format set written in the flat format: per-epoch running loss average, RMSE,
    MAE and both scores equal the oracle's, which is driven exactly like trainer.py drives the reference
    (DataLoader(shuffle=True) order from the same torch.manual_seed, ragged last batch, eval after every epoch).
    Dropout is off on both sides (the streams differ by construction; tests/test_training_rmse_gpu.py covers it)."""
    from oracle import fc_stgnn_oracle as orc
    from gnn_rul_benchmarking_b200 import data, trainer
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS
    dev = torch.device("cuda:0")
    cfg = CONFIGS["FD004"]
    gen = torch.Generator().manual_seed(77)
    Xtr, ytr = _synthetic(530, gen)                     # 5 batches of 100 + a tail of 30
    Xte, yte = _synthetic(130, gen)
    for name, (X, y) in (("train", (Xtr, ytr)), ("test", (Xte, yte))):
        data.save_flat(str(tmp_path / f"{name}.stgw"), {"samples": X.permute(0, 2, 1).numpy(), "labels": y.numpy(),
                                                        "max_ruls": 125})
    torch.manual_seed(0)
    alg = get_algorithm_class("FC_STGNN")(cfg, TRAIN_PARAMS, dev)
    alg.model.positional_encoding.dropout.p = 0.0
    sd = {k: v.detach().clone() for k, v in alg.model.state_dict().items()}
    oracle = orc.OracleAlgorithm(cfg, TRAIN_PARAMS, sd={k: v.clone() for k, v in sd.items()})
    keep = torch.full((1,), 1.0 - orc.PE_DROPOUT)       # the oracle applies keep / (1 - p): this mask is the identity
    alg = alg.to(dev)
    tr, te, max_rul = data.data_generator(str(tmp_path), Cfg, {"batch_size": 100}, dev)
    t = trainer.DeviceTrainer(alg, tr, te, max_rul, num_epochs=3, use_cuda_graph=use_graph)

    torch.manual_seed(123)
    hist = t.fit()
    # the oracle, driven like trainer.py:101-126 + 134-177 + 189-260
    torch.manual_seed(123)
    ds = torch.utils.data.TensorDataset(Xtr, ytr)
    dl = torch.utils.data.DataLoader(ds, batch_size=100, shuffle=True, drop_last=False)
    lsum, lcnt, best = 0.0, 0, float("inf")
    for ep in range(3):
        for X, y in dl:
            l = oracle.update(X, y, dropout_keep=keep.expand(X.shape[0] * 14, cfg["num_patch"], 2 * cfg["hidden_dim"]))["loss"]
            lsum += l * X.shape[0]
            lcnt += X.shape[0]
        pred = oracle.predict(Xte).view(-1)
        s1, s2, mae, rmse = orc.calc_metrics(pred, yte.view(-1), 125.0)
        h = hist[ep]
        assert abs(h["loss"] - lsum / lcnt) < 2e-4 * (lsum / lcnt), (ep, h["loss"], lsum / lcnt)
        assert abs(h["test"]["RMSE"] - rmse) <= 0.05, (ep, h["test"]["RMSE"], rmse)          # north_star bound
        assert abs(h["test"]["MAE"] - mae) <= 0.05
        assert abs(h["test"]["Score_v1"] - s1) <= 2e-3 * abs(s1) + 1e-3
        assert abs(h["test"]["Score_v2"] - s2) <= 2e-3 * abs(s2) + 1e-6
        best = min(best, rmse)
    assert abs(t.best_result[3][-1] - best) <= 0.05 and len(t.best_result[3]) >= 2
