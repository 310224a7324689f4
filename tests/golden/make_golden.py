"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned on
outputs of the reference's own classes imported read-only from /root/reference through
oracle/ref_shims.py:
  * model-level:  models/FC_STGNN/Model.py FC_STGNN_RUL -- eval forward, train forward with
    the positional-encoding dropout mask pinned, running statistics after one train forward,
    d(mse)/d(param) for every parameter and dX.
  * block-level:  models/FC_STGNN/Model_Base.py GraphConvpoolMPNN_block_v6 -- same, on random
    [B,T,N,C] inputs, strides 1 and 2.
Seeds follow utils.py:63-69 fix_randomness(seed) (torch.manual_seed) before construction.
The .npz files are small (float32, tiny batches) and are committed; this script is the
provenance record.  /root/reference does not travel to the GPU box; the fixtures do.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle.fc_stgnn_oracle import CONFIGS  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class PinnedDropout(torch.nn.Module):
    """Stands in for nn.Dropout(p) on a reference *instance* so the mask is reproducible."""

    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        if not self.training:
            return x
        return x * self.keep / (1.0 - self.p)


def _np(t):
    return t.detach().cpu().numpy().copy()


def perturb_bn_(module, gen):
    """Fresh BN layers have weight=1,bias=0,mean=0,var=1 -- give them non-trivial values so
    the goldens exercise the affine + running-stat paths."""
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=gen))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=gen))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=gen))


def model_golden(name, cfg, bs, seed):
    FC_STGNN_RUL, _ = ref_shims.fc_stgnn_classes()
    torch.manual_seed(seed)
    model = FC_STGNN_RUL(**cfg)
    gen = torch.Generator().manual_seed(seed + 1000)
    perturb_bn_(model, gen)
    N, Lraw = cfg["num_node"], cfg["num_patch"] * cfg["patch_size"]
    X = torch.rand(bs, N, Lraw, generator=gen)
    y = torch.rand(bs, 1, generator=gen)
    C = 2 * cfg["hidden_dim"]
    keep = (torch.rand(bs * N, cfg["num_patch"], C, generator=gen) >= 0.1).float()
    model.positional_encoding.dropout = PinnedDropout(keep, 0.1)

    out = {"X": _np(X), "y": _np(y), "keep": _np(keep).astype(np.uint8)}
    for k, v in model.state_dict().items():
        if k == "positional_encoding.pe":
            out["pe_head"] = _np(v[0, :64])
            continue
        out["sd0/" + k] = _np(v)
    model.eval()
    with torch.no_grad():
        out["y_eval"] = _np(model(X))
    model.train()
    Xg = X.clone().requires_grad_(True)
    pred = model(Xg)
    loss = torch.nn.functional.mse_loss(pred, y)
    loss.backward()
    out["y_train"] = _np(pred)
    out["loss"] = _np(loss)
    out["grad/X"] = _np(Xg.grad)
    for k, p in model.named_parameters():
        out["grad/" + k] = _np(p.grad)
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["sd1/" + k] = _np(v)
    path = os.path.join(OUT, f"model_{name}_b{bs}_s{seed}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def block_golden(tag, B, T, N, C, H, stride, seed):
    _, Block = ref_shims.fc_stgnn_classes()
    torch.manual_seed(seed)
    blk = Block(C, H, N, 10, time_window_size=2, stride=stride, decay=0.7, pool_choice="mean")
    gen = torch.Generator().manual_seed(seed + 2000)
    perturb_bn_(blk, gen)
    x = torch.randn(B, T, N, C, generator=gen)
    out = {"x": _np(x), "stride": np.int64(stride)}
    for k, v in blk.state_dict().items():
        out["sd0/" + k] = _np(v)
    blk.eval()
    with torch.no_grad():
        out["out_eval"] = _np(blk(x))
    blk.train()
    xg = x.clone().requires_grad_(True)
    o = blk(xg)
    dout = torch.randn(o.shape, generator=gen)
    (o * dout).sum().backward()
    out["out_train"] = _np(o)
    out["dout"] = _np(dout)
    out["grad/x"] = _np(xg.grad)
    for k, p in blk.named_parameters():
        out["grad/" + k] = _np(p.grad)
    for k, v in blk.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["sd1/" + k] = _np(v)
    out["mask"] = _np(blk.pre_relation)
    path = os.path.join(OUT, f"block_{tag}_s{stride}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def aux_golden():
    """Reference metrics (utils.py:136-201) and reference data path (dataloader/dataloader.py) outputs."""
    ref_shims.install()
    import utils as ref_utils                                    # noqa: E402  (reference module)
    from dataloader.dataloader import Load_Dataset               # noqa: E402
    rng = np.random.default_rng(0)
    out = {}
    for i, n in enumerate((1, 7, 100, 257)):
        real = rng.uniform(0.0, 1.0, n).astype(np.float32)
        pred = (real + rng.normal(0, 0.15, n)).astype(np.float32)
        s1, s2, mae, rmse = ref_utils._calc_metrics(pred, real, 125)
        sa, avg, rmse_a = ref_utils._calc_metrics_aeroengine(pred, real, 125)
        sb, mae_b, rmse_b = ref_utils._calc_metrics_bearing(pred, real, 125)
        out[f"m{i}/pred"], out[f"m{i}/real"] = pred, real
        out[f"m{i}/all"] = np.array([s1, s2, mae, rmse], dtype=np.float64)
        out[f"m{i}/aero"] = np.array([sa, avg, rmse_a], dtype=np.float64)
        out[f"m{i}/bearing"] = np.array([sb, mae_b, rmse_b], dtype=np.float64)
    # data path: C-MAPSS on-disk format (list of [L, C] float32 windows) -> channel-first tensors, and the
    # batch order of DataLoader(shuffle=True) after torch.manual_seed(3)
    samples = [rng.uniform(0, 1, (50, 14)).astype(np.float32) for _ in range(23)]
    labels = rng.uniform(0, 1, (23, 1)).astype(np.float32)
    ds = Load_Dataset(samples, labels, False)
    torch.manual_seed(3)
    dl = torch.utils.data.DataLoader(dataset=ds, batch_size=5, shuffle=True, drop_last=False, num_workers=0)
    xs, ys = zip(*[(x, y) for x, y in dl])
    out["data/samples"] = np.stack(samples)
    out["data/labels"] = labels
    out["data/x_batches"] = _np(torch.cat(xs))
    out["data/y_batches"] = _np(torch.cat(ys))
    # sibling adjacency builders: reference functions, outputs and input gradients
    from models.ST_GCN.Model import pcc_graph_construction           # noqa: E402
    from models.HAGCN.Model import cosine_distance                   # noqa: E402
    from models.ASTGCNN.Model import construct_graph                 # noqa: E402
    from models.STGNN.Model import compute_adjacency_matrix          # noqa: E402
    tg = torch.Generator().manual_seed(11)

    def adj_case(tag, fn, shape):
        x = torch.randn(*shape, generator=tg).requires_grad_(True)
        a = fn(x)
        da = torch.randn(a.shape, generator=tg)
        (a * da).sum().backward()
        out[f"adj/{tag}/x"], out[f"adj/{tag}/a"] = _np(x), _np(a)
        out[f"adj/{tag}/da"], out[f"adj/{tag}/dx"] = _np(da), _np(x.grad)

    adj_case("pcc", pcc_graph_construction, (5, 10, 40))                 # ST_GCN: 10 statistic nodes x 40 patches
    adj_case("cosine", cosine_distance, (4, 14, 60))                     # HAGCN: 14 sensors x 60 features
    adj_case("cosine_big", cosine_distance, (2, 160, 40))                # SAGCN: 160 patch nodes
    cg = construct_graph(50)
    with torch.no_grad():
        cg.P.weight.copy_(torch.eye(50))                                 # P = I: pins exp(-cdist) itself
    adj_case("gauss", cg, (3, 20, 50))                                   # ASTGCNN N-CMAPSS
    adj_case("gauss2_top10", lambda t: compute_adjacency_matrix(t, 10), (2, 3, 14, 5))   # STGNN
    adj_case("gauss2_top3", lambda t: compute_adjacency_matrix(0.3 * t, 3), (2, 1, 14, 50))
    # graph aggregation layers: reference modules, outputs and all gradients
    from models.STMSGCN.Model import GCNLayer                        # noqa: E402
    from models.ASTGCNN.Model import ChebNet                         # noqa: E402

    def layer_case(tag, layer, x_shape, pos_adj):
        x = torch.randn(*x_shape, generator=tg).requires_grad_(True)
        n = x_shape[1]
        a = torch.rand(x_shape[0], n, n, generator=tg) if pos_adj else 0.3 * torch.randn(x_shape[0], n, n, generator=tg)
        a.requires_grad_(True)
        y = layer(x, a)
        dy = torch.randn(y.shape, generator=tg)
        (y * dy).sum().backward()
        out[f"layer/{tag}/x"], out[f"layer/{tag}/a"], out[f"layer/{tag}/y"] = _np(x), _np(a), _np(y)
        out[f"layer/{tag}/dy"], out[f"layer/{tag}/dx"], out[f"layer/{tag}/da"] = _np(dy), _np(x.grad), _np(a.grad)
        for k, p in layer.named_parameters():
            out[f"layer/{tag}/p/{k}"] = _np(p)
            out[f"layer/{tag}/g/{k}"] = _np(p.grad)

    torch.manual_seed(5)
    layer_case("gcn", GCNLayer(5, 7), (4, 9, 5), True)                 # STMSGCN-style, positive degrees
    layer_case("gcn_wide", GCNLayer(40, 100), (2, 32, 40), True)       # SAGCN-sized
    layer_case("cheb", ChebNet(50, 64, 3), (3, 14, 50), False)         # ASTGCNN FD004
    layer_case("cheb_small", ChebNet(5, 8, 3), (6, 20, 5), False)      # STGNN N-CMAPSS patch features
    # ASTGCNN (BASELINE configs[2]): whole reference model, train + eval forward, all gradients, running stats
    from models.ASTGCNN.Model import ASTGCNN_model                   # noqa: E402
    for tag, cfg, bs in (("astgcnn_c", dict(num_nodes=14, time_length=50, encoder_out_dim=50, output_dim=64, K=3), 5),
                         ("astgcnn_n", dict(num_nodes=20, time_length=50, encoder_out_dim=50, output_dim=64, K=3), 3)):
        torch.manual_seed(2)
        mdl = ASTGCNN_model(**cfg)
        perturb_bn_(mdl, tg)
        for k, v in mdl.state_dict().items():
            out[f"{tag}/sd0/{k}"] = _np(v)
        X = torch.rand(bs, cfg["num_nodes"], 50, generator=tg)
        yt = torch.rand(bs, 1, generator=tg)
        mdl.eval()
        with torch.no_grad():
            out[f"{tag}/y_eval"] = _np(mdl(X))
        mdl.train()
        pred = mdl(X)
        loss = torch.nn.functional.mse_loss(pred, yt)
        loss.backward()
        out[f"{tag}/X"], out[f"{tag}/y"], out[f"{tag}/y_train"] = _np(X), _np(yt), _np(pred)
        for k, p in mdl.named_parameters():
            if p.grad is not None:
                out[f"{tag}/grad/{k}"] = _np(p.grad)
        for k, v in mdl.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out[f"{tag}/sd1/{k}"] = _np(v)
    # ST_GCN (BASELINE configs[2]; PHM2012 hparams with dropout disabled so that train mode is reproducible)
    from models.ST_GCN.Model import ST_GCN_model, segment_and_compute_features   # noqa: E402
    xs_ = torch.rand(37, 64, generator=tg) * 2 - 0.7
    out["stats/x"], out["stats/f"] = _np(xs_), _np(segment_and_compute_features(xs_))
    for tag, cfg, bs in (("stgcn_40", dict(num_patch=40, patch_size=64, dropout=0.2), 4),
                         ("stgcn_160", dict(num_patch=160, patch_size=16, dropout=0.2), 2)):
        torch.manual_seed(4)
        mdl = ST_GCN_model(**cfg)
        perturb_bn_(mdl, tg)
        for li, layer in enumerate(mdl.sg_tcn.layers):           # pin the dropout masks (train mode)
            keep = (torch.rand(bs, 10, cfg["num_patch"], generator=tg) >= 0.2).float()
            layer[2] = PinnedDropout(keep, 0.2)
            out[f"{tag}/keep{li}"] = _np(keep)
        for k, v in mdl.state_dict().items():
            out[f"{tag}/sd0/{k}"] = _np(v)
        X = torch.rand(bs, cfg["num_patch"] * cfg["patch_size"], generator=tg)
        yt = torch.rand(bs, 1, generator=tg)
        mdl.eval()
        with torch.no_grad():
            out[f"{tag}/y_eval"] = _np(mdl(X))
        mdl.train()
        pred = mdl(X)
        torch.nn.functional.mse_loss(pred, yt).backward()
        out[f"{tag}/X"], out[f"{tag}/y"], out[f"{tag}/y_train"] = _np(X), _np(yt), _np(pred)
        for k, p in mdl.named_parameters():
            if p.grad is not None:
                out[f"{tag}/grad/{k}"] = _np(p.grad)
        for k, v in mdl.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out[f"{tag}/sd1/{k}"] = _np(v)
    # STGNN (FD004 / N-CMAPSS hparams) and STMSGCN (PHM2012 hparams): whole reference models
    from models.STGNN.Model import STGNN_model                       # noqa: E402
    from models.STMSGCN.Model import STMSGCN_model                   # noqa: E402
    cases = (("stgnn_fd4", STGNN_model, dict(patch_size=50, num_patch=1, num_nodes=14, hidden_dim=64, K=3, top_k=10), (4, 14, 50)),
             ("stgnn_nc", STGNN_model, dict(patch_size=5, num_patch=10, num_nodes=20, hidden_dim=64, K=3, top_k=10), (3, 20, 50)),
             ("stmsgcn", STMSGCN_model, dict(num_patch=40, patch_size=64, interval=4, band_width=10,
                                             gcn_dims=[16, 64, 16, 1], gru_hidden_dim=8), (3, 2560)))
    for tag, cls, cfg, xshape in cases:
        torch.manual_seed(6)
        mdl = cls(**cfg)
        for k, v in mdl.state_dict().items():
            out[f"{tag}/sd0/{k}"] = _np(v)
        X = torch.rand(*xshape, generator=tg)
        yt = torch.rand(xshape[0], 1, generator=tg)
        mdl.train()
        pred = mdl(X)
        torch.nn.functional.mse_loss(pred, yt).backward()
        assert torch.isfinite(pred).all(), tag
        out[f"{tag}/X"], out[f"{tag}/y"], out[f"{tag}/y_train"] = _np(X), _np(yt), _np(pred)
        for k, p in mdl.named_parameters():
            if p.grad is not None:
                out[f"{tag}/grad/{k}"] = _np(p.grad)
    # GAT_LSTM (BASELINE configs[3], PHM2012 hparams): 11 patch statistics, one attention layer, the whole model.
    # The reference draws its attention dropout with F.dropout; the module-level name F of the imported reference
    # module is re-bound (in this process only) to a namespace whose dropout applies the recorded masks.
    import types
    import models.GAT_LSTM.Model as gat_ref                         # noqa: E402
    Fn = torch.nn.functional
    masks = []

    def pinned_dropout(att, p, training=True):
        if not training:
            return att
        keep = (torch.rand(att.shape, generator=tg) >= p).float()
        masks.append(keep)
        return att * keep / (1.0 - p)
    gat_ref.F = types.SimpleNamespace(softmax=Fn.softmax, dropout=pinned_dropout, leaky_relu=Fn.leaky_relu)
    xs = torch.randn(50, 64, generator=tg) * 0.7 + 0.2
    out["gat/stats_x"], out["gat/stats_f"] = _np(xs), _np(gat_ref.extract_features(xs))
    for tag, N, fin, fout, p, training in (("gat_layer_eval", 12, 11, 20, 0.2, False), ("gat_layer_train", 40, 30, 50, 0.2, True)):
        torch.manual_seed(8)
        layer = gat_ref.GraphAttentionLayer(fin, fout, p, 0.1)
        layer.train(training)
        h = torch.randn(3, N, fin, generator=tg).requires_grad_()
        adj = torch.eye(N).unsqueeze(0).repeat(3, 1, 1)
        idx = torch.arange(N - 1)
        adj[:, idx, idx + 1] = 1
        adj[:, idx + 1, idx] = 1
        wgt = torch.randn(3, N, fout, generator=tg)
        del masks[:]
        yl = layer(h, adj)
        (yl * wgt).sum().backward()
        for k, v in layer.state_dict().items():
            out[f"{tag}/sd/{k}"] = _np(v)
        out[f"{tag}/h"], out[f"{tag}/adj"], out[f"{tag}/w"], out[f"{tag}/y"], out[f"{tag}/dh"] = _np(h), _np(adj[0]), _np(wgt), _np(yl), _np(h.grad)
        if training:
            out[f"{tag}/keep"] = _np(masks[0])
        for k, prm in layer.named_parameters():
            out[f"{tag}/grad/{k}"] = _np(prm.grad)
    cfg = dict(num_patch=40, patch_size=64, hidden_dim=[300, 200, 100], lstm_hidden_dim=[30, 20], dropout=0.2)
    torch.manual_seed(9)
    mdl = gat_ref.GAT_LSTM_model(**cfg)
    for k, v in mdl.state_dict().items():
        out[f"gatlstm/sd0/{k}"] = _np(v)
    X = torch.rand(3, 2560, generator=tg)
    yt = torch.rand(3, 1, generator=tg)
    mdl.eval()
    with torch.no_grad():
        out["gatlstm/y_eval"] = _np(mdl(X))
    mdl.train()
    del masks[:]
    pred = mdl(X)
    torch.nn.functional.mse_loss(pred, yt).backward()
    assert torch.isfinite(pred).all() and len(masks) == 3
    out["gatlstm/X"], out["gatlstm/y"], out["gatlstm/y_train"] = _np(X), _np(yt), _np(pred)
    for li, keep in enumerate(masks):
        out[f"gatlstm/keep{li}"] = _np(keep)
    for k, prm in mdl.named_parameters():
        out[f"gatlstm/grad/{k}"] = _np(prm.grad)
    gat_ref.F = Fn
    # HAGCN (BASELINE configs[4], C-MAPSS hparams): whole reference model, train=True (prediction + KL term),
    # the two active encoder dropouts pinned
    from models.HAGCN.Model import HAGCN_model                       # noqa: E402
    for tag, cfg, xshape in (("hagcn_p10", dict(patch_size=10, num_patch=5, encoder_hidden_dim=60, hidden_dim=64, output_dim=32), (4, 14, 50)),):
        torch.manual_seed(10)
        mdl = HAGCN_model(**cfg)
        for k, v in mdl.state_dict().items():
            out[f"{tag}/sd0/{k}"] = _np(v)
        X = torch.rand(*xshape, generator=tg)
        yt = torch.rand(xshape[0], 1, generator=tg)
        mdl.eval()
        with torch.no_grad():
            out[f"{tag}/y_eval"] = _np(mdl(X))
        mdl.train()
        bsn, tl = xshape[0] * xshape[1], cfg["num_patch"]
        keeps = [(torch.rand(tl, bsn, 120, generator=tg) >= 0.2).float(), (torch.rand(tl, bsn, 60, generator=tg) >= 0.2).float()]
        mdl.TD.drop2, mdl.TD.drop3 = PinnedDropout(keeps[0], 0.2), PinnedDropout(keeps[1], 0.2)
        pred, kl = mdl(X, train=True)
        (torch.nn.functional.mse_loss(pred, yt) + 100.0 * kl).backward()
        assert torch.isfinite(pred).all() and torch.isfinite(kl)
        out[f"{tag}/X"], out[f"{tag}/y"], out[f"{tag}/y_train"], out[f"{tag}/kl"] = _np(X), _np(yt), _np(pred), _np(kl)
        out[f"{tag}/keep0"], out[f"{tag}/keep1"] = _np(keeps[0]).astype(np.uint8), _np(keeps[1]).astype(np.uint8)
        for k, prm in mdl.named_parameters():
            if prm.grad is not None:
                out[f"{tag}/grad/{k}"] = _np(prm.grad)
    # SAGCN (PHM2012 hparams: 160 patches of 16 samples): feature extraction and the whole reference model
    import models.SAGCN.Model as sag_ref                            # noqa: E402
    xs = torch.randn(60, 16, generator=tg) * 0.6
    out["sagcn/stats_x"], out["sagcn/stats_t"] = _np(xs), _np(sag_ref.extract_temporal_features(xs))
    out["sagcn/stats_f"] = _np(sag_ref.extract_frequency_features(xs))
    cfg = dict(num_patch=160, patch_size=16, gcn_hidden_dim=100, attention_hidden_dim=100)
    torch.manual_seed(12)
    mdl = sag_ref.SAGCN_model(**cfg)
    for k, v in mdl.state_dict().items():
        out[f"sagcn/sd0/{k}"] = _np(v)
    X = torch.randn(3, 2560, generator=tg) * 0.5
    yt = torch.rand(3, 1, generator=tg)
    mdl.train()
    out["sagcn/feat"] = _np(sag_ref.extract_features(X.reshape(3, 160, 16)))
    pred = mdl(X)
    torch.nn.functional.mse_loss(pred, yt).backward()
    assert torch.isfinite(pred).all()
    out["sagcn/X"], out["sagcn/y"], out["sagcn/y_train"] = _np(X), _np(yt), _np(pred)
    for k, prm in mdl.named_parameters():
        out[f"sagcn/grad/{k}"] = _np(prm.grad)
    path = os.path.join(OUT, "aux_metrics_data.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    aux_golden()
    model_golden("FD004", CONFIGS["FD004"], 6, 0)
    model_golden("FD004", CONFIGS["FD004"], 3, 1)
    model_golden("FD001", CONFIGS["FD001"], 4, 0)
    model_golden("NCMAPSS", CONFIGS["NCMAPSS"], 3, 0)
    model_golden("S2", CONFIGS["S2"], 2, 0)
    for stride in (1, 2):
        block_golden("S1_B5_T25_N14_C16_H8", 5, 25, 14, 16, 8, stride, 0)     # FD004 shapes
        block_golden("S2_B3_T50_N21_C14_H7", 3, 50, 21, 14, 7, stride, 1)     # north_star synthetic
        block_golden("NC_B3_T25_N20_C16_H8", 3, 25, 20, 16, 8, stride, 2)     # N-CMAPSS
        block_golden("FD3_B2_T12_N14_C48_H24", 2, 12, 14, 48, 24, stride, 3)  # FD003 widths
        block_golden("min_B2_T2_N14_C16_H8", 2, 2, 14, 16, 8, stride, 4)      # FD001: T == w (one window)
        block_golden("odd_B2_T7_N3_C4_H2", 2, 7, 3, 4, 2, stride, 5)          # ragged tail (T-w not /stride)


if __name__ == "__main__":
    main()
