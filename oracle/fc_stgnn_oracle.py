"""TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") of the reference FC_STGNN hot path.

A from-scratch functional restatement, in plain PyTorch CPU ops, of what the reference
computes on the path named by BASELINE.json `north_star`:

    models/FC_STGNN/Model.py:43-85          FC_STGNN_RUL.forward          -> model_forward()
    models/FC_STGNN/Model_Base.py:12-41     Feature_extractor_1DCNN_RUL   -> encoder_forward()
    models/FC_STGNN/Model_Base.py:111-134   PositionalEncoding            -> positional_table(), pe_dropout()
    models/FC_STGNN/Model_Base.py:137-148   Conv_GraphST (unfold)         -> windows()
    models/FC_STGNN/Model_Base.py:44-67     Dot_Graph_Construction_weights-> block_forward() "A"
    models/FC_STGNN/Model_Base.py:150-170   Mask_Matrix                   -> decay_mask()
    models/FC_STGNN/Model_Base.py:72-107    MPNN_mk_v2 (k=1)              -> block_forward() "Y"
    models/FC_STGNN/Model_Base.py:175-225   GraphConvpoolMPNN_block_v6    -> block_forward()
    algorithms/algorithms.py:67-76          FC_STGNN.update               -> OracleAlgorithm.update()
    utils.py:148-151                        rmse_value                    -> rmse()

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the *unmodified reference itself*, run in
the build container through oracle/ref_shims.py by tests/golden/make_golden.py and
committed as tests/golden/*.npz.  tests/test_oracle_golden.py checks this file against
every one of them (forward eval/train, running statistics, all parameter grads, dX).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
leg may import this module.  The product package never does: it fails loudly when its
CUDA library is missing.

All functions work on a flat dict `sd` with the reference's state_dict key names
(SURVEY.md section 8b), so reference checkpoints load verbatim.  dtype follows the inputs
(float32 for parity/timing, float64 for gradient cross-checks).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

LEAKY = 0.01          # F.leaky_relu default slope (Model_Base.py:60,107)
BN_EPS = 1e-5         # nn.BatchNorm1d defaults
BN_MOMENTUM = 0.1
DECAY = 0.7           # Model.py:12
WINDOW = (2, 2)       # Model.py:13
STRIDE = (1, 2)       # Model.py:14
PE_DROPOUT = 0.1      # Model.py:24


# ----------------------------------------------------------------------------- helpers
def num_windows(T: int, w: int, s: int) -> int:
    return (T - w) // s + 1


def decay_mask(N: int, w: int, decay: float, dtype=torch.float32, device=None) -> torch.Tensor:
    """mask[i,k] = decay^|i//N - k//N|  (closed form of Mask_Matrix, Model_Base.py:150-170)."""
    t = torch.arange(w * N) // N
    return torch.as_tensor(decay, dtype=torch.float64).pow((t[:, None] - t[None, :]).abs().double()).to(dtype).to(device)


def windows(x: torch.Tensor, w: int, s: int) -> torch.Tensor:
    """[B,T,N,C] -> [B,L,w*N,C]; node m = j*N+n (time-major inside the window).
    Same elements as Conv_GraphST + the transpose/reshape at Model_Base.py:194-198."""
    B, T, N, C = x.shape
    L = num_windows(T, w, s)
    idx = (torch.arange(L, device=x.device)[:, None] * s + torch.arange(w, device=x.device)[None, :])       # [L,w]
    return x[:, idx].reshape(B, L, w * N, C)


def _bn_train(v: torch.Tensor, weight, bias, run_mean, run_var, nbt, update: bool):
    """BatchNorm over all leading dims, features last.  Biased var for the normalisation,
    unbiased for the running estimate (PyTorch semantics)."""
    flat = v.reshape(-1, v.shape[-1])
    R = flat.shape[0]
    mean = flat.mean(0)
    var = flat.var(0, unbiased=False)
    if update:
        with torch.no_grad():
            run_mean.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().to(run_mean.dtype))
            run_var.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (var.detach() * (R / max(R - 1, 1))).to(run_var.dtype))
            nbt.add_(1)
    return (v - mean) / torch.sqrt(var + BN_EPS) * weight + bias


def _bn_eval(v, weight, bias, run_mean, run_var):
    return (v - run_mean.to(v.dtype)) / torch.sqrt(run_var.to(v.dtype) + BN_EPS) * weight + bias


# ----------------------------------------------------------------------------- the graph-conv block
def block_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, stride: int,
                  w: int = 2, decay: float = DECAY, training: bool = False,
                  update_stats: bool = True, return_intermediates: bool = False):
    """GraphConvpoolMPNN_block_v6.forward (Model_Base.py:190-225), pool_choice='mean'.

    x [B,T,N,C] -> out [B,L,N,H].  SURVEY.md section 9.1 symbol for symbol.
    """
    B, T, N, C = x.shape
    M = w * N
    Wm, bm = sd[prefix + "graph_construction.mapping.weight"], sd[prefix + "graph_construction.mapping.bias"]
    Wt, bt = sd[prefix + "MPNN.theta.0.weight"], sd[prefix + "MPNN.theta.0.bias"]
    H = Wt.shape[0]

    G = windows(x, w, stride)                                  # [B,L,M,C]
    L = G.shape[1]
    # --- adjacency (Model_Base.py:49-63): learned map, Gram, diag out, lrelu, softmax, +I
    Fm = G @ Wm.t() + bm
    S = Fm @ Fm.transpose(-1, -2)                              # [B,L,M,M]
    eye = torch.eye(M, dtype=x.dtype, device=x.device)
    Lam = F.leaky_relu(S - 1e8 * eye, LEAKY)
    P = torch.softmax(Lam, dim=-1)
    A = (P + eye) * decay_mask(N, w, decay, x.dtype, x.device)           # Model_Base.py:203
    # --- BN over the unfolded rows (Model_Base.py:206-208)
    k0 = prefix + "BN."
    if training:
        Xb = _bn_train(G, sd[k0 + "weight"], sd[k0 + "bias"], sd[k0 + "running_mean"],
                       sd[k0 + "running_var"], sd[k0 + "num_batches_tracked"], update_stats)
    else:
        Xb = _bn_eval(G, sd[k0 + "weight"], sd[k0 + "bias"], sd[k0 + "running_mean"], sd[k0 + "running_var"])
    # --- MPNN_mk_v2, k=1 (Model_Base.py:85-107)
    Z = A @ Xb
    Yp = Z @ Wt.t() + bt                                       # [B,L,M,H]
    k1 = prefix + "MPNN.bn1."
    if training:
        Yn = _bn_train(Yp, sd[k1 + "weight"], sd[k1 + "bias"], sd[k1 + "running_mean"],
                       sd[k1 + "running_var"], sd[k1 + "num_batches_tracked"], update_stats)
    else:
        Yn = _bn_eval(Yp, sd[k1 + "weight"], sd[k1 + "bias"], sd[k1 + "running_mean"], sd[k1 + "running_var"])
    Ya = F.leaky_relu(Yn, LEAKY)
    out = Ya.reshape(B, L, w, N, H).mean(2)                    # Model_Base.py:212-216
    if return_intermediates:
        return out, dict(G=G, F=Fm, S=S, P=P, A=A, Xb=Xb, Z=Z, Yp=Yp, Yn=Yn)
    return out


def block_backward_manual(x, sd, prefix, stride, dout, w=2, decay=DECAY):
    """Hand-derived training-mode gradient of block_forward, written
    in the re-associated form the CUDA kernels use:  Y' = A.(Xb.Wt^T) + bt  (V = Xb.Wt^T is
    computed once per time step; dV, dF are folded back per time step before the
    per-feature tail).  Returns dict of grads keyed like the state_dict plus 'x'.
    Used by tests to validate the derivation against autograd in float64.
    """
    B, T, N, C = x.shape
    M = w * N
    Wm, bm = sd[prefix + "graph_construction.mapping.weight"], sd[prefix + "graph_construction.mapping.bias"]
    Wt = sd[prefix + "MPNN.theta.0.weight"]
    g0, b0 = sd[prefix + "BN.weight"], sd[prefix + "BN.bias"]
    g1 = sd[prefix + "MPNN.bn1.weight"]
    H = Wt.shape[0]
    L = num_windows(T, w, stride)
    R = B * L * M
    dt = x.dtype
    cnt = torch.zeros(T, dtype=dt)                              # windows covering each time step
    for l in range(L):
        cnt[l * stride:l * stride + w] += 1

    # ---- forward recompute, per time step
    sw = cnt.view(1, T, 1, 1)
    mu0 = (x * sw).sum((0, 1, 2)) / R
    var0 = ((x - mu0) ** 2 * sw).sum((0, 1, 2)) / R
    r0 = 1.0 / torch.sqrt(var0 + BN_EPS)
    Xh = (x - mu0) * r0                                         # [B,T,N,C]
    Xb = Xh * g0 + b0
    Fm = x @ Wm.t() + bm                                        # [B,T,N,C]
    V = Xb @ Wt.t()                                             # [B,T,N,H]  (no bias)
    Fw, Vw = windows(Fm, w, stride), windows(V, w, stride)      # [B,L,M,*]
    S = Fw @ Fw.transpose(-1, -2)
    eye = torch.eye(M, dtype=dt)
    offd = 1.0 - eye
    Lam = torch.where(eye.bool(), torch.full_like(S, -float("inf")), F.leaky_relu(S, LEAKY))
    P = torch.softmax(Lam, -1)
    mask = decay_mask(N, w, decay, dt)
    A = (P + eye) * mask
    Yp = A @ Vw + sd[prefix + "MPNN.theta.0.bias"]
    mu1 = Yp.mean((0, 1, 2))
    var1 = Yp.var((0, 1, 2), unbiased=False)
    r1 = 1.0 / torch.sqrt(var1 + BN_EPS)
    Yh = (Yp - mu1) * r1
    Yn = Yh * g1 + sd[prefix + "MPNN.bn1.bias"]

    # ---- backward
    dYa = (dout / w).unsqueeze(2).expand(B, L, w, N, H).reshape(B, L, M, H)
    dYn = dYa * torch.where(Yn > 0, torch.ones_like(Yn), torch.full_like(Yn, LEAKY))
    grads = {}
    grads[prefix + "MPNN.bn1.weight"] = (dYn * Yh).sum((0, 1, 2))
    grads[prefix + "MPNN.bn1.bias"] = dYn.sum((0, 1, 2))
    dYh = dYn * g1
    dYp = r1 * (dYh - dYh.mean((0, 1, 2)) - Yh * (dYh * Yh).mean((0, 1, 2)))
    grads[prefix + "MPNN.theta.0.bias"] = dYp.sum((0, 1, 2))
    dA = dYp @ Vw.transpose(-1, -2)                             # [B,L,M,M]
    dVw = A.transpose(-1, -2) @ dYp                             # [B,L,M,H]
    dP = dA * mask
    dLam = P * (dP - (dP * P).sum(-1, keepdim=True))
    dS = dLam * torch.where(S > 0, torch.ones_like(S), torch.full_like(S, LEAKY)) * offd
    dFw = (dS + dS.transpose(-1, -2)) @ Fw                      # [B,L,M,C]

    def fold(dw):                                               # [B,L,M,K] -> [B,T,N,K]
        K = dw.shape[-1]
        acc = torch.zeros(B, T, N, K, dtype=dt)
        d5 = dw.reshape(B, L, w, N, K)
        for l in range(L):
            acc[:, l * stride:l * stride + w] += d5[:, l]
        return acc

    dV, dF = fold(dVw), fold(dFw)                               # per time step
    grads[prefix + "MPNN.theta.0.weight"] = torch.einsum("btnh,btnc->hc", dV, Xb)
    dXb = dV @ Wt                                               # folded over windows
    grads[prefix + "BN.weight"] = (dXb * Xh).sum((0, 1, 2))
    grads[prefix + "BN.bias"] = dXb.sum((0, 1, 2))
    dXh = dXb * g0
    m1 = dXh.sum((0, 1, 2)) / R
    m2 = (dXh * Xh).sum((0, 1, 2)) / R
    grads[prefix + "graph_construction.mapping.weight"] = torch.einsum("btno,btni->oi", dF, x)
    grads[prefix + "graph_construction.mapping.bias"] = dF.sum((0, 1, 2))
    grads["x"] = dF @ Wm + r0 * (dXh - sw * (m1 + Xh * m2))
    return grads


# ----------------------------------------------------------------------------- encoder / PE / head
def positional_table(T: int, d_model: int, dtype=torch.float32) -> torch.Tensor:
    """pe[:T] of PositionalEncoding (Model_Base.py:118-126); note ln(100), not ln(10000)."""
    pos = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(100.0) / d_model))
    pe = torch.zeros(T, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)[:, : d_model // 2]
    return pe.to(dtype)


def encoder_forward(rows: torch.Tensor, sd, training: bool, kernel: int) -> torch.Tensor:
    """Feature_extractor_1DCNN_RUL (Model_Base.py:33-41) + nonlin_map2 (Model.py:18-22,57-59).
    rows [R, P] (one patch of one sensor per row) -> [R, 2h]."""
    mom = BN_MOMENTUM

    def bn(v, key):
        if training:
            sd[key + "num_batches_tracked"].add_(1)
        return F.batch_norm(v, sd[key + "running_mean"], sd[key + "running_var"], sd[key + "weight"],
                            sd[key + "bias"], training, mom, BN_EPS)

    h = rows.unsqueeze(1)                                       # [R,1,P]
    h = F.conv1d(h, sd["nonlin_map.conv_block1.0.weight"], None, 1, kernel // 2)
    h = torch.relu(bn(h, "nonlin_map.conv_block1.1."))
    h = F.conv1d(h, sd["nonlin_map.conv_block2.0.weight"], None, 1, 1)
    h = torch.relu(bn(h, "nonlin_map.conv_block2.1."))
    h = h.reshape(h.shape[0], -1)
    h = F.linear(h, sd["nonlin_map2.0.weight"], sd["nonlin_map2.0.bias"])
    return bn(h, "nonlin_map2.1.")


def model_forward(X: torch.Tensor, sd, cfg: dict, training: bool = False,
                  dropout_keep: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FC_STGNN_RUL.forward (Model.py:43-85).  X [bs,N,L] -> [bs,1].

    dropout_keep: optional 0/1 keep mask laid out like the reference's dropout input
    [bs*N, T, 2h] (Model.py:64-65); train mode multiplies by keep/(1-p).  None in train
    mode draws it from torch's CPU generator with the same call the reference makes.
    """
    bs, N, _ = X.shape
    T, Pz = cfg["num_patch"], cfg["patch_size"]
    x = X.reshape(bs, N, T, Pz).transpose(1, 2)                 # [bs,T,N,P]
    enc = encoder_forward(x.reshape(bs * T * N, Pz), sd, training, cfg["encoder_conv_kernel"])
    C = enc.shape[-1]
    enc = enc.reshape(bs, T, N, C)
    pe = sd["positional_encoding.pe"][0, :T].to(enc.dtype)      # [T,C]
    h = enc + pe[None, :, None, :]
    if training:
        hb = h.transpose(1, 2).reshape(bs * N, T, C)
        if dropout_keep is None:
            hb = F.dropout(hb, PE_DROPOUT, True)
        else:
            hb = hb * dropout_keep.to(hb.dtype) / (1.0 - PE_DROPOUT)
        h = hb.reshape(bs, N, T, C).transpose(1, 2)
    o1 = block_forward(h, sd, "MPNN1.", STRIDE[0], WINDOW[0], DECAY, training)
    o2 = block_forward(h, sd, "MPNN2.", STRIDE[1], WINDOW[1], DECAY, training)
    f = torch.cat([o1.reshape(bs, -1), o2.reshape(bs, -1)], -1)
    f = torch.relu(F.linear(f, sd["fc.fc1.weight"], sd["fc.fc1.bias"]))
    f = torch.relu(F.linear(f, sd["fc.fc2.weight"], sd["fc.fc2.bias"]))
    f = torch.relu(F.linear(f, sd["fc.fc3.weight"], sd["fc.fc3.bias"]))
    return F.linear(f, sd["fc.fc4.weight"], sd["fc.fc4.bias"])


# ----------------------------------------------------------------------------- init + update rule
PARAM_SUFFIXES = ("weight", "bias")


def init_state(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """A state dict with the reference's keys/shapes and PyTorch-default-like init
    (kaiming-uniform a=sqrt(5) for conv/linear, BN weight 1 / bias 0).  Not bit-identical to
    the reference constructor's RNG stream -- parity tests load reference state dicts."""
    g = torch.Generator().manual_seed(seed)
    h, E, Eh, k = cfg["hidden_dim"], cfg["encoder_out_dim"], cfg["encoder_hidden_dim"], cfg["encoder_conv_kernel"]
    C, N = 2 * h, cfg["num_node"]
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return ((torch.rand(shape, generator=g) * 2 - 1) * b).to(dtype)

    def bn(key, n):
        sd[key + "weight"] = torch.ones(n, dtype=dtype)
        sd[key + "bias"] = torch.zeros(n, dtype=dtype)
        sd[key + "running_mean"] = torch.zeros(n, dtype=dtype)
        sd[key + "running_var"] = torch.ones(n, dtype=dtype)
        sd[key + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def lin(key, o, i):
        sd[key + "weight"] = uni((o, i), i)
        sd[key + "bias"] = uni((o,), i)

    sd["nonlin_map.conv_block1.0.weight"] = uni((Eh, 1, k), k)
    bn("nonlin_map.conv_block1.1.", Eh)
    sd["nonlin_map.conv_block2.0.weight"] = uni((E, Eh, k), Eh * k)
    bn("nonlin_map.conv_block2.1.", E)
    lin("nonlin_map2.0.", C, E * cfg["encoder_time_out"])
    bn("nonlin_map2.1.", C)
    sd["positional_encoding.pe"] = positional_table(5000, C, dtype).unsqueeze(0)
    for p in ("MPNN1.", "MPNN2."):
        lin(p + "graph_construction.mapping.", C, C)
        bn(p + "BN.", C)
        lin(p + "MPNN.theta.0.", h, C)
        bn(p + "MPNN.bn1.", h)
    lin("fc.fc1.", C, h * cfg["num_windows"] * N)
    lin("fc.fc2.", C, C)
    lin("fc.fc3.", h, C)
    lin("fc.fc4.", 1, h)
    return sd


def is_param(key: str) -> bool:
    return key.endswith(PARAM_SUFFIXES)


class OracleAlgorithm:
    """algorithms/algorithms.py:51-76 (class FC_STGNN): Adam(lr, weight_decay) + MSE;
    update(X, y) = forward -> mse -> zero_grad -> backward -> step -> {'loss': float}."""

    def __init__(self, cfg: dict, hparams: dict, sd: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0,
                 device=None):
        """device: None / cpu = the reference's CPU path; "cuda:0" = the same eager ATen ops on the GPU (the
        reference's default device, main.py:34) -- bench.py's eager_cuda_baseline."""
        self.cfg = dict(cfg)
        self.sd = sd if sd is not None else init_state(cfg, seed)
        if device is not None:
            self.sd = {k: v.detach().to(device) for k, v in self.sd.items()}
        self.params = [v.requires_grad_(True) for k, v in self.sd.items() if is_param(k)]
        self.optimizer = torch.optim.Adam(self.params, lr=hparams["learning_rate"],
                                          weight_decay=hparams["weight_decay"])

    def forward_backward(self, X, y, dropout_keep=None):
        pred = model_forward(X, self.sd, self.cfg, True, dropout_keep)
        loss = F.mse_loss(pred, y)
        self.optimizer.zero_grad()
        loss.backward()
        return loss

    def update(self, X, y, epoch=None, dropout_keep=None):
        loss = self.forward_backward(X, y, dropout_keep)
        self.optimizer.step()
        return {"loss": loss.item()}

    @torch.no_grad()
    def predict(self, X):
        return model_forward(X, self.sd, self.cfg, False)


def rmse(pred, real, max_rul: float) -> float:
    """utils.py:148-151: sqrt(MSE) * max_rul."""
    pred = torch.as_tensor(pred, dtype=torch.float64).reshape(-1)
    real = torch.as_tensor(real, dtype=torch.float64).reshape(-1)
    return float(torch.sqrt(torch.mean((pred - real) ** 2)) * max_rul)


def score_v1(pred, real, max_rul: float) -> Tuple[float, float]:
    """utils.py:136-146 (asymmetric exponential score), vectorised."""
    pred = torch.as_tensor(pred, dtype=torch.float64).reshape(-1) * max_rul
    real = torch.as_tensor(real, dtype=torch.float64).reshape(-1) * max_rul
    d = pred - real
    s = torch.where(d < 0, torch.exp(-d / 13) - 1, torch.exp(d / 10) - 1).sum()
    return float(s), float(s / pred.numel())


def score_v2(pred, real) -> float:
    """utils.py:157-169 (relative-error score, averaged), vectorised."""
    pred = torch.as_tensor(pred, dtype=torch.float64).reshape(-1)
    real = torch.as_tensor(real, dtype=torch.float64).reshape(-1)
    err = (real - pred) / (real + 1e-8) * 100.0
    s = torch.where(err <= 0, torch.exp(-math.log(0.5) * (err / 5.0)), torch.exp(math.log(0.5) * (err / 20.0)))
    return float(s.mean())


def mae(pred, real, max_rul: float) -> float:
    """utils.py:153-155: mean absolute error * max_rul."""
    pred = torch.as_tensor(pred, dtype=torch.float64).reshape(-1)
    real = torch.as_tensor(real, dtype=torch.float64).reshape(-1)
    return float((pred - real).abs().mean() * max_rul)


def calc_metrics(pred, real, max_rul: float):
    """utils.py:191-201 _calc_metrics -> (Scores_v1, Scores_v2, MAE, RMSE)."""
    return score_v1(pred, real, max_rul)[0], score_v2(pred, real), mae(pred, real, max_rul), rmse(pred, real, max_rul)


CONFIGS = {
    # configs/hparams.py:149-151 (CMAPSS FD004) -- the metric config "S1"
    "FD004": dict(patch_size=2, num_patch=25, encoder_time_out=4, encoder_hidden_dim=8, encoder_out_dim=6,
                  encoder_conv_kernel=2, hidden_dim=8, num_sequential=10, num_node=14, num_windows=36),
    # configs/hparams.py:32-34 (FD001)
    "FD001": dict(patch_size=25, num_patch=2, encoder_time_out=27, encoder_hidden_dim=8, encoder_out_dim=32,
                  encoder_conv_kernel=2, hidden_dim=8, num_sequential=6, num_node=14, num_windows=2),
    # configs/hparams.py:69-71 (FD002)
    "FD002": dict(patch_size=1, num_patch=50, encoder_time_out=3, encoder_hidden_dim=8, encoder_out_dim=12,
                  encoder_conv_kernel=2, hidden_dim=8, num_sequential=10, num_node=14, num_windows=74),
    # configs/hparams.py:109-111 (FD003)
    "FD003": dict(patch_size=1, num_patch=50, encoder_time_out=3, encoder_hidden_dim=8, encoder_out_dim=6,
                  encoder_conv_kernel=2, hidden_dim=24, num_sequential=25, num_node=14, num_windows=74),
    # configs/hparams.py:196-198 (N-CMAPSS)
    "NCMAPSS": dict(patch_size=2, num_patch=25, encoder_time_out=4, encoder_hidden_dim=8, encoder_out_dim=32,
                    encoder_conv_kernel=2, hidden_dim=8, num_sequential=6, num_node=20, num_windows=36),
    # SURVEY.md section 8d "S2": BASELINE north_star synthetic [B,T=50,N=21,C=14]
    "S2": dict(patch_size=1, num_patch=50, encoder_time_out=3, encoder_hidden_dim=8, encoder_out_dim=6,
               encoder_conv_kernel=2, hidden_dim=7, num_sequential=10, num_node=21, num_windows=74),
}
TRAIN_HPARAMS = dict(num_epochs=81, batch_size=100, weight_decay=1e-4, learning_rate=1e-3)  # hparams.py:133
