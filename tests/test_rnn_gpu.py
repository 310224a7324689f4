"""Native recurrent layer (csrc/stg_rnn.cu behind rnn.LSTM / rnn.GRU) vs torch's own nn.LSTM / nn.GRU on the CPU in
fp32 -- the arithmetic the reference models run (models/HAGCN/Model.py:33-53, GAT_LSTM/Model.py:129-132,
STGNN/Model.py:72, STMSGCN/Model.py:55).  Shapes are the reference call sites' (batch, sequence, input, hidden),
including HAGCN's long-sequence / tiny-batch layout and the 120-wide layer that runs on a 4-CTA cluster."""
import pytest
import torch
import torch.nn as nn

TOL_OUT = 2e-5        # fp32 recurrences, different summation order
TOL_GRAD = 2e-4       # relative to the tensor's largest entry


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-7)


# (cell, B, T, I, H, bidirectional, batch_first)
CASES = [
    ("lstm", 5, 700, 10, 60, True, True),      # HAGCN bi_lstm1 (sequence = bs*N)
    ("lstm", 5, 300, 60, 120, True, True),     # HAGCN bi_lstm2: cluster of 4 CTAs
    ("lstm", 2, 257, 25, 60, True, True),      # HAGCN FD002 hparams (num_patch 2)
    ("lstm", 1, 64, 50, 60, True, True),       # HAGCN FD004 hparams (num_patch 1)
    ("lstm", 128, 40, 100, 30, False, True),   # GAT_LSTM layer 1
    ("lstm", 37, 40, 30, 20, False, True),     # GAT_LSTM layer 2, ragged batch tile
    ("gru", 300, 5, 64, 64, False, True),      # STGNN (N-CMAPSS: 5 patches)
    ("gru", 70, 1, 64, 64, False, True),       # STGNN (C-MAPSS: one patch)
    ("gru", 64, 160, 97, 8, False, True),      # STMSGCN
    ("gru", 3, 50, 7, 12, True, False),        # generic: bidirectional GRU, time-major input
    ("gru", 4, 33, 20, 100, True, True),       # generic: GRU on the cluster path
]


@pytest.mark.gpu
@pytest.mark.parametrize("cell,B,T,I,H,bi,bf", CASES)
def test_recurrence_matches_torch_cpu(cell, B, T, I, H, bi, bf):
    from gnn_rul_benchmarking_b200 import rnn
    torch.manual_seed(B * 1000 + T + H)
    ref = (nn.LSTM if cell == "lstm" else nn.GRU)(I, H, num_layers=1, batch_first=bf, bidirectional=bi)
    mine = (rnn.LSTM if cell == "lstm" else rnn.GRU)(I, H, num_layers=1, batch_first=bf, bidirectional=bi)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda()
    x = torch.randn((B, T, I) if bf else (T, B, I))
    w = torch.randn((B, T, (2 if bi else 1) * H) if bf else (T, B, (2 if bi else 1) * H))
    xr = x.clone().requires_grad_(True)
    out_r, hn_r = ref(xr)
    (out_r * w).sum().backward()
    xm = x.cuda().requires_grad_(True)
    out_m, hn_m = mine(xm)
    (out_m * w.cuda()).sum().backward()
    assert _rel(out_m.detach().cpu(), out_r.detach()) < TOL_OUT
    hr = hn_r[0] if cell == "lstm" else hn_r
    hm = hn_m[0] if cell == "lstm" else hn_m
    assert _rel(hm.detach().cpu(), hr.detach()) < TOL_OUT
    assert _rel(xm.grad.cpu(), xr.grad) < TOL_GRAD, "dx"
    for (n, pr), (_, pm) in zip(ref.named_parameters(), mine.named_parameters()):
        assert _rel(pm.grad.cpu(), pr.grad) < TOL_GRAD, n

    # inference path (no saved activations)
    with torch.no_grad():
        assert _rel(mine(x.cuda())[0].cpu(), out_r.detach()) < TOL_OUT


def test_drop_in_state_dict_and_guards():
    from gnn_rul_benchmarking_b200 import rnn
    a, b = nn.LSTM(10, 60, batch_first=True, bidirectional=True), rnn.LSTM(10, 60, batch_first=True, bidirectional=True)
    assert list(a.state_dict()) == list(b.state_dict())
    g = rnn.GRU(4, 8, batch_first=True)
    assert list(g.state_dict()) == list(nn.GRU(4, 8, batch_first=True).state_dict())
    with pytest.raises(RuntimeError):
        g(torch.zeros(2, 3, 4))                      # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        rnn.GRU(4, 8, num_layers=2)(torch.zeros(2, 3, 4))


def test_rnn_argument_validation():
    from gnn_rul_benchmarking_b200 import _lib
    lib = _lib.load()
    assert lib.stg_rnn_batch_tile(1) == 8 and lib.stg_rnn_batch_tile(100) == 8
    assert lib.stg_rnn_saved_floats(0, 10, 5, 60, 2) == 2 * 1 * 10 * 6 * 60 * 8
    assert lib.stg_rnn_saved_floats(1, 10, 9, 8, 1) == 1 * 2 * 10 * 6 * 8 * 8
    assert lib.stg_rnn_forward(7, None, 0, 0, None, None, 1, 1, 1, 1, None, 0, 0, None, None) == -1
    assert b"unknown cell" in lib.stg_last_error()
    assert lib.stg_rnn_forward(0, None, 0, 0, None, None, 1, 1, 200, 1, None, 0, 0, None, None) != 0
    assert b"hidden size" in lib.stg_last_error()
