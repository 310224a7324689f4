"""nn.LSTM / nn.GRU drop-ins whose time recurrence runs in libstgconv_b200.so (csrc/stg_rnn.cu, primitive T2 of
SURVEY.md 2.2) instead of cuDNN: same constructors, parameter names (`weight_ih_l0`, `weight_hh_l0_reverse`, ...) and
return values `(output, (h_n, c_n))` / `(output, h_n)`, so state dicts interchange with the reference's layers
(models/HAGCN/Model.py:33-53, models/GAT_LSTM/Model.py:129-132, models/STGNN/Model.py:72, models/STMSGCN/Model.py:55).

Split of the work: the input projection x.W_ih^T + b and the weight gradients are plain GEMMs (library calls, autograd
through torch); the recurrence h_t = cell(xg_t + W_hh h_{t-1}) -- the part that cuDNN runs as one small launch chain per
time step -- is ONE persistent kernel per direction pair and per pass.  One layer, zero initial state (every reference
call site).  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .primitives import _TallLinear, col_sum as _col_sum, tall_gemm_t as _tall_gemm_t

CELL_LSTM, CELL_GRU = 0, 1


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Recurrence(torch.autograd.Function):
    """xg [B, T, ndir*G*H] (batch-major), whh [ndir, G*H, H], bhn [ndir, H] (GRU) -> out [B, T, ndir*H]."""

    @staticmethod
    def forward(ctx, xg, whh, bhn, cell):
        if not xg.is_cuda:
            raise RuntimeError("the recurrent kernels run on the device (no CPU fallback)")
        lib = _lib.load()
        B, T, _ = xg.shape
        ndir, GH, H = whh.shape
        xg, whh = xg.contiguous(), whh.contiguous()
        out = torch.empty(B, T, ndir * H, device=xg.device, dtype=torch.float32)
        need = any(ctx.needs_input_grad[:3])          # forward() itself always runs with grad mode off
        saved = None
        if need:
            saved = torch.empty(lib.stg_rnn_saved_floats(cell, T, B, H, ndir), device=xg.device, dtype=torch.float32)
        with torch.cuda.device(xg.device):
            _lib.check(lib.stg_rnn_forward(cell, xg.data_ptr(), T * ndir * GH, ndir * GH, whh.data_ptr(),
                                           bhn.data_ptr() if cell == CELL_GRU else None, T, B, H, ndir,
                                           out.data_ptr(), T * ndir * H, ndir * H,
                                           saved.data_ptr() if need else None, _stream()), "stg_rnn_forward")
        ctx.cell, ctx.dims = cell, (B, T, H, ndir, GH)
        ctx.save_for_backward(whh, saved, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        whh, saved, out = ctx.saved_tensors
        B, T, H, ndir, GH = ctx.dims
        cell = ctx.cell
        lib = _lib.load()
        dout = dout.contiguous()
        dxg = torch.empty(B, T, ndir * GH, device=dout.device, dtype=torch.float32)
        dhn = torch.empty(B, T, ndir * H, device=dout.device, dtype=torch.float32) if cell == CELL_GRU else None
        with torch.cuda.device(dout.device):
            _lib.check(lib.stg_rnn_backward(cell, whh.data_ptr(), saved.data_ptr(), dout.data_ptr(), T * ndir * H,
                                            ndir * H, T, B, H, ndir, dxg.data_ptr(), T * ndir * GH, ndir * GH,
                                            dhn.data_ptr() if dhn is not None else None, _stream()),
                       "stg_rnn_backward")
        # dW_hh[d] = sum_{b,t} dgates_h[b,t,d,:] (x) h_prev[b,t,d,:]  -- a (tall, split-K) GEMM against the shifted outputs;
        # GRU: the n rows use dhn (recurrent side of the n gate), not dn
        dg2 = dxg.view(B * T, ndir * GH)
        o = out.view(B, T, ndir, H)
        dwhh = torch.empty_like(whh)
        dbhn = None
        if cell == CELL_GRU:
            dhn2 = dhn.view(B * T, ndir * H)
            dbhn = _col_sum(dhn2).view(ndir, H)
        for d in range(ndir):
            hp = torch.zeros(B, T, H, device=out.device, dtype=torch.float32)
            if T > 1:
                if d == 0:
                    hp[:, 1:] = o[:, :-1, 0]
                else:
                    hp[:, :-1] = o[:, 1:, 1]
            hp2 = hp.view(B * T, H)
            if cell == CELL_GRU:
                dwhh[d, :2 * H] = _tall_gemm_t(dg2[:, d * GH:d * GH + 2 * H], hp2)
                dwhh[d, 2 * H:] = _tall_gemm_t(dhn2[:, d * H:(d + 1) * H], hp2)
            else:
                dwhh[d] = _tall_gemm_t(dg2[:, d * GH:(d + 1) * GH], hp2)
        return dxg, dwhh, dbhn, None


def _layer_weights(mod: nn.RNNBase, G: int):
    """Stacked per-direction tensors in the layout the kernel reads (autograd flows back to the nn parameters)."""
    H = mod.hidden_size
    sfx = [""] + (["_reverse"] if mod.bidirectional else [])
    wih = torch.cat([getattr(mod, "weight_ih_l0" + s) for s in sfx], 0)              # [ndir*G*H, I]
    whh = torch.stack([getattr(mod, "weight_hh_l0" + s) for s in sfx], 0)            # [ndir, G*H, H]
    bias, bhn = None, None
    if mod.bias:
        parts = []
        for s in sfx:
            bi, bh = getattr(mod, "bias_ih_l0" + s), getattr(mod, "bias_hh_l0" + s)
            if G == 3:                                                               # b_hn stays inside r * (...)
                parts.append(bi + torch.cat([bh[:2 * H], torch.zeros_like(bh[2 * H:])]))
            else:
                parts.append(bi + bh)
        bias = torch.cat(parts, 0)
        if G == 3:
            bhn = torch.stack([getattr(mod, "bias_hh_l0" + s)[2 * H:] for s in sfx], 0)
    elif G == 3:
        bhn = torch.zeros(len(sfx), H, device=whh.device, dtype=whh.dtype)
    return wih, whh, bias, bhn


def _run(mod: nn.RNNBase, x: torch.Tensor, cell: int):
    if mod.num_layers != 1 or getattr(mod, "proj_size", 0):
        raise NotImplementedError("native recurrence: one layer without projection (every reference call site)")
    if x.dim() != 3:
        raise ValueError("expected a [batch, seq, feature] / [seq, batch, feature] tensor")
    if not mod.batch_first:
        x = x.transpose(0, 1)
    G = 4 if cell == CELL_LSTM else 3
    wih, whh, bias, bhn = _layer_weights(mod, G)
    Bx, Tx, Ix = x.shape
    xg = _TallLinear.apply(x.reshape(Bx * Tx, Ix), wih, bias).view(Bx, Tx, -1)          # plain GEMM
    out = _Recurrence.apply(xg, whh, bhn, cell)
    H = mod.hidden_size
    h_n = out[:, -1, :H].unsqueeze(0)
    if mod.bidirectional:
        h_n = torch.cat([h_n, out[:, 0, H:].unsqueeze(0)], 0)
    if not mod.batch_first:
        out = out.transpose(0, 1)
    return out, h_n


class LSTM(nn.LSTM):
    """nn.LSTM with the recurrence in k_rnn_fwd / k_rnn_bwd.  Returns (output, (h_n, None)): no reference call site
    reads c_n."""

    def forward(self, x, hx=None):
        if hx is not None:
            raise NotImplementedError("initial states are always zero in the reference")
        out, h_n = _run(self, x, CELL_LSTM)
        return out, (h_n, None)


class GRU(nn.GRU):
    def forward(self, x, hx=None):
        if hx is not None:
            raise NotImplementedError("initial states are always zero in the reference")
        return _run(self, x, CELL_GRU)
