"""torch.autograd bindings of the op-level C ABI (device pointers, caller's current stream).

PyTorch is plumbing here: it owns the device memory and the stream; all arithmetic of the
graph-conv block happens in libstgconv_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# order of the per-block tensors passed to graph_blocks()
BLOCK_TENSORS = ("Wm", "bm", "bn0_w", "bn0_b", "bn0_rm", "bn0_rv", "Wt", "bt", "bn1_w", "bn1_b", "bn1_rm", "bn1_rv")
_GRAD_FIELDS = {"Wm": "dWm", "bm": "dbm", "bn0_w": "dbn0_w", "bn0_b": "dbn0_b", "Wt": "dWt", "bt": "dbt",
                "bn1_w": "dbn1_w", "bn1_b": "dbn1_b"}


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_dev(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the sm_100a extension has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def stats_doubles(C: int, H: int, T: int) -> int:
    """STG_BLOCK_STATS_DOUBLES(C,H,T) of include/stgconv_b200.h."""
    p16, p8 = (C + 15) // 16 * 16, (H + 7) // 8 * 8
    coef = 4 * p16 + (p16 + p8) + 4 + p16 * (p16 + p8) + T
    return (4 * H + 2 * C + 1) // 2 * 2 + (coef + 1) // 2


def num_windows(T: int, w: int, s: int) -> int:
    return (T - w) // s + 1


def saved_floats(windows: int, M: int, H: int) -> int:
    """STG_BLOCK_SAVED_FLOATS (include/stgconv_b200.h): Y' rows, then the F | V rows and the softmax rows the
    tcgen05 forward keeps for the backward."""
    rows = windows * M
    return (rows * H + 3) // 4 * 4 + rows * 24 + rows * (M + 1)


class _GraphBlocks(torch.autograd.Function):
    """nblk GraphConvpoolMPNN_block_v6 instances reading the same x -> concatenated features."""

    @staticmethod
    def forward(ctx, x, hyper, training, *tensors):
        lib = _lib.load()
        nblk = len(hyper)
        assert len(tensors) == nblk * len(BLOCK_TENSORS)
        _check_dev(x, "x")
        B, T, N, Cc = x.shape
        per_blk = [dict(zip(BLOCK_TENSORS, tensors[z * len(BLOCK_TENSORS):(z + 1) * len(BLOCK_TENSORS)]))
                   for z in range(nblk)]
        sizes, Ls = [], []
        for z, hp in enumerate(hyper):
            if T < hp["w"]:
                raise ValueError(f"time_length {T} shorter than the window {hp['w']}")
            L = num_windows(T, hp["w"], hp["stride"])
            Ls.append(L)
            sizes.append(L * N * hp["H"])
            for k, t in per_blk[z].items():
                _check_dev(t, k)
        Ftot = sum(sizes)
        feat = torch.empty(B, Ftot, device=x.device, dtype=torch.float32)
        descs = (_lib.StgBlockDesc * nblk)()
        saved_yp, saved_stats = [], []
        xmom = None
        with torch.cuda.device(x.device):
            if training:
                xmom = torch.empty(2 * T * Cc, device=x.device, dtype=torch.float64)
                _lib.check(lib.stg_block_xmoments(x.data_ptr(), B, T, N, Cc, xmom.data_ptr(), _stream_ptr()),
                           "stg_block_xmoments")
            off = 0
            for z, hp in enumerate(hyper):
                d, p = descs[z], per_blk[z]
                d.H, d.w, d.stride, d.decay = hp["H"], hp["w"], hp["stride"], hp["decay"]
                for k in BLOCK_TENSORS:
                    setattr(d, k, p[k].data_ptr())
                d.out = feat.data_ptr() + 4 * off
                d.out_bstride = Ftot
                if training:
                    yp = torch.empty(saved_floats(B * Ls[z], hp["w"] * N, hp["H"]), device=x.device, dtype=torch.float32)
                    st = torch.empty(stats_doubles(Cc, hp["H"], T), device=x.device, dtype=torch.float64)
                    d.yp, d.stats = yp.data_ptr(), st.data_ptr()
                    saved_yp.append(yp)
                    saved_stats.append(st)
                off += sizes[z]
            _lib.check(lib.stg_block_forward(x.data_ptr(), B, T, N, Cc, descs, nblk,
                                             xmom.data_ptr() if training else None, int(training),
                                             BN_MOMENTUM, BN_EPS, _stream_ptr()), "stg_block_forward")
        ctx.hyper, ctx.training, ctx.sizes, ctx.nblk = hyper, training, sizes, nblk
        ctx.save_for_backward(x, xmom if training else x.new_empty(0), *tensors, *saved_yp, *saved_stats)
        ctx.mark_non_differentiable()
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        if not ctx.training:
            raise RuntimeError("stgconv backward is implemented for training mode (batch statistics) only; "
                               "the reference never differentiates in eval mode (trainer.py:134-153)")
        lib = _lib.load()
        nblk, hyper = ctx.nblk, ctx.hyper
        saved = ctx.saved_tensors
        x, xmom = saved[0], saved[1]
        nt = len(BLOCK_TENSORS)
        tensors = saved[2:2 + nblk * nt]
        yps = saved[2 + nblk * nt:2 + nblk * nt + nblk]
        stats = saved[2 + nblk * nt + nblk:]
        B, T, N, Cc = x.shape
        dfeat = dfeat.contiguous()
        Ftot = dfeat.shape[1]
        descs = (_lib.StgBlockDesc * nblk)()
        grads = (_lib.StgBlockGrads * nblk)()
        # per-block dx scratch: [B,T,N,C], or one row per (window, node) [B,L,w*N,C] on the tcgen05 path
        # (STG_BLOCK_DXP_FLOATS in include/stgconv_b200.h)
        dxp_floats = max(B * max(T * N, ((T - hp["w"]) // hp["stride"] + 1) * hp["w"] * N) * Cc for hp in hyper)
        dxp = torch.empty(nblk, dxp_floats, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        out_grads: List[torch.Tensor] = []
        off = 0
        for z, hp in enumerate(hyper):
            d, g = descs[z], grads[z]
            p = dict(zip(BLOCK_TENSORS, tensors[z * nt:(z + 1) * nt]))
            d.H, d.w, d.stride, d.decay = hp["H"], hp["w"], hp["stride"], hp["decay"]
            for k in BLOCK_TENSORS:
                setattr(d, k, p[k].data_ptr())
            d.yp, d.stats = yps[z].data_ptr(), stats[z].data_ptr()
            d.out, d.out_bstride = None, Ftot
            g.dout = dfeat.data_ptr() + 4 * off
            g.dout_bstride = Ftot
            g.dxp = dxp[z].data_ptr()
            for k in BLOCK_TENSORS:
                if k in _GRAD_FIELDS:
                    gt = torch.zeros_like(p[k])
                    setattr(g, _GRAD_FIELDS[k], gt.data_ptr())
                    out_grads.append(gt)
                else:
                    out_grads.append(None)
            off += ctx.sizes[z]
        with torch.cuda.device(x.device):
            _lib.check(lib.stg_block_backward(x.data_ptr(), B, T, N, Cc, descs, grads, nblk, xmom.data_ptr(),
                                              BN_EPS, dx.data_ptr(), _stream_ptr()), "stg_block_backward")
        return (dx, None, None, *out_grads)


def graph_blocks(x: torch.Tensor, hyper: Sequence[dict], tensors: Sequence[Sequence[torch.Tensor]],
                 training: bool) -> torch.Tensor:
    """x [B,T,N,C] -> features [B, sum_k L_k*N*H_k] of len(hyper) graph-conv blocks (flattened
    and concatenated exactly like FC_STGNN_RUL.forward, Model.py:74-81).

    hyper[k] = dict(H=, w=, stride=, decay=); tensors[k] = the 12 tensors in BLOCK_TENSORS order
    (running statistics are updated in place when training)."""
    flat = [t for blk in tensors for t in blk]
    x = x.contiguous()
    if x.data_ptr() % 16:                       # a contiguous view at an odd offset: the kernels want 16-B alignment
        x = x.clone()
    return _GraphBlocks.apply(x, tuple(dict(h) for h in hyper), bool(training), *flat)
