"""Adjacency builders of the sibling models as device ops (primitives A2-A4), same
names and call signatures as the reference functions they replace; forward and backward run in
libstgconv_b200.so (csrc/stg_adj.cu).  No CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ADJ_PCC, ADJ_COSINE, ADJ_GAUSS, ADJ_GAUSS2 = 0, 1, 2, 3


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Adjacency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind, top_k):
        if not x.is_cuda:
            raise RuntimeError("adjacency builders run on the device (no CPU fallback)")
        if x.dtype != torch.float32:
            raise TypeError("x must be float32")
        lead, (N, F) = x.shape[:-2], x.shape[-2:]
        xc = x.reshape(-1, N, F).contiguous()
        G = xc.shape[0]
        adj = torch.empty(G, N, N, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().stg_adj_forward(kind, xc.data_ptr(), G, N, F, int(top_k), adj.data_ptr(), None,
                                                   _stream()), "stg_adj_forward")
        ctx.save_for_backward(xc, adj)
        ctx.kind, ctx.shape = kind, x.shape
        return adj.view(*lead, N, N)

    @staticmethod
    def backward(ctx, dadj):
        xc, adj = ctx.saved_tensors
        G, N, F = xc.shape
        dadj = dadj.reshape(G, N, N).contiguous()
        dx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.load().stg_adj_backward(ctx.kind, xc.data_ptr(), adj.data_ptr(), dadj.data_ptr(), G, N, F,
                                                    dx.data_ptr(), _stream()), "stg_adj_backward")
        return dx.view(ctx.shape), None, None


def pcc_graph_construction(data: torch.Tensor) -> torch.Tensor:
    """models/ST_GCN/Model.py:53-71 (also ST_Conv:10-28, LOGO:17-35): data [bs, N, f] -> [bs, N, N]."""
    return _Adjacency.apply(data, ADJ_PCC, 0)


def cosine_distance(matrix1: torch.Tensor) -> torch.Tensor:
    """models/HAGCN/Model.py:122-127, models/SAGCN/Model.py:74-79: [..., N, f] -> [..., N, N]."""
    return _Adjacency.apply(matrix1, ADJ_COSINE, 0)


def gaussian_adjacency(PX: torch.Tensor) -> torch.Tensor:
    """exp(-cdist(PX, PX, p=2)) of models/ASTGCNN/Model.py:193-194 (apply the layer's Linear P first)."""
    return _Adjacency.apply(PX, ADJ_GAUSS, 0)


def compute_adjacency_matrix(input: torch.Tensor, top_k: int) -> torch.Tensor:
    """models/STGNN/Model.py:8-25: input [bs, L, N, f] -> exp(-cdist^2) with the top_k entries of each row kept."""
    return _Adjacency.apply(input, ADJ_GAUSS2, top_k)
