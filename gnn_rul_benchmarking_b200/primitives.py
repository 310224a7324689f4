"""Adjacency builders of the sibling models as device ops (primitives A2-A4), same
names and call signatures as the reference functions they replace; forward and backward run in
libstgconv_b200.so (csrc/stg_adj.cu).  No CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ADJ_PCC, ADJ_COSINE, ADJ_GAUSS, ADJ_GAUSS2, ADJ_GRAM = 0, 1, 2, 3, 4


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


def gram_adjacency(x: torch.Tensor) -> torch.Tensor:
    """torch.bmm(x, x^T) of models/STMSGCN/Model.py:96: [G, N, f] -> [G, N, N]."""
    return _Adjacency.apply(x, ADJ_GRAM, 0)


def compute_adjacency_matrix(input: torch.Tensor, top_k: int) -> torch.Tensor:
    """models/STGNN/Model.py:8-25: input [bs, L, N, f] -> exp(-cdist^2) with the top_k entries of each row kept."""
    return _Adjacency.apply(input, ADJ_GAUSS2, top_k)


# ------------------------------------------------------------------------------------------ tall-skinny weight gradients
def tall_gemm_t(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a^T b for a [R, M], b [R, N] with R (= graphs x nodes or batch x time, 1e4 .. 1e5) >> M, N -- the shape of every
    weight gradient of these models.  The GEMM library walks all of R in a handful of CTAs for it (ncu: 513 us for three
    such products in one STGNN update), so R is cut into chunks that run as ONE batched product and the partial results
    are added (split-K)."""
    R = a.shape[0]
    S = min(256, R // 256)
    if S < 2:
        return a.t() @ b
    per = R // S
    main = per * S
    out = torch.bmm(a[:main].view(S, per, a.shape[1]).transpose(1, 2), b[:main].view(S, per, b.shape[1])).sum(0)
    if main < R:
        out = out + a[main:].t() @ b[main:]
    return out


def col_sum(a: torch.Tensor) -> torch.Tensor:
    """Column sums of a tall [R, M] matrix in two stages (chunks first)."""
    R = a.shape[0]
    S = min(1024, R // 64)
    if S < 2:
        return a.sum(0)
    per = R // S
    main = per * S
    out = a[:main].view(S, per, a.shape[1]).sum(1).sum(0)
    if main < R:
        out = out + a[main:].sum(0)
    return out


class _TallLinear(torch.autograd.Function):
    """y = x W^T + b over [R, in] rows (a plain library GEMM); weight / bias gradients through the split-K helpers."""

    @staticmethod
    def forward(ctx, x2, w, b):
        ctx.save_for_backward(x2, w)
        ctx.has_b = b is not None
        return torch.nn.functional.linear(x2, w, b)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ w if ctx.needs_input_grad[0] else None
        dw = tall_gemm_t(dy, x2) if ctx.needs_input_grad[1] else None
        db = col_sum(dy) if (ctx.has_b and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def tall_linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """F.linear(x, weight, bias) for x [..., in] with many leading rows."""
    lead = x.shape[:-1]
    return _TallLinear.apply(x.reshape(-1, x.shape[-1]), weight, bias).view(*lead, weight.shape[0])


class _ChebProject(torch.autograd.Function):
    """einsum("bknf,kfo->bno", T, filters) (models/ASTGCNN/Model.py:226-228); the filter gradient is a contraction over
    bs*N rows and goes through tall_gemm_t."""

    @staticmethod
    def forward(ctx, T, filters):
        ctx.save_for_backward(T, filters)
        return torch.einsum("bknf,kfo->bno", T, filters)

    @staticmethod
    def backward(ctx, dout):
        T, filters = ctx.saved_tensors
        bs, K, N, f = T.shape
        dout = dout.contiguous()
        dT = torch.einsum("bno,kfo->bknf", dout, filters) if ctx.needs_input_grad[0] else None
        dfil = None
        if ctx.needs_input_grad[1]:
            Tp = T.permute(0, 2, 1, 3).reshape(bs * N, K * f)                # rows (b, n), columns (k, f)
            dfil = tall_gemm_t(Tp, dout.view(bs * N, -1)).view(K, f, -1)
        return dT, dfil


# ------------------------------------------------------------------------------------------ M2 / M3
AGG_GCN, AGG_CHEB3, AGG_AX = 0, 1, 2


class _Aggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj, kind):
        if not (x.is_cuda and adj.is_cuda):
            raise RuntimeError("graph aggregation runs on the device (no CPU fallback)")
        G, N, F = x.shape
        xc, ac = x.contiguous(), adj.contiguous()
        out = torch.empty((G, 3, N, F) if kind == AGG_CHEB3 else (G, N, F), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().stg_agg_forward(kind, xc.data_ptr(), ac.data_ptr(), G, N, F, out.data_ptr(), _stream()),
                       "stg_agg_forward")
        ctx.save_for_backward(xc, ac)
        ctx.kind = kind
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, ac = ctx.saved_tensors
        G, N, F = xc.shape
        dout = dout.contiguous()
        dx, dadj = torch.empty_like(xc), torch.empty_like(ac)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.load().stg_agg_backward(ctx.kind, xc.data_ptr(), ac.data_ptr(), dout.data_ptr(), G, N, F,
                                                    dx.data_ptr(), dadj.data_ptr(), _stream()), "stg_agg_backward")
        return dx, dadj, None


def gcn_aggregate(X: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """D^-1/2 (A+I) D^-1/2 X of GCNLayer.forward (models/STMSGCN/Model.py:39-47, SAGCN:86-93, RGCNU:12-19)."""
    return _Aggregate.apply(X, A, AGG_GCN)


def cheb_terms(x: torch.Tensor, adj_matrix: torch.Tensor) -> torch.Tensor:
    """[T0, T1, T2] = [x, A x, 2 A (A x) - x] of ChebNet.forward, K=3 (models/ASTGCNN/Model.py:218-228) -> [bs,3,N,f]."""
    return _Aggregate.apply(x, adj_matrix, AGG_CHEB3)


class GCNLayer(torch.nn.Module):
    """models/STMSGCN/Model.py:34-49: leaky_relu(Linear(D^-1/2 (A+I) D^-1/2 X)); same parameter names."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear = torch.nn.Linear(in_features, out_features)

    def forward(self, X, A):
        return torch.nn.functional.leaky_relu(tall_linear(gcn_aggregate(X, A), self.linear.weight, self.linear.bias))


class ChebNet(torch.nn.Module):
    """models/ASTGCNN/Model.py:198-230 (K = 3, the only value the reference configures); same parameter name."""

    def __init__(self, in_channels, out_channels, K):
        super().__init__()
        if K != 3:
            raise NotImplementedError("the reference hyper-parameters use K = 3 (configs/hparams.py)")
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.filters = torch.nn.Parameter(torch.empty(K, in_channels, out_channels))
        torch.nn.init.xavier_uniform_(self.filters)

    def forward(self, x, adj_matrix):
        T = cheb_terms(x, adj_matrix)                                  # [bs, 3, N, f]
        return _ChebProject.apply(T, self.filters)


def graph_matmul(A: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """torch.bmm(A, X) of MPNN_mk.forward (k = 1; models/ST_GCN/Model.py:85-88): A [bs,N,N], X [bs,N,f]."""
    return _Aggregate.apply(X, A, AGG_AX)


class MPNN_mk(torch.nn.Module):
    """models/ST_GCN/Model.py:74-90 (also ST_Conv, HierCorrPool, LOGO, AGCN_TF): leaky_relu(sum_k theta_k(A^k X)).
    The reference only ever instantiates k = 1."""

    def __init__(self, input_dimension, output_dimension, k):
        super().__init__()
        if k != 1:
            raise NotImplementedError("the reference configurations use k = 1")
        self.k = k
        self.theta = torch.nn.ModuleList([torch.nn.Linear(input_dimension, output_dimension) for _ in range(k)])

    def forward(self, X, A):
        return torch.nn.functional.leaky_relu(tall_linear(graph_matmul(A, X), self.theta[0].weight, self.theta[0].bias))


def segment_and_compute_features(data: torch.Tensor) -> torch.Tensor:
    """models/ST_GCN/Model.py:7-37: data [rows, patch_size] -> [rows, 10] statistics (forward only)."""
    if not data.is_cuda:
        raise RuntimeError("patch statistics run on the device (no CPU fallback)")
    x = data.detach().contiguous().float()
    R, P = x.shape
    out = torch.empty(R, 10, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().stg_patch_stats(x.data_ptr(), R, P, out.data_ptr(), _stream()), "stg_patch_stats")
    return out


def extract_features(data: torch.Tensor) -> torch.Tensor:
    """models/GAT_LSTM/Model.py:6-70: data [rows, patch_size] -> [rows, 11] statistics (forward only)."""
    if not data.is_cuda:
        raise RuntimeError("patch statistics run on the device (no CPU fallback)")
    x = data.detach().contiguous().float()
    R, P = x.shape
    out = torch.empty(R, 11, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().stg_patch_stats11(x.data_ptr(), R, P, out.data_ptr(), _stream()), "stg_patch_stats11")
    return out


def extract_temporal_features(signals: torch.Tensor) -> torch.Tensor:
    """models/SAGCN/Model.py:21-38: signals [rows, patch_size] -> [rows, 12] statistics (forward only)."""
    if not signals.is_cuda:
        raise RuntimeError("patch statistics run on the device (no CPU fallback)")
    x = signals.detach().contiguous().float()
    R, P = x.shape
    out = torch.empty(R, 12, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().stg_patch_stats12(x.data_ptr(), R, P, out.data_ptr(), _stream()), "stg_patch_stats12")
    return out


class _GatAttention(torch.autograd.Function):
    """stg_gat_forward / stg_gat_backward: everything of GraphAttentionLayer.forward after its nn.Linear."""

    @staticmethod
    def forward(ctx, Wh, att_w, att_b, adj, keep, pdrop, alpha, out_slope):
        if not Wh.is_cuda:
            raise RuntimeError("graph attention runs on the device (no CPU fallback)")
        Wh = Wh.contiguous().float()
        G, N, F = Wh.shape
        aw, ab = att_w.contiguous().float().view(-1), att_b.contiguous().float().view(-1)
        adj = adj.detach().contiguous().float()
        per_graph = 1 if adj.dim() == 3 else 0
        if adj.shape[-2:] != (N, N) or (per_graph and adj.shape[0] != G) or aw.numel() != 2 * F:
            raise ValueError("graph attention: inconsistent shapes")
        keep = None if keep is None else keep.detach().contiguous().float()
        out = torch.empty_like(Wh)
        with torch.cuda.device(Wh.device):
            _lib.check(_lib.load().stg_gat_forward(Wh.data_ptr(), aw.data_ptr(), ab.data_ptr(), adj.data_ptr(), per_graph,
                                                   0 if keep is None else keep.data_ptr(), float(pdrop), float(alpha),
                                                   float(out_slope), G, N, F, out.data_ptr(), _stream()),
                       "stg_gat_forward")
        ctx.save_for_backward(Wh, aw, ab, adj, keep if keep is not None else Wh.new_empty(0), out)
        ctx.cfg = (per_graph, keep is not None, float(pdrop), float(alpha), float(out_slope), att_w.shape, att_b.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        Wh, aw, ab, adj, keep, out = ctx.saved_tensors
        per_graph, has_keep, pdrop, alpha, out_slope, wshape, bshape = ctx.cfg
        G, N, F = Wh.shape
        dout = dout.contiguous().float()
        dWh = torch.empty_like(Wh)
        daw, dab = torch.zeros_like(aw), torch.zeros_like(ab)
        with torch.cuda.device(Wh.device):
            _lib.check(_lib.load().stg_gat_backward(Wh.data_ptr(), aw.data_ptr(), ab.data_ptr(), adj.data_ptr(), per_graph,
                                                    keep.data_ptr() if has_keep else 0, pdrop, alpha, out_slope, G, N, F,
                                                    out.data_ptr(), dout.data_ptr(), dWh.data_ptr(), daw.data_ptr(),
                                                    dab.data_ptr(), _stream()), "stg_gat_backward")
        return dWh, daw.view(wshape), dab.view(bshape), None, None, None, None, None


def gat_attention(Wh, att_w, att_b, adj, keep=None, pdrop=0.0, alpha=0.1, out_slope=0.01):
    """leaky_relu((dropout(softmax_j(leaky_relu_alpha(a.[Wh_i || Wh_j] + b))) * adj) Wh): Wh [G,N,F], att_w [1,2F]
    (nn.Linear(2F,1).weight), att_b [1], adj [N,N] or [G,N,N], keep = 0/1 mask [G,N,N] of the attention dropout."""
    return _GatAttention.apply(Wh, att_w, att_b, adj, keep, pdrop, alpha, out_slope)
