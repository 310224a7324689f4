"""ASTGCNN (BASELINE.json configs[2]) assembled from the native sibling primitives -- a drop-in for
models/ASTGCNN/Model.py: same class names, constructor arguments, sub-module / parameter / buffer names
(state dicts interchange with the reference's) and forward signatures.

Native (libstgconv_b200.so): the temporal conv net (stg_tcn_*), the Gaussian adjacency (stg_adj_*) and
the Chebyshev graph convolution's aggregation (stg_agg_*).  Plain GEMMs + pointwise ops (gate Linear+tanh,
the adjacency's Linear P, the Chebyshev filter products, the output Linear) stay library calls.
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm

from . import _lib
from .primitives import ChebNet, gaussian_adjacency

BN_MOMENTUM, BN_EPS = 0.1, 1e-5


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _bn_struct(bn: nn.BatchNorm1d, with_stats: bool = True) -> _lib.StgBN:
    s = _lib.StgBN()
    s.weight, s.bias = bn.weight.data_ptr(), bn.bias.data_ptr()
    if with_stats:
        s.running_mean, s.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        s.num_batches_tracked = bn.num_batches_tracked.data_ptr()
    return s


class _TcnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tcn, w1, g1, b1, w2, g2, b2):
        if not x.is_cuda:
            raise RuntimeError("TemporalConvNet runs on the device (no CPU fallback)")
        x = x.contiguous()
        B, Cc, L = x.shape
        conv1, bn1, conv2, bn2 = tcn.conv_block1[0], tcn.conv_block1[2], tcn.conv_block2[0], tcn.conv_block2[2]
        p = _lib.StgTcnParams()
        p.conv1_w, p.conv2_w = w1.data_ptr(), w2.data_ptr()
        p.bn1, p.bn2 = _bn_struct(bn1), _bn_struct(bn2)
        K = conv1.kernel_size[0]
        out = torch.empty_like(x)
        scratch = torch.empty(8 * Cc, device=x.device, dtype=torch.float64)
        # raw conv outputs kept between the phase kernels and for the backward (recomputed when absent)
        saved = torch.empty(2 * B * Cc * L, device=x.device, dtype=torch.float32) if tcn.training else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().stg_tcn_forward(x.data_ptr(), B, Cc, L, K, C.byref(p), int(tcn.training), BN_MOMENTUM,
                                                   BN_EPS, scratch.data_ptr(), saved.data_ptr() if saved is not None else None,
                                                   out.data_ptr(), _stream()),
                       "stg_tcn_forward")
        ctx.save_for_backward(x, scratch, saved, w1, g1, b1, w2, g2, b2)
        ctx.tcn, ctx.K = tcn, K
        return out

    @staticmethod
    def backward(ctx, dout):
        x, scratch, saved, w1, g1, b1, w2, g2, b2 = ctx.saved_tensors
        tcn = ctx.tcn
        if not tcn.training:
            raise RuntimeError("TemporalConvNet backward is implemented for training mode (batch statistics)")
        B, Cc, L = x.shape
        bn1, bn2 = tcn.conv_block1[2], tcn.conv_block2[2]
        p, g = _lib.StgTcnParams(), _lib.StgTcnParams()
        p.conv1_w, p.conv2_w = w1.data_ptr(), w2.data_ptr()
        p.bn1, p.bn2 = _bn_struct(bn1), _bn_struct(bn2)
        # one zero fill for all six gradient tensors (the kernels accumulate into them)
        srcs = (w1, g1, b1, w2, g2, b2)
        sizes = [(t.numel() + 3) // 4 * 4 for t in srcs]
        flat = torch.zeros(sum(sizes), device=x.device, dtype=torch.float32)
        grads, o = [], 0
        for t, n in zip(srcs, sizes):
            grads.append(flat[o:o + t.numel()].view_as(t))
            o += n
        g.conv1_w, g.conv2_w = grads[0].data_ptr(), grads[3].data_ptr()
        g.bn1.weight, g.bn1.bias = grads[1].data_ptr(), grads[2].data_ptr()
        g.bn2.weight, g.bn2.bias = grads[4].data_ptr(), grads[5].data_ptr()
        dx = torch.empty_like(x)
        dout = dout.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().stg_tcn_backward(x.data_ptr(), dout.data_ptr(), B, Cc, L, ctx.K, C.byref(p), C.byref(g),
                                                    BN_EPS, scratch.data_ptr(),
                                                    saved.data_ptr() if saved is not None else None, dx.data_ptr(),
                                                    _stream()),
                       "stg_tcn_backward")
        return (dx, None, *grads)


class Chomp1d(nn.Module):
    """Marker kept for state-dict / module-index parity (models/ASTGCNN/Model.py:65-71); the causal
    truncation happens inside the kernel."""

    def __init__(self, chomp_size):
        super().__init__()
        self.chomp_size = chomp_size


class TemporalConvNet(nn.Module):
    """models/ASTGCNN/Model.py:72-146.  `net0`, `net1` (and the weight-norm parametrisation of net0) are
    never used by the reference's forward either; they exist so that checkpoints interchange."""

    def __init__(self, input_channels, tcn_layers, kernel_size):
        super().__init__()
        c_in, c0, c1 = input_channels, tcn_layers[1], tcn_layers[1]
        if not (c_in == c0 == c1 == tcn_layers[0]):
            raise NotImplementedError("the reference configurations use equal channel counts (no downsample path)")
        pad0, pad1 = (kernel_size - 1), (kernel_size - 1) * 2
        self.net0 = nn.Sequential(weight_norm(nn.Conv1d(c_in, c0, kernel_size, padding=pad0)), nn.ReLU(),
                                  weight_norm(nn.Conv1d(c0, c0, kernel_size, padding=pad0)), nn.ReLU())
        self.downsample0 = None
        self.relu = nn.ReLU()
        self.net1 = nn.Sequential(nn.Conv1d(c_in, c1, kernel_size, padding=pad1, dilation=2), nn.ReLU(),
                                  nn.Conv1d(c1, c1, kernel_size, padding=pad1, dilation=2), nn.ReLU())
        self.downsample1 = None
        self.conv_block1 = nn.Sequential(nn.Conv1d(c_in, c0, kernel_size, bias=False, padding=pad0), Chomp1d(pad0),
                                         nn.BatchNorm1d(c0), nn.ReLU())
        self.conv_block2 = nn.Sequential(nn.Conv1d(c0, c1, kernel_size, bias=False, padding=pad1, dilation=2),
                                         Chomp1d(pad1), nn.BatchNorm1d(c1), nn.ReLU())

    def forward(self, inputs):
        b1, b2 = self.conv_block1, self.conv_block2
        return _TcnFn.apply(inputs, self, b1[0].weight, b1[2].weight, b1[2].bias, b2[0].weight, b2[2].weight, b2[2].bias)


class GatingMechanism(nn.Module):
    """models/ASTGCNN/Model.py:169-181: tanh(Linear(x) + bias) * tcn_output."""

    def __init__(self, num_channels, out_channels):
        super().__init__()
        self.theta = nn.Linear(num_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, tcn_output):
        return torch.tanh(self.theta(x) + self.bias) * tcn_output


class construct_graph(nn.Module):
    """models/ASTGCNN/Model.py:184-195: exp(-cdist(P X, P X))."""

    def __init__(self, num_features):
        super().__init__()
        self.P = nn.Linear(num_features, num_features, bias=False)

    def forward(self, X):
        return gaussian_adjacency(self.P(X))


class ASTGCNN_model(nn.Module):
    """models/ASTGCNN/Model.py:233-254.  forward(X[bs, N, L]) -> [bs, 1]."""

    def __init__(self, num_nodes, time_length, encoder_out_dim, output_dim, K):
        super().__init__()
        self.tcn = TemporalConvNet(num_nodes, [num_nodes, num_nodes], kernel_size=6)
        self.gate = GatingMechanism(time_length, encoder_out_dim)
        self.distance_module = construct_graph(encoder_out_dim)
        self.chebnet = ChebNet(encoder_out_dim, output_dim, K)
        self.fc = nn.Linear(output_dim, 1)

    def forward(self, X):
        gated = self.gate(X, self.tcn(X))
        out = self.chebnet(gated, self.distance_module(gated))
        return self.fc(out.mean(dim=1))
