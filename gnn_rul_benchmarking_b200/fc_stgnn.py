"""Host-side mirror of the reference FC_STGNN model interface (drop-in for
models/FC_STGNN/Model.py + Model_Base.py): same class names, constructor arguments,
sub-module / parameter / buffer names (so `checkpoint.pt` state dicts are interchangeable,
utils.py:111-120) and forward() signatures, so trainer.py / algorithms/algorithms.py drive it
unchanged.  The graph-conv block runs in the sm_100a extension (csrc/stg_block.cu); there is no
CPU implementation -- tensors must live on a CUDA device.

Deliberate differences from the reference (behaviour-preserving):
  * no hard-coded `.cuda()` (Model_Base.py:58,119,151): buffers follow `module.to(device)`.
  * `pre_relation` stays a plain attribute that is NOT in the state_dict (Model_Base.py:187);
    the kernel evaluates decay^|dt| in place of the materialised mask.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import engine as ENG
from . import functional as SF

DECAY = 0.7            # Model.py:12
MOVING_WINDOW = (2, 2)  # Model.py:13
STRIDE = (1, 2)        # Model.py:14


class Feature_extractor_1DCNN_RUL(nn.Module):
    """Model_Base.py:12-41: Conv1d(k,pad k//2,no bias)+BN+ReLU+Dropout -> Conv1d(k,pad 1)+BN+ReLU."""

    def __init__(self, input_channels, num_hidden, out_dim, kernel_size=8, stride=1, dropout=0):
        super().__init__()
        self.conv_block1 = nn.Sequential(
            nn.Conv1d(input_channels, num_hidden, kernel_size=kernel_size, stride=stride, bias=False,
                      padding=kernel_size // 2),
            nn.BatchNorm1d(num_hidden), nn.ReLU(), nn.Dropout(dropout))
        self.conv_block2 = nn.Sequential(
            nn.Conv1d(num_hidden, out_dim, kernel_size=kernel_size, stride=1, bias=False, padding=1),
            nn.BatchNorm1d(out_dim), nn.ReLU())

    def forward(self, x_in):
        return self.conv_block2(self.conv_block1(torch.transpose(x_in, -1, -2)))


class Dot_Graph_Construction_weights(nn.Module):
    """Parameter container for the learned dot-product adjacency (Model_Base.py:44-67).  Its
    arithmetic is fused into the block kernel; calling it on its own is not part of the path."""

    def __init__(self, input_dim):
        super().__init__()
        self.mapping = nn.Linear(input_dim, input_dim)


class MPNN_mk_v2(nn.Module):
    """Parameter container of MPNN_mk_v2 (Model_Base.py:72-107); only k=1 exists in the reference."""

    def __init__(self, input_dimension, outpuut_dinmension, k):
        super().__init__()
        if k != 1:
            raise NotImplementedError("the reference only instantiates MPNN_mk_v2 with k=1 (Model_Base.py:185)")
        self.k = k
        self.theta = nn.ModuleList([nn.Linear(input_dimension, outpuut_dinmension) for _ in range(k)])
        self.bn1 = nn.BatchNorm1d(outpuut_dinmension)


class PositionalEncoding(nn.Module):
    """Model_Base.py:111-134 -- note ln(100), not ln(10000)."""

    def __init__(self, d_model, dropout, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(100.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)[:, : d_model // 2]
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.dropout(x + self.pe[:, : x.size(1)])


def Mask_Matrix(num_node, time_length, decay_rate):
    """Closed form of Model_Base.py:150-170: mask[i,k] = decay^|i//N - k//N| (CPU tensor)."""
    t = torch.arange(num_node * time_length) // num_node
    return torch.tensor(float(decay_rate)).pow((t[:, None] - t[None, :]).abs().float())


def _block_tensors(blk: "GraphConvpoolMPNN_block_v6"):
    m, bn0, th, bn1 = blk.graph_construction.mapping, blk.BN, blk.MPNN.theta[0], blk.MPNN.bn1
    return (m.weight, m.bias, bn0.weight, bn0.bias, bn0.running_mean, bn0.running_var,
            th.weight, th.bias, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var)


def _block_hyper(blk: "GraphConvpoolMPNN_block_v6"):
    return dict(H=blk.output_dim, w=blk.time_window_size, stride=blk.stride, decay=blk.decay)


def _tick(blk: "GraphConvpoolMPNN_block_v6"):
    blk.BN.num_batches_tracked += 1
    blk.MPNN.bn1.num_batches_tracked += 1


class GraphConvpoolMPNN_block_v6(nn.Module):
    """Model_Base.py:175-225.  forward(input[bs,T,N,C]) -> [bs, L, N, output_dim]."""

    def __init__(self, input_dim, output_dim, num_sensors, time_length, time_window_size, stride, decay, pool_choice):
        super().__init__()
        if pool_choice != "mean":
            raise NotImplementedError("FC_STGNN_RUL hard-wires pool_choice='mean' (Model.py:11)")
        self.num_sensors = num_sensors
        self.time_window_size = time_window_size
        self.stride = stride
        self.output_dim = output_dim
        self.decay = decay
        self.graph_construction = Dot_Graph_Construction_weights(input_dim)
        self.BN = nn.BatchNorm1d(input_dim)
        self.MPNN = MPNN_mk_v2(input_dim, output_dim, k=1)
        self.pre_relation = Mask_Matrix(num_sensors, time_window_size, decay)   # attribute, not a buffer
        self.pool_choice = pool_choice

    def forward(self, input):
        bs, T, N, _ = input.shape
        feat = SF.graph_blocks(input, [_block_hyper(self)], [_block_tensors(self)], self.training)
        if self.training:
            _tick(self)
        return feat.view(bs, -1, N, self.output_dim)


class FC_STGNN_RUL(nn.Module):
    """Model.py:5-85.  forward(X[bs, num_node, L]) -> [bs, 1]."""

    def __init__(self, patch_size, num_patch, encoder_time_out, encoder_hidden_dim, encoder_out_dim,
                 encoder_conv_kernel, hidden_dim, num_sequential, num_node, num_windows):
        super().__init__()
        self.patch_size = patch_size
        self.num_patch = num_patch
        self.nonlin_map = Feature_extractor_1DCNN_RUL(1, encoder_hidden_dim, encoder_out_dim,
                                                      kernel_size=encoder_conv_kernel)
        self.nonlin_map2 = nn.Sequential(nn.Linear(encoder_out_dim * encoder_time_out, 2 * hidden_dim),
                                         nn.BatchNorm1d(2 * hidden_dim))
        self.positional_encoding = PositionalEncoding(2 * hidden_dim, 0.1, max_len=5000)
        self.MPNN1 = GraphConvpoolMPNN_block_v6(2 * hidden_dim, hidden_dim, num_node, num_sequential,
                                                time_window_size=MOVING_WINDOW[0], stride=STRIDE[0], decay=DECAY,
                                                pool_choice="mean")
        self.MPNN2 = GraphConvpoolMPNN_block_v6(2 * hidden_dim, hidden_dim, num_node, num_sequential,
                                                time_window_size=MOVING_WINDOW[1], stride=STRIDE[1], decay=DECAY,
                                                pool_choice="mean")
        self.fc = nn.Sequential(OrderedDict([
            ("fc1", nn.Linear(hidden_dim * num_windows * num_node, 2 * hidden_dim)),
            ("relu1", nn.ReLU(inplace=True)),
            ("fc2", nn.Linear(2 * hidden_dim, 2 * hidden_dim)),
            ("relu2", nn.ReLU(inplace=True)),
            ("fc3", nn.Linear(2 * hidden_dim, hidden_dim)),
            ("relu3", nn.ReLU(inplace=True)),
            ("fc4", nn.Linear(hidden_dim, 1)),
        ]))

    def encode(self, X):
        """Model.py:45-68: patch encoder + positional encoding -> [bs, T, N, 2h]."""
        bs, num_node, _ = X.size()
        X = torch.reshape(X, [bs, num_node, self.num_patch, self.patch_size]).transpose(1, 2)
        bs, tlen, num_node, dimension = X.size()
        A = self.nonlin_map(torch.reshape(X, [bs * tlen * num_node, dimension, 1]))
        A = self.nonlin_map2(torch.reshape(A, [bs * tlen * num_node, -1]))
        A = torch.reshape(A, [bs, tlen, num_node, -1]).transpose(1, 2)
        A = self.positional_encoding(torch.reshape(A, [bs * num_node, tlen, -1]))
        return torch.reshape(A, [bs, num_node, tlen, -1]).transpose(1, 2).contiguous()

    @property
    def engine(self) -> "ENG.ModelEngine":
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = ENG.ModelEngine(self)
            self.__dict__["_engine"] = eng
        return eng

    def forward(self, X):
        """Whole model in the sm_100a engine (csrc/stg_encoder.cu, stg_block.cu, stg_head.cu): one
        autograd node; BatchNorm running statistics and num_batches_tracked are updated on the
        device in training mode."""
        return ENG.model_forward(self.engine, X)

    def forward_torch_encoder(self, X):
        """Earlier composition kept for cross-checks: torch encoder/head around the native blocks."""
        h = self.encode(X)
        # MPNN1 and MPNN2 read the same tensor (Model.py:74-75): one fused launch sequence that
        # writes straight into the concatenated feature layout of Model.py:78-81.
        feat = SF.graph_blocks(h, [_block_hyper(self.MPNN1), _block_hyper(self.MPNN2)],
                               [_block_tensors(self.MPNN1), _block_tensors(self.MPNN2)], self.training)
        if self.training:
            _tick(self.MPNN1)
            _tick(self.MPNN2)
        return self.fc(feat)
