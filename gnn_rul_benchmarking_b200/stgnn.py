"""STGNN and STMSGCN (north_star model list) assembled from the native sibling primitives -- drop-ins for
models/STGNN/Model.py and models/STMSGCN/Model.py (same class names, constructor arguments, parameter names,
forward signatures; state dicts interchange).

Native (libstgconv_b200.so): Gaussian top-k / outer-product adjacency (stg_adj_*), Chebyshev and sym-norm GCN
aggregation (stg_agg_*).  The recurrent layer is nn.GRU (cuDNN), the spectral features use torch.fft (cuFFT),
the projections are library GEMMs.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .primitives import ChebNet, GCNLayer, compute_adjacency_matrix, gram_adjacency
from .rnn import GRU


class STGNN_model(nn.Module):
    """models/STGNN/Model.py:64-107.  forward(x[bs, num_nodes, num_patch*patch_size]) -> [bs, 1]."""

    def __init__(self, patch_size, num_patch, num_nodes, hidden_dim, K, top_k):
        super().__init__()
        self.num_patch, self.patch_size, self.top_k = num_patch, patch_size, top_k
        self.chebnet = ChebNet(patch_size, hidden_dim, K)
        self.gru = GRU(hidden_dim, hidden_dim, batch_first=True)
        self.fc = nn.Linear(hidden_dim * num_patch * num_nodes, 1)

    def forward(self, x):
        bs, N, _ = x.shape
        L, f = self.num_patch, self.patch_size
        g = x.reshape(bs, N, L, f).transpose(1, 2).contiguous()            # [bs, L, N, f]: one graph per patch
        adj = compute_adjacency_matrix(g, self.top_k)                      # [bs, L, N, N]
        h = self.chebnet(g.view(bs * L, N, f), adj.view(bs * L, N, N))     # [bs*L, N, hidden]
        seq = h.view(bs, L, N, -1).permute(0, 2, 1, 3).reshape(bs * N, L, -1)
        out, _ = self.gru(seq)
        return self.fc(out.reshape(bs, -1))


def SED_features(input_data, interval, band_width):
    """models/STMSGCN/Model.py:7-31: spectral energy difference per frequency band (torch.fft = cuFFT)."""
    bs = input_data.size(0)
    spec = torch.fft.fft(input_data, dim=-1)
    sd = spec[:, interval:] - spec[:, :-interval]
    return (sd.real ** 2 + sd.imag ** 2).view(bs, -1, band_width).sum(dim=-1)


class GRULayer(nn.Module):
    """models/STMSGCN/Model.py:52-60."""

    def __init__(self, input_dim, hidden_dim, num_layers):
        super().__init__()
        self.gru = GRU(input_dim, hidden_dim, num_layers, batch_first=True)

    def forward(self, x):
        return self.gru(x)[0]


class STMSGCN_model(nn.Module):
    """models/STMSGCN/Model.py:63-111.  forward(x[bs, (1,) num_patch*patch_size]) -> [bs, 1]."""

    def __init__(self, num_patch, patch_size, interval, band_width, gcn_dims, gru_hidden_dim):
        super().__init__()
        self.num_patch, self.patch_size, self.interval, self.band_width = num_patch, patch_size, interval, band_width
        dims = [1] + list(gcn_dims)
        self.gcn_dims = dims
        self.gcn_layers = nn.ModuleList([GCNLayer(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])
        self.gru_layer = GRULayer(sum(dims), gru_hidden_dim, 1)
        self.fc = nn.Linear(gru_hidden_dim * num_patch, 1)

    def forward(self, x):
        bs = x.size(0)
        sed = SED_features(x.reshape(bs * self.num_patch, self.patch_size), self.interval, self.band_width)
        h = sed.reshape(bs * self.num_patch, -1, 1).contiguous()
        N = h.size(1)
        feats = [h]
        for gcn in self.gcn_layers:                    # adjacency re-built from the current node features
            h = gcn(h, gram_adjacency(h))
            feats.append(h)
        z = torch.cat(feats, dim=-1).reshape(bs, self.num_patch, N, -1).transpose(1, 2).reshape(bs * N, self.num_patch, -1)
        out = self.gru_layer(z).reshape(bs, N, self.num_patch, -1).mean(1)
        return self.fc(out.reshape(bs, -1))
