"""GAT_LSTM (BASELINE.json configs[3]) assembled from the native primitives -- drop-in for
models/GAT_LSTM/Model.py (same class names, constructor arguments, parameter names, forward signature; state
dicts interchange).

Native (libstgconv_b200.so): the 11 per-patch statistics (stg_patch_stats11) and the whole dense graph attention
after the layer's projection (stg_gat_*: scores without the [N*N, 2F] concat, softmax, dropout mask, adjacency,
aggregation, leaky_relu, and their backward).  The projections are library GEMMs, the recurrent layers cuDNN
nn.LSTM.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .primitives import extract_features, gat_attention, tall_linear
from .rnn import LSTM


class GraphAttentionLayer(nn.Module):
    """models/GAT_LSTM/Model.py:74-109."""

    def __init__(self, in_features, out_features, dropout, alpha=0.1):
        super().__init__()
        self.in_features, self.out_features, self.dropout, self.alpha = in_features, out_features, dropout, alpha
        self.linear = nn.Linear(in_features, out_features)
        self.attention = nn.Linear(2 * out_features, 1)
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.keep_mask = None          # tests pin the attention-dropout mask here ([bs, N, N] of 0/1)

    def forward(self, h, adj):
        Wh = tall_linear(h, self.linear.weight, self.linear.bias)
        keep = None
        if self.training and self.dropout > 0:
            keep = self.keep_mask
            if keep is None:           # same distribution as F.dropout on the [bs, N, N] attention matrix
                keep = (torch.rand(Wh.shape[0], Wh.shape[1], Wh.shape[1], device=Wh.device) >= self.dropout).float()
        return gat_attention(Wh, self.attention.weight, self.attention.bias, adj, keep, self.dropout, self.alpha, 0.01)


class GAT_LSTM_model(nn.Module):
    """models/GAT_LSTM/Model.py:112-168.  forward(x[bs, (1,) num_patch*patch_size]) -> [bs, 1]."""

    def __init__(self, num_patch, patch_size, hidden_dim, lstm_hidden_dim, dropout=0.1, alpha=0.1):
        super().__init__()
        self.num_patch, self.patch_size = num_patch, patch_size
        hidden_dim = [11] + list(hidden_dim)
        lstm_hidden_dim = [hidden_dim[-1]] + list(lstm_hidden_dim)
        self.gat_layers = nn.ModuleList([GraphAttentionLayer(hidden_dim[i], hidden_dim[i + 1], dropout, alpha)
                                         for i in range(len(hidden_dim) - 1)])
        self.lstm_layers = nn.ModuleList([LSTM(lstm_hidden_dim[i], lstm_hidden_dim[i + 1], num_layers=1, batch_first=True)
                                          for i in range(len(lstm_hidden_dim) - 1)])
        self.fc = nn.Linear(lstm_hidden_dim[-1] * num_patch, 1)
        # Model.py:144-148: identity + first off-diagonals (path graph over the patches).  Constant, so it is
        # built once (one [N,N] for all graphs; not part of the state dict, like the reference's local tensor)
        adj = torch.eye(num_patch)
        idx = torch.arange(num_patch - 1)
        adj[idx, idx + 1] = 1
        adj[idx + 1, idx] = 1
        self.register_buffer("path_adj", adj, persistent=False)

    def forward(self, x):
        bs = x.size(0)
        x = extract_features(x.reshape(bs * self.num_patch, self.patch_size)).reshape(bs, self.num_patch, -1)
        adj = self.path_adj
        for gat in self.gat_layers:
            x = gat(x, adj)
        for lstm in self.lstm_layers:
            x, _ = lstm(x)
        return self.fc(x.reshape(bs, -1))
