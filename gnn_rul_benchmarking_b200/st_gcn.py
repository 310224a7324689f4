"""ST_GCN (BASELINE.json configs[2]) assembled from the native sibling primitives -- a drop-in for
models/ST_GCN/Model.py (same class names, constructor arguments, parameter names, forward signature).

Native (libstgconv_b200.so): per-patch statistics (stg_patch_stats), Pearson adjacency (stg_adj_*), the
message-passing aggregation A.X (stg_agg_*), the temporal conv net (stg_tcn_*).  Linear layers, dropout and
the global max pool are library / pointwise calls.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .astgcnn import TemporalConvNet
from .primitives import MPNN_mk, pcc_graph_construction, segment_and_compute_features


class SG_TCN(nn.Module):
    """models/ST_GCN/Model.py:176-195: num_layers x [MPNN_mk -> TemporalConvNet -> Dropout] with residuals."""

    def __init__(self, in_features, num_patch, num_layers=5, dropout=0.2, k=1):
        super().__init__()
        self.layers = nn.ModuleList(
            nn.ModuleList([MPNN_mk(num_patch, num_patch, k),
                           TemporalConvNet(in_features, [in_features, in_features], kernel_size=2),
                           nn.Dropout(dropout)]) for _ in range(num_layers))

    def forward(self, x, adj):
        out = x
        for mpnn, tcn, dropout in self.layers:
            out = dropout(tcn(mpnn(out, adj))) + out
        return out


class ST_GCN_model(nn.Module):
    """models/ST_GCN/Model.py:197-222.  forward(x[bs, (1,) num_patch*patch_size]) -> [bs, 1]."""

    def __init__(self, num_patch, patch_size, num_layers=2, dropout=0.5, k=1):
        super().__init__()
        self.num_patch, self.patch_size = num_patch, patch_size
        self.sg_tcn = SG_TCN(10, num_patch, num_layers, dropout, k)
        self.global_max_pool = nn.AdaptiveMaxPool1d(1)
        self.fc1 = nn.Linear(num_patch, num_patch)
        self.fc2 = nn.Linear(num_patch, 1)

    def forward(self, x):
        bs = x.size(0)
        feats = segment_and_compute_features(x.reshape(bs * self.num_patch, self.patch_size))
        nodes = feats.reshape(bs, self.num_patch, 10).transpose(-1, -2).contiguous()   # nodes = the 10 statistics
        out = self.sg_tcn(nodes, pcc_graph_construction(nodes))
        out = self.global_max_pool(out.permute(0, 2, 1)).squeeze(-1)
        return self.fc2(F.relu(self.fc1(out)))
