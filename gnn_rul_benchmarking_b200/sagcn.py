"""SAGCN assembled from the native primitives -- drop-in for models/SAGCN/Model.py (same class names, constructor
arguments, parameter names, forward signature; state dicts interchange).

Native (libstgconv_b200.so): the 12 temporal patch statistics (stg_patch_stats12), the cosine adjacency (stg_adj_*)
and the sym-norm GCN aggregation (stg_agg_*).  The 8 spectral statistics use cuFFT / device sort through torch, the
cumulative features are a closed form of the reference's O(L^2 f) Python loop (one cumsum), the projections are
library GEMMs.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .primitives import cosine_distance, extract_temporal_features, gcn_aggregate, tall_linear


def generate_cumulative_features(signals):
    """models/SAGCN/Model.py:7-19: x'_p = cumsum_p / sqrt(|cumsum_p|) (clamped at 1e-12), without the loops."""
    cs = torch.cumsum(signals, dim=1)
    return cs / torch.sqrt(cs.abs().clamp_min(1e-12))


def extract_frequency_features(signals, fs=1.0):
    """models/SAGCN/Model.py:41-57 on the device (cuFFT).  Ties between the mirrored bins k / n-k of a real signal are
    broken towards the lower index (stable sort, first maximum), which is what the reference's CPU run does for the
    patch sizes it configures."""
    n = signals.shape[-1]
    freqs = torch.fft.fftfreq(n, d=1 / fs, device=signals.device)
    fft_vals = torch.fft.fft(signals, dim=-1)
    amp = torch.abs(fft_vals)
    psd = amp ** 2 / n
    tot = torch.sum(psd, dim=-1)
    mean_freq = torch.sum(freqs * psd, dim=-1) / tot
    median_freq = freqs[torch.argsort(psd, dim=-1, stable=True)[:, n // 2]]
    occupied_bw = torch.sum(psd * (freqs < fs / 2), dim=-1) / tot
    power_bw = torch.sqrt(torch.sum(psd ** 2, dim=-1) / tot)
    return torch.stack([mean_freq, median_freq, tot, occupied_bw, power_bw, torch.max(psd, dim=-1)[0],
                        torch.max(amp, dim=-1)[0], freqs[torch.argmax(amp, dim=-1)]], dim=-1)


def extract_features(signals, fs=1.0):
    """models/SAGCN/Model.py:60-72: [bs, num_patch, patch_size] -> [bs, num_patch, 40], Frobenius-normalised."""
    bs, num_patch, _ = signals.size()
    rows = signals.reshape(bs * num_patch, -1)
    feats = torch.cat([extract_temporal_features(rows), extract_frequency_features(rows, fs)], dim=-1)
    feats = feats.reshape(bs, num_patch, -1)
    feats = torch.cat([feats, generate_cumulative_features(feats)], -1)
    return feats / torch.norm(feats, dim=(1, 2), keepdim=True)


class GCNLayer(nn.Module):
    """models/SAGCN/Model.py:81-95: relu(Linear(D^-1/2 (A+I) D^-1/2 X))."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features)

    def forward(self, X, A):
        return F.relu(tall_linear(gcn_aggregate(X, A), self.linear.weight, self.linear.bias))


class GraphProjectionLayer(nn.Module):
    """models/SAGCN/Model.py:99-112."""

    def __init__(self, in_features, out_features, num_nodes):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features)
        self.project_matrices = nn.Linear(num_nodes, num_nodes)

    def forward(self, x):
        p = tall_linear(x.transpose(-1, -2), self.project_matrices.weight, self.project_matrices.bias)
        return F.relu(tall_linear(p.transpose(-1, -2), self.linear.weight, self.linear.bias))


class SelfAttentionLayer(nn.Module):
    """models/SAGCN/Model.py:115-124."""

    def __init__(self, num_nodes, attention_hidden_dim):
        super().__init__()
        self.tanh_layer = nn.Linear(num_nodes, attention_hidden_dim)
        self.softmax_layer = nn.Linear(attention_hidden_dim, num_nodes)

    def forward(self, x):
        scores = torch.tanh(tall_linear(x.transpose(-1, -2), self.tanh_layer.weight, self.tanh_layer.bias))
        return F.softmax(tall_linear(scores, self.softmax_layer.weight, self.softmax_layer.bias), dim=-1).transpose(-1, -2)


class SAGCN_model(nn.Module):
    """models/SAGCN/Model.py:127-163.  forward(x[bs, (1,) num_patch*patch_size]) -> [bs, 1]."""

    def __init__(self, num_patch, patch_size, gcn_hidden_dim, attention_hidden_dim):
        super().__init__()
        self.num_patch, self.patch_size = num_patch, patch_size
        self.gcn1 = GCNLayer(40, gcn_hidden_dim)
        self.proj1 = GraphProjectionLayer(gcn_hidden_dim, gcn_hidden_dim, num_patch)
        self.proj2 = GraphProjectionLayer(gcn_hidden_dim, gcn_hidden_dim, num_patch)
        self.attn = SelfAttentionLayer(num_patch, attention_hidden_dim)
        self.fc = nn.Linear(gcn_hidden_dim * num_patch, 1)

    def forward(self, x):
        bs = x.size(0)
        feats = extract_features(x.reshape(bs, self.num_patch, self.patch_size))
        h = self.proj2(self.proj1(self.gcn1(feats, cosine_distance(feats))))
        h = h * self.attn(h)
        return self.fc(h.reshape(bs, -1))
