"""HAGCN (BASELINE.json configs[4]) assembled from the native primitives -- drop-in for models/HAGCN/Model.py
(same class names, constructor arguments, parameter names, forward signature incl. the `train` flag that adds
the KL term; state dicts interchange).

Native (libstgconv_b200.so): the cosine adjacency (stg_adj_*) and every dense aggregation A.X of the GIN and
SAGPool layers (stg_agg_*, forward and backward wrt both A and X), and the recurrence of the three bidirectional
LSTMs (stg_rnn_*, rnn.LSTM: persistent kernels with W_hh in registers -- the reference's layout makes the SEQUENCE
bs*N steps long with a batch of num_patch, the worst case for a launch-per-step library).  Projections are library
GEMMs; node ranking (sort / gather) stays index arithmetic in torch.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .primitives import cosine_distance, graph_matmul, tall_linear
from .rnn import LSTM


class GINLayer(nn.Module):
    """models/HAGCN/Model.py:6-24: mlp(A x + (1 + eps) x)."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.eps = nn.Parameter(torch.Tensor([0]))
        self.mlp = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))

    def forward(self, x, adj):
        h = graph_matmul(adj, x) + (1 + self.eps) * x
        h = F.relu(tall_linear(h, self.mlp[0].weight, self.mlp[0].bias))        # split-K weight gradients (bs*tlen*N rows)
        return tall_linear(h, self.mlp[2].weight, self.mlp[2].bias)


class Bi_LSTM_Standard(nn.Module):
    """models/HAGCN/Model.py:26-73 (three bidirectional LSTMs, halves summed; drop1 exists but is never applied)."""

    def __init__(self, input_dim, num_hidden, time_length):
        super().__init__()
        self.num_hidden, self.input_dim, self.time_length = 16, input_dim, time_length
        self.bi_lstm1 = LSTM(input_size=input_dim, hidden_size=num_hidden, num_layers=1, batch_first=True,
                             dropout=0, bidirectional=True)
        self.drop1 = nn.Dropout(p=0.2)
        self.bi_lstm2 = LSTM(input_size=num_hidden, hidden_size=num_hidden * 2, num_layers=1, batch_first=True,
                             dropout=0, bidirectional=True)
        self.drop2 = nn.Dropout(p=0.2)
        self.bi_lstm3 = LSTM(input_size=num_hidden * 2, hidden_size=num_hidden, num_layers=1, batch_first=True,
                             bidirectional=True)
        self.drop3 = nn.Dropout(p=0.2)

    @staticmethod
    def _fold(x):
        a, b = torch.split(x, x.shape[2] // 2, 2)
        return a + b

    def forward(self, x):
        x = self._fold(self.bi_lstm1(x)[0])
        x = self.drop2(self._fold(self.bi_lstm2(x)[0]))
        x = self.drop3(self._fold(self.bi_lstm3(x)[0]))
        return F.leaky_relu(x)


class SAGPool(nn.Module):
    """models/HAGCN/Model.py:75-120: self-attention pooling to the n best-ranked nodes + KL(prior || rank)."""

    def __init__(self, input_dimension, output_dimension, n):
        super().__init__()
        self.rank = nn.Linear(input_dimension, 1)
        self.model = nn.Linear(input_dimension, output_dimension)
        self.n = n
        self.mlp = nn.Sequential(nn.Linear(input_dimension, input_dimension // 2), nn.ReLU(),
                                 nn.Linear(input_dimension // 2, 1))

    def forward(self, X, A):
        AX = graph_matmul(A, X)                       # the reference computes bmm(A, X) twice; once is enough
        x_out = F.leaky_relu(tall_linear(AX, self.model.weight, self.model.bias))
        hm = F.relu(tall_linear(X, self.mlp[0].weight, self.mlp[0].bias))
        P = torch.softmax(tall_linear(hm, self.mlp[2].weight, self.mlp[2].bias), dim=1).squeeze()
        score = torch.softmax(tall_linear(AX, self.rank.weight, self.rank.bias), 1).squeeze()
        kl_div = F.kl_div(P.log(), score, reduction='batchmean')
        _, idx = torch.sort(score, descending=True, dim=1)
        topk = idx[:, :self.n]
        bat_id = torch.arange(X.size(0), device=X.device).unsqueeze(1)
        x_out = x_out[bat_id, topk]
        A_out = torch.transpose(A[bat_id, topk], 1, 2)[bat_id, topk]
        return x_out, A_out, kl_div


class HAGCN_model(nn.Module):
    """models/HAGCN/Model.py:129-195.  forward(X[bs, N, num_patch*patch_size], train=False) -> [bs,1] (, kl)."""

    def __init__(self, patch_size, num_patch, encoder_hidden_dim, hidden_dim, output_dim):
        super().__init__()
        self.patch_size, self.num_patch = patch_size, num_patch
        self.TD = Bi_LSTM_Standard(patch_size, encoder_hidden_dim, None)
        self.gin1 = GINLayer(encoder_hidden_dim, hidden_dim)
        self.gnn1 = SAGPool(hidden_dim, hidden_dim, 10)
        self.gin2 = GINLayer(hidden_dim, hidden_dim)
        self.gnn2 = SAGPool(hidden_dim, hidden_dim, 5)
        self.gin3 = GINLayer(hidden_dim, hidden_dim)
        self.gnn3 = SAGPool(hidden_dim, hidden_dim, 1)
        self.fc = nn.Sequential(nn.Linear(hidden_dim * 3 * num_patch, output_dim), nn.ReLU(inplace=True),
                                nn.Linear(output_dim, 1))

    def forward(self, X, train=False):
        bs, num_node, _ = X.size()
        tlen = self.num_patch
        # the reference feeds the LSTM [tlen, bs*N, patch]: batch = patches, SEQUENCE = bs*N (Model.py:155-161)
        seq = X.reshape(bs * num_node, tlen, self.patch_size).transpose(1, 0).contiguous()
        td = self.TD(seq).transpose(1, 0).reshape(bs, num_node, tlen, -1).transpose(1, 2)
        g = td.reshape(bs * tlen, num_node, -1)
        adj0 = cosine_distance(g)
        out1, adj1, kl1 = self.gnn1(self.gin1(g, adj0), adj0)
        out2, adj2, kl2 = self.gnn2(self.gin2(out1, adj1), adj1)
        out3, _, kl3 = self.gnn3(self.gin3(out2, adj2), adj2)
        out = torch.cat([out1.mean(1), out2.mean(1), out3.mean(1)], dim=-1).squeeze().reshape(bs, -1)
        output = self.fc(out)
        return (output, kl1 + kl2 + kl3) if train else output
