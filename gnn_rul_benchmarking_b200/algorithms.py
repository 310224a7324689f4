"""Host-side mirror of the reference's algorithm wrapper for the FC_STGNN path
(algorithms/algorithms.py:29-76): `get_algorithm_class(name)(configs, hparams, device)` with
`.model`, `.optimizer`, `.update(X, y, epoch) -> {'loss': float}`, so trainer.py:96-110 drives
it unchanged.  The model is the drop-in FC_STGNN_RUL whose graph-conv blocks run in the sm_100a
extension; there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .fc_stgnn import FC_STGNN_RUL


def get_algorithm_class(algorithm_name):
    """algorithms.py:29-33: unknown names raise NotImplementedError."""
    if algorithm_name not in _ALGORITHMS:
        raise NotImplementedError("Algorithm not found: {}".format(algorithm_name))
    return _ALGORITHMS[algorithm_name]


class Algorithm(nn.Module):
    """algorithms.py:36-48."""

    def __init__(self, configs):
        super().__init__()
        self.configs = configs
        self.mse = nn.MSELoss()

    def update(self, *args, **kwargs):
        raise NotImplementedError


class FC_STGNN(Algorithm):
    """algorithms.py:51-76: Adam(lr, weight_decay) + MSE; update = forward -> mse -> zero_grad ->
    backward -> step -> {'loss': loss.item()}."""

    def __init__(self, configs, hparams, device):
        super().__init__(configs)
        self.model = FC_STGNN_RUL(**configs)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=hparams["learning_rate"],
                                          weight_decay=hparams["weight_decay"])
        self.hparams = hparams

    def step(self, X, y):
        """One optimisation step with everything left on the device; returns the loss tensor."""
        predicted_RUL = self.model(X)
        loss = self.mse(predicted_RUL, y)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss

    def update(self, X, y, epoch=None):
        return {"loss": self.step(X, y).item()}


_ALGORITHMS = {"FC_STGNN": FC_STGNN}
