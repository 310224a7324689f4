"""Host-side mirror of the reference's algorithm wrapper for the FC_STGNN path
(algorithms/algorithms.py:29-76): `get_algorithm_class(name)(configs, hparams, device)` with
`.model`, `.optimizer`, `.update(X, y, epoch) -> {'loss': float}`, so trainer.py:96-110 drives
it unchanged.  The model is the drop-in FC_STGNN_RUL whose graph-conv blocks run in the sm_100a
extension; there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import StgAdam
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
        # same update rule as torch.optim.Adam(lr, weight_decay) (algorithms.py:60-64), one kernel
        self.optimizer = StgAdam(self.model.engine, lr=hparams["learning_rate"],
                                 weight_decay=hparams["weight_decay"])
        self.hparams = hparams
        self._dp_group, self._dp_world = None, 1

    def attach_data_parallel(self, group=None, broadcast=True):
        """Shard windows across ranks, all-reduce ONE flat gradient buffer per step."""
        import torch.distributed as dist
        self._dp_group, self._dp_world = group, dist.get_world_size(group)
        if broadcast:
            with torch.no_grad():
                for t in list(self.model.parameters()) + list(self.model.buffers()):
                    dist.broadcast(t, 0, group=group)
        self.optimizer.grad_scale = 1.0 / self._dp_world

    def step(self, X, y):
        """One optimisation step (forward -> MSE -> backward -> [all-reduce] -> Adam) with everything
        left on the device; returns the loss as a 0-dim device tensor."""
        eng = self.model.engine
        loss = eng.loss_backward(X, y, zero_grad=True)
        if self._dp_world > 1:
            import torch.distributed as dist
            dist.all_reduce(eng.flat["grad"], op=dist.ReduceOp.SUM, group=self._dp_group)
        self.optimizer.step()
        return loss

    def update(self, X, y, epoch=None):
        return {"loss": self.step(X, y).item()}


_ALGORITHMS = {"FC_STGNN": FC_STGNN}
