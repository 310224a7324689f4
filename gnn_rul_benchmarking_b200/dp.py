"""Data-parallel plumbing for the unchanged caller.

`Algorithm.update` runs zero_grad -> backward -> step back to back (algorithms/algorithms.py:72-74)
with no hook point, so the gradient exchange fires from inside backward: the first post-accumulate-grad
hook of a backward pass queues an end-of-backward callback on the autograd engine, which all-reduces ONE
flat fp32 buffer (FC_STGNN FD004: 66 429 floats = 266 KB) and scatters the mean back into the .grad
tensors.  Parameters that took no part in the pass (the reference keeps never-used TemporalConvNet.net0/net1
modules, models/ST_GCN/Model.py:110-132) contribute zeros and keep grad = None, exactly as without the hook.
BatchNorm statistics stay per rank (PyTorch-DDP default); parameters and buffers are broadcast
from rank 0 once at attach time.  Works with any torch.distributed backend (nccl on the GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int, drop_tail: bool = True) -> range:
    """Strided split of a (shuffled) window index across ranks.  With drop_tail every rank gets
    the same count so no rank waits in a collective for a batch another rank does not have
    (SURVEY.md 8e caveat 3)."""
    per = n_items // world if drop_tail else -(-n_items // world)
    return range(rank, min(n_items, per * world) if drop_tail else n_items, world)


class FlatGradAllReduce:
    """Attach to a module: one all-reduce (mean) of all gradients per backward pass."""

    def __init__(self, module: torch.nn.Module, group=None, broadcast: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self._pending = 0
        self._handles = []
        if broadcast:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, 0, group=group)
        for p in self.params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self.n_allreduce = 0

    def _hook(self, _p):
        # first gradient of this backward pass: run the exchange once the whole pass is done (every parameter that
        # is going to get a gradient has it by then; unused parameters never fire a hook and must not be waited for)
        if not self._pending:
            self._pending = 1
            torch.autograd.Variable._execution_engine.queue_callback(self._exchange)

    def _exchange(self):
        self._pending = 0
        used = [p.grad is not None for p in self.params]
        grads = [p.grad if u else torch.zeros_like(p) for p, u in zip(self.params, used)]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        off = 0
        dst, views = [], []
        for g, u in zip(grads, used):
            if u:                                   # same set on every rank: the model and its inputs' shapes agree
                dst.append(g)
                views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        if dst:
            torch._foreach_copy_(dst, views)
        self.n_allreduce += 1

    def detach(self):
        for h in self._handles:
            h.remove()
        self._handles = []
