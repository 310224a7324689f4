"""Flat parameter / gradient buffers and the one-kernel Adam for ANY nn.Module (the sibling models' update rule,
algorithms/algorithms.py:139-163 and its clones: torch.optim.Adam(lr, weight_decay) over model.parameters()).

`FlatParams` moves the parameters that take part in training into ONE flat fp32 buffer (they become views of it) with a
matching flat gradient buffer (p.grad views), so that the optimizer is one launch of stg_adam_step and the
data-parallel exchange ONE all-reduce -- or, with NVLink symmetric memory, no separate collective at all
(stg_allreduce_adam reads the peers' gradient buffers inside the optimizer kernel, csrc/stg_p2p.cu).

Parameters the forward never uses (the reference keeps TemporalConvNet.net0 / net1 / downsample* for checkpoint
compatibility, models/ST_GCN/Model.py:110-132, models/ASTGCNN/Model.py:72-146) get no gradient in the reference, so
torch's Adam never touches them -- no weight decay either.  They are therefore left OUT of the flat buffers: `used`
is found by one probe backward on a sample batch.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def find_used_parameters(module: torch.nn.Module, loss_fn: Callable[[], torch.Tensor]) -> List[torch.nn.Parameter]:
    """Parameters that receive a gradient from `loss_fn()` (one probe forward + backward; leaves .grad = None).
    Buffers the forward updates (BatchNorm running statistics) are restored afterwards."""
    params = [p for p in module.parameters() if p.requires_grad]
    saved = [p.grad for p in params]
    bufs = [(b, b.detach().clone()) for b in module.buffers()]
    for p in params:
        p.grad = None
    loss_fn().backward()
    used = [p for p in params if p.grad is not None]
    for p, g in zip(params, saved):
        p.grad = g
    with torch.no_grad():
        for b, v in bufs:
            b.copy_(v)
    return used


class FlatParams:
    def __init__(self, params: List[torch.nn.Parameter]):
        if not params:
            raise ValueError("no parameters to flatten")
        dev = params[0].device
        self.params = list(params)
        self.offsets, o = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("parameters must be float32 on one device")
            self.offsets.append(o)
            o += (p.numel() + 3) // 4 * 4                      # 16-byte aligned slots
        self.n = o
        self.param = torch.zeros(o, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(o, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                v = self.param[off:off + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
        self.bind_grads()

    def bind_grads(self) -> None:
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view_as(p)

    def replace_grad_buffer(self, gflat: torch.Tensor) -> None:
        if gflat.numel() != self.n or gflat.dtype != torch.float32 or gflat.device != self.param.device:
            raise ValueError("gradient buffer must be float32, on the parameters' device, with the flat size")
        gflat.zero_()
        self.grad = gflat
        for p in self.params:
            p.grad = None
        self.bind_grads()

    def check_bound(self) -> None:
        """The parameters must still be views of the flat buffer: `module.to(...)` / `load_state_dict(assign=True)` after
        flattening re-allocates them, and the kernel would then update a buffer nobody reads."""
        base = self.param.data_ptr()
        for i in (0, len(self.params) - 1):          # a moved module re-allocates every parameter: two probes are enough
            if self.params[i].data_ptr() != base + 4 * self.offsets[i]:
                raise RuntimeError("a parameter is no longer a view of the flat buffer (module moved or re-assigned after "
                                   "use_flat_optimizer / attach_data_parallel): build the flat optimizer again")

    def gather_stray_grads(self) -> None:
        """autograd may have replaced a .grad view by a fresh tensor (first accumulation into None): copy it back."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is not None and p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                self.grad[off:off + p.numel()].view_as(p).copy_(p.grad)


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics as ONE kernel over FlatParams; the step counter lives
    on the device, so the update can be captured in a CUDA graph.  grad_scale folds the 1/world of data parallelism."""

    def __init__(self, flat: FlatParams, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(flat.params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.flat = flat
        dev = flat.param.device
        self.exp_avg = torch.zeros(flat.n, device=dev)
        self.exp_avg_sq = torch.zeros(flat.n, device=dev)
        self.step_dev = torch.zeros((), device=dev, dtype=torch.int64)
        for p, off in zip(flat.params, flat.offsets):
            self.state[p] = dict(step=self.step_dev, exp_avg=self.exp_avg[off:off + p.numel()].view_as(p),
                                 exp_avg_sq=self.exp_avg_sq[off:off + p.numel()].view_as(p))
        self.grad_scale = 1.0
        self._p2p = None

    def attach_p2p(self, grad_ptrs, flag_ptrs, rank, world, keepalive) -> None:
        gp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in grad_ptrs])
        fp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in flag_ptrs])
        self._p2p = (gp, fp, int(rank), int(world), keepalive)

    def state_tensors(self):
        return [self.flat.param, self.exp_avg, self.exp_avg_sq, self.step_dev]

    def zero_grad(self, set_to_none: bool = False):
        self.flat.grad.zero_()
        self.flat.bind_grads()

    @torch.no_grad()
    def step(self, closure=None):
        fl, g = self.flat, self.param_groups[0]
        fl.check_bound()
        fl.gather_stray_grads()
        lib = _lib.load()
        with torch.cuda.device(fl.param.device):
            if self._p2p is not None:
                gp, fp, rank, world, _ = self._p2p
                _lib.check(lib.stg_allreduce_adam(fl.param.data_ptr(), self.exp_avg.data_ptr(),
                                                  self.exp_avg_sq.data_ptr(), fl.n, self.step_dev.data_ptr(), gp, fp,
                                                  rank, world, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                                                  g["weight_decay"], _stream()), "stg_allreduce_adam")
            else:
                _lib.check(lib.stg_adam_step(fl.param.data_ptr(), fl.grad.data_ptr(), self.exp_avg.data_ptr(),
                                             self.exp_avg_sq.data_ptr(), fl.n, self.step_dev.data_ptr(), g["lr"],
                                             g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                                             float(self.grad_scale), _stream()), "stg_adam_step")
        return None
