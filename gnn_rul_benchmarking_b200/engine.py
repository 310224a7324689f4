"""Host side of the whole-model engine: binds an FC_STGNN_RUL module's tensors to the C ABI
(stg_model_* / stg_adam_step in include/stgconv_b200.h).  torch supplies device memory, the
current stream and autograd bookkeeping; every arithmetic op of the model runs in
libstgconv_b200.so.  No CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


StgModelDims, StgBN, StgModelBlock = _lib.StgModelDims, _lib.StgBN, _lib.StgModelBlock
StgModelParams, StgDropout = _lib.StgModelParams, _lib.StgDropout


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def param_tensors(model) -> List[Tuple[str, torch.Tensor]]:
    """(struct path, tensor) for every learnable tensor, in a fixed order."""
    out = [("conv1_w", model.nonlin_map.conv_block1[0].weight),
           ("bn1.weight", model.nonlin_map.conv_block1[1].weight), ("bn1.bias", model.nonlin_map.conv_block1[1].bias),
           ("conv2_w", model.nonlin_map.conv_block2[0].weight),
           ("bn2.weight", model.nonlin_map.conv_block2[1].weight), ("bn2.bias", model.nonlin_map.conv_block2[1].bias),
           ("lin_w", model.nonlin_map2[0].weight), ("lin_b", model.nonlin_map2[0].bias),
           ("bn3.weight", model.nonlin_map2[1].weight), ("bn3.bias", model.nonlin_map2[1].bias)]
    for z, blk in enumerate((model.MPNN1, model.MPNN2)):
        m, th = blk.graph_construction.mapping, blk.MPNN.theta[0]
        out += [(f"blk.{z}.Wm", m.weight), (f"blk.{z}.bm", m.bias),
                (f"blk.{z}.bn0.weight", blk.BN.weight), (f"blk.{z}.bn0.bias", blk.BN.bias),
                (f"blk.{z}.Wt", th.weight), (f"blk.{z}.bt", th.bias),
                (f"blk.{z}.bn1.weight", blk.MPNN.bn1.weight), (f"blk.{z}.bn1.bias", blk.MPNN.bn1.bias)]
    for i, name in enumerate(("fc1", "fc2", "fc3", "fc4")):
        lin = getattr(model.fc, name)
        out += [(f"fc_w.{i}", lin.weight), (f"fc_b.{i}", lin.bias)]
    return out


def buffer_tensors(model) -> List[Tuple[str, torch.Tensor]]:
    out = []
    bns = [("bn1", model.nonlin_map.conv_block1[1]), ("bn2", model.nonlin_map.conv_block2[1]),
           ("bn3", model.nonlin_map2[1])]
    for z, blk in enumerate((model.MPNN1, model.MPNN2)):
        bns += [(f"blk.{z}.bn0", blk.BN), (f"blk.{z}.bn1", blk.MPNN.bn1)]
    for path, bn in bns:
        out += [(path + ".running_mean", bn.running_mean), (path + ".running_var", bn.running_var),
                (path + ".num_batches_tracked", bn.num_batches_tracked)]
    out.append(("pe", model.positional_encoding.pe))
    return out


def _set_path(struct, path: str, value: int) -> None:
    parts = path.split(".")
    obj = struct
    for k in parts[:-1]:
        obj = obj[int(k)] if k.isdigit() else getattr(obj, k)
    last = parts[-1]
    if last.isdigit():
        obj[int(last)] = value
    else:
        setattr(obj, last, value)


def _check(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the sm_100a engine has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


class ModelEngine:
    """Per-module binding: pointer structs, workspaces, flat gradient / Adam buffers."""

    def __init__(self, model):
        self.model = model
        self._fp = None            # fingerprint of the bound tensors
        self._params = None        # StgModelParams
        self._ws: Dict[Tuple, torch.Tensor] = {}
        self._gen = 0              # bumped by every training forward (detects a clobbered workspace)
        self.flat: Optional[dict] = None
        self.graph_seed = None     # (base seed, device step counter) while a CUDA graph drives the dropout

    # ---------------------------------------------------------------- binding
    def _tensors(self):
        return param_tensors(self.model) + buffer_tensors(self.model)

    def bind(self) -> StgModelParams:
        tens = self._tensors()
        fp = tuple(t.data_ptr() for _, t in tens)
        if fp == self._fp:
            return self._params
        dev = tens[0][1].device
        ps = StgModelParams()
        for path, t in tens:
            if path.endswith("num_batches_tracked"):
                _check(t, path, torch.int64)
            else:
                _check(t, path)
            if t.device != dev:
                raise RuntimeError("all model tensors must be on one device")
            _set_path(ps, path, t.data_ptr())
        self._fp, self._params = fp, ps
        return ps

    def dims(self, B: int) -> StgModelDims:
        m = self.model
        d = StgModelDims()
        conv1 = m.nonlin_map.conv_block1[0]
        d.B, d.N, d.T, d.P = B, m.MPNN1.num_sensors, m.num_patch, m.patch_size
        d.K, d.EH, d.E = conv1.kernel_size[0], conv1.out_channels, m.nonlin_map.conv_block2[0].out_channels
        d.H = m.MPNN1.output_dim
        for z, blk in enumerate((m.MPNN1, m.MPNN2)):
            d.w[z], d.stride[z] = blk.time_window_size, blk.stride
        d.decay = m.MPNN1.decay
        drop = m.positional_encoding.dropout
        d.pe_dropout = float(getattr(drop, "p", 0.0)) if m.training else 0.0
        d.bn_momentum, d.bn_eps = BN_MOMENTUM, BN_EPS
        return d

    def workspace(self, d: StgModelDims, device, slot: str) -> torch.Tensor:
        key = (slot, d.B, str(device))
        ws = self._ws.get(key)
        if ws is None:
            n = _lib.load().stg_model_workspace_bytes(C.byref(d))
            if n == 0:
                raise ValueError("invalid model dimensions for the sm_100a engine")
            ws = torch.empty(n + 256, dtype=torch.uint8, device=device)
            off = (-ws.data_ptr()) % 256
            ws = ws[off:off + n]
            self._ws[key] = ws
        return ws

    def dropout(self, d: StgModelDims, X: torch.Tensor) -> Tuple[StgDropout, Optional[torch.Tensor]]:
        dr = StgDropout()
        keep = getattr(self.model.positional_encoding.dropout, "keep", None)
        if keep is not None and d.pe_dropout > 0:
            keep = keep.to(device=X.device, dtype=torch.float32).contiguous()
            dr.keep = keep.data_ptr()
        elif d.pe_dropout > 0 and self.graph_seed is not None:
            dr.seed = self.graph_seed[0]                               # + device counter, ticked per step
            dr.step_dev = self.graph_seed[1].data_ptr()
        elif d.pe_dropout > 0:
            # like the reference's nn.Dropout on cuda, consume the CUDA generator (seed + Philox offset, advanced on the
            # host without launching anything): follows torch.manual_seed, leaves the CPU stream -- and with it the
            # DataLoader / DeviceLoader permutations -- exactly where the reference run would have it
            dr.seed = draw_cuda_seed(X.device)
        return dr, keep

    # ---------------------------------------------------------------- forward / backward
    def forward(self, X: torch.Tensor, training: bool):
        """-> (pred [B,1], saved) ; saved is what backward() needs (None in eval mode)."""
        lib = _lib.load()
        _check(X, "X")
        m = self.model
        B, N, L = X.shape
        if L != m.num_patch * m.patch_size:
            raise ValueError(f"time length {L} != num_patch*patch_size = {m.num_patch * m.patch_size}")
        d = self.dims(B)
        if N != d.N:
            raise ValueError(f"X has {N} sensors, the model was built for {d.N}")
        ps = self.bind()
        ws = self.workspace(d, X.device, "train" if training else "eval")
        dr, keep = self.dropout(d, X)
        pred = torch.empty(B, 1, device=X.device, dtype=torch.float32)
        with torch.cuda.device(X.device):
            _lib.check(lib.stg_model_forward(C.byref(d), C.byref(ps), X.data_ptr(), ws.data_ptr(), ws.numel(),
                                             int(training), C.byref(dr), pred.data_ptr(), _stream()),
                       "stg_model_forward")
        if not training:
            return pred, None
        self._gen += 1
        return pred, (d, dr, keep, ws, self._gen)

    def backward(self, X: torch.Tensor, saved, dpred: torch.Tensor, grads: StgModelParams) -> None:
        lib = _lib.load()
        d, dr, keep, ws, gen = saved
        if gen != self._gen:
            raise RuntimeError("stgconv: another training forward ran on this module before backward(); "
                               "the saved activations were overwritten")
        ps = self.bind()
        dpred = dpred.contiguous()
        with torch.cuda.device(X.device):
            _lib.check(lib.stg_model_backward(C.byref(d), C.byref(ps), C.byref(grads), X.data_ptr(), ws.data_ptr(),
                                              ws.numel(), C.byref(dr), dpred.data_ptr(), _stream()),
                       "stg_model_backward")

    def grad_struct(self, flat: torch.Tensor, offsets: List[int]) -> StgModelParams:
        gs = StgModelParams()
        base = flat.data_ptr()
        for (path, _), off in zip(param_tensors(self.model), offsets):
            _set_path(gs, path, base + 4 * off)
        return gs

    # ---------------------------------------------------------------- flat parameter / optimizer state
    def flatten(self) -> dict:
        """Moves every parameter into ONE flat fp32 buffer (parameters become views of it) with a
        matching flat gradient buffer (p.grad views), so the optimizer is one kernel and the
        data-parallel exchange one all-reduce.  Re-done automatically if the module was moved."""
        pts = param_tensors(self.model)
        fl = self.flat
        if fl is not None and all(t.data_ptr() == fl["base"] + 4 * o for (_, t), o in zip(pts, fl["offsets"])):
            return fl
        dev = pts[0][1].device
        offsets, o = [], 0
        for _, t in pts:
            offsets.append(o)
            o += (t.numel() + 3) // 4 * 4          # 16-byte aligned slots
        flat = torch.zeros(o, device=dev, dtype=torch.float32)
        gflat = torch.zeros(o, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for (_, t), off in zip(pts, offsets):
                _check(t, "parameter")
                view = flat[off:off + t.numel()].view_as(t)
                view.copy_(t)
                t.data = view
                t.grad = gflat[off:off + t.numel()].view_as(t)
        self.flat = dict(base=flat.data_ptr(), param=flat, grad=gflat, offsets=offsets, n=o,
                         gstruct=None)
        self.flat["gstruct"] = self.grad_struct(gflat, offsets)
        self._fp = None
        return self.flat

    def replace_grad_buffer(self, gflat: torch.Tensor) -> None:
        """Use `gflat` (e.g. NVLink symmetric memory, so peers can read it) as the flat gradient buffer."""
        fl = self.flatten()
        if gflat.numel() != fl["n"] or gflat.dtype != torch.float32 or gflat.device != fl["param"].device:
            raise ValueError("gradient buffer must be float32, on the parameters' device, with the flat size")
        gflat.zero_()
        for (_, t), off in zip(param_tensors(self.model), fl["offsets"]):
            t.grad = gflat[off:off + t.numel()].view_as(t)
        fl["grad"] = gflat
        fl["gstruct"] = self.grad_struct(gflat, fl["offsets"])

    def loss_backward(self, X: torch.Tensor, y: torch.Tensor, zero_grad: bool = True):
        """forward -> MSE -> backward in one C call; gradients land in the flat gradient buffer.
        Returns the loss as a 0-dim device tensor (no host synchronisation)."""
        lib = _lib.load()
        _check(X, "X")
        _check(y, "y")
        m = self.model
        if not m.training:
            raise RuntimeError("loss_backward needs the module in training mode")
        B, N, L = X.shape
        if L != m.num_patch * m.patch_size:
            raise ValueError(f"time length {L} != num_patch*patch_size = {m.num_patch * m.patch_size}")
        if y.numel() != B:
            raise ValueError("y must hold one target per window")
        fl = self.flatten()
        d = self.dims(B)
        if N != d.N:
            raise ValueError(f"X has {N} sensors, the model was built for {d.N}")
        ps = self.bind()
        ws = self.workspace(d, X.device, "train")
        dr, keep = self.dropout(d, X)
        loss = torch.empty((), device=X.device, dtype=torch.float32)
        if zero_grad:
            fl["grad"].zero_()
        self._gen += 1
        with torch.cuda.device(X.device):
            _lib.check(lib.stg_model_loss_backward(C.byref(d), C.byref(ps), C.byref(fl["gstruct"]), X.data_ptr(),
                                                   y.data_ptr(), ws.data_ptr(), ws.numel(), C.byref(dr), None,
                                                   loss.data_ptr(), _stream()), "stg_model_loss_backward")
        return loss


class _ModelFn(torch.autograd.Function):
    """FC_STGNN_RUL.forward as one autograd node (the unchanged-caller path: the reference's
    algorithms.py runs mse -> zero_grad -> backward -> optimizer.step around it)."""

    @staticmethod
    def forward(ctx, engine: ModelEngine, X, *params):
        pred, saved = engine.forward(X, True)
        ctx.engine, ctx.saved = engine, saved
        ctx.save_for_backward(X)
        ctx.shapes = [p.shape for p in params]
        return pred

    @staticmethod
    def backward(ctx, dpred):
        engine = ctx.engine
        (X,) = ctx.saved_tensors
        offsets, o = [], 0
        for s in ctx.shapes:
            offsets.append(o)
            o += (s.numel() + 3) // 4 * 4
        gflat = torch.zeros(o, device=X.device, dtype=torch.float32)
        engine.backward(X, ctx.saved, dpred, engine.grad_struct(gflat, offsets))
        grads = [gflat[off:off + s.numel()].view(s) for off, s in zip(offsets, ctx.shapes)]
        return (None, None, *grads)


def model_forward(engine: ModelEngine, X: torch.Tensor) -> torch.Tensor:
    m = engine.model
    X = X.contiguous()
    if m.training and torch.is_grad_enabled():
        return _ModelFn.apply(engine, X, *[t for _, t in param_tensors(m)])
    pred, _ = engine.forward(X, m.training)
    return pred


def draw_cuda_seed(device) -> int:
    """A 62-bit seed from the device's default CUDA generator: (seed, Philox offset), offset advanced on the host."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    off = int(gen.get_offset())
    gen.set_offset(off + 4)
    return (int(gen.initial_seed()) * 0x9E3779B97F4A7C15 + off * 0xD1B54A32D192ED03 + 1) % (2 ** 62)


class StgAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (algorithms.py:60-64) as ONE kernel
    over the engine's flat parameter / gradient buffers (stg_adam_step)."""

    def __init__(self, engine: ModelEngine, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [t for _, t in param_tensors(engine.model)]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.engine = engine
        self._st = None
        self.grad_scale = 1.0
        self._p2p = None           # (grad_ptrs, flag_ptrs, rank, world, keepalive) once attach_p2p() ran

    def attach_p2p(self, grad_ptrs, flag_ptrs, rank, world, keepalive) -> None:
        """Exchange gradients inside the optimizer kernel: peers' flat gradient buffers are read over
        NVLink (stg_allreduce_adam) instead of an NCCL all-reduce followed by the Adam kernel."""
        gp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in grad_ptrs])
        fp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in flag_ptrs])
        self._p2p = (gp, fp, int(rank), int(world), keepalive)

    def _state(self):
        fl = self.engine.flatten()
        st = self._st
        if st is None or st["base"] != fl["base"]:
            dev = fl["param"].device
            old = st
            st = dict(base=fl["base"], exp_avg=torch.zeros(fl["n"], device=dev),
                      exp_avg_sq=torch.zeros(fl["n"], device=dev),
                      step=torch.zeros((), device=dev, dtype=torch.int64))
            if old is not None:                       # module moved: carry the moments over
                st["exp_avg"].copy_(old["exp_avg"])
                st["exp_avg_sq"].copy_(old["exp_avg_sq"])
                st["step"].copy_(old["step"])
            self._st = st
            for (_, p), off in zip(param_tensors(self.engine.model), fl["offsets"]):
                self.state[p] = dict(step=st["step"], exp_avg=st["exp_avg"][off:off + p.numel()].view_as(p),
                                     exp_avg_sq=st["exp_avg_sq"][off:off + p.numel()].view_as(p))
        return fl, st

    def state_dict(self):
        """Standard torch.optim layout; the per-parameter exp_avg / exp_avg_sq entries are views of the flat moment
        buffers the kernel updates, `step` is the shared device counter."""
        self._state()
        return super().state_dict()

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Copies the loaded moments and step INTO the flat buffers (torch's implementation would rebind
        self.state[p] to fresh tensors the kernel never reads)."""
        fl, st = self._state()
        groups = state_dict["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        params = [t for _, t in param_tensors(self.engine.model)]
        if len(ids) != len(params):
            raise ValueError("loaded optimizer state has a different number of parameters")
        step = None
        for pid, p, off in zip(ids, params, fl["offsets"]):
            ent = state_dict["state"].get(pid)
            if ent is None:
                continue
            st["exp_avg"][off:off + p.numel()].copy_(ent["exp_avg"].reshape(-1))
            st["exp_avg_sq"][off:off + p.numel()].copy_(ent["exp_avg_sq"].reshape(-1))
            step = ent["step"] if step is None else step
        if step is not None:
            st["step"].fill_(int(step))
        for g, new in zip(self.param_groups, groups):
            for k, v in new.items():
                if k != "params":
                    g[k] = v

    def zero_grad(self, set_to_none: bool = False):
        fl = self.engine.flatten()
        fl["grad"].zero_()
        for (_, p), off in zip(param_tensors(self.engine.model), fl["offsets"]):
            if p.grad is None or p.grad.data_ptr() != fl["grad"].data_ptr() + 4 * off:
                p.grad = fl["grad"][off:off + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self, closure=None):
        fl, st = self._state()
        g = self.param_groups[0]
        # gradients produced by autograd may have replaced the flat views: gather them back
        for (_, p), off in zip(param_tensors(self.engine.model), fl["offsets"]):
            if p.grad is not None and p.grad.data_ptr() != fl["grad"].data_ptr() + 4 * off:
                fl["grad"][off:off + p.numel()].view_as(p).copy_(p.grad)
        lib = _lib.load()
        if self._p2p is not None:
            gp, fp, rank, world, _ = self._p2p
            with torch.cuda.device(fl["param"].device):
                _lib.check(lib.stg_allreduce_adam(fl["param"].data_ptr(), st["exp_avg"].data_ptr(),
                                                  st["exp_avg_sq"].data_ptr(), fl["n"], st["step"].data_ptr(), gp, fp,
                                                  rank, world, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                                                  g["weight_decay"], _stream()), "stg_allreduce_adam")
            return None
        with torch.cuda.device(fl["param"].device):
            _lib.check(lib.stg_adam_step(fl["param"].data_ptr(), fl["grad"].data_ptr(), st["exp_avg"].data_ptr(),
                                         st["exp_avg_sq"].data_ptr(), fl["n"], st["step"].data_ptr(), g["lr"],
                                         g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                                         float(self.grad_scale), _stream()), "stg_adam_step")
        return None
