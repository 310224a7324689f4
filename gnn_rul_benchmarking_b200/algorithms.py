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

    def attach_data_parallel(self, group=None, broadcast=True, p2p="auto"):
        """Shard windows across ranks; ONE exchange of the flat gradient buffer per step.
        p2p: True / "auto" -> gradients live in NVLink symmetric memory and the optimizer kernel reads the
        peers' buffers itself (stg_allreduce_adam); False (or when symmetric memory is unavailable with
        "auto") -> NCCL all-reduce followed by the Adam kernel."""
        import torch.distributed as dist
        self._dp_group, self._dp_world = group, dist.get_world_size(group)
        if broadcast:
            with torch.no_grad():
                for t in list(self.model.parameters()) + list(self.model.buffers()):
                    dist.broadcast(t, 0, group=group)
        self.optimizer.grad_scale = 1.0 / self._dp_world
        self._dp_p2p = False
        if p2p and self._dp_world > 1 and dist.get_backend(group) == "nccl":
            try:
                self._attach_p2p(group)
                self._dp_p2p = True
            except Exception:
                if p2p is True:
                    raise

    def _attach_p2p(self, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        eng = self.model.engine
        fl = eng.flatten()
        dev = fl["param"].device
        pg = group if group is not None else dist.group.WORLD
        gsym = symm_mem.empty(fl["n"], dtype=torch.float32, device=dev)
        hg = symm_mem.rendezvous(gsym, pg)
        flags = symm_mem.empty(64, dtype=torch.int32, device=dev)
        flags.zero_()
        flags[34] = 1            # epoch of the flag protocol: monotonic, never restored (stg_p2p.cu)
        hf = symm_mem.rendezvous(flags, pg)
        eng.replace_grad_buffer(gsym)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                       # every rank's flags are zero before anyone signals
        self.optimizer.attach_p2p(list(hg.buffer_ptrs), list(hf.buffer_ptrs), hg.rank, hg.world_size,
                                  (gsym, flags, hg, hf))
        self._p2p_flags = flags

    def p2p_timed_out(self) -> bool:
        """True if a peer failed to show up inside the fused exchange kernel (bounded wait)."""
        return bool(getattr(self, "_p2p_flags", None) is not None and int(self._p2p_flags[33]) != 0)

    def check_exchange(self) -> None:
        """Raises if the fused NVLink exchange ever timed out on this rank (the kernel then skipped the
        parameter update instead of applying a partial sum).  Synchronises the device."""
        if self.p2p_timed_out():
            raise RuntimeError("stg_allreduce_adam: a peer did not publish its gradients in time; "
                               "the parameter update of that step was skipped on this rank")

    # ------------------------------------------------------------------ CUDA-graph replay of the step
    def enable_cuda_graph(self, batch_size):
        """Captures one optimisation step (all library launches + the gradient exchange) for windows of
        `batch_size` into a CUDA graph; later step()/update() calls with that batch size copy X, y into
        static buffers and replay it.  Model/optimizer state is saved before the warm-up and capture
        and restored afterwards, so enabling the graph does not change training."""
        eng = self.model.engine
        fl = eng.flatten()
        dev = fl["param"].device
        _, st = self.optimizer._state()
        m = self.model
        N, L = m.MPNN1.num_sensors, m.num_patch * m.patch_size
        self._gX = torch.zeros(batch_size, N, L, device=dev)
        self._gy = torch.zeros(batch_size, 1, device=dev)
        from .engine import draw_cuda_seed
        eng.graph_seed = (draw_cuda_seed(dev), torch.zeros(1, dtype=torch.int64, device=dev))
        snap = [t.clone() for t in (fl["param"], st["exp_avg"], st["exp_avg_sq"], st["step"])]
        bufs = [b for b in m.buffers()]
        snap_b = [b.clone() for b in bufs]
        was_training = self.training
        self.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager_step(self._gX, self._gy)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._hloss = torch.zeros(1, dtype=torch.float32).pin_memory()    # update()'s loss lands here
        self._herr = torch.zeros(1, dtype=torch.int32).pin_memory()       # ... and the exchange's timeout word
        graph = torch.cuda.CUDAGraph()                  # step(): loss stays on the device
        with torch.cuda.graph(graph):
            self._gloss = self._eager_step(self._gX, self._gy)
        graph_u = torch.cuda.CUDAGraph()                # update(): the D2H read of the loss is a node of the graph
        with torch.cuda.graph(graph_u):
            self._gloss_u = self._eager_step(self._gX, self._gy)
            self._hloss.copy_(self._gloss_u.reshape(1), non_blocking=True)
            if getattr(self, "_dp_p2p", False):
                self._herr.copy_(self._p2p_flags[33:34], non_blocking=True)
        with torch.no_grad():
            for t, s0 in zip((fl["param"], st["exp_avg"], st["exp_avg_sq"], st["step"]), snap):
                t.copy_(s0)
            for b, s0 in zip(bufs, snap_b):
                b.copy_(s0)
            eng.graph_seed[1].zero_()
        self.train(was_training)
        self._graph, self._graph_u, self._graph_bs = graph, graph_u, batch_size

    def disable_cuda_graph(self):
        self._graph = self._graph_u = None
        self.model.engine.graph_seed = None

    def _graph_ready(self, X):
        return getattr(self, "_graph", None) is not None and X.shape[0] == self._graph_bs and self.training

    def step(self, X, y):
        """One optimisation step (forward -> MSE -> backward -> [all-reduce] -> Adam) with everything
        left on the device; returns the loss as a 0-dim device tensor.  X, y may be device tensors or
        (pinned) host tensors: with the graph enabled they are copied straight into its input buffers."""
        if self._graph_ready(X):
            self._gX.copy_(X, non_blocking=True)
            self._gy.copy_(y.reshape(self._gy.shape), non_blocking=True)
            self._graph.replay()
            return self._gloss
        return self._eager_step(X, y)

    def _eager_step(self, X, y):
        eng = self.model.engine
        dev = next(self.model.parameters()).device
        if X.device != dev:
            X, y = X.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
        loss = eng.loss_backward(X, y, zero_grad=True)
        if self._dp_world > 1 and not getattr(self, "_dp_p2p", False):
            import torch.distributed as dist
            dist.all_reduce(eng.flat["grad"], op=dist.ReduceOp.SUM, group=self._dp_group)
        self.optimizer.step()          # p2p: reads the peers' gradients itself
        return loss

    def update(self, X, y, epoch=None):
        if self._graph_ready(X):
            self._gX.copy_(X, non_blocking=True)
            self._gy.copy_(y.reshape(self._gy.shape), non_blocking=True)
            self._graph_u.replay()
            torch.cuda.current_stream(self._gX.device).synchronize()
            if int(self._herr[0]) != 0:
                self.check_exchange()
            return {"loss": float(self._hloss[0])}
        loss = self.step(X, y).item()
        if getattr(self, "_dp_p2p", False):
            self.check_exchange()
        return {"loss": loss}


class _ModelAlgorithm(Algorithm):
    """Common shape of the reference's other Algorithm subclasses (algorithms.py:139-163 and its clones):
    `self.model = <Model>(**configs)`, `torch.optim.Adam(lr, weight_decay)`, `update` = forward -> loss ->
    zero_grad -> backward -> step -> {'loss': loss.item()}.  Subclasses name the model class and may override
    `_loss`.

    `enable_cuda_graph(X, y)` captures that whole update (model forward with the library's kernels, autograd
    backward, Adam) for batches of X's shape into one CUDA graph; later `update` calls with that shape copy the
    batch into the graph's input buffers and replay it.  These models are small and launch-bound in eager mode, so
    the replay is what a training loop should use.  Parameters, buffers and optimizer state are snapshotted before the
    warm-up / capture and restored afterwards: enabling the graph does not change training."""

    MODEL = None            # (module name, class name)

    def __init__(self, configs, hparams, device):
        super().__init__(configs)
        import importlib
        mod, cls = self.MODEL
        self.model = getattr(importlib.import_module("." + mod, __package__), cls)(**configs)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=hparams["learning_rate"],
                                          weight_decay=hparams["weight_decay"])
        self.hparams = hparams
        self._graph = None
        self._dp_group, self._dp_world, self._dp_p2p = None, 1, False

    def _loss(self, X, y):
        return self.mse(self.model(X), y)

    def _eager_update(self, X, y):
        loss = self._loss(X, y)
        self.optimizer.zero_grad()
        loss.backward()
        if self._dp_world > 1 and not self._dp_p2p:
            import torch.distributed as dist
            self.optimizer.flat.gather_stray_grads()
            dist.all_reduce(self.optimizer.flat.grad, op=dist.ReduceOp.SUM, group=self._dp_group)
        self.optimizer.step()          # fused exchange: reads the peers' gradients itself
        return loss

    def update(self, X, y, epoch=None):
        if self._graph is not None and self.training and X.shape == self._gX.shape:
            self._gX.copy_(X, non_blocking=True)
            self._gy.copy_(y.reshape(self._gy.shape), non_blocking=True)
            self._graph.replay()
            loss = self._gloss.item()
        else:
            loss = self._eager_update(X, y).item()
        if self._dp_p2p and int(self._p2p_flags[33]) != 0:
            raise RuntimeError("stg_allreduce_adam: a peer did not publish its gradients in time; "
                               "the parameter update of that step was skipped on this rank")
        return {"loss": loss}

    # ------------------------------------------------------------------ flat one-kernel optimizer / data parallel
    def use_flat_optimizer(self, X, y):
        """Replaces torch.optim.Adam by the library's flat Adam (one launch over one buffer, flat_optim.FlatAdam;
        same update rule).  (X, y) is a sample batch: one probe backward finds the parameters that train at all --
        the reference keeps never-used modules (TemporalConvNet.net0 / net1, models/ST_GCN/Model.py:110-132) whose
        parameters torch's Adam never touches either.  Moments accumulated so far are carried over."""
        from .flat_optim import FlatAdam, FlatParams, find_used_parameters
        if isinstance(self.optimizer, FlatAdam):
            return self.optimizer
        was_training = self.training
        self.train()
        gen_state = torch.cuda.get_rng_state(X.device) if X.is_cuda else None
        used = find_used_parameters(self.model, lambda: self._loss(X, y))
        if gen_state is not None:
            torch.cuda.set_rng_state(gen_state, X.device)          # the probe must not shift the dropout stream
        self.train(was_training)
        old = self.optimizer
        hp = self.hparams
        opt = FlatAdam(FlatParams(used), lr=hp["learning_rate"], weight_decay=hp["weight_decay"])
        with torch.no_grad():
            for p in used:
                st = old.state.get(p)
                if st:
                    opt.state[p]["exp_avg"].copy_(st["exp_avg"])
                    opt.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                    opt.step_dev.fill_(int(st["step"]))
        self.optimizer = opt
        return opt

    def attach_data_parallel(self, X, y, group=None, broadcast=True, p2p="auto"):
        """Shard windows across ranks: parameters / buffers broadcast from rank 0, then per step ONE
        exchange of the flat gradient buffer -- fused into the optimizer kernel over NVLink symmetric memory
        (p2p True / "auto") or an NCCL all-reduce followed by the Adam kernel (p2p False).  BatchNorm statistics stay
        per rank.  (X, y): sample batch for use_flat_optimizer()."""
        import torch.distributed as dist
        self._dp_group, self._dp_world = group, dist.get_world_size(group)
        if broadcast:
            with torch.no_grad():
                for t in list(self.model.parameters()) + list(self.model.buffers()):
                    dist.broadcast(t, 0, group=group)
        opt = self.use_flat_optimizer(X, y)
        opt.grad_scale = 1.0 / self._dp_world
        self._dp_p2p = False
        if p2p and self._dp_world > 1 and dist.get_backend(group) == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                dev = opt.flat.param.device
                pg = group if group is not None else dist.group.WORLD
                gsym = symm_mem.empty(opt.flat.n, dtype=torch.float32, device=dev)
                hg = symm_mem.rendezvous(gsym, pg)
                flags = symm_mem.empty(64, dtype=torch.int32, device=dev)
                flags.zero_()
                flags[34] = 1            # epoch of the flag protocol: monotonic, never restored (stg_p2p.cu)
                hf = symm_mem.rendezvous(flags, pg)
                opt.flat.replace_grad_buffer(gsym)
                torch.cuda.synchronize(dev)
                dist.barrier(group=group)
                opt.attach_p2p(list(hg.buffer_ptrs), list(hf.buffer_ptrs), hg.rank, hg.world_size, (gsym, flags, hg, hf))
                self._p2p_flags = flags
                self._dp_p2p = True
            except Exception:
                if p2p is True:
                    raise

    def p2p_timed_out(self) -> bool:
        return bool(self._dp_p2p and int(self._p2p_flags[33]) != 0)

    def enable_cuda_graph(self, X, y):
        from .flat_optim import FlatAdam
        if isinstance(self.optimizer, FlatAdam):
            return self._capture_flat(X, y)
        params = [p for p in self.model.parameters()]
        dev = params[0].device
        hp = self.hparams
        old = self.optimizer.state_dict()
        # graph-capturable Adam (step counter on the device); carries over any state accumulated so far
        self.optimizer = torch.optim.Adam(params, lr=hp["learning_rate"], weight_decay=hp["weight_decay"], capturable=True)
        if old["state"]:
            for st in old["state"].values():
                if not torch.is_tensor(st["step"]) or st["step"].device != dev:
                    st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=dev)
            self.optimizer.load_state_dict(old)
            for grp in self.optimizer.param_groups:          # load_state_dict brought the old (eager) flag back
                grp["capturable"] = True
        self._gX, self._gy = X.detach().clone(), y.detach().clone()
        snap_m = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        had_state = bool(old["state"])
        snap_o = None
        if had_state:
            snap_o = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.optimizer.state[p].items()}
                      for p in params]
        was_training = self.training
        self.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.optimizer.zero_grad(set_to_none=True)
                self._loss(self._gX, self._gy).backward()
                self.optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            self._gloss = self._loss(self._gX, self._gy)
            self._gloss.backward()
            self.optimizer.step()
        with torch.no_grad():                       # undo the warm-up steps: weights, buffers, Adam moments
            sd = self.model.state_dict()
            for k, v in snap_m.items():
                sd[k].copy_(v)
            for i, p in enumerate(params):
                st = self.optimizer.state[p]
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(snap_o[i][k]) if had_state else v.zero_()
        self.train(was_training)
        self._graph = graph

    def _capture_flat(self, X, y):
        """Capture of the update with the flat optimizer (and, data parallel, the gradient exchange) as graph nodes."""
        opt = self.optimizer
        dev = opt.flat.param.device
        self._gX, self._gy = X.detach().clone(), y.detach().clone()
        snap_o = [t.clone() for t in opt.state_tensors()]
        snap_b = [(b, b.detach().clone()) for b in self.model.buffers()]
        unused = [(p, p.detach().clone()) for p in self.model.parameters() if p not in opt.state]
        was_training = self.training
        self.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._eager_update(self._gX, self._gy)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._gloss = self._eager_update(self._gX, self._gy)
        with torch.no_grad():                       # undo the warm-up steps: weights, buffers, Adam moments
            for t, v in zip(opt.state_tensors(), snap_o):
                t.copy_(v)
            for t, v in snap_b + unused:
                t.copy_(v)
        self.train(was_training)
        self._graph = graph

    def disable_cuda_graph(self):
        self._graph = None


class ASTGCNN(_ModelAlgorithm):
    """algorithms.py:139-163 around ASTGCNN_model (native TCN / adjacency / Chebyshev aggregation, astgcnn.py)."""
    MODEL = ("astgcnn", "ASTGCNN_model")


class ST_GCN(_ModelAlgorithm):
    """algorithms.py:465-491 around ST_GCN_model (native statistics / Pearson adjacency / aggregation / TCN, st_gcn.py)."""
    MODEL = ("st_gcn", "ST_GCN_model")


class STGNN(_ModelAlgorithm):
    """reference class STGNN around STGNN_model (stgnn.py)."""
    MODEL = ("stgnn", "STGNN_model")


class STMSGCN(_ModelAlgorithm):
    """reference class STMSGCN around STMSGCN_model (stgnn.py)."""
    MODEL = ("stgnn", "STMSGCN_model")


class SAGCN(_ModelAlgorithm):
    """reference class SAGCN around SAGCN_model (sagcn.py)."""
    MODEL = ("sagcn", "SAGCN_model")


class GAT_LSTM(_ModelAlgorithm):
    """reference class GAT_LSTM around GAT_LSTM_model (gat_lstm.py)."""
    MODEL = ("gat_lstm", "GAT_LSTM_model")


class HAGCN(_ModelAlgorithm):
    """algorithms.py:222-248: Adam + MSE + alpha * KL of the three SAGPool layers, around HAGCN_model (hagcn.py)."""
    MODEL = ("hagcn", "HAGCN_model")

    def __init__(self, configs, hparams, device):
        super().__init__(configs, hparams, device)
        self.alpha = hparams["alpha"]

    def _loss(self, X, y):
        pred, kl = self.model(X, train=True)
        return self.mse(pred, y) + self.alpha * kl


_ALGORITHMS = {"FC_STGNN": FC_STGNN, "ASTGCNN": ASTGCNN, "ST_GCN": ST_GCN, "STGNN": STGNN, "STMSGCN": STMSGCN,
               "GAT_LSTM": GAT_LSTM, "HAGCN": HAGCN, "SAGCN": SAGCN}
