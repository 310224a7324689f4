"""2-GPU data parallel (NCCL): windows sharded across ranks, ONE all-reduce of the flat gradient
buffer per step; replicas stay bit-identical and the first step equals the single-GPU step on the
concatenated batch when BatchNorm sees the same statistics (eval of the averaged gradient)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, p2p=False, graph=False):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    from gnn_rul_benchmarking_b200.configs import CONFIGS, TRAIN_PARAMS
    torch.manual_seed(50 + rank)                       # different init per rank: the broadcast must fix it
    alg = get_algorithm_class("FC_STGNN")(CONFIGS["FD004"], TRAIN_PARAMS, dev).to(dev)
    alg.model.positional_encoding.dropout.p = 0.0
    alg.train()
    alg.attach_data_parallel(p2p=p2p)
    assert alg._dp_p2p == bool(p2p)
    if graph:          # the order bench.py uses: exchange attached first, then the step graph captured
        alg.enable_cuda_graph(8)
    g = torch.Generator().manual_seed(3)
    X, y = torch.rand(16, 14, 50, generator=g), torch.rand(16, 1, generator=g)
    losses = []
    for it in range(3):
        out = alg.update(X[rank::world].contiguous().to(dev), y[rank::world].contiguous().to(dev), it)
        losses.append(out["loss"])
    flat = alg.model.engine.flat["param"].clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert not alg.p2p_timed_out()
    if rank == 0:
        q.put((losses, [t.cpu() for t in gathered]))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_replicas_stay_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    losses, params = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(l == l for l in losses)                 # finite
    assert torch.equal(params[0], params[1])           # same averaged gradient -> same Adam update


def _params_agree(a, b, steps=3, lr=1e-3):
    """Two runs that differ only by the order of floating-point atomics: nearly every parameter agrees to a few ulp;
    where a gradient entry is itself at the noise floor Adam (m / sqrt(v) ~ sign(g)) may move it by up to lr per step."""
    d = (a - b).abs()
    assert float(d.max()) <= steps * lr * 1.05, float(d.max())
    assert float((d > 2e-6 + 1e-4 * b.abs()).float().mean()) <= 0.02


def _run(p2p, graph=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + (7 if p2p else 0) + (13 if graph else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, p2p, graph)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_fused_nvlink_exchange_matches_nccl():
    """stg_allreduce_adam (peers' gradients read over NVLink inside the Adam kernel) == NCCL all-reduce +
    Adam.  Replicas of one run are bit-identical (same sums in the same rank order); two separate runs
    agree to float-atomic summation noise of the forward/backward kernels."""
    l_nccl, p_nccl = _run(False)
    l_p2p, p_p2p = _run(True)
    assert torch.equal(p_p2p[0], p_p2p[1])             # replicas identical
    assert all(abs(a - b) < 1e-5 for a, b in zip(l_nccl, l_p2p))
    _params_agree(p_nccl[0], p_p2p[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_fused_exchange_inside_cuda_graph_keeps_replicas_identical():
    """attach_data_parallel(p2p) followed by enable_cuda_graph (bench.py's order): the capture's warm-up steps are
    rolled back, the flag protocol's epoch is not -- the replayed steps must still wait for their peers.  Replicas
    stay bit-identical, nobody times out, and the result equals the eager fused exchange."""
    l_e, p_e = _run(True, graph=False)
    l_g, p_g = _run(True, graph=True)
    assert torch.equal(p_g[0], p_g[1])
    assert all(abs(a - b) < 1e-5 for a, b in zip(l_e, l_g))
    _params_agree(p_e[0], p_g[0])


def _sibling_worker(rank, world, port, q, name, cfg, shape, p2p, graph):
    import warnings
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    from gnn_rul_benchmarking_b200.configs import TRAIN_PARAMS
    torch.manual_seed(70 + rank)                       # different init per rank: the broadcast must fix it
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alg = get_algorithm_class(name)(cfg, dict(TRAIN_PARAMS, alpha=100), dev).to(dev)
    alg.train()
    g = torch.Generator().manual_seed(5)
    X, y = torch.rand(*shape, generator=g), torch.rand(shape[0], 1, generator=g)
    Xr, yr = X[rank::world].contiguous().to(dev), y[rank::world].contiguous().to(dev)
    alg.attach_data_parallel(Xr, yr, p2p=p2p)
    assert alg._dp_p2p == bool(p2p)
    if graph:
        alg.enable_cuda_graph(Xr, yr)
    losses = [alg.update(Xr, yr, it)["loss"] for it in range(3)]
    flat = alg.optimizer.flat.param.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert not alg.p2p_timed_out()
    if rank == 0:
        q.put((losses, [t.cpu() for t in gathered]))
    dist.destroy_process_group()


def _run_sibling(name, cfg, shape, p2p, graph=False, salt=0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 1500) + salt
    procs = [ctx.Process(target=_sibling_worker, args=(r, 2, port, q, name, cfg, shape, p2p, graph)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sibling_models_data_parallel_fused_exchange_matches_nccl():
    """BASELINE configs[2] model ST_GCN (dropout off; it keeps never-used TemporalConvNet parameters, which stay out
    of the flat exchange) sharded over 2 GPUs: the fused NVLink exchange + Adam kernel, eager and inside the captured
    update, against NCCL all-reduce + Adam.  Replicas bit-identical."""
    cfg, shape = dict(num_patch=20, patch_size=50, dropout=0.0), (32, 20, 50)
    l_n, p_n = _run_sibling("ST_GCN", cfg, shape, False, salt=1)
    l_p, p_p = _run_sibling("ST_GCN", cfg, shape, True, salt=2)
    l_g, p_g = _run_sibling("ST_GCN", cfg, shape, True, graph=True, salt=3)
    for ps in (p_n, p_p, p_g):
        assert torch.equal(ps[0], ps[1])
    assert all(abs(a - b) < 1e-5 * (abs(a) + 1e-6) for a, b in zip(l_n, l_p))
    assert all(abs(a - b) < 1e-5 * (abs(a) + 1e-6) for a, b in zip(l_n, l_g))
    _params_agree(p_n[0], p_p[0])
    _params_agree(p_n[0], p_g[0])
