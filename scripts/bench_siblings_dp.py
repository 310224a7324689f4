"""BASELINE.json configs[2] (ST_GCN + ASTGCNN, N-CMAPSS shape, batch 512 per GPU) and configs[4] (HAGCN + STMSGCN on the
synthetic [.,T=50,N=21] shape, global batch 1024 at 8 GPUs = 128 per GPU) as data-parallel training steps:
    python scripts/bench_siblings_dp.py                                  (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_siblings_dp.py
One JSON line per model from rank 0: global windows/s (weak scaling, per-GPU batch fixed), ms per step = max over ranks
of CUDA-event time, the exchange used, replicas' parameter checksums equal.  The update is the reference's
(algorithms.py:139-163: forward, MSE (+ alpha KL for HAGCN), backward, Adam) captured in one CUDA graph per rank with the
flat gradient exchange fused into the Adam kernel over NVLink (flat_optim.FlatAdam / stg_allreduce_adam)."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
from gnn_rul_benchmarking_b200.configs import ASTGCNN_CONFIGS, TRAIN_PARAMS

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
STEPS, WARM = int(os.environ.get("STEPS", "20")), 5
CASES = [
    ("ASTGCNN", ASTGCNN_CONFIGS["NCMAPSS"], (512, 20, 50), "configs[2]"),
    ("ST_GCN", dict(num_patch=20, patch_size=50, dropout=0.2), (512, 20, 50), "configs[2] (sensor-as-patch, SURVEY 8d)"),
    ("HAGCN", dict(patch_size=10, num_patch=5, encoder_hidden_dim=60, hidden_dim=64, output_dim=32), (128, 21, 50), "configs[4]"),
    ("STMSGCN", dict(num_patch=21, patch_size=50, interval=5, band_width=5, gcn_dims=[16, 64, 16, 1], gru_hidden_dim=8),
     (128, 1050), "configs[4] (sensor-as-patch, SURVEY 8d)"),
]
for name, cfg, shape, what in CASES:
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alg = get_algorithm_class(name)(cfg, dict(TRAIN_PARAMS, alpha=100), dev).to(dev)
    alg.train()
    g = torch.Generator().manual_seed(100 + rank)
    X, y = torch.rand(*shape, generator=g).to(dev), torch.rand(shape[0], 1, generator=g).to(dev)
    if world > 1:
        alg.attach_data_parallel(X, y, p2p="auto" if os.environ.get("NCCL_ONLY") != "1" else False)
    else:
        alg.use_flat_optimizer(X, y)
    graph = True
    try:
        alg.enable_cuda_graph(X, y)
    except Exception as e:
        graph = False
        torch.cuda.synchronize()
    for _ in range(WARM):
        alg.update(X, y, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for e0, e1 in evs:
        e0.record()
        alg.update(X, y, 0)
        e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs) / STEPS], device=dev)
    csum = alg.optimizer.flat.param.double().sum().reshape(1)
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        allc = [torch.zeros_like(csum) for _ in range(world)]
        dist.all_gather(allc, csum)
        same = all(bool(torch.equal(c, allc[0])) for c in allc)
    if rank == 0:
        print(json.dumps({"model": name, "config": what, "n_gpus": world, "per_gpu_batch": shape[0], "x_shape": list(shape),
                          "ms_per_step": round(float(ms), 4), "value": round(shape[0] * world / float(ms) * 1e3, 1),
                          "unit": "windows/s", "scaling": "weak", "cuda_graph": graph,
                          "exchange": None if world == 1 else ("fused NVLink" if alg._dp_p2p else "nccl"),
                          "replicas_identical": same, "timed_out": alg.p2p_timed_out()}), flush=True)
    del alg
if world > 1:
    dist.destroy_process_group()
