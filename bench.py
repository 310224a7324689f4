#!/usr/bin/env python
"""bench.py -- windows/s of one FC_STGNN training step (forward + MSE + backward + Adam, i.e. one
`Algorithm.update`, algorithms/algorithms.py:67-76) on the BASELINE.json metric config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload S1|S2]

Workload S1 (default) = BASELINE.json configs[1] "FC_STGNN on CMAPSS FD004, batch=256, 1xB200" with the
reference's own FD004 hyper-parameters (configs/hparams.py:149-151: 14 sensors, 25 patches of 2,
C=16, H=8); S2 = the north_star synthetic block shape [B=256,T=50,N=21,C=14].  Synthetic U(0,1)
windows, random-init weights (no dataset ships with the reference).

One JSON line on stdout (rank 0):
  value     windows/s over all ranks, inputs already resident in HBM
  e2e       same step through the reference-facing API (FC_STGNN.update) with pinned HOST buffers:
            H2D of X,y and D2H of the loss inside the timed region
  roofline  dominant kernel of libstgconv_b200.so, timed with CUDA events on its stream during a
            second pass over the same steps (stg_profile_*), algorithmic bytes from DESIGN.md; block_kernels = HBM-side
            and compute-side fraction of both graph-conv block kernels
  extra     the same step on the north_star shape S2 (B=256) and on S1 at B=4096 (1 GPU only)
  eager_cuda_baseline  the reference's eager ATen path on this GPU (oracle port with tensors on cuda:0)
  dp_check  N>1: replicas bit-identical, fused NVLink exchange not timed out, fused == NCCL over 3 steps
  cpu_baseline  the CPU oracle port (oracle/fc_stgnn_oracle.py) on this box's host cores, bounded sample
`--impl reference` times that CPU port alone (all host threads) and prints the same line shape.
Timing: per-step CUDA-event pairs, L2 flushed (256 MiB memset) between steps outside the pairs,
barrier + synchronize on both sides of the loop, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "windows/s (fwd+bwd) FC_STGNN CMAPSS-FD004"
UNIT = "windows/s"
L2_FLUSH_BYTES = 256 << 20

WORKLOADS = {   # name -> (oracle.CONFIGS key, description)
    "S1": ("FD004", "FC_STGNN FD004 hparams: X[256,14,50] -> T=25,N=14,C=16,H=8"),
    "S2": ("S2", "FC_STGNN synthetic north_star shape: X[256,21,50] -> T=50,N=21,C=14,H=7"),
}
# configs/hparams.py:133 (FD004 train_params; batch overridden by BASELINE.json to 256)
HPARAMS = dict(num_epochs=81, batch_size=256, weight_decay=1e-4, learning_rate=1e-3)


def model_cfg(name):
    # kept in sync with oracle.fc_stgnn_oracle.CONFIGS by tests/test_host_logic.py
    from gnn_rul_benchmarking_b200.configs import CONFIGS
    return dict(CONFIGS[WORKLOADS[name][0]])


def algorithmic_bytes(cfg, B):
    """fp32 bytes the two fused graph-conv blocks must move per launch."""
    T, N, h = cfg["num_patch"], cfg["num_node"], cfg["hidden_dim"]
    C, H = 2 * h, h
    L = (T - 2) // 1 + 1 + (T - 2) // 2 + 1
    fwd = 4 * (T * N * C + L * N * H) * B                 # read x, write out1,out2
    bwd = 4 * (2 * T * N * C + L * N * H) * B             # read x, read dout, write dx
    return {"fwd": fwd, "bwd": bwd}


KERNEL_BYTES_KIND = {"k_block_fwd": "fwd", "k_block_bwd": "bwd"}


def algorithmic_flops(cfg, B):
    """FLOPs of the two blocks as the reference computes them: per graph 2MC^2 (mapping) + 2M^2C
    (Gram) + 2M^2C (A.X) + 2MCH (theta) + ~8M^2 (softmax / mask); backward ~ 2x forward."""
    T, N, h = cfg["num_patch"], cfg["num_node"], cfg["hidden_dim"]
    C, H, M = 2 * h, h, 2 * N
    per_graph = 2 * M * C * C + 4 * M * M * C + 2 * M * C * H + 8 * M * M
    L = (T - 2) // 1 + 1 + (T - 2) // 2 + 1
    fwd = per_graph * L * B
    return {"fwd": fwd, "bwd": 2 * fwd}


def measured_peak_tf32():
    """Dense TF32 tensor peak.  MEASURED_PEAKS.json holds the measured dense bf16 figure; TF32 runs at half the
    bf16 rate on this part (1.1 vs 2.25 PFLOP/s nominal), so the denominator is bf16_tflops / 2."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["bf16_tflops"]) / 2.0, "measured bf16 / 2"
    except Exception:
        return 1590.0 / 2.0, "fallback bf16 / 2"


def ncu_metric(kernel, key):
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(p) as fh:
            return json.load(fh)["kernels"][kernel].get(key)
    except Exception:
        return None


def kernel_rooflines(kern, cfg, B, peak):
    """HBM-side and compute-side fraction of each graph-conv block kernel from its live CUDA-event time."""
    ab, fl = algorithmic_bytes(cfg, B), algorithmic_flops(cfg, B)
    tpeak, tsrc = measured_peak_tf32()
    out = {}
    for name, kind in KERNEL_BYTES_KIND.items():
        if name not in kern:
            continue
        tot_ms, n = kern[name]
        avg_s = tot_ms / n / 1e3
        gbs = ab[kind] / avg_s / 1e9
        tfs = fl[kind] / avg_s / 1e12
        out[name] = {"avg_launch_us": round(avg_s * 1e6, 2), "algorithmic_bytes": ab[kind], "hbm_gbs": round(gbs, 1),
                     "hbm_frac": gbs / peak, "algorithmic_flops": fl[kind], "tflops": round(tfs, 3),
                     "tensor_frac": tfs / tpeak, "tensor_peak_tflops": tpeak, "tensor_peak_source": tsrc,
                     "sm__pipe_tensor_pct_ncu": ncu_metric(name, "pipe_tensor_pct")}
    return out


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/ncu_summary.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(p) as fh:
            k = json.load(fh)["kernels"][kernel]
        return k["dram_bytes_read"] + k["dram_bytes_write"]
    except Exception:
        return None


def measured_peak_gbs():
    """HBM GB/s from the driver-written MEASURED_PEAKS.json (any key naming hbm / copy bandwidth; the sustained
    figure when both are given -- the kernel is timed inside a long step), else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                doc = json.load(fh)
            found = []

            def walk(node, path):
                if isinstance(node, dict):
                    for k, v in node.items():
                        walk(v, path + "/" + str(k).lower())
                elif isinstance(node, (int, float)) and not isinstance(node, bool):
                    if ("hbm" in path or "copy" in path or "dram" in path) and 500.0 < float(node) < 20000.0:
                        found.append((path, float(node)))
            walk(doc, "")
            if found:
                sustained = [v for k, v in found if "sustain" in k]
                return (sustained[0] if sustained else found[0][1]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self._ready = threading.Event()       # set after the first sample: NVML initialisation can take 100s of ms

    def _run_nvml(self):
        """NVML directly (a sample every few ms: the timed region is only tens of ms long)."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = (("hw_slowdown", int(getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8))),
                ("hw_thermal_slowdown", int(getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40))),
                ("sw_thermal_slowdown", int(getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20))),
                ("sw_power_cap", int(getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4))))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
            self._ready.set()
            self._stop.wait(0.002)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._ready.set()
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        self._ready.wait(timeout=5.0)         # the loop it brackets is only tens of ms long: sample from its first step
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(cfg, B, steps, warmup, budget_s=None):
    """The CPU oracle port's update (fwd + MSE + bwd + Adam) on this box's host cores.
    Returns (ms_per_step, steps_done, threads)."""
    from oracle import fc_stgnn_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    alg = orc.OracleAlgorithm(cfg, HPARAMS, seed=0)
    X, y = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"]), torch.rand(B, 1)
    for _ in range(warmup):
        alg.update(X, y)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        alg.update(X, y)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 3:
            break
    dt = time.perf_counter() - t0
    return 1e3 * dt / done, done, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_cfg(args.workload)
    B = args.batch * max(1, args.gpus)          # the native arm's GLOBAL batch at this N (weak scaling: per-GPU batch fixed)
    ms, done, threads = cpu_reference_steps(cfg, B, args.steps, args.warmup, budget_s=150.0)   # bounded: a few minutes
    val = B / (ms / 1e3)
    sample = f"{done} update steps of B={B} ({args.workload}) on {threads} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][1]}", "global_batch": B,
                   "per_gpu_batch": args.batch, "step": "fwd+mse+bwd+adam", "device": "cpu"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def time_config(workload, B, dev, steps, warmup, flush):
    """value / per-kernel times of one more (workload, batch) on a single GPU: the `extra` block of the line."""
    from gnn_rul_benchmarking_b200 import _lib
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    cfg = model_cfg(workload)
    torch.manual_seed(0)
    alg = get_algorithm_class("FC_STGNN")(cfg, HPARAMS, dev).to(dev)
    alg.train()
    alg.enable_cuda_graph(B)
    g = torch.Generator().manual_seed(99)
    X = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], generator=g).to(dev)
    y = torch.rand(B, 1, generator=g).to(dev)
    for _ in range(warmup):
        alg.step(X, y)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        alg.step(X, y)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    alg.disable_cuda_graph()
    with _lib.kernel_profile() as prof:
        for _ in range(steps):
            flush.zero_()
            alg.step(X, y)
        torch.cuda.synchronize()
    kern = prof.result()
    peak, _ = measured_peak_gbs()
    return {"workload": f"{workload}: {WORKLOADS[workload][1]}", "batch": B, "value": B / (ms / 1e3), "unit": UNIT,
            "ms_per_step": ms, "steps": steps, "kernels_us": {k: round(1e3 * t / n, 2) for k, (t, n) in kern.items()},
            "block_kernels": kernel_rooflines(kern, cfg, B, peak)}


def eager_cuda_baseline(cfg, B, dev, steps=10, warmup=3):
    """The reference's own eager-ATen path on this GPU (main.py:34 defaults to cuda:0): the oracle port's update with
    every tensor on the device -- cuBLAS bmm, cuDNN conv / BatchNorm, ~500 launches per step."""
    from oracle import fc_stgnn_oracle as orc
    torch.manual_seed(0)
    alg = orc.OracleAlgorithm(cfg, HPARAMS, seed=0, device=dev)
    X = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], device=dev)
    y = torch.rand(B, 1, device=dev)
    for _ in range(warmup):
        alg.update(X, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        alg.update(X, y)            # .item() on the loss synchronises every step, as the reference does
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    return {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "what": "oracle port of algorithms.py:67-76 with all tensors on cuda:0 (eager ATen: cuBLAS / cuDNN), wall clock"}


def dp_check(alg, dev, world, rank, cfg, B):
    """Correctness evidence for the multi-GPU line: replicas bit-identical after the timed loop, the fused NVLink
    exchange never timed out, and 3 steps of the fused exchange equal 3 steps of NCCL all-reduce + Adam."""
    import torch.distributed as dist
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    flat = alg.model.engine.flat["param"]
    mine = flat.double().sum().reshape(1)
    mine2 = (flat.double() * torch.arange(1, flat.numel() + 1, device=dev, dtype=torch.float64)).sum().reshape(1)
    both = torch.cat([mine, mine2])
    gathered = [torch.zeros_like(both) for _ in range(world)]
    dist.all_gather(gathered, both)
    identical = all(torch.equal(gathered[0], t) for t in gathered)
    timed_out = bool(alg.p2p_timed_out()) if getattr(alg, "_dp_p2p", False) else False
    # fused exchange vs NCCL on fresh replicas and identical data
    res = []
    for p2p in ("auto", False):
        torch.manual_seed(7)
        a2 = get_algorithm_class("FC_STGNN")(cfg, HPARAMS, dev).to(dev)
        a2.model.positional_encoding.dropout.p = 0.0
        a2.train()
        a2.attach_data_parallel(p2p=p2p)
        g = torch.Generator().manual_seed(4321 + rank)
        for it in range(3):
            X = torch.rand(B, cfg["num_node"], cfg["num_patch"] * cfg["patch_size"], generator=g).to(dev)
            y = torch.rand(B, 1, generator=g).to(dev)
            a2.update(X, y, it)
        res.append((a2.model.engine.flat["param"].clone(), bool(getattr(a2, "_dp_p2p", False))))
    dmax = float((res[0][0] - res[1][0]).abs().max())
    return {"replicas_bit_identical": bool(identical), "p2p_timed_out": timed_out,
            "fused_vs_nccl_max_abs_dparam_3_steps": dmax, "fused_exchange_used": res[0][1]}


def run_native(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the native arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from gnn_rul_benchmarking_b200 import _lib
    from gnn_rul_benchmarking_b200.algorithms import get_algorithm_class
    _lib.load()

    cfg = model_cfg(args.workload)
    B = args.batch                                   # per-GPU batch (weak scaling)
    L_in = cfg["num_patch"] * cfg["patch_size"]
    torch.manual_seed(0)
    alg = get_algorithm_class("FC_STGNN")(cfg, HPARAMS, dev).to(dev)
    alg.train()
    if world > 1:
        alg.attach_data_parallel(p2p=False if args.nccl else "auto")   # gradient exchange fused into the Adam kernel
    if not args.no_graph:
        alg.enable_cuda_graph(B)        # the step's launches replayed from one CUDA graph

    g = torch.Generator().manual_seed(1234 + rank)
    nbuf = 4
    Xh = [torch.rand(B, cfg["num_node"], L_in, generator=g).pin_memory() for _ in range(nbuf)]
    yh = [torch.rand(B, 1, generator=g).pin_memory() for _ in range(nbuf)]
    Xd = [t.to(dev) for t in Xh]
    yd = [t.to(dev) for t in yh]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        return alg.step(Xd[i % nbuf], yd[i % nbuf])

    def step_e2e(i):
        # the reference-facing call with HOST buffers: update() copies them to the device (H2D), runs the
        # step and reads the loss back (D2H) before returning
        return alg.update(Xh[i % nbuf], yh[i % nbuf], 1)["loss"]

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # the sampler thread starts with the warm-up (NVML initialisation takes longer than a short timed region): every
    # sample is taken while this rank's GPU runs the step
    with ClockSampler(local) as clk:
        for i in range(args.warmup):
            step_resident(i)
        for i in range(max(3, args.warmup // 2)):
            step_e2e(i)
        ms_res = timed(step_resident, args.steps)
        ms_e2e = timed(step_e2e, args.steps)
    # second pass: same steps with per-kernel events of our library (roofline of the dominant kernel)
    alg.disable_cuda_graph()            # event pairs cannot be recorded inside a graph replay
    with _lib.kernel_profile() as prof:
        timed(step_resident, args.steps)
    kern = prof.result()
    dpc = dp_check(alg, dev, world, rank, cfg, B) if world > 1 else None
    if world > 1 and (dpc["p2p_timed_out"] or not dpc["replicas_bit_identical"]):
        raise SystemExit(f"bench.py: data-parallel replicas diverged / exchange timed out: {dpc}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = measured_peak_gbs()
    ab = algorithmic_bytes(cfg, B)
    roof = None
    launches = sum(n for _, n in kern.values())
    cand = {k: v for k, v in kern.items() if k in KERNEL_BYTES_KIND}
    if cand:
        name = max(cand, key=lambda k: cand[k][0])
        tot_ms, n = cand[name]
        avg_s = tot_ms / n / 1e3
        nbytes = ab[KERNEL_BYTES_KIND[name]]
        achieved = nbytes / avg_s / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "peak_source": peak_kind,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(name) if args.workload == "S1" and B == 256 else None,
                "algorithmic_bytes": nbytes,
                "avg_launch_us": avg_s * 1e6,
                "share_of_lib_time": tot_ms / max(1e-9, sum(t for t, _ in kern.values())),
                "kernels_us": {k: round(1e3 * t / n2, 2) for k, (t, n2) in kern.items()},
                "block_kernels": kernel_rooflines(kern, cfg, B, peak)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ms_cpu, done, threads = cpu_reference_steps(cfg, B, 10_000, 2, budget_s=args.cpu_budget)
        cpu = {"value": B / (ms_cpu / 1e3), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{done} update steps of B={B} ({args.workload}), {ms_cpu:.1f} ms/step, oracle port"}
    extra, eager = None, None
    if world == 1 and not args.no_extra:
        extra = []
        for wl, bb in (("S2", 256), ("S1", 4096)):
            if (wl, bb) != (args.workload, B):
                extra.append(time_config(wl, bb, dev, max(5, min(args.steps, 20)), 5, flush))
        eager = eager_cuda_baseline(cfg, B, dev)
    h2d = Xh[0].numel() * 4 + yh[0].numel() * 4
    line = {
        "metric": METRIC, "value": world * B / (ms_res / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][1]}", "global_batch": world * B,
                   "per_gpu_batch": B, "step": "fwd+mse+bwd+adam", "parallelism": f"dp{world}",
                   "cuda_graph": not args.no_graph,
                   "grad_exchange": ("nvlink-p2p fused with adam" if getattr(alg, "_dp_p2p", False) else "nccl all-reduce") if world > 1 else None,
                   "l2": "flushed between steps (256 MiB memset outside the event pairs)"},
        "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(round(launches / args.steps)) * args.steps,
        "roofline": roof, "cpu_baseline": cpu, "clocks": clk.summary(),
        "extra": extra, "eager_cuda_baseline": eager, "dp_check": dpc,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="S1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=256, help="windows per GPU per step")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra S2 / B=4096 / eager-CUDA measurements")
    ap.add_argument("--nccl", action="store_true", help="N>1: NCCL all-reduce + Adam instead of the fused NVLink kernel")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
