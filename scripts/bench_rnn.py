"""Native recurrence (rnn.LSTM / rnn.GRU, csrc/stg_rnn.cu) vs cuDNN nn.LSTM / nn.GRU at the reference call sites'
shapes: forward + backward per layer, CUDA events, and the three-layer HAGCN encoder as the reference stacks it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gnn_rul_benchmarking_b200 import rnn

torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
CASES = [("lstm", 5, 3584, 10, 60, True, "HAGCN bi_lstm1 [256x14 steps]"),
         ("lstm", 5, 3584, 60, 120, True, "HAGCN bi_lstm2"),
         ("lstm", 5, 3584, 120, 60, True, "HAGCN bi_lstm3"),
         ("lstm", 5, 21504, 60, 120, True, "HAGCN bi_lstm2, config 5 [1024x21 steps]"),
         ("lstm", 128, 40, 100, 30, False, "GAT_LSTM layer 1"),
         ("gru", 10240, 5, 64, 64, False, "STGNN N-CMAPSS"),
         ("gru", 128, 160, 97, 8, False, "STMSGCN")]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ONLY = os.environ.get("RNN_CASES")          # e.g. "0,1": restrict (ncu captures)
for ci, (cell, B, T, I, H, bi, what) in enumerate(CASES):
    if ONLY and str(ci) not in ONLY.split(","):
        continue
    x = torch.randn(B, T, I, device=dev, requires_grad=True)
    res = {}
    for tag, cls in (("cudnn", nn.LSTM if cell == "lstm" else nn.GRU), ("native", rnn.LSTM if cell == "lstm" else rnn.GRU)):
        m = cls(I, H, batch_first=True, bidirectional=bi).to(dev)

        def fb():
            out = m(x)[0]
            out.sum().backward()

        def f():
            with torch.no_grad():
                m(x)
        res[tag] = (timeit(f), timeit(fb))
    print(f"{what:45s} B={B:5d} T={T:5d} I={I:3d} H={H:3d}  fwd cudnn {res['cudnn'][0]:8.3f} ms native {res['native'][0]:8.3f} ms"
          f" | fwd+bwd cudnn {res['cudnn'][1]:8.3f} ms native {res['native'][1]:8.3f} ms"
          f" | native fwd {res['native'][0] / T * 1e3:6.3f} us/step", flush=True)
