"""Evaluation metrics of the reference (utils.py:136-201) on the device: one kernel reduces
Score_v1, Score_v2, |e| and e^2 over the predictions (stg_metrics), 32 bytes come back.
The reference loops over samples in Python every epoch (trainer.py:119-121)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


def calc_metrics(pred_labels: torch.Tensor, true_labels: torch.Tensor, max_rul: float):
    """_calc_metrics (utils.py:191-201): -> (Scores_v1, Scores_v2, MAE, RMSE) as Python floats."""
    s1, s2, sa, sq, n = _sums(pred_labels, true_labels, max_rul)
    return s1, s2 / n, sa / n * max_rul, math.sqrt(sq / n) * max_rul


def calc_metrics_aeroengine(pred_labels, true_labels, max_rul):
    """_calc_metrics_aeroengine (utils.py:171-178): -> (Scores, AvgScores, RMSE)."""
    s1, _, _, sq, n = _sums(pred_labels, true_labels, max_rul)
    return s1, s1 / n, math.sqrt(sq / n) * max_rul


def calc_metrics_bearing(pred_labels, true_labels, max_rul):
    """_calc_metrics_bearing (utils.py:180-189): -> (Scores_v2, MAE, RMSE)."""
    _, s2, sa, sq, n = _sums(pred_labels, true_labels, max_rul)
    return s2 / n, sa / n * max_rul, math.sqrt(sq / n) * max_rul


def _sums(pred, real, max_rul):
    lib = _lib.load()
    pred = pred.reshape(-1).contiguous()
    real = real.reshape(-1).contiguous()
    if not pred.is_cuda or not real.is_cuda:
        raise RuntimeError("metrics run on the device: pass CUDA tensors (no CPU fallback)")
    if pred.dtype != torch.float32 or real.dtype != torch.float32:
        raise TypeError("predictions and labels must be float32")
    if pred.numel() != real.numel() or pred.numel() == 0:
        raise ValueError("predictions and labels must be non-empty and of equal length")
    out = torch.zeros(4, device=pred.device, dtype=torch.float64)
    with torch.cuda.device(pred.device):
        _lib.check(lib.stg_metrics(pred.data_ptr(), real.data_ptr(), pred.numel(), float(max_rul), out.data_ptr(),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "stg_metrics")
    s1, s2, sa, sq = out.tolist()
    return s1, s2, sa, sq, pred.numel()
