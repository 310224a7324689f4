"""Device-resident epoch loop around the fused update: what trainer.py:101-177 and
trainer.py:189-260 do per run -- train over the shuffled loader, evaluate every epoch, keep the best RMSE -- with the
window set, the predictions and the metric reductions all staying in HBM.

The reference moves every batch host -> device synchronously (trainer.py:108), calls `.item()` on every loss
(algorithms.py:76), appends predictions to numpy arrays batch by batch (trainer.py:150-151) and scores them in
per-sample Python loops (utils.py:136-169).  Here an epoch is: batches gathered on the device (data.DeviceLoader, same
shuffling stream as the reference DataLoader), one `Algorithm.step` per batch whose loss stays on the device, ONE
synchronisation at the end of the epoch for the running loss average, predictions written into one preallocated
device vector, and the four indicators from one reduction kernel (metrics.calc_metrics, stg_metrics).
Logging / csv / checkpoint writing stay with the caller (out of scope); `history` holds what
trainer.calc_results_per_run would have written.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch

from . import metrics as _metrics
from .data import DeviceLoader


class DeviceTrainer:
    def __init__(self, algorithm, train_loader: DeviceLoader, test_loader: Union[DeviceLoader, Dict], max_ruls,
                 num_epochs: int, use_cuda_graph: bool = True):
        self.algorithm, self.train_dl, self.test_dl, self.max_ruls = algorithm, train_loader, test_loader, max_ruls
        self.num_epochs = int(num_epochs)
        self.use_cuda_graph = use_cuda_graph
        # trainer.py:91-96: best_result = [[inf], [inf], [inf], [inf]] (per key for dict test sets)
        inf = float("inf")
        if isinstance(test_loader, dict):
            self.best_result = {k: [[inf], [inf], [inf], [inf]] for k in test_loader}
        else:
            self.best_result = [[inf], [inf], [inf], [inf]]
        self.history: List[dict] = []
        # AverageMeter of trainer.py:100: never reset between epochs (the logged loss is the average since the run began)
        self._loss_sum = None
        self._loss_cnt = 0
        self._graph_bs = None

    # ------------------------------------------------------------------------------------------ training
    def _step(self, X, y):
        alg = self.algorithm
        if hasattr(alg, "step"):                       # FC_STGNN: loss stays on the device
            return alg.step(X, y).reshape(())
        return torch.as_tensor(alg.update(X, y, 0)["loss"], device=X.device)

    def train_epoch(self, epoch: int) -> float:
        """One pass over the training loader; returns the running loss average the reference logs."""
        alg = self.algorithm
        alg.train()
        bs = self.train_dl.batch_size // max(1, self.train_dl.world)
        if self.use_cuda_graph and self._graph_bs is None and hasattr(alg, "step"):
            alg.enable_cuda_graph(bs)                  # batches of other sizes (the tail) run eagerly
            self._graph_bs = bs
        for X, y in self.train_dl:
            loss = self._step(X, y)
            w = float(X.shape[0])
            # the graph's loss tensor is overwritten by the next replay: accumulate right away (stream-ordered)
            self._loss_sum = loss * w if self._loss_sum is None else self._loss_sum + loss * w
            self._loss_cnt += X.shape[0]
        return float(self._loss_sum) / max(1, self._loss_cnt)          # the epoch's only synchronisation

    # ------------------------------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def test_base(self, loader: DeviceLoader):
        """trainer.py:134-153: -> (pred [n], true [n]) device vectors and the mean of the per-batch MSE losses."""
        model = self.algorithm.model
        n = len(loader.dataset)
        dev = loader.dataset.x_data.device
        pred, true = torch.empty(n, device=dev), torch.empty(n, device=dev)
        losses, o = [], 0
        for X, y in loader:
            p = model(X).view(-1)
            yv = y.view(-1)
            pred[o:o + p.numel()] = p
            true[o:o + p.numel()] = yv
            losses.append(torch.nn.functional.mse_loss(p, yv))
            o += p.numel()
        return pred[:o], true[:o], torch.stack(losses).mean()

    def test_prediction(self):
        """trainer.py:154-177."""
        self.algorithm.model.eval()
        if isinstance(self.test_dl, dict):
            out = {k: self.test_base(dl) for k, dl in self.test_dl.items()}
            self.pred_labels = {k: v[0] for k, v in out.items()}
            self.true_labels = {k: v[1] for k, v in out.items()}
            self.total_loss = {k: v[2] for k, v in out.items()}
        else:
            self.pred_labels, self.true_labels, self.total_loss = self.test_base(self.test_dl)

    def calc_results(self) -> dict:
        """trainer.py:189-260 without the file writes: the four indicators of this epoch; best_result grows when the RMSE
        improves (per key for dict test sets)."""
        def one(pred, true, max_rul, best):
            s1, s2, mae, rmse = _metrics.calc_metrics(pred, true, float(max_rul))
            if rmse < best[3][-1]:
                for lst, v in zip(best, (s1, s2, mae, rmse)):
                    lst.append(v)
            return dict(Score_v1=s1, Score_v2=s2, MAE=mae, RMSE=rmse)
        if isinstance(self.pred_labels, dict):
            return {k: one(self.pred_labels[k], self.true_labels[k], self.max_ruls[k], self.best_result[k])
                    for k in self.pred_labels}
        return one(self.pred_labels, self.true_labels, self.max_ruls, self.best_result)

    def fit(self, num_epochs: Optional[int] = None) -> List[dict]:
        for epoch in range(1, (num_epochs or self.num_epochs) + 1):
            avg = self.train_epoch(epoch)
            self.test_prediction()
            res = self.calc_results()
            self.history.append(dict(epoch=epoch, loss=avg, test=res))
        return self.history
