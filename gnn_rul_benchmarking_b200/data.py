"""Device-resident mirror of the reference data path (dataloader/dataloader.py:13-94).

The reference indexes a CPU dataset sample by sample, collates, and copies every batch to the GPU
synchronously (trainer.py:108).  Once a training step takes < 0.5 ms that loop is the bottleneck, so
here the whole window set lives in HBM (C-MAPSS FD004 is ~140 MB) and a batch is one device-side
gather.  Same on-disk format, same channel-first fix-up, same shuffling stream as
`DataLoader(shuffle=True)`, so a run sees the batches the reference would see.
"""
from __future__ import annotations

import os
from typing import Dict, Iterator, Tuple, Union

import numpy as np
import torch

from .dp import shard_indices


class DeviceWindowDataset:
    """Load_Dataset (dataloader.py:13-57) with the tensors kept on `device` as float32."""

    def __init__(self, X, y, device, normalize: bool = False):
        X = torch.as_tensor(np.array(X))
        y = torch.as_tensor(np.array(y))
        if X.dim() < 3:
            X = X.unsqueeze(2)
        # make sure the channels are the second dim (dataloader.py:27-28)
        if X.shape.index(min(X.shape[1], X.shape[2])) != 1:
            X = X.permute(0, 2, 1)
        if y.dim() == 1:
            y = y.unsqueeze(-1)
        # `normalize` in the reference is Normalize(mean=0, std=1): the identity (dataloader.py:36-41)
        self.x_data = X.float().contiguous().to(device)
        self.y_data = y.float().contiguous().to(device)
        self.num_channels = self.x_data.shape[1]
        self.len = self.x_data.shape[0]

    def __len__(self):
        return self.len

    def __getitem__(self, index):
        return self.x_data[index], self.y_data[index]


class DeviceLoader:
    """Iterates (X, y) device batches.  shuffle=True draws the permutation exactly like
    torch.utils.data.RandomSampler (seed taken from the default CPU generator, fresh generator,
    randperm), so the batch order equals the reference DataLoader's for the same torch.manual_seed.
    rank/world shard every batch's window indices across data-parallel ranks (equal counts)."""

    def __init__(self, dataset: DeviceWindowDataset, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 rank: int = 0, world: int = 1):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.rank, self.world = rank, world

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def order(self) -> torch.Tensor:
        n = len(self.dataset)
        if not self.shuffle:
            return torch.arange(n)
        # DataLoader.__iter__ first draws its base seed from the default generator, then RandomSampler
        # draws the seed of the permutation generator: consume both, in that order
        torch.empty((), dtype=torch.int64).random_()
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        perm = self.order()
        n = perm.numel()
        dev = self.dataset.x_data.device
        for i in range(0, n, self.batch_size):
            idx = perm[i:i + self.batch_size]
            if idx.numel() < self.batch_size and self.drop_last:
                break
            if self.world > 1:
                idx = idx[list(shard_indices(idx.numel(), self.rank, self.world))]
                if idx.numel() == 0:
                    continue
            idx = idx.to(dev, non_blocking=True)
            yield self.dataset.x_data.index_select(0, idx), self.dataset.y_data.index_select(0, idx)


def data_generator(data_path, dataset_configs, hparams, device, rank: int = 0, world: int = 1):
    """dataloader.py:60-94 with device-resident loaders: -> (train_loader, test_loader | {key: loader}, max_RUL)."""
    train = torch.load(os.path.join(data_path, "train.pt"), weights_only=False)
    test = torch.load(os.path.join(data_path, "test.pt"), weights_only=False)
    bs = hparams["batch_size"]
    train_loader = DeviceLoader(DeviceWindowDataset(train["samples"], train["labels"], device, dataset_configs.normalize),
                                bs, shuffle=dataset_configs.shuffle, drop_last=dataset_configs.drop_last,
                                rank=rank, world=world)
    test_x, test_y = test["samples"], test["labels"]
    if isinstance(test_x, dict):
        test_loader: Union[DeviceLoader, Dict] = {
            k: DeviceLoader(DeviceWindowDataset(test_x[k], test_y[k], device, dataset_configs.normalize), bs,
                            shuffle=False, drop_last=dataset_configs.drop_last) for k in test_x}
    else:
        test_loader = DeviceLoader(DeviceWindowDataset(test_x, test_y, device, dataset_configs.normalize), bs,
                                   shuffle=False, drop_last=dataset_configs.drop_last)
    return train_loader, test_loader, train["max_ruls"]
