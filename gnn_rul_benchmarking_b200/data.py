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


# ------------------------------------------------------------------------------------------ flat on-disk format
# The reference stores a split as a pickled dict of Python lists of numpy arrays (Data_read_CMAPSS.py:323-324), which
# needs torch.load(weights_only=False) -- arbitrary code execution on load -- and one Python object per window.  The
# flat format is a header plus raw little-endian arrays: nothing is unpickled, and a split maps straight into one
# contiguous float32 tensor that goes to the device in a single copy.
#   bytes 0..7   b"STGW1\0\0\0"
#   bytes 8..15  uint64 header length h
#   bytes 16..   h bytes of JSON: {"max_ruls": number | [[key, number], ...],
#                                  "segments": [{"name": "samples" | "labels", "key": null | str | number,
#                                                "shape": [...], "offset": byte offset from the data start}]}
#   data start = 16 + h rounded up to 64; every segment is float32, C order, 64-byte aligned.
_MAGIC = b"STGW1\0\0\0"


def _as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def save_flat(path: str, split: dict) -> None:
    """split: the reference's dict {"samples": array-like | {key: array-like}, "labels": ..., "max_ruls": ...}."""
    import json
    segs, blobs, off = [], [], 0

    def add(name, key, arr):
        nonlocal off
        arr = _as_f32(arr)
        segs.append(dict(name=name, key=key, shape=list(arr.shape), offset=off))
        blobs.append(arr)
        off += (arr.nbytes + 63) // 64 * 64

    def jkey(k):
        return k if isinstance(k, str) else float(k)

    for name in ("samples", "labels"):
        v = split[name]
        if isinstance(v, dict):
            for k in v:
                add(name, jkey(k), v[k])
        else:
            add(name, None, v)
    mr = split.get("max_ruls")
    mr = [[jkey(k), float(v)] for k, v in mr.items()] if isinstance(mr, dict) else (None if mr is None else float(mr))
    head = json.dumps(dict(max_ruls=mr, segments=segs)).encode()
    start = (16 + len(head) + 63) // 64 * 64
    with open(path, "wb") as fh:
        fh.write(_MAGIC)
        fh.write(np.uint64(len(head)).tobytes())
        fh.write(head)
        fh.write(b"\0" * (start - 16 - len(head)))
        for seg, arr in zip(segs, blobs):
            assert fh.tell() == start + seg["offset"]
            fh.write(arr.tobytes())
            fh.write(b"\0" * ((arr.nbytes + 63) // 64 * 64 - arr.nbytes))


def load_flat(path: str) -> dict:
    """-> the same dict layout torch.load gives for the reference's .pt files, arrays as read-only memory maps."""
    import json
    with open(path, "rb") as fh:
        if fh.read(8) != _MAGIC:
            raise ValueError(f"{path}: not a flat window file")
        hlen = int(np.frombuffer(fh.read(8), dtype=np.uint64)[0])
        head = json.loads(fh.read(hlen).decode())
    start = (16 + hlen + 63) // 64 * 64
    size = os.path.getsize(path)
    out: dict = {}
    for seg in head["segments"]:
        n = int(np.prod(seg["shape"])) if seg["shape"] else 1
        if start + seg["offset"] + 4 * n > size:
            raise ValueError(f"{path}: segment {seg['name']} runs past the end of the file")
        arr = np.memmap(path, dtype="<f4", mode="r", offset=start + seg["offset"], shape=tuple(seg["shape"]))
        if seg["key"] is None:
            out[seg["name"]] = arr
        else:
            out.setdefault(seg["name"], {})[seg["key"]] = arr
    mr = head["max_ruls"]
    out["max_ruls"] = {k: v for k, v in mr} if isinstance(mr, list) else mr
    return out


def convert_pt(pt_path: str, flat_path: str) -> None:
    """One-off conversion of a reference split (train.pt / test.pt) to the flat format."""
    save_flat(flat_path, torch.load(pt_path, weights_only=False))


def _load_split(data_path: str, name: str) -> dict:
    flat = os.path.join(data_path, name + ".stgw")
    if os.path.exists(flat):
        return load_flat(flat)
    return torch.load(os.path.join(data_path, name + ".pt"), weights_only=False)


def data_generator(data_path, dataset_configs, hparams, device, rank: int = 0, world: int = 1):
    """dataloader.py:60-94 with device-resident loaders: -> (train_loader, test_loader | {key: loader}, max_RUL).
    Reads train.stgw / test.stgw (flat format above) when present, else the reference's train.pt / test.pt."""
    train = _load_split(data_path, "train")
    test = _load_split(data_path, "test")
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
