"""Batch assembly on the device.

Reference: a ``DataLoader`` over ``SplitDataset`` with PIL ``RandomCrop(32, padding=4)`` + ``RandomHorizontalFlip``
+ ``ToTensor`` + ``Normalize`` on the CPU, one sample at a time, then a host->device copy per step
(``src/data.py:19-27,113-133``).  Here a client's shard is a uint8 tensor already on the device and one fused op
gathers a batch, applies crop/flip/normalise and emits channels-last activations in the compute dtype.
"""
from __future__ import annotations

from typing import Optional

import torch

from .datasets import VisionData


class VisionAugment:
    """Crop (zero padding 4) + horizontal flip + normalise for CIFAR train batches; normalise only otherwise."""

    def __init__(self, ds: VisionData, train: bool, device, dtype):
        self.augment = bool(train and ds.data_name in ("CIFAR10", "CIFAR100"))
        self.pad = 4
        self.mean = torch.tensor(ds.mean, device=device, dtype=torch.float32).view(1, -1, 1, 1)
        self.inv_std = (1.0 / torch.tensor(ds.std, device=device, dtype=torch.float32)).view(1, -1, 1, 1)
        self.device, self.dtype = device, dtype

    def draw(self, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Random augmentation parameters ``[n, 3]`` int32 = (dy, dx, flip) -- drawn once per epoch, consumed by
        index inside captured graphs."""
        if not self.augment:
            return torch.zeros(n, 3, dtype=torch.int32, device=self.device)
        r = torch.empty(n, 3, dtype=torch.int32)
        r[:, 0] = torch.randint(0, 2 * self.pad + 1, (n,), generator=generator)
        r[:, 1] = torch.randint(0, 2 * self.pad + 1, (n,), generator=generator)
        r[:, 2] = torch.randint(0, 2, (n,), generator=generator)
        return r.to(self.device)

    def __call__(self, src_u8: torch.Tensor, idx: Optional[torch.Tensor] = None,
                 params: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Gather ``src_u8[idx]`` (``[n, H, W, C]`` uint8; all of it when ``idx`` is None), augment with
        ``params [B, 3]`` and return ``[B, C, H, W]`` (channels-last strides on GPU) normalised."""
        if src_u8.is_cuda:
            from ..ops import cuda as cu
            if cu.available():
                return cu.augment_batch(src_u8, idx, params if self.augment else None, self.mean, self.inv_std,
                                        self.pad, self.dtype)
        img_u8 = src_u8 if idx is None else src_u8[idx]
        B, H, W, C = img_u8.shape
        x = img_u8
        if self.augment and params is not None:
            p = self.pad
            xp = torch.zeros(B, H + 2 * p, W + 2 * p, C, dtype=img_u8.dtype, device=img_u8.device)
            xp[:, p:p + H, p:p + W] = img_u8
            ar_h = torch.arange(H, device=img_u8.device)
            ar_w = torch.arange(W, device=img_u8.device)
            rows = params[:, 0].long().view(B, 1) + ar_h.view(1, H)
            cols = params[:, 1].long().view(B, 1) + ar_w.view(1, W)
            flip = params[:, 2].bool().view(B, 1)
            cols = torch.where(flip, cols.flip(1), cols)
            bidx = torch.arange(B, device=img_u8.device).view(B, 1, 1)
            x = xp[bidx, rows.view(B, H, 1), cols.view(B, 1, W)]
        x = x.permute(0, 3, 1, 2).float().div_(255.0)
        x = (x - self.mean) * self.inv_std
        x = x.to(self.dtype)
        return x.contiguous(memory_format=torch.channels_last) if x.is_cuda else x


def lm_windows(T: int, bptt: int):
    """Window starts/lengths of ``BatchDataset`` (reference ``src/data.py:136-150``)."""
    return [(s, min(bptt, T - s)) for s in range(0, T, bptt)]
