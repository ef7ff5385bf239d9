"""Output side of the render path: the step AFTER `render_rays*` in the reference's Lightning systems.

    reference                                              here
    --------------------------------------------------     ------------------------------------------------------------
    LitModel.alter_gather_cat  models/interface.py:30-50    gather_images(): NCCL all-gather of every rank's ray range into frames
    LitModel.psnr_each         models/interface.py:53-61    psnr_each(): clipped squared error reduced on the GPU (neo_clipped_sq_err)
    store_image / store_depth_raw  models/utils.py:21-53    store_image() / store_depth_raw() (same file naming)
"""
from __future__ import annotations

import math
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import sharding


def psnr(pred: torch.Tensor, gt: torch.Tensor) -> float:
    """-10 log10(mean((clip(pred) - clip(gt))^2)); the reduction runs in the library's CUDA kernel (no CPU fallback)."""
    if not (pred.is_cuda and gt.is_cuda):
        raise RuntimeError("neo360_b200 needs CUDA tensors (no CPU fallback)")
    a, b = pred.contiguous().float(), gt.contiguous().float()
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    out = torch.zeros(1, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        L.check(L.load().neo_clipped_sq_err(a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    mse = float(out.item()) / a.numel()
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def psnr_each(preds: Sequence[torch.Tensor], gts: Sequence[torch.Tensor]) -> torch.Tensor:
    """models/interface.py:53-61"""
    return torch.tensor([psnr(p, g) for p, g in zip(preds, gts)])


def gather_images(local: torch.Tensor, image_sizes: Sequence[Tuple[int, int]], world: int, chunk: int, group=None) -> List[torch.Tensor]:
    """alter_gather_cat (models/interface.py:30-50): `local` holds this rank's rays (rows) of the concatenated frames, sharded with
    sharding.shard_range; returns the frames [(h,w,3) | (h,w)] on every rank.  world == 1 needs no process group."""
    n = sum(h * w for h, w in image_sizes)
    allr = local if world == 1 else sharding.gather_rays(local, n, world, chunk, group)
    if allr.dim() == 2 and allr.shape[-1] == 1:
        allr = allr.squeeze(-1)
    ret, cur = [], 0
    for (h, w) in image_sizes:
        ret.append(allr[cur:cur + h * w].reshape(h, w, 3) if allr.dim() == 2 else allr[cur:cur + h * w].reshape(h, w))
        cur += h * w
    return ret


def _to8b(x: np.ndarray) -> np.ndarray:
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def store_image(dirpath: str, rgbs: Sequence[torch.Tensor], name: str) -> List[str]:
    """models/utils.py:21-27: one `<name><idx:03d>.jpg` per frame (PPM when PIL is unavailable)."""
    os.makedirs(dirpath, exist_ok=True)
    paths = []
    for i, rgb in enumerate(rgbs):
        img = _to8b(rgb.detach().cpu().numpy())
        try:
            from PIL import Image
            path = os.path.join(dirpath, f"{name}{str(i).zfill(3)}.jpg")
            Image.fromarray(img).save(path)
        except ImportError:
            path = os.path.join(dirpath, f"{name}{str(i).zfill(3)}.ppm")
            with open(path, "wb") as f:
                f.write(b"P6 %d %d 255\n" % (img.shape[1], img.shape[0]))
                f.write(img.tobytes())
        paths.append(path)
    return paths


def store_depth_raw(dirpath: str, depths: Sequence[torch.Tensor], name: str) -> List[str]:
    """models/utils.py:45-53: compressed npz per frame."""
    os.makedirs(dirpath, exist_ok=True)
    paths = []
    for i, d in enumerate(depths):
        path = os.path.join(dirpath, f"{name}{str(i).zfill(3)}.npz")
        np.savez_compressed(path, d.detach().cpu().numpy())
        paths.append(path)
    return paths
