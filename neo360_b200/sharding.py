"""Ray sharding across ranks.

The path shards embarrassingly by rays; the only cross-ray coupling in the reference is quirk Q1 (a ray is conditioned on the
view direction of another ray of its own `chunk`), so shards must start on chunk boundaries.  No data-path collective is needed:
every rank renders its range with the replicated scene, and the caller gathers pixels at the end (the reference's
`alter_gather_cat`, models/interface.py:30-50)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def shard_range(n_rays: int, world: int, rank: int, chunk: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rays for `rank`: ceil(n_chunks / world) whole chunks per rank (the tail ranks may be empty)."""
    if chunk <= 0:
        chunk = n_rays
    n_chunks = (n_rays + chunk - 1) // chunk
    per = (n_chunks + world - 1) // world
    start = min(rank * per * chunk, n_rays)
    stop = min((rank + 1) * per * chunk, n_rays)
    return start, stop


def shard_batch(batch: Dict[str, torch.Tensor], world: int, rank: int, chunk: int) -> Tuple[Dict[str, torch.Tensor], Tuple[int, int]]:
    """Slice every per-ray tensor of a reference batch dict; the `src_*` / scene entries are replicated (model.py:827-838)."""
    n = batch["rays_o"].shape[0]
    a, b = shard_range(n, world, rank, chunk)
    out = {}
    for k, v in batch.items():
        per_ray = torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == n and not k.startswith("src_") and k not in (
            "planes_xz", "planes_xy", "planes_yz", "latent")
        out[k] = v[a:b] if per_ray else v
    return out, (a, b)


def gather_rays(local: torch.Tensor, n_rays: int, world: int, chunk: int, group=None) -> torch.Tensor:
    """All-gather per-rank results (rows = rays of that rank's shard) back into frame order."""
    import torch.distributed as dist
    sizes = [shard_range(n_rays, world, r, chunk) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], 0)
