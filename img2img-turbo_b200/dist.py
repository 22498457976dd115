"""Multi-GPU data parallelism for the path: images are independent units (no BatchNorm, per-sample GroupNorm /
LayerNorm / attention), so a batch shards contiguously over ranks with NO data-path collective, and the only exchange
is one all-gather of the outputs (BASELINE.json north_star).  One process per GPU, torch.distributed
(NCCL over NVLink on the GPU box, gloo in CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split [lo, hi) of `batch` images for `rank`; the first (batch % world) ranks get one extra image."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def all_gather_outputs(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """One collective: every rank ends up with the full [batch, ...] output.  Even splits use all_gather_into_tensor
    (a single ncclAllGather); ragged splits pad to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if batch % world == 0:
        out = torch.empty((batch,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: sizes[rank]] = local
    buf = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)


def sharded_forward(model, x: torch.Tensor, *args, group=None, **kwargs) -> torch.Tensor:
    """Run `model` (Pix2Pix_Turbo / CycleGAN_Turbo) on this rank's shard of the batch and all-gather the images.
    Per-sample tensors in kwargs (eps, noise_map) are sharded the same way."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = x.shape[0]
    kw = {k: (shard(v, rank, world) if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == B else v)
          for k, v in kwargs.items()}
    y = model(shard(x, rank, world), *args, **kw)
    return all_gather_outputs(y, B, group)
