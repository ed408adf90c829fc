"""Batch sharding across the GPUs of one node: one process per GPU, full weight replica per
GPU, the image batch split contiguously, NO collective on the data path.  The only communication is the optional
gather of the per-image outputs ((3 + N) * 4 bytes per image) over NCCL / NVSwitch, off the critical path.

The reference has no multi-device code at all (grep for distributed|DataParallel|nccl in /root/reference -> none);
this module is the B200 addition around the same `forward_with_guidance` surface.

RNG note: in guided mode the reference draws `randn(B,192)`, `randn(B,768)` and a fresh
`nn.Linear(768,64)` from the global CPU generator on every call (src/model.py:609,744,1421).  For a sharded run to
return exactly what the un-sharded run returns, every rank must replay those draws with the GLOBAL batch size under
the same seed; `ShardedInference` sets `model.rng_replay_batch` accordingly.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch


def shard_bounds(n: int, world: int) -> Sequence[Tuple[int, int]]:
    """Contiguous, balanced split of n images over `world` ranks: the first n % world ranks get one extra image."""
    if world <= 0:
        raise ValueError("world must be positive")
    if n < 0:
        raise ValueError("n must be non-negative")
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return shard_bounds(n, world)[rank]


def shard_batch(images: torch.Tensor, exif: Optional[Dict[str, torch.Tensor]], rank: int, world: int):
    """Slice a global batch (and its EXIF dict, entries [B] or [B,1]) down to this rank's shard (views, no copy)."""
    lo, hi = shard_range(images.shape[0], rank, world)
    ex = None if exif is None else {k: v[lo:hi] for k, v in exif.items()}
    return images[lo:hi], ex


def gather_outputs(local: Sequence[torch.Tensor], n_global: int, group=None) -> Sequence[torch.Tensor]:
    """all_gather of per-image outputs (each [b_local, ...]) into [n_global, ...] on every rank.  Shards may be
    ragged (n_global not divisible by the world size): every rank pads to the largest shard, the pad rows are
    dropped after the collective.  Works on NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bounds = shard_bounds(n_global, world)
    bmax = max(hi - lo for lo, hi in bounds)
    outs = []
    for t in local:
        lo, hi = bounds[dist.get_rank(group)]
        if t.shape[0] != hi - lo:
            raise ValueError(f"local output has {t.shape[0]} rows, this rank's shard has {hi - lo}")
        pad = t.new_zeros((bmax,) + tuple(t.shape[1:]))
        pad[: t.shape[0]] = t
        buf = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(buf, pad.contiguous(), group=group)
        outs.append(torch.cat([b[: h - l] for b, (l, h) in zip(buf, bounds)], dim=0))
    return outs


class ShardedInference:
    """Runs `model.forward_with_guidance` / `model.forward` on this rank's contiguous shard of a global batch.

    model : a CognitiveAimModel replica on this rank's GPU
    gather: all_gather the outputs so every rank returns the global [B, ...] tensors (default: local shard only)
    """

    def __init__(self, model, rank: int, world: int, gather: bool = False, group=None):
        if not 0 <= rank < world:
            raise ValueError(f"rank {rank} outside world of {world}")
        self.model, self.rank, self.world, self.gather, self.group = model, rank, world, gather, group

    def _run(self, fn, images, exif_data, *args, **kw):
        n = images.shape[0]
        x, ex = shard_batch(images, exif_data, self.rank, self.world)
        if x.shape[0] == 0:
            raise ValueError(f"global batch of {n} leaves rank {self.rank} of {self.world} without images")
        prev = (getattr(self.model, "rng_replay_batch", None), getattr(self.model, "rng_replay_offset", 0))
        # replay the reference's CPU-generator draws at the GLOBAL batch size and use this shard's rows of them
        self.model.rng_replay_batch = n
        self.model.rng_replay_offset = shard_range(n, self.rank, self.world)[0]
        try:
            out = fn(x, ex, *args, **kw)
        finally:
            self.model.rng_replay_batch, self.model.rng_replay_offset = prev
        if self.gather and self.world > 1:
            out = tuple(gather_outputs(list(out), n, self.group))
        return out

    def forward_with_guidance(self, images, exif_data=None, attention_guidance=None, return_attention=False):
        return self._run(self.model.forward_with_guidance, images, exif_data, attention_guidance,
                         return_attention=return_attention)

    def forward(self, images, exif_data=None, return_attention=False):
        return self._run(self.model.forward, images, exif_data, return_attention=return_attention)
