"""Batch sharding across ranks: synthesis is embarrassingly data-parallel per utterance.

One process per GPU (torchrun), weights replicated (13.9 M parameters), no collective on the
hot path.  ``plan_shards`` deals utterances to ranks by length so that
each rank's padded batch has about the same number of frames; ``gather_waveforms`` is the
optional final exchange (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def plan_shards(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Snake-deal utterance indices (longest first) over ``world_size`` ranks.

    Returns ``world_size`` index lists; every index appears exactly once.  Sorting keeps
    utterances of similar length together (less padding inside a rank's batch) and the
    boustrophedon order balances the total frames per rank.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for pos, idx in enumerate(order):
        lap, off = divmod(pos, world_size)
        rank = off if lap % 2 == 0 else world_size - 1 - off
        shards[rank].append(idx)
    return shards


def shard_batch(tensors: Sequence[torch.Tensor], lengths: Sequence[int], rank: int, world_size: int):
    """Select this rank's rows of each (B, ...) tensor.  Returns (index list, sharded tensors)."""
    idx = plan_shards(lengths, world_size)[rank]
    sel = torch.as_tensor(idx, dtype=torch.long)
    return idx, [t.index_select(0, sel.to(t.device)) for t in tensors]


def gather_waveforms(
    wav: torch.Tensor, wav_len: torch.Tensor, index: Sequence[int], total: int,
    group=None, dst: Optional[int] = None,
) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Collect every rank's (b_r, 1, L_r) waveforms into original utterance order.

    Ranks may hold different batch sizes and lengths: shapes are exchanged first, then the
    payload is all-gathered padded to the global maximum.  Returns ``(wav (total,1,Lmax),
    wav_len (total,))`` on every rank (or only on ``dst`` if given; other ranks get None).
    """
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = wav.device
    meta = torch.tensor([wav.shape[0], wav.shape[-1]], dtype=torch.long, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    bmax = int(max(m[0] for m in metas))
    lmax = int(max(m[1] for m in metas))
    pad_w = torch.zeros((bmax, 1, lmax), dtype=wav.dtype, device=dev)
    pad_w[: wav.shape[0], :, : wav.shape[-1]] = wav
    pad_l = torch.zeros(bmax, dtype=torch.long, device=dev)
    pad_l[: wav.shape[0]] = wav_len.to(torch.long)
    pad_i = torch.full((bmax,), -1, dtype=torch.long, device=dev)
    pad_i[: len(index)] = torch.as_tensor(list(index), dtype=torch.long, device=dev)
    ws = [torch.empty_like(pad_w) for _ in range(world)]
    ls = [torch.empty_like(pad_l) for _ in range(world)]
    ix = [torch.empty_like(pad_i) for _ in range(world)]
    dist.all_gather(ws, pad_w, group=group)
    dist.all_gather(ls, pad_l, group=group)
    dist.all_gather(ix, pad_i, group=group)
    if dst is not None and rank != dst:
        return None
    out_w = torch.zeros((total, 1, lmax), dtype=wav.dtype, device=dev)
    out_l = torch.zeros(total, dtype=torch.long, device=dev)
    for w, l, i in zip(ws, ls, ix):
        valid = i >= 0
        out_w[i[valid]] = w[valid]
        out_l[i[valid]] = l[valid]
    return out_w, out_l
