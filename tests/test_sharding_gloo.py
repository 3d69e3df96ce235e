"""Multi-rank host logic on CPU: world_size-2 gloo (the GPU path uses NCCL with the same code)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vtts_b200 import gather_waveforms, plan_shards, shard_batch


def test_plan_shards_partitions_and_balances():
    lengths = [500, 120, 480, 300, 310, 90, 700, 650, 20, 400, 410, 50, 600, 30, 220, 230]
    for world in (1, 2, 4, 8):
        shards = plan_shards(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lengths)))
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        totals = [sum(lengths[i] for i in s) for s in shards]
        assert max(totals) - min(totals) <= max(lengths)
    assert plan_shards([], 2) == [[], []]
    assert plan_shards([5], 4) == [[0], [], [], []]
    with pytest.raises(ValueError):
        plan_shards([1], 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = len(lengths)
        g = torch.Generator().manual_seed(0)
        hs = torch.randn(total, 7, 4, generator=g)
        idx, (mine,) = shard_batch([hs], lengths, rank, world)
        assert torch.equal(mine, hs[idx])
        # stand-in "synthesis": waveform row i is filled with i, valid length = lengths[i] * 2
        lmax = max(lengths[i] for i in idx) * 2 if idx else 0
        wav = torch.zeros(len(idx), 1, lmax)
        wl = torch.zeros(len(idx), dtype=torch.long)
        for r, i in enumerate(idx):
            wav[r, 0, : lengths[i] * 2] = float(i + 1)
            wl[r] = lengths[i] * 2
        out = gather_waveforms(wav, wl, idx, total)
        assert out is not None
        ow, ol = out
        assert ow.shape == (total, 1, max(lengths) * 2)
        for i in range(total):
            assert int(ol[i]) == lengths[i] * 2
            assert torch.all(ow[i, 0, : lengths[i] * 2] == float(i + 1))
            assert torch.all(ow[i, 0, lengths[i] * 2:] == 0)
        only0 = gather_waveforms(wav, wl, idx, total, dst=0)
        assert (only0 is not None) == (rank == 0)
    finally:
        dist.destroy_process_group()


def test_gather_waveforms_world2_gloo():
    lengths = [9, 3, 7, 5, 1]
    mp.spawn(_worker, args=(2, _free_port(), lengths), nprocs=2, join=True)
