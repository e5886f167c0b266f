"""CPU tests of the multi-GPU host logic: utterance sharding + token-id gather, world_size 2 and 3 on gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200_whisper.runtime import gather_token_ids, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 16, 64, 65):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(64, 8, 3) == (24, 32)  # BASELINE configs[3]: 64 utterances over 8 GPUs
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, T, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_bounds(n_total, world, rank)
        # a deterministic stand-in for the per-utterance decode: token t of utterance u is 1000 * u + t
        local = (torch.arange(b, e, dtype=torch.int32)[:, None] * 1000 + torch.arange(T, dtype=torch.int32)[None, :])
        full = gather_token_ids(local, n_total)
        expect = torch.arange(n_total, dtype=torch.int32)[:, None] * 1000 + torch.arange(T, dtype=torch.int32)[None, :]
        ret[rank] = bool(torch.equal(full, expect))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 5), (2, 16), (3, 2), (2, 1)])
def test_gather_token_ids_gloo(world, n_total):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, 7, ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values())
