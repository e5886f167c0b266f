"""Utterance sharding across GPUs (the only multi-GPU mode of this path).

Utterances are independent: each rank decodes a contiguous slice of the batch with its own replica of the weights and
its own KV caches; nothing is exchanged during encode or decode.  The single collective is a gather of the int32 token
ids after the decode (<= 64 x 64 x 4 B), on NCCL over NVLink for GPU tensors and gloo for the CPU tests.
The reference has no multi-GPU mode for Whisper (world_size = 1 is hard-coded, T/examples/whisper/run.py:38-41).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_total, world_size, rank):
    """Contiguous, balanced split: the first (n_total % world_size) ranks get one extra utterance.
    Returns (begin, end); empty slices are allowed (more ranks than utterances)."""
    if world_size < 1 or not (0 <= rank < world_size) or n_total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_token_ids(local_tokens, n_total, group=None, pad_id=-1):
    """local_tokens: int32 [n_local, T] of this rank's slice -> int32 [n_total, T] on every rank, in utterance order.
    Ragged slices are padded to the largest slice for the all_gather and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local, T = local_tokens.shape
    b, e = shard_bounds(n_total, world, rank)
    assert e - b == n_local, f"rank {rank} holds {n_local} utterances, expected {e - b}"
    n_max = -(-n_total // world) if n_total else 0
    buf = torch.full((max(n_max, 1), T), pad_id, dtype=torch.int32, device=local_tokens.device)
    buf[:n_local] = local_tokens
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for r in range(world):
        rb, re = shard_bounds(n_total, world, r)
        parts.append(out[r][: re - rb])
    return torch.cat(parts, dim=0) if parts else buf[:0]
