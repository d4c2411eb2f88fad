"""Multi-GPU sharding of the batched decode.

Sequences are independent units, so the batch is cut into one contiguous slice per rank, balanced by
forward steps sum(T_b - 1); the model is replicated; no collective sits on the data path.  The only
collectives are the optional all-gather of the decoded paths/scores (`gather=True`) over NCCL (or gloo on
CPU-only hosts for the host-logic tests).  One process per GPU, `torch.distributed` for the plumbing."""
from __future__ import annotations

import numpy as np


def shard_bounds(seq_off, world_size: int):
    """Contiguous sequence ranges [b0, b1) per rank, balanced by sum(T_b - 1) (+1 per sequence so that
    length-1 sequences still count).  Returns int64 array of world_size + 1 boundaries."""
    seq_off = np.asarray(seq_off, dtype=np.int64)
    B = len(seq_off) - 1
    w = np.diff(seq_off)                      # (T_b - 1) + 1
    cum = np.concatenate([[0], np.cumsum(w)])
    total = cum[-1]
    bounds = np.zeros(world_size + 1, dtype=np.int64)
    for r in range(1, world_size):
        bounds[r] = int(np.searchsorted(cum, total * r / world_size, side="left"))
    bounds[world_size] = B
    return np.maximum.accumulate(bounds)


def local_slice(obs_flat, seq_off, rank: int, world_size: int):
    """This rank's (obs, seq_off rebased to 0, b0, b1)."""
    seq_off = np.asarray(seq_off, dtype=np.int64)
    bnd = shard_bounds(seq_off, world_size)
    b0, b1 = int(bnd[rank]), int(bnd[rank + 1])
    e0, e1 = int(seq_off[b0]), int(seq_off[b1])
    return np.asarray(obs_flat)[e0:e1], seq_off[b0:b1 + 1] - e0, b0, b1


def decode_batch_sharded(hmm, obs_flat, seq_off, device: int = -1, gather: bool = True, decode_fn=None):
    """Decode this rank's slice on its GPU; with gather=True every rank returns the full (paths, scores).

    `decode_fn(hmm, obs, off)` defaults to the GPU path (`viterbi.decode_batch`); the CPU-only tests of the
    host logic inject a stand-in."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if decode_fn is None:
        from .viterbi import decode_batch

        def decode_fn(h, o, f):
            return decode_batch(h, o, f, device=device)
    seq_off = np.asarray(seq_off, dtype=np.int64)
    obs_l, off_l, b0, b1 = local_slice(obs_flat, seq_off, rank, world)
    paths_l, scores_l = decode_fn(hmm, obs_l, off_l) if b1 > b0 else (np.zeros(0, np.uint32), np.zeros(0))
    if not gather or world == 1:
        return paths_l, scores_l, (b0, b1)
    bnd = shard_bounds(seq_off, world)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    # all-gather with padding to the largest slice (paths u32 as int32 bit patterns, scores f64)
    n_el = [int(seq_off[bnd[r + 1]] - seq_off[bnd[r]]) for r in range(world)]
    n_sq = [int(bnd[r + 1] - bnd[r]) for r in range(world)]
    pbuf = torch.zeros(max(n_el), dtype=torch.int32, device=dev)
    pbuf[: len(paths_l)] = torch.from_numpy(np.ascontiguousarray(paths_l).view(np.int32)).to(dev)
    sbuf = torch.zeros(max(n_sq), dtype=torch.float64, device=dev)
    sbuf[: len(scores_l)] = torch.from_numpy(np.ascontiguousarray(scores_l)).to(dev)
    pall = [torch.empty_like(pbuf) for _ in range(world)]
    sall = [torch.empty_like(sbuf) for _ in range(world)]
    dist.all_gather(pall, pbuf)
    dist.all_gather(sall, sbuf)
    paths = np.concatenate([pall[r][: n_el[r]].cpu().numpy().view(np.uint32) for r in range(world)])
    scores = np.concatenate([sall[r][: n_sq[r]].cpu().numpy() for r in range(world)])
    return paths, scores, (b0, b1)
