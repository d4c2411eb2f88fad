"""Multi-GPU sharding of the batched decode (BASELINE.json north_star: "each GPU taking a
contiguous slice ... NCCL over NVLink is used only to all-gather decoded paths").

Sequences are independent units (`viterbi::decode` is called once per sequence, viterbi.rs:5), so the batch is cut
into one contiguous slice per rank, balanced by forward steps; the model is replicated.  One process per GPU,
`torch.distributed` (NCCL) for the plumbing.  The data path never leaves the device between the decode and the
collective: `cv_decode_batch_dev(_u8)` writes this rank's scores and paths straight into its row of ONE padded
gather buffer and one in-place `all_gather_into_tensor` (ncclAllGather over NVLink) completes it on every rank.  The
collective is the only exchange; there is no reduction on this path.  The gathered product is the padded buffer
(row r = rank r's slice: `gscores[r, :n_sq[r]]`, `gpaths[r, :n_el[r]]`); `paths()` / `scores()` concatenate the rows
into batch order for callers that want one flat array."""
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


class ShardedDecoder:
    """The sharded batched decode of ONE batch layout (seq_off), reusable across steps with new observations.

    Every rank builds it with the same `seq_off`.  Buffers (torch tensors on this rank's device):
      obs_l   [n_el(rank)]  int32   this rank's slice of the observations (fill with `load_obs` or write in place)
      gbuf    [world, 8*S + P*e] uint8  the gather buffer, one row per rank: scores first, then paths
      gscores [world, S]    float64 view of gbuf, S = max_r n_seq(r)
      gpaths  [world, P]    uint8 / int32 view of gbuf, P = max_r n_el(r); row r = paths of rank r's slice
                            (u8 states when K <= 64 -- a quarter of the all-gather bytes -- else u32 bit patterns)
    `step()` enqueues decode + ONE in-place all-gather on the current stream; `paths()` / `scores()` return the
    batch-ordered results (device tensors, one concatenation of the un-padded rows).

    `decode_fn(obs_np, off_np) -> (paths, scores)` replaces the GPU call in the CPU-only (gloo) host-logic tests."""

    def __init__(self, hmm, seq_off, device: int = -1, decode_fn=None, group=None, narrow_paths=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.hmm, self.decode_fn = hmm, decode_fn
        seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        self.seq_off = seq_off
        self.bounds = shard_bounds(seq_off, self.world)
        self.n_el = [int(seq_off[self.bounds[r + 1]] - seq_off[self.bounds[r]]) for r in range(self.world)]
        self.n_sq = [int(self.bounds[r + 1] - self.bounds[r]) for r in range(self.world)]
        self.b0, self.b1 = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.e0, self.e1 = int(seq_off[self.b0]), int(seq_off[self.b1])
        on_gpu = decode_fn is None
        if on_gpu:
            if not torch.cuda.is_available():
                raise RuntimeError("ShardedDecoder needs a CUDA device (no CPU fallback)")
            self.device_index = torch.cuda.current_device() if device < 0 else device
            self.dev = torch.device("cuda", self.device_index)
        else:
            self.device_index, self.dev = -1, torch.device("cpu")
        off_l = seq_off[self.b0:self.b1 + 1] - self.e0
        self.off_l_np = off_l
        self.max_len = int(np.diff(off_l).max()) if self.b1 > self.b0 else 0
        if narrow_paths is None:
            narrow_paths = on_gpu and hmm.nstates() <= 64
        self.narrow_paths = bool(narrow_paths)
        P, S = max(max(self.n_el), 1), max(max(self.n_sq), 1)
        # rows padded to 512 bytes (scores and paths each): every rank's row starts on a 512-byte boundary of the gather
        # buffer
        P, S = (P + 511) // 512 * 512, (S + 63) // 64 * 64
        self.off_l = torch.from_numpy(off_l.copy()).to(self.dev)
        self.obs_l = torch.zeros(max(self.n_el[self.rank], 1), dtype=torch.int32, device=self.dev)
        pbytes = P * (1 if self.narrow_paths else 4)
        self.gbuf = torch.zeros((self.world, 8 * S + pbytes), dtype=torch.uint8, device=self.dev)
        self.gscores = self.gbuf[:, : 8 * S].view(torch.float64)
        self.gpaths = self.gbuf[:, 8 * S:] if self.narrow_paths else self.gbuf[:, 8 * S:].view(torch.int32)
        self.handle = hmm.device_handle(self.device_index) if on_gpu else None

    # -- inputs ------------------------------------------------------------------------------------------------
    def load_obs(self, obs_flat, non_blocking: bool = False):
        """Copy this rank's slice of the FULL observation array (numpy u32, or a pinned torch int32 tensor)."""
        torch = self.torch
        if isinstance(obs_flat, np.ndarray):
            src = torch.from_numpy(np.ascontiguousarray(obs_flat[self.e0:self.e1]).view(np.int32))
        else:
            src = obs_flat[self.e0:self.e1]
        self.obs_l[: self.e1 - self.e0].copy_(src, non_blocking=non_blocking)

    # -- one step ------------------------------------------------------------------------------------------------
    def decode_local(self):
        """This rank's slice: paths / scores land directly in row `rank` of the gather buffers."""
        n_el, n_sq = self.n_el[self.rank], self.n_sq[self.rank]
        if n_sq == 0:
            return
        if self.decode_fn is not None:
            p, s = self.decode_fn(self.obs_l[:n_el].numpy().view(np.uint32), self.off_l_np)
            pp = np.ascontiguousarray(p)
            self.gpaths[self.rank, :n_el] = self.torch.from_numpy(pp.astype(np.uint8) if self.narrow_paths else pp.view(np.int32))
            self.gscores[self.rank, :n_sq] = self.torch.from_numpy(np.ascontiguousarray(s))
            return
        from . import _lib
        st = self.torch.cuda.current_stream(self.dev)
        fn = _lib.lib().cv_decode_batch_dev_u8 if self.narrow_paths else _lib.lib().cv_decode_batch_dev
        rc = fn(self.handle, self.obs_l.data_ptr(), self.off_l.data_ptr(), n_sq, n_el, self.max_len,
                self.gpaths[self.rank].data_ptr(), self.gscores[self.rank].data_ptr(), st.cuda_stream, 0)
        _lib.check(rc)

    def gather(self):
        """In-place all-gather of the buffer (ncclAllGather: row r of every rank's buffer <- rank r's row)."""
        if self.world == 1:
            return
        self.dist.all_gather_into_tensor(self.gbuf.view(-1), self.gbuf[self.rank], group=self.group)

    def step(self):
        self.decode_local()
        self.gather()

    # -- results ---------------------------------------------------------------------------------------------------
    def paths(self):
        """[N] states in batch order on this rank's device: uint8, or int32 bit patterns of u32 when not narrow."""
        return self.torch.cat([self.gpaths[r, : self.n_el[r]] for r in range(self.world)])

    def scores(self):
        return self.torch.cat([self.gscores[r, : self.n_sq[r]] for r in range(self.world)])


def decode_batch_sharded(hmm, obs_flat, seq_off, device: int = -1, gather: bool = True, decode_fn=None):
    """Decode this rank's slice on its GPU; with gather=True every rank returns the full (paths, scores) as numpy
    arrays.  One H2D of the rank's observations, one D2H of the results; everything between stays on the device
    (see ShardedDecoder).  Returns (paths u32, scores f64, (b0, b1))."""
    sd = ShardedDecoder(hmm, seq_off, device=device, decode_fn=decode_fn)
    sd.load_obs(np.ascontiguousarray(obs_flat, dtype=np.uint32))
    sd.decode_local()
    if gather and sd.world > 1:
        sd.gather()
        paths, scores = sd.paths(), sd.scores()
    else:
        paths, scores = sd.gpaths[sd.rank, : sd.n_el[sd.rank]], sd.gscores[sd.rank, : sd.n_sq[sd.rank]]
    pn = paths.cpu().numpy()
    return (pn.astype(np.uint32) if sd.narrow_paths else pn.view(np.uint32)), scores.cpu().numpy(), (sd.b0, sd.b1)
