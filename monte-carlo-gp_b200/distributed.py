"""Multi-GPU scale-out of the race simulation: shard sims, one integer all-reduce.

Simulated races are independent units and the native RNG is keyed by the *global* sim index, so rank g of G
simply runs the contiguous range ``shard_range(n, g, G)`` on its own GPU; the only exchange is ONE
``all_reduce(SUM)`` of the ``[n_races, n, n]`` int64 count tables (3.2 KB per race: pure latency on
NVLink/NVSwitch).  The result is bit-identical for any G.  One process per GPU (torchrun), NCCL backend on GPUs;
the same code runs over gloo on CPU tensors for the host-logic tests (with an injected count producer -- the
product path itself has no CPU implementation).
"""
from __future__ import annotations

import os
from typing import Callable

import numpy as np


def shard_range(n_sims: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous global sim range of `rank`: returns (begin, count).  Ranges tile [0, n_sims) exactly."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    begin = n_sims * rank // world
    end = n_sims * (rank + 1) // world
    return begin, end - begin


def all_reduce_counts(hist, group=None):
    """In-place SUM all-reduce of an int64 count tensor over the process group (no-op without one)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def run_sharded(n_sims: int, n_races: int, n_drivers: int, producer: Callable, rank: int | None = None,
                world: int | None = None, group=None, device=None):
    """Generic driver: ``producer(begin, count, hist)`` adds this rank's counts into the int64 tensor
    ``hist[n_races, n, n]`` (on `device`); then one all-reduce.  Returns the global table on every rank."""
    import torch
    import torch.distributed as dist
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    hist = torch.zeros((n_races, n_drivers, n_drivers), dtype=torch.int64, device=device)
    begin, count = shard_range(n_sims, rank, world)
    if count:
        producer(begin, count, hist)
    return all_reduce_counts(hist, group)


class ShardedSimulator:
    """Device-resident multi-GPU front end: parameters uploaded once, counts stay on the GPU until the
    all-reduce.  Usage (one process per GPU, torch.distributed initialised with the nccl backend):

        sims = ShardedSimulator([params...])        # mcgp_race_params blocks
        hist = sims.run(n_sims, seed)               # torch.int64 [n_races, n, n] on the GPU, same on all ranks
    """

    def __init__(self, races, device: int | None = None, flags: int = 0):
        import torch
        from . import capi
        self.torch = torch
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        torch.cuda.set_device(self.device)
        # An engine of its own: the resident race blocks and claim counters must not be the ones RaceSimulator's
        # host-buffer calls (capi.get_engine) re-upload on every call.  Launches of one engine are ordered by the
        # library whatever stream they are given (include/mcgp.h), so overlapping launches cannot share counters.
        self.engine = capi.Engine(self.device)
        self.engine.upload_races(races)
        self.n_races, self.n = self.engine.n_races, self.engine.n_drivers
        self.flags = flags

    def launch(self, begin: int, count: int, seed: int, hist) -> None:
        """Asynchronous kernel launch on torch's current stream, accumulating into `hist` (int64 on this GPU)."""
        assert hist.is_cuda and hist.dtype == self.torch.int64 and hist.is_contiguous()
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.engine.launch_native(count, begin, seed, hist.data_ptr(), self.flags, stream=stream)

    def run_by_lap(self, n_sims: int, seed: int, group=None):
        """Like run(), plus the per-lap position histogram (include/mcgp.h: mcgp_launch_native_laphist).  Both tables
        travel in ONE all-reduce (a flat int64 buffer [n_races * (n*n + laps*n*n)]); returns (hist, laphist) views."""
        torch = self.torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device)
        laps = self.engine._lib.mcgp_lap_histogram_laps(self.engine._h)
        n_h, n_l = self.n_races * self.n * self.n, self.n_races * laps * self.n * self.n
        buf = torch.zeros(n_h + n_l, dtype=torch.int64, device=dev)
        hist, laphist = buf[:n_h].view(self.n_races, self.n, self.n), buf[n_h:].view(self.n_races, laps, self.n, self.n)
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
        begin, count = shard_range(n_sims, rank, world)
        if count:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.engine.launch_native_laphist(count, begin, seed, hist.data_ptr(), laphist.data_ptr(), self.flags, stream=stream)
        all_reduce_counts(buf, group)
        return hist, laphist

    def run(self, n_sims: int, seed: int, group=None):
        dev = self.torch.device("cuda", self.device)
        return run_sharded(n_sims, self.n_races, self.n, lambda b, c, h: self.launch(b, c, seed, h), group=group,
                           device=dev)


def tallies(hist: np.ndarray) -> dict:
    """win / podium / points-finish counts per driver from a count table hist[..., driver, pos]."""
    h = np.asarray(hist)
    return dict(win=h[..., 0], podium=h[..., : min(3, h.shape[-1])].sum(-1), points=h[..., : min(10, h.shape[-1])].sum(-1))
