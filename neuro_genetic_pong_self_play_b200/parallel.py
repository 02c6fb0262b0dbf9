"""Population sharding over the GPUs of one box and the per-generation exchange.

The reference scatters individuals with ``toolbox.map = scoop.futures.map`` (ga.py:83), gathers every fitness on the
master and lets eaSimple update one hall of fame from the whole population (main.py:165-170).  Here rank r owns a
contiguous shard (``shard_bounds``); self-play pairings, selection, crossover and mutation stay inside the shard (island
model); once per generation every rank contributes ONE fixed-size record -- its fitness values and its k best genomes,
packed by ``ngp_pack_elites`` -- to a single NCCL all-gather.  The gathered records are unpacked on the device
(``ngp_unpack_elites``) into the global fitness vector (statistics) and the merged elites, which every rank feeds to its
hall of fame in the same order, so all ranks hold the same hall of fame.

The all-gather is issued asynchronously on a side stream and consumed one generation later (``Exchange.start`` /
``Exchange.finish``): a rank whose longest episode is short does not wait for the slowest rank inside the generation.
torch.distributed is plumbing only (process group, the collective itself)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; the first n_total % world ranks hold one genome more."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_total: int, world: int):
    return [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]


def elite_counts(n_total: int, world: int, k_total: int):
    """How many elites each rank contributes: k_total // world (at least 1), never more than its shard."""
    per = max(1, k_total // world) if k_total > 0 else 0
    return [min(per, n) for n in shard_sizes(n_total, world)]


def gather_records(record: torch.Tensor, group=None, async_op: bool = False):
    """All-gather one equally sized uint8 record per rank -> (gathered [world * len(record)], work handle or None).
    Works on CUDA tensors (NCCL) and on CPU tensors (gloo: host-logic tests)."""
    if not dist.is_initialized():
        return record.clone(), None
    world = dist.get_world_size(group)
    gathered = torch.empty(world * record.numel(), dtype=record.dtype, device=record.device)
    if record.is_cuda:
        work = dist.all_gather_into_tensor(gathered, record, group=group, async_op=async_op)
    else:
        work = dist.all_gather(list(gathered.view(world, -1).unbind(0)), record, group=group, async_op=async_op)
    return gathered, work


class Exchange:
    """Per-generation exchange of one rank: pack on the compute stream, all-gather on a side stream, unpack when consumed."""

    def __init__(self, engine, n_total: int, k_total: int, group=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        sizes = shard_sizes(n_total, self.world)
        counts = elite_counts(n_total, self.world, k_total)
        self.n_local, self.k_local = sizes[self.rank], counts[self.rank]
        self.n_max, self.k_max = max(sizes), max(counts)
        self.k_sum = sum(counts)
        self.side = torch.cuda.Stream(device=engine.device)
        self._pending = None

    def start(self, genomes: torch.Tensor, fitness: torch.Tensor):
        """Pack this rank's record (current stream) and launch the all-gather on the side stream."""
        assert genomes.shape[0] == self.n_local and self._pending is None
        record = self.engine.pack_elites(genomes, fitness, self.k_local, self.n_max, self.k_max)
        ready = torch.cuda.Event()
        ready.record()
        record.record_stream(self.side)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            gathered, work = gather_records(record, self.group, async_op=True)
        self._pending = (record, gathered, work)

    def finish(self):
        """Wait for the pending all-gather and unpack it (side stream).  Returns (fitness_all f64[N], elite_genomes
        f32[sum k][G], elite_fitness f64[sum k]); the current stream is made to wait for the results."""
        record, gathered, work = self._pending
        self._pending = None
        with torch.cuda.stream(self.side):
            if work is not None:
                work.wait()
            out = self.engine.unpack_elites(gathered, self.world, self.n_max, self.k_max, self.n_total, self.k_sum)
            done = torch.cuda.Event()
            done.record()
        torch.cuda.current_stream().wait_event(done)
        for t in (record, gathered) + tuple(out):
            t.record_stream(torch.cuda.current_stream())
        return out

    @property
    def pending(self) -> bool:
        return self._pending is not None
