"""Multi-GPU plumbing: the population is sharded over ranks (one process per GPU); self-play
pairings stay inside a shard; per generation the ranks exchange only the fitness vector
(all-gather) and their best genomes (elite all-gather = every owner broadcasting its elites).
This replaces the reference's scoop.futures.map scatter/gather (ga.py:83, main.py:165).

All functions work on whatever backend the process group uses (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of the population."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_elites(fitness: torch.Tensor, genomes: torch.Tensor, k: int):
    """The k best local genomes, best first; ties keep the lower index first (stable)."""
    k = min(k, fitness.numel())
    order = torch.argsort(fitness, descending=True, stable=True)[:k]
    return genomes.index_select(0, order), fitness.index_select(0, order), order


def exchange_generation(fitness: torch.Tensor, genomes: torch.Tensor, k_elite: int):
    """fitness f64[n_local], genomes f32[n_local, G] on this rank's device.
    Returns (global_fitness f64[world*n_local], elite_genomes f32[world*k, G], elite_fitness f64[world*k])
    identical on every rank; elites are merged best-first (ties: lower rank, then lower local index)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        eg, ef, _ = local_elites(fitness, genomes, k_elite)
        return fitness, eg, ef
    world = dist.get_world_size()
    n_local = fitness.numel()
    global_fitness = torch.empty(world * n_local, dtype=fitness.dtype, device=fitness.device)
    dist.all_gather_into_tensor(global_fitness, fitness.contiguous())
    eg, ef, _ = local_elites(fitness, genomes, k_elite)
    k = eg.shape[0]
    all_g = torch.empty((world * k, genomes.shape[1]), dtype=genomes.dtype, device=genomes.device)
    all_f = torch.empty(world * k, dtype=fitness.dtype, device=fitness.device)
    dist.all_gather_into_tensor(all_g, eg.contiguous())
    dist.all_gather_into_tensor(all_f, ef.contiguous())
    order = torch.argsort(all_f, descending=True, stable=True)
    return global_fitness, all_g.index_select(0, order), all_f.index_select(0, order)


def global_stats(global_fitness: torch.Tensor):
    """avg / std (ddof=0) / min / max over the whole population (main.py:158-162)."""
    f = global_fitness.double()
    return f.mean().item(), f.std(unbiased=False).item(), f.min().item(), f.max().item()
