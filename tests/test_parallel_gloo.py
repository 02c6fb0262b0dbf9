"""World-size-2 and -3 gloo tests of the multi-GPU host logic: sharding (even and uneven), elite quotas and the
all-gather of the fixed-size per-rank records.  The records themselves are packed / unpacked by CUDA kernels in the product
(ngp_pack_elites / ngp_unpack_elites, covered by the -m gpu tests); here the numpy restatement of the record layout in
oracle/ stands in for them, so that the collective plumbing (parallel.gather_records) runs on CPU tensors over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuro_genetic_pong_self_play_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _population(n_total, G):
    rng = np.random.RandomState(0)
    return np.round(rng.standard_normal(n_total), 1), rng.random_sample((n_total, G)).astype(np.float32)


def _worker(rank, world, port, n_total, k_total, q):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = 5
    fitness_all, genomes_all = _population(n_total, G)
    lo, hi = parallel.shard_bounds(n_total, world, rank)
    sizes = parallel.shard_sizes(n_total, world); counts = parallel.elite_counts(n_total, world, k_total)
    n_max, k_max = max(sizes), max(counts)
    rec = oracle.pack_record(genomes_all[lo:hi], fitness_all[lo:hi], counts[rank], n_max, k_max)
    gathered, work = parallel.gather_records(torch.from_numpy(rec), async_op=True)
    work.wait()
    fit, eg, ef = oracle.unpack_records(gathered.numpy(), world, n_max, k_max, G)
    q.put((rank, fit, eg, ef))
    dist.destroy_process_group()


def test_shard_bounds_cover_population():
    for n in (1, 7, 64, 1024, 16385):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == parallel.shard_sizes(n, w)
            counts = parallel.elite_counts(n, w, n // 4)
            assert all(0 <= c <= s for c, s in zip(counts, sizes))


@pytest.mark.parametrize("world,n_total,k_total", [(2, 10, 4), (3, 10, 4), (3, 11, 9)])
def test_record_allgather(world, n_total, k_total):
    """Every rank ends with the same global fitness vector (rank order = population order) and the same merged elites (each
    shard's best first), also when the shards differ in size (10 genomes on 3 ranks)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, k_total, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    fitness_all, genomes_all = _population(n_total, 5)
    counts = parallel.elite_counts(n_total, world, k_total)
    expect = []
    for r in range(world):
        lo, hi = parallel.shard_bounds(n_total, world, r)
        expect += list(np.argsort(-fitness_all[lo:hi], kind="stable")[:counts[r]] + lo)
    for rank, fit, eg, ef in res:
        assert np.array_equal(fit, fitness_all)
        assert np.array_equal(ef, fitness_all[expect]) and np.array_equal(eg, genomes_all[expect])


def test_gather_records_without_process_group():
    rec = torch.arange(32, dtype=torch.uint8)
    gathered, work = parallel.gather_records(rec)
    assert work is None and torch.equal(gathered, rec)
