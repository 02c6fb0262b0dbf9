"""World-size-2 gloo tests of the multi-GPU host logic (sharding, fitness all-gather, elite merge)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuro_genetic_pong_self_play_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total, G, k = 10, 5, 2
    rng = np.random.RandomState(0)
    fitness_all = torch.from_numpy(np.round(rng.standard_normal(n_total), 1))
    genomes_all = torch.from_numpy(rng.random_sample((n_total, G)).astype(np.float32))
    lo, hi = parallel.shard_bounds(n_total, world, rank)
    gf, eg, ef = parallel.exchange_generation(fitness_all[lo:hi].clone(), genomes_all[lo:hi].clone(), k)
    q.put((rank, gf.numpy(), eg.numpy(), ef.numpy()))
    dist.destroy_process_group()


def test_shard_bounds_cover_population():
    for n in (1, 7, 64, 1024, 16385):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_exchange_generation_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    rng = np.random.RandomState(0)
    fitness_all = np.round(rng.standard_normal(10), 1)
    genomes_all = rng.random_sample((10, 5)).astype(np.float32)
    # every rank sees the same global fitness and the same merged elites
    for rank, gf, eg, ef in res:
        assert np.array_equal(gf, fitness_all)
        assert np.array_equal(eg, res[0][2]) and np.array_equal(ef, res[0][3])
    # elites = top-2 of each shard, merged best-first
    expect = []
    for r in range(2):
        lo, hi = parallel.shard_bounds(10, 2, r)
        order = np.argsort(-fitness_all[lo:hi], kind="stable")[:2] + lo
        expect += list(order)
    expect = sorted(expect, key=lambda i: -fitness_all[i])
    assert np.array_equal(res[0][3], fitness_all[expect])
    assert np.array_equal(res[0][2], genomes_all[expect])


def test_exchange_generation_single_process():
    f = torch.tensor([0.1, 0.9, 0.5], dtype=torch.float64)
    g = torch.arange(6, dtype=torch.float32).reshape(3, 2)
    gf, eg, ef = parallel.exchange_generation(f, g, 2)
    assert torch.equal(gf, f) and ef.tolist() == [0.9, 0.5] and eg.tolist() == [[2.0, 3.0], [4.0, 5.0]]
    assert parallel.global_stats(f)[2:] == (0.1, 0.9)
