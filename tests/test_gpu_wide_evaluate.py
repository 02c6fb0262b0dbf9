"""ngp_evaluate with networks wider than the fused rollout's 32-unit limit (per-frame stepwise driver, csrc/ngp_stepwise.cu;
BASELINE config 4 = [6,512,512,2] with 64 environments per genome) against the CPU oracle.

The oracle's policy is numpy_nn.NeuralNetwork.run in FP64; the wide CUDA layers compute in FP32 / 3xTF32 with an FP64 action
layer (north_star: rtol 1e-5, argmax agreement).  An episode is reproduced bit for bit as long as every argmax agrees, which is
what is asserted here on seeded genomes: frame counts and FP64 rewards of every game equal the oracle's."""
import concurrent.futures as cf
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ngp():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import neuro_genetic_pong_self_play_b200 as m
    return m


def _selfplay(args):
    import oracle
    nodes, right, left, seed, env_id = args
    r = oracle.selfplay_game(list(nodes), right, left, seed=seed, env_id=env_id)
    return r.reward, r.frames


def _evaluate(args):
    import oracle
    nodes, genome, hof, hof_fit, pick, seed, gid = args
    return oracle.evaluate(list(nodes), genome, hof, hof_fit, pick, seed=seed, genome_id=gid)


def _genomes(rng, n, G, nodes):
    """Weights scaled by 1/sqrt(fan-in) per layer so that hidden units do not all saturate and decisions depend on the inputs."""
    g = np.zeros((n, G), np.float32)
    off = 0
    for a, b in zip(nodes[:-1], nodes[1:]):
        cnt = (a + 1) * b
        g[:, off:off + cnt] = (rng.standard_normal((n, cnt)) * 3.0 / np.sqrt(a + 1)).astype(np.float32)
        off += cnt
    return g


@pytest.mark.parametrize("nodes,n,games", [((6, 48, 2), 6, 3), ((6, 512, 512, 2), 5, 4), ((6, 64, 40, 2), 4, 2)])
def test_wide_round_robin_parity(ngp, nodes, n, games):
    cfg = ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, GAMES_TO_PLAY=games, NETWORK_SHAPE=nodes, POPULATION_SIZE=n)
    eng = ngp.Engine(cfg, device=0)
    rng = np.random.RandomState(sum(nodes))
    genomes = _genomes(rng, n, eng.gene_size, nodes)
    jobs = [(nodes, genomes[g], genomes[(g + k + 1) % n], 21, g * games + k) for g in range(n) for k in range(games)]
    with cf.ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        ref = list(ex.map(_selfplay, jobs))
    out = eng.evaluate(torch.from_numpy(genomes).cuda(), seed=21, want_detail=True)
    frames = out["frames"].cpu().numpy().reshape(-1); rewards = out["rewards"].cpu().numpy().reshape(-1)
    assert np.array_equal(frames, np.array([f for _, f in ref])), (frames, [f for _, f in ref])
    assert np.array_equal(rewards, np.array([r for r, _ in ref]))
    assert out["frames_total"] == int(frames.sum())
    acc = np.zeros(n)
    for k in range(games):
        acc = acc + out["rewards"].cpu().numpy()[:, k]
    assert np.array_equal(out["fitness"].cpu().numpy(), acc / games)
    # same bits on a second call (per-handle scratch reused) and through the interpreter core
    again = eng.evaluate(torch.from_numpy(genomes).cuda(), seed=21, want_detail=True)
    assert torch.equal(again["rewards"], out["rewards"]) and torch.equal(again["frames"], out["frames"])
    eng.close()


def test_wide_reference_schedule_parity(ngp):
    """Bots, the 1-player cartridge-robot game and hall-of-fame opponents (gathered weights, one MLP row each)."""
    nodes = (6, 40, 2)
    eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=nodes), device=0)
    rng = np.random.RandomState(8)
    n, G = 5, eng.gene_size
    genomes = _genomes(rng, n, G, nodes)
    hof = _genomes(rng, 3, G, nodes); hof_fit = np.array([1.25, 0.5, -0.125])
    pick = rng.randint(0, 3, size=(n, 3)).astype(np.int32)
    with cf.ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        ref = list(ex.map(_evaluate, [(nodes, genomes[g], hof, hof_fit, pick[g], 4, g) for g in range(n)]))
    out = eng.evaluate(torch.from_numpy(genomes).cuda(), torch.from_numpy(hof).cuda(), torch.from_numpy(hof_fit).cuda(),
                       torch.from_numpy(pick).cuda(), seed=4, want_detail=True)
    for g in range(n):
        fit, r, f = ref[g]
        assert np.array_equal(out["frames"][g].cpu().numpy(), f), (g, out["frames"][g], f)
        assert np.array_equal(out["rewards"][g].cpu().numpy(), r)
        assert out["fitness"][g].item() == fit
    # without a hall of fame games 3..5 fall back to HardcodedAi (main.py:43-53)
    out2 = eng.evaluate(torch.from_numpy(genomes[:2]).cuda(), seed=4, want_detail=True)
    for g in range(2):
        fit, r, f = _evaluate((nodes, genomes[g], None, None, (0, 0, 0), 4, g))
        assert np.array_equal(out2["frames"][g].cpu().numpy(), f) and np.array_equal(out2["rewards"][g].cpu().numpy(), r)
    eng.close()
