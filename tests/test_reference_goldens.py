"""Episode-level parity against the REFERENCE'S OWN code.

tests/golden/episode_vectors.npz was produced by tools/make_golden_episode.py: /root/reference/main.py imported unmodified
(perform_episode, get_actions, calculate_timeout_and_frames, evaluate -- main.py:28-154 -- with utils.py, numpy_nn.py,
dumb_ais.py, config.py underneath) and driven against a gym-style env backed by the oracle emulator, with np.random.choice
and random.shuffle replaced by the Philox draws / injected hall-of-fame picks the product uses.  It holds, for 20 genomes x 6
games (bots, the 1-player cartridge-robot game, hall-of-fame opponents, hall_of_fame None / empty, 2000-frame timeouts):
the fitness main.evaluate returned, perform_episode's reward per game, the number of env.step calls per game and the button
vector passed to every env.step.

CPU test: the C restatement (oracle/episode_oracle.c) reproduces all of it bit for bit (a subset by default to stay within the
CPU suite's budget; NGP_FULL_GOLDENS=1 runs all 20).  GPU test: the fused CUDA rollout behind ngp_evaluate reproduces rewards,
frame counts and fitness of all 20 genomes bit for bit."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ep():
    return np.load(os.path.join(ROOT, "tests", "golden", "episode_vectors.npz"))


def _plain_mean(rewards):
    """main.py:65 `sum(all_rewards) / float(GAMES_TO_PLAY)` as the reference's own interpreter computes it.  The goldens were
    generated under this image's CPython 3.12, whose built-in sum() of floats is Neumaier-compensated (new in 3.12); gym-retro
    only exists for CPython <= 3.8, where sum() is the plain left-to-right chain of IEEE additions starting from int 0.  The
    product and the oracle follow the reference's interpreter; the two differ by at most one ulp (checked)."""
    out = []
    for r in np.asarray(rewards, np.float64):
        acc = 0.0
        for v in r.tolist():
            acc = acc + v
        out.append(acc / 6.0)
    return np.array(out, np.float64)


def _hof(ep, mode):
    return (ep["hof_genomes"], ep["hof_fitness"]) if mode == 2 else (None, None)


def test_goldens_cover_the_cases(ep):
    frames = ep["frames"]
    assert frames.shape == (20, 6) and ep["rewards"].shape == (20, 6)
    assert (frames > 2000).any(), "a 2000-frame timeout (main.py:106-107)"
    assert (ep["rewards"] > 0).any() and (ep["rewards"] < 0).any(), "games won and lost by the right player"
    assert set(ep["hof_mode"].tolist()) == {0, 1, 2}
    assert np.array_equal(ep["fitness"], np.array([sum(r) / 6.0 for r in ep["rewards"].tolist()]))       # main.py:65 as run here
    np.testing.assert_array_max_ulp(_plain_mean(ep["rewards"]), ep["fitness"], maxulp=1)
    assert ep["action_offsets"][-1] == frames.sum() == len(ep["action_bits"])


def _check_genome(i):
    import oracle
    ep = np.load(os.path.join(ROOT, "tests", "golden", "episode_vectors.npz"))
    mode = int(ep["hof_mode"][i]); gid = int(ep["genome_id"][i])
    seed, gen = int(ep["seed"]), int(ep["generation"])
    hg, hf = _hof(ep, mode)
    fit, rewards, frames = oracle.evaluate([6, 2, 2], ep["genomes"][i], hg, hf, ep["hof_pick"][i], seed=seed, genome_id=gid, generation=gen)
    assert np.array_equal(frames, ep["frames"][i]), (i, frames, ep["frames"][i])
    assert np.array_equal(rewards, ep["rewards"][i]), (i, rewards, ep["rewards"][i])
    assert fit == _plain_mean(ep["rewards"][i:i + 1])[0]
    return i


def _check_actions(ep, i, game):
    """Every button vector the reference passed to env.step in one game against the restated episode loop's decisions."""
    import oracle
    mode = int(ep["hof_mode"][i]); gid = int(ep["genome_id"][i])
    seed, gen = int(ep["seed"]), int(ep["generation"])
    sh = oracle.Shape.make([6, 2, 2])
    left, mult, state = ("hardcoded", None), 1.0, oracle.STATE_START_2P
    if game == 1:
        state = oracle.STATE_START_1P
    elif game == 2:
        left = ("score", None)
    elif game >= 3 and mode == 2:
        h = int(ep["hof_pick"][i][game - 3])
        left, mult = ("mlp", ep["hof_genomes"][h]), float(ep["hof_fitness"][h])
    env = oracle.Atari()
    env.reset_to_state(state)
    n = int(ep["frames"][i][game])
    res, trace = env.episode(sh, left, ("mlp", ep["genomes"][i]), mult=mult, seed=seed, env_id=gid * 6 + game, generation=gen, trace_cap=n + 8)
    assert res.frames == n and res.reward == ep["rewards"][i][game]
    off = int(ep["action_offsets"][i * 6 + game])
    bits = ep["action_bits"][off:off + n]                 # a[4] | a[5] << 1 | a[6] << 2 | a[7] << 3 of the vector given to step t
    enc = {oracle.ACT_NONE: 0, oracle.ACT_UP: 1, oracle.ACT_DOWN: 2}
    mine = np.zeros(n, np.uint8)                          # step 0 gets BLANK_ACTION; step t+1 the decisions made after step t
    for t in range(n - 1):
        mine[t + 1] = enc[int(trace[t, 129])] | (enc[int(trace[t, 128])] << 2)
    assert np.array_equal(mine, bits), (i, game, np.nonzero(mine != bits)[0][:5])


def test_oracle_reproduces_reference_episodes(ep):
    full = os.environ.get("NGP_FULL_GOLDENS") == "1"
    order = np.argsort(ep["frames"].sum(axis=1))
    # the cheapest genome of each hall-of-fame mode by default (every branch of main.evaluate's schedule)
    idx = list(range(20)) if full else [int([i for i in order if ep["hof_mode"][i] == m][0]) for m in (2, 0, 1)]
    import concurrent.futures as cf
    with cf.ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:     # one oracle per process (it is not re-entrant)
        assert sorted(ex.map(_check_genome, idx)) == sorted(idx)


def test_oracle_reproduces_reference_button_vectors(ep):
    order = np.argsort(ep["frames"].sum(axis=1))
    i = int([j for j in order if ep["hof_mode"][j] == 2][0])
    for game in (0, 1, 2, 3):          # HardcodedAi, cartridge robot (players=1), ScoreHardcodedAi, hall-of-fame MLP
        _check_actions(ep, i, game)


@pytest.mark.gpu
def test_cuda_reproduces_reference_episodes(ep):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import neuro_genetic_pong_self_play_b200 as ngp
    eng = ngp.Engine(ngp.Config(), device=0)
    seed, gen = int(ep["seed"]), int(ep["generation"])
    for mode in (2, 0, 1):
        idx = np.nonzero(ep["hof_mode"] == mode)[0]
        assert np.array_equal(ep["genome_id"][idx], np.arange(len(idx)))
        g = torch.from_numpy(ep["genomes"][idx]).cuda()
        if mode == 2:
            out = eng.evaluate(g, torch.from_numpy(ep["hof_genomes"]).cuda(), torch.from_numpy(ep["hof_fitness"]).cuda(),
                               torch.from_numpy(ep["hof_pick"][idx]).cuda(), seed=seed, generation=gen, want_detail=True)
        else:
            out = eng.evaluate(g, seed=seed, generation=gen, want_detail=True)
        assert np.array_equal(out["frames"].cpu().numpy(), ep["frames"][idx])
        assert np.array_equal(out["rewards"].cpu().numpy(), ep["rewards"][idx])
        assert np.array_equal(out["fitness"].cpu().numpy(), _plain_mean(ep["rewards"][idx]))
    eng.close()
