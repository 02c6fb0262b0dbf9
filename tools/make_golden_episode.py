#!/usr/bin/env python3
"""Generate tests/golden/episode_vectors.npz by running the REFERENCE's own, unmodified episode logic.

/root/reference/main.py is imported as it is (perform_episode, get_actions, calculate_timeout_and_frames, evaluate:
main.py:28-154, plus everything it pulls from utils.py / numpy_nn.py / dumb_ais.py / config.py) under stub modules for
the third-party packages that are absent from the image:

  retro          make(game, state=, players=) -> a gym-style env (step/reset/close, info = {"score1","score2"}) backed
                 by the in-repo oracle emulator (oracle.Atari); the gym-retro layer (8 buttons per player, players=1
                 consumes action[0:8], FILTERED cancels opposite directions, button names -> console events) is stated
                 here in Python, independently of the C oracle's eo_action_to_input
  deap, scoop    inert (main.py only needs the names at import time for the functions used here)
  ga             toolbox with register(), hall_of_fame object with .items (DEAP HallOfFame's public attribute that
                 utils.create_model_from_hall_of_fame reads and shuffles, utils.py:90-101)
  human_control  inert
plus two numpy-version shims: np.int = int (config.py:21,26) and the pre-1.24 ragged np.array behaviour for utils.find_stuff.

Two sources of randomness are replaced so that the run is reproducible and comparable with the CUDA path:
  np.random.choice (utils.get_random_action, two draws per frame, main.py:139-140) -> Philox4x32-10 keyed by the seed with
      counter (env, frame, 2*generation + player, 'PONG')   [oracle.philox4x32 is pinned to the Random123 vectors]
  random.shuffle  (utils.py:94) -> puts the injected hall-of-fame pick first (canonical order otherwise; the reference's
      in-place shuffle of DEAP's items list is a side effect the product does not reproduce)

Only outputs are stored (rewards, frame counts, per-frame button vectors); nothing of the reference is copied.
Runs only in the build container (needs /root/reference).  ~10 minutes (the reference path costs ~3.5 ms per frame).
"""
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "episode_vectors.npz")

import oracle  # noqa: E402  (test infrastructure; this tool is a test-fixture generator)

SEED = 20261018
GENERATION = 3
BUTTONS = ["BUTTON", None, "SELECT", "RESET", "UP", "DOWN", "LEFT", "RIGHT"]     # gym-retro Atari2600 layout, per player
# (player, button) -> console event, following the reference's own names (config.py:15-20, main.py:91-92): player 1's BUTTON
# (index 0, RIGHT_PLAYER_START_BUTTON) is the right player's = paddle 1's fire button, index -1 = player 2's RIGHT
# (LEFT_PLAYER_START_BUTTON) the left player's = paddle 0's; action[4:6] = player 1's UP/DOWN move the right paddle (paddle 1),
# action[6:8] = player 1's LEFT/RIGHT the left paddle (paddle 0).  The product's default button map says the same in codes.
CONSOLE_EVENTS = {(0, "BUTTON"): ("fire", 1), (1, "RIGHT"): ("fire", 0), (0, "UP"): ("up", 1), (0, "DOWN"): ("down", 1),
                  (0, "LEFT"): ("up", 0), (0, "RIGHT"): ("down", 0), (0, "SELECT"): ("select", 0), (1, "SELECT"): ("select", 0),
                  (0, "RESET"): ("reset", 0), (1, "RESET"): ("reset", 0)}


class Ctx:
    """What the stubs need to know about the evaluation in progress."""
    genome_id = 0
    game = -1
    calls = 0
    hof_pick = (0, 0, 0)
    hof_calls = 0
    hof_canonical = []
    log = None          # list of per-game lists of action vectors


ctx = Ctx()


class FakeRetroEnv:
    def __init__(self, game, state=None, players=1):
        assert game == "Pong-Atari2600"
        self.players = players
        self.state_id = oracle.STATE_START_2P if state == "Start.2P" else oracle.STATE_START_1P
        self.emu = oracle.Atari()
        self.use_restricted_actions = None
        self.closed = False
        ctx.game += 1
        ctx.calls = 0
        ctx.log.append([])

    def reset(self):
        self.emu.reset_to_state(self.state_id)

    def close(self):
        self.closed = True

    def step(self, action):
        a = np.asarray(action)
        assert a.shape == (16,)
        swchb, fire, dec, inc = 0x3F, 0, 0, 0
        for p in range(self.players):
            names = {BUTTONS[i] for i in range(8) if a[8 * p + i] and BUTTONS[i]}
            for x, y in (("UP", "DOWN"), ("LEFT", "RIGHT")):          # retro.Actions.FILTERED
                if x in names and y in names:
                    names -= {x, y}
            for name in names:
                ev = CONSOLE_EVENTS.get((p, name))
                if ev is None: continue
                kind, paddle = ev
                if kind == "fire": fire |= 1 << paddle
                elif kind == "up": dec |= 1 << paddle                  # up = lower resistance
                elif kind == "down": inc |= 1 << paddle
                elif kind == "select": swchb &= ~0x02
                elif kind == "reset": swchb &= ~0x01
        fb = self.emu.run_frame(swchb, fire, dec, inc)
        ram = self.emu.ram
        ctx.log[-1].append(np.array(a, np.uint8))
        return oracle.fb_to_rgb(fb), 0.0, False, {"score1": int(ram[13]), "score2": int(ram[14])}


def philox_choice(n, size=None, replace=True):
    assert n == 2 and size is None
    frame, player = divmod(ctx.calls, 2)
    ctx.calls += 1
    env_id = ctx.genome_id * 6 + ctx.game
    o = oracle.philox4x32([env_id, frame, ((GENERATION << 1) | player) & 0xFFFFFFFF, 0x504F4E47], [SEED & 0xFFFFFFFF, SEED >> 32])
    return int(o[0] & 1)


def picked_shuffle(items):
    pick = ctx.hof_pick[ctx.hof_calls]
    ctx.hof_calls += 1
    canon = ctx.hof_canonical
    items[:] = [canon[pick]] + [x for j, x in enumerate(canon) if j != pick]


class _Fitness:
    def __init__(self, v):
        self.valid = v is not None
        self.values = (v,)


class Individual(list):
    def __init__(self, genes, fit=None):
        super().__init__(genes)
        self.fitness = _Fitness(fit)


class FakeHoF:
    def __init__(self, items):
        self.items = items


def import_reference_main():
    np.int = int                                               # config.py:21,26 on numpy >= 1.24

    class _L:
        def __getattr__(self, n):
            return lambda *a, **k: None

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("scoop", logger=_L(), futures=types.SimpleNamespace(map=map))
    mod("retro", make=FakeRetroEnv, Actions=types.SimpleNamespace(FILTERED="FILTERED"))
    deap = mod("deap")
    deap.algorithms = mod("deap.algorithms")
    deap.tools = mod("deap.tools")
    mod("human_control", HumanInput=object)

    class _Toolbox:
        def register(self, name, fn, *a, **k):
            setattr(self, name, fn)

    mod("ga", toolbox=_Toolbox(), hall_of_fame=None, population=[])
    sys.path.insert(0, REF)
    import main  # noqa: the reference's main.py, unmodified
    import utils

    # utils.find_stuff ends with np.array([ball, left, right]) (utils.py:19).  With one object missing (None) the numpy the
    # reference was written for (< 1.24) built an object array that main.py:81 unpacks; numpy >= 1.24 raises ValueError.
    # The legacy behaviour is restored for utils' view of numpy only.
    class _LegacyNumpy:
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def array(seq, *a, **k):
            try:
                return np.array(seq, *a, **k)
            except ValueError:
                out = np.empty(len(seq), dtype=object)
                for i, v in enumerate(seq):
                    out[i] = v
                return out

    utils.np = _LegacyNumpy()
    return main


def run_evaluate(main, genome, genome_id, hof_mode, hof_items, pick):
    ctx.genome_id = genome_id; ctx.game = -1; ctx.log = []; ctx.hof_calls = 0; ctx.hof_pick = tuple(int(p) for p in pick)
    if hof_mode == 0:
        main.hall_of_fame = None
    elif hof_mode == 1:
        main.hall_of_fame = FakeHoF([])
    else:
        ctx.hof_canonical = list(hof_items)
        main.hall_of_fame = FakeHoF(list(hof_items))
    rewards = []
    orig = main.perform_episode

    def spy(*a, **k):
        r = orig(*a, **k)
        rewards.append(float(r))
        return r

    main.perform_episode = spy
    try:
        fit, = main.evaluate([float(x) for x in genome], render=False)
    finally:
        main.perform_episode = orig
    return float(fit), rewards, ctx.log


def main_():
    oracle.build()
    main = import_reference_main()
    np.random.choice = philox_choice
    random.shuffle = picked_shuffle
    rng = np.random.RandomState(77)

    def tracker(k, a, noise):
        """A genome that follows the ball: h0 = sigmoid(k (ball_y - me_y)), out0/out1 = sigmoid(-/+ a (h0 - 1/2)), perturbed."""
        w0 = np.zeros((2, 7)); w0[0, 1] = k; w0[0, 4] = -k
        w1 = np.array([[-a, 0.0, a / 2], [a, 0.0, -a / 2]])
        g = np.concatenate([w0.ravel(), w1.ravel()])
        return (g + rng.standard_normal(20) * noise).astype(np.float32)

    def draw(kind):
        if kind == 0: return rng.random_sample(20).astype(np.float32)                 # the GA's initial distribution
        if kind == 1: return (rng.standard_normal(20) * 2).astype(np.float32)
        if kind == 2: return (rng.random_sample(20) * 8 - 4).astype(np.float32)
        return tracker(rng.uniform(10, 60), rng.uniform(4, 12), rng.uniform(0.0, 1.5))

    # hall of fame: 4 members with valid fitness (values as a GA would hold them): two ball trackers, two random nets
    hof_genomes = np.stack([draw(3), draw(1), draw(3), draw(2)])
    hof_fitness = np.array([1.75, 0.9, 0.4375, -0.25], np.float64)
    hof_items = [Individual([float(x) for x in g], float(f)) for g, f in zip(hof_genomes, hof_fitness)]

    # 14 genomes against the hall of fame, 3 with hall_of_fame = None and 3 with an empty hall (games 3..5 then fall back to
    # HardcodedAi with multiplier 1, main.py:43-53).  Ball trackers of varying skill (rallies, points for either side) and
    # random nets of three distributions (stationary players make the 2000-frame timeout, main.py:106-107, common; the run
    # asserts that one is present).
    kinds = [3, 3, 1, 3, 0, 3, 2, 3, 1, 3, 0, 3, 2, 3]
    groups = [(2, [(draw(k), rng.randint(0, 4, 3)) for k in kinds]),
              (0, [(draw(k), np.zeros(3, np.int64)) for k in (3, 1, 3)]),
              (1, [(draw(k), np.zeros(3, np.int64)) for k in (3, 2, 0)])]

    out = {"seed": np.uint64(SEED), "generation": np.uint64(GENERATION), "hof_genomes": hof_genomes, "hof_fitness": hof_fitness}
    genomes, modes, picks, gids, fits, rewards, frames, acts, offs = [], [], [], [], [], [], [], [], [0]
    for mode, sel in groups:
        for gid, (g, pick) in enumerate(sel):
            fit, rw, log = run_evaluate(main, g, gid, mode, hof_items, pick)
            assert len(rw) == 6 and len(log) == 6
            genomes.append(g); modes.append(mode); picks.append(pick); gids.append(gid); fits.append(fit); rewards.append(rw)
            frames.append([len(l) for l in log])
            for l in log:
                a = np.stack(l)
                assert (a[:, 0] == 1).all() and (a[:, 15] == 1).all() and not a[:, [1, 2, 3, 8, 9, 10, 11, 12, 13, 14]].any()
                acts.append((a[:, 4] | (a[:, 5] << 1) | (a[:, 6] << 2) | (a[:, 7] << 3)).astype(np.uint8))
                offs.append(offs[-1] + len(l))
            print(f"mode {mode} genome {gid}: fitness {fit:.6f} frames {frames[-1]}", flush=True)
    out.update(genomes=np.stack(genomes), hof_mode=np.array(modes, np.int32), hof_pick=np.stack(picks).astype(np.int32),
               genome_id=np.array(gids, np.int32), fitness=np.array(fits, np.float64), rewards=np.array(rewards, np.float64),
               frames=np.array(frames, np.int32), action_bits=np.concatenate(acts), action_offsets=np.array(offs, np.int64))
    assert (out["frames"] > 2000).any(), "no 2000-frame timeout among the goldens"
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", int(np.sum(frames)), "frames")


if __name__ == "__main__":
    main_()
