#!/usr/bin/env python3
"""Generate tests/golden/reference_vectors.npz by importing the reference's own numpy code.

Runs only in the build container (needs /root/reference).  The reference targets numpy < 1.24
and SCOOP, so two shims are installed first: ``np.int = int`` (config.py:21,26) and a stub
``scoop`` module exposing ``logger`` (utils.py:7, numpy_nn.py:2).  Nothing from the reference
is copied: only its *outputs* on seeded inputs are stored.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_vectors.npz")


def import_reference():
    np.int = int  # noqa: shim for numpy >= 1.24
    scoop = types.ModuleType("scoop")

    class _L:
        def __getattr__(self, n):
            return lambda *a, **k: None

    scoop.logger = _L()
    sys.modules["scoop"] = scoop
    sys.path.insert(0, REF)
    import config, utils, numpy_nn, dumb_ais  # noqa
    return config, utils, numpy_nn, dumb_ais


def synth_frame(rng, with_ball=True, with_left=True, with_right=True, extra_noise=False):
    """obs.npy-shaped frame: BG + walls + optional ball / paddles at random places."""
    f = np.zeros((210, 160, 3), np.uint8)
    f[:] = (144, 72, 17)
    f[24:34] = (236, 236, 236)
    f[194:210] = (236, 236, 236)
    if with_ball:
        r, c = rng.randint(34, 190), rng.randint(0, 158)
        f[r:r + 4, c:c + 2] = (236, 236, 236)
    if with_left:
        r = rng.randint(34, 178)
        f[r:r + 16, 16:20] = (213, 130, 74)
    if with_right:
        r = rng.randint(34, 178)
        f[r:r + 16, 140:144] = (92, 186, 92)
    if extra_noise:  # single-channel matches: exercises the per-channel semantics of get_rect_quickly
        for _ in range(20):
            r, c = rng.randint(34, 194), rng.randint(0, 160)
            f[r, c] = (236, rng.randint(0, 255), 74)
    return f


def main():
    config, utils, numpy_nn, dumb_ais = import_reference()
    rng = np.random.RandomState(1234)
    out = {}

    # --- find_stuff / get_rect_quickly (utils.py:14-19, 60-68) ---------------------------
    frames, locs, valids = [], [], []
    obs = np.load(os.path.join(REF, "obs.npy"))
    cases = [obs, np.zeros_like(obs)]
    for i in range(30):
        cases.append(synth_frame(rng, with_ball=i % 5 != 0, with_left=i % 7 != 0, with_right=i % 11 != 0,
                                 extra_noise=i % 3 == 0))
    for f in cases:
        chopped = f[config.GAME_TOP:config.GAME_BOTTOM, :]
        loc = np.zeros((3, 2)); valid = np.zeros(3, np.uint8)
        for t, col in enumerate((config.BALL_COLOUR, config.LEFT_GUY_COLOUR, config.RIGHT_GUY_COLOUR)):
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    v = utils.get_rect_quickly(chopped, col)
            if v is not None:
                loc[t] = v; valid[t] = 1
        frames.append(f); locs.append(loc); valids.append(valid)
    out["fs_frames"] = np.stack(frames); out["fs_loc"] = np.stack(locs); out["fs_valid"] = np.stack(valids)

    # --- NeuralNetwork.run (numpy_nn.py:120-137) ------------------------------------------
    for name, nodes, n_genomes, scale in (("mlp_default", [6, 2, 2], 64, 1.0), ("mlp_default_wide_range", [6, 2, 2], 64, 8.0),
                                          ("mlp_mid", [6, 16, 16, 2], 16, 1.0)):
        G = sum((nodes[i] + 1) * nodes[i + 1] for i in range(len(nodes) - 1))
        genomes = ((rng.random_sample((n_genomes, G)) - (0.0 if scale == 1.0 else 0.5)) * scale).astype(np.float32)
        xs = rng.random_sample((n_genomes, 8, 6))
        outs = np.zeros((n_genomes, 8, nodes[-1])); acts = np.zeros((n_genomes, 8), np.uint8)
        for g in range(n_genomes):
            nn = numpy_nn.NeuralNetwork(nodes=nodes, weights=[float(v) for v in genomes[g]], bias=True)
            for e in range(8):
                a = nn.run(list(xs[g, e]))
                outs[g, e] = nn.list_of_transitional_arrays[-1][:-1]
                acts[g, e] = 1 if a == [1, 0] else 2
        out[name + "_nodes"] = np.array(nodes); out[name + "_genomes"] = genomes; out[name + "_x"] = xs
        out[name + "_out"] = outs; out[name + "_act"] = acts
    # wide net: genomes regenerated from the seed in the test (267 266 genes each)
    nodes = [6, 512, 512, 2]
    G = sum((nodes[i] + 1) * nodes[i + 1] for i in range(len(nodes) - 1))
    wide_out = np.zeros((4, 4, 2)); wide_act = np.zeros((4, 4), np.uint8); wide_x = rng.random_sample((4, 4, 6))
    for g in range(4):
        genome = (np.random.RandomState(9000 + g).standard_normal(G) * 0.05).astype(np.float32)
        nn = numpy_nn.NeuralNetwork(nodes=nodes, weights=[float(v) for v in genome], bias=True)
        for e in range(4):
            a = nn.run(list(wide_x[g, e]))
            wide_out[g, e] = nn.list_of_transitional_arrays[-1][:-1]
            wide_act[g, e] = 1 if a == [1, 0] else 2
    out["mlp_wide_x"] = wide_x; out["mlp_wide_out"] = wide_out; out["mlp_wide_act"] = wide_act
    # saturation tie rule
    nn = numpy_nn.NeuralNetwork(nodes=[6, 2, 2], weights=[0.0] * 14 + [50.0] * 3 + [60.0] * 3, bias=True)
    out["mlp_saturation_act"] = np.array(1 if nn.run([0.5] * 6) == [1, 0] else 2)

    # --- inference vector on obs.npy (utils.py:139-153) -----------------------------------
    class Spy:
        def run(self, v):
            self.v = list(v); return [0, 0]
    spy = Spy()
    ball, left, right = out["fs_loc"][0]
    utils.inference(ball, ball, right, left, spy)
    out["inference_right_on_obs"] = np.array(spy.v)

    # --- reward / clamp / bots ------------------------------------------------------------
    rw_in = np.array([[1, 1234.0, 3, 1], [0.37, 500.0, 1, 3], [2.5, 77.0, 3, 0], [1, 2001.0, 0, 3]], np.float64)
    out["reward_in"] = rw_in
    out["reward_out"] = np.array([utils.calculate_reward(m, t, int(a), int(b)) for m, t, a, b in rw_in])
    ys = np.array([0.0, 15.5, 16.0, 16.5, 80.0, 143.5, 144.0, 144.5, 159.0])
    cl = np.zeros((len(ys), 3), np.uint8)
    enc = lambda a: 0 if list(a) == [0, 0] else (1 if list(a) == [1, 0] else 2)
    for i, y in enumerate(ys):
        for j, act in enumerate(([0, 0], [1, 0], [0, 1])):
            cl[i, j] = enc(utils.keep_within_game_bounds_please([y, 17.5], act))
    out["clamp_y"] = ys; out["clamp_out"] = cl
    xs = rng.random_sample((32, 6)); xs[:4, 1] = xs[:4, 4]
    hb = np.array([enc(dumb_ais.HardcodedAi().run(list(x))) for x in xs], np.uint8)
    sb = np.zeros((32, 2), np.uint8)
    for i, x in enumerate(xs):
        for j, sc in enumerate(({"score1": 0, "score2": 1}, {"score1": 2, "score2": 1})):
            ai = dumb_ais.ScoreHardcodedAi(); ai.set_score(sc); sb[i, j] = enc(ai.run(list(x)))
    out["bots_x"] = xs; out["bots_hard"] = hb; out["bots_score"] = sb
    out["gene_size_default"] = np.array(utils.calculate_gene_size())

    np.savez_compressed(OUT, **out)
    print("wrote", os.path.abspath(OUT), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
