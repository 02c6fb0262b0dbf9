"""oracle/numpy_path.py (the reference's per-frame numpy path restated for bench.py's "port-numpy" CPU baseline) against
the goldens generated from the imported reference."""
import os

import numpy as np

import oracle
from oracle import numpy_path as npp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_find_stuff_and_mlp_match_reference(golden, obs_npy):
    for f, loc, valid in zip(golden["fs_frames"][:12], golden["fs_loc"], golden["fs_valid"]):
        got = npp.find_stuff(f)
        for t in range(3):
            assert (got[t] is not None) == bool(valid[t])
            if valid[t]:
                assert np.array_equal(got[t], loc[t])
    assert [v.tolist() for v in npp.find_stuff(obs_npy)] == [[111.5, 64.5], [122.5, 17.5], [127.5, 141.5]]
    for name in ("mlp_default", "mlp_mid"):
        nodes = [int(n) for n in golden[name + "_nodes"]]
        for g, xs, acts in zip(golden[name + "_genomes"][:8], golden[name + "_x"], golden[name + "_act"]):
            nn = npp.NeuralNetwork(nodes, [float(v) for v in g])
            for x, a in zip(xs, acts):
                assert (1 if nn.run(list(x)) == [1, 0] else 2) == a


def test_perform_episode_matches_reference_run():
    """One whole game of the reference-run goldens (tests/golden/episode_vectors.npz): same number of env.step calls, same reward."""
    ep = np.load(os.path.join(ROOT, "tests", "golden", "episode_vectors.npz"))
    i = int(np.argmin(ep["frames"][:, 0] + 10000 * (ep["hof_mode"] != 2)))
    seed, gen, gid = int(ep["seed"]), int(ep["generation"]), int(ep["genome_id"][i])
    calls = [0]

    def rnd():
        frame, player = divmod(calls[0], 2)
        calls[0] += 1
        return int(oracle.philox4x32([gid * 6, frame, (gen << 1) | player, 0x504F4E47], [seed & 0xFFFFFFFF, seed >> 32])[0] & 1)

    emu = oracle.Atari()
    emu.reset_to_state(oracle.STATE_START_2P)
    steps, reward = npp.perform_episode(emu, npp.HardcodedAi(), npp.NeuralNetwork([6, 2, 2], [float(v) for v in ep["genomes"][i]]), rnd=rnd)
    assert steps == ep["frames"][i][0] and reward == ep["rewards"][i][0]
