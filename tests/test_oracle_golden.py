"""The CPU oracle against the reference's own outputs (tests/golden/, made by tools/make_golden.py
from the imported reference code) and against obs.npy, the only emulator-side fixture the
reference ships."""
import hashlib
import os

import numpy as np

import oracle


def test_rom_fixture_hashes():
    rom = oracle.load_rom()
    assert hashlib.sha1(rom).hexdigest() == "1ffe89d79d55adabc0916b95cc37e18619ef7830"
    assert hashlib.md5(rom).hexdigest() == "60e0ea3cbe0913d39803477945e9e5ec"
    assert rom[-6:] == bytes([0x11, 0x11, 0x00, 0xF0, 0x38, 0xF4])


def test_find_stuff_matches_reference(golden, obs_npy):
    for f, loc, valid in zip(golden["fs_frames"], golden["fs_loc"], golden["fs_valid"]):
        l, v = oracle.find_stuff(f)
        assert np.array_equal(v, valid)
        assert np.array_equal(l[valid == 1], loc[valid == 1])      # bit-exact float64
    l, v = oracle.find_stuff(obs_npy)
    assert l.tolist() == [[111.5, 64.5], [122.5, 17.5], [127.5, 141.5]] and v.tolist() == [1, 1, 1]
    l, v = oracle.find_stuff(np.zeros_like(obs_npy))
    assert v.tolist() == [0, 0, 0]


def test_mlp_matches_numpy_nn(golden):
    for name in ("mlp_default", "mlp_default_wide_range", "mlp_mid"):
        nodes = [int(n) for n in golden[name + "_nodes"]]
        for g, xs, outs, acts in zip(golden[name + "_genomes"], golden[name + "_x"], golden[name + "_out"], golden[name + "_act"]):
            for x, o, a in zip(xs, outs, acts):
                out, act = oracle.mlp_forward(nodes, g, x)
                np.testing.assert_allclose(out, o, rtol=1e-12, atol=0)
                assert act == a
    # saturation tie: both outputs == 1.0 -> argmax 0 -> up
    _, act = oracle.mlp_forward([6, 2, 2], [0.0] * 14 + [50.0] * 3 + [60.0] * 3, [0.5] * 6)
    assert act == int(golden["mlp_saturation_act"]) == oracle.ACT_UP


def test_mlp_wide_matches_numpy_nn(golden):
    nodes = [6, 512, 512, 2]
    G = sum((nodes[i] + 1) * nodes[i + 1] for i in range(3))
    assert G == 267266
    for g in range(4):
        genome = (np.random.RandomState(9000 + g).standard_normal(G) * 0.05).astype(np.float32)
        for e in range(4):
            out, act = oracle.mlp_forward(nodes, genome, golden["mlp_wide_x"][g, e])
            np.testing.assert_allclose(out, golden["mlp_wide_out"][g, e], rtol=1e-10)
            assert act == golden["mlp_wide_act"][g, e]


def test_det_exp_accuracy():
    xs = np.concatenate([np.linspace(-700, 700, 20001), np.random.RandomState(0).uniform(-40, 40, 20000)])
    got = np.array([oracle.det_exp(x) for x in xs])
    np.testing.assert_allclose(got, np.exp(xs), rtol=4e-16 * 4)
    assert oracle.det_exp(800.0) == np.inf and oracle.det_exp(-800.0) == 0.0


def test_reward_clamp_bots(golden):
    L = oracle.lib()
    for (m, t, a, b), r in zip(golden["reward_in"], golden["reward_out"]):
        assert L.eo_reward(m, t, int(a), int(b)) == r
    assert L.eo_reward(1.0, 1234.0, 3, 1) == 0.4051863857374392
    for y, row in zip(golden["clamp_y"], golden["clamp_out"]):
        for act in range(3):
            assert L.eo_clamp(1, float(y), act) == row[act]
            assert L.eo_clamp(0, float(y), act) == act
    sh = oracle.Shape.make([6, 2, 2])
    assert L.eo_gene_size(sh) == int(golden["gene_size_default"]) == 20
    # dumb_ais.HardcodedAi / ScoreHardcodedAi on the reference's own outputs (32 vectors, 4 of them with ball_y == me_y;
    # scores 0-1 = the bot acts, 2-1 = it stands still)
    for x, hb, sb in zip(golden["bots_x"], golden["bots_hard"], golden["bots_score"]):
        assert oracle.bot_act(oracle.POLICY_HARDCODED, x) == hb
        assert oracle.bot_act(oracle.POLICY_SCORE_HARDCODED, x, 0, 1) == sb[0]
        assert oracle.bot_act(oracle.POLICY_SCORE_HARDCODED, x, 2, 1) == sb[1]
    assert set(golden["bots_hard"].tolist()) == {0, 1, 2} and set(golden["bots_score"][:, 1].tolist()) == {0}


def test_inference_vector_matches_reference(golden, obs_npy):
    """utils.inference's six-vector for the right player on obs.npy (last ball = ball), bit-exact float64."""
    loc, valid = oracle.find_stuff(obs_npy)
    x = oracle.inference_vector(loc[0], loc[0], loc[2][0], loc[1][0])
    assert np.array_equal(x, golden["inference_right_on_obs"])
    assert x.tolist() == [0.403125, 0.696875, 0.403125, 0.696875, 0.796875, 0.765625]


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert oracle.philox4x32([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_emulator_frame_geometry_matches_obs_npy(obs_npy):
    """obs.npy is the one frame the reference ships: every static row (score digits, walls)
    must be pixel-identical and the three objects must have its sizes and columns."""
    env = oracle.Atari()
    env.reset_to_state(oracle.STATE_START_2P)
    for _ in range(80):
        fb = env.step([1] + [0] * 14 + [1])
    rgb = oracle.fb_to_rgb(fb)
    assert np.array_equal(rgb[:34], obs_npy[:34])
    assert np.array_equal(rgb[194:], obs_npy[194:])
    cols = {(236, 236, 236): (4, 2, None), (213, 130, 74): (16, 4, 16), (92, 186, 92): (16, 4, 140)}
    for col, (h, w, x0) in cols.items():
        for img in (rgb, obs_npy):
            ys, xs = np.nonzero(np.all(img[34:194] == col, axis=-1))
            assert ys.max() - ys.min() + 1 == h and xs.max() - xs.min() + 1 == w and len(ys) == h * w
            if x0 is not None:
                assert xs.min() == x0
    u1, c1 = np.unique(rgb.reshape(-1, 3), axis=0, return_counts=True)
    u2, c2 = np.unique(obs_npy.reshape(-1, 3), axis=0, return_counts=True)
    assert np.array_equal(u1, u2) and np.array_equal(c1, c2)


def test_emulator_frame_timing_and_states():
    env = oracle.Atari()
    env.reset_to_state(oracle.STATE_START_1P)
    assert env.ram[0x16] == 0 and env.ram[13] == 0 and env.ram[14] == 0 and env.ram[0x10] == 0xFF
    env.reset_to_state(oracle.STATE_START_2P)
    assert env.ram[0x16] == 2 and env.ram[13] == 0 and env.ram[14] == 0 and env.ram[0x10] == 0xFF
    c0 = env.cycles
    env.step([0] * 16, want_frame=False)
    assert env.cycles - c0 == 262 * 76          # one NTSC frame


def test_episode_bots_play_to_win_score():
    env = oracle.Atari()
    env.reset_to_state(oracle.STATE_START_2P)
    res, trace = env.episode(oracle.Shape.make([6, 2, 2]), ("hardcoded", None), ("hardcoded", None), trace_cap=4096)
    assert max(res.score1, res.score2) == 3 and res.frames == len(trace)
    sc = trace[:, 130:132].astype(int)
    assert np.all(np.diff(sc, axis=0) >= 0)
    assert res.reward == oracle.lib().eo_reward(1.0, res.total_frames, res.score2, res.score1)


def test_emulator_reproduces_obs_npy_pixel_for_pixel(obs_npy):
    """tests/golden/obs_trace.npz (found by tools/search_obs_trace.py): 353 env.step calls from 'Start.2P' after which the
    emulated frame equals the reference's real gym-retro frame obs.npy in every one of its 210 x 160 x 3 bytes -- rendering,
    colours, score digits, and ball / paddle positions reachable by the emulated game's own dynamics."""
    tr = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs_trace.npz"))
    env = oracle.Atari()
    env.reset_to_state(int(tr["state"]))
    for a in tr["actions"]:
        fb = env.step(a)
    assert np.array_equal(oracle.fb_to_rgb(fb), obs_npy)
    assert env.ram[13] == 0 and env.ram[14] == 0
    loc, valid = oracle.find_stuff(oracle.fb_to_rgb(fb))
    assert valid.tolist() == [1, 1, 1] and loc.tolist() == [[111.5, 64.5], [122.5, 17.5], [127.5, 141.5]]
