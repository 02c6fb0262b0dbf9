"""Replay of recorded gym-retro traces (tools/record_retro_trace.py: actions, RAM, scores, frame CRCs and sampled frames of
env.step, main.py:77) through the oracle emulator (CPU) and the CUDA core (GPU): RAM and every pixel must match bit for bit.

gym-retro cannot be installed in the build image, so tests/golden/retro_traces/ ships EMPTY and both tests skip themselves;
they are the harness that turns "emulator parity unpinned"  into a checked fact as soon as somebody with a
gym-retro install drops traces into that directory."""
import glob
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "retro_traces", "*.npz")))
needs_traces = pytest.mark.skipif(not TRACES, reason="no gym-retro traces recorded (tests/golden/retro_traces is empty)")


def _state_id(tr):
    import oracle
    return oracle.STATE_START_1P if int(tr["players"]) == 1 else oracle.STATE_START_2P


def _replay(path):
    import oracle
    tr = np.load(path)
    env = oracle.Atari()
    env.reset_to_state(_state_id(tr))
    keep = {int(f): i for i, f in enumerate(tr["frame_index"])}
    for f, a in enumerate(tr["actions"]):
        rgb = oracle.fb_to_rgb(env.step(a))
        assert np.array_equal(env.ram, tr["ram"][f]), f"RAM differs at frame {f}: bytes {np.nonzero(env.ram != tr['ram'][f])[0]}"
        assert tuple(env.ram[13:15]) == tuple(tr["score"][f]), f"score at frame {f}"
        assert zlib.crc32(rgb.tobytes()) == int(tr["frame_crc"][f]), f"frame pixels differ at frame {f}"
        if f in keep:
            assert np.array_equal(rgb, tr["frames"][keep[f]])


@needs_traces
@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p) for p in TRACES])
def test_oracle_replays_retro_trace(path):
    _replay(path)


@needs_traces
@pytest.mark.gpu
@pytest.mark.parametrize("core", [0, 1])
def test_cuda_replays_retro_traces(core):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import neuro_genetic_pong_self_play_b200 as ngp
    eng = ngp.Engine(ngp.Config(), device=0)
    for players in (1, 2):
        traces = [np.load(p) for p in TRACES]
        traces = [t for t in traces if int(t["players"]) == players]
        if not traces:
            continue
        T = min(len(t["actions"]) for t in traces)
        eng.env_reset(len(traces), ngp.STATE_START_1P if players == 1 else ngp.STATE_START_2P)
        for f in range(T):
            a = torch.from_numpy(np.stack([t["actions"][f] for t in traces])).cuda()
            out = eng.env_step(a, core=core)
            ram = out["ram"].cpu().numpy(); frames = out["frames"].cpu().numpy()
            for j, t in enumerate(traces):
                assert np.array_equal(ram[j], t["ram"][f]), f"RAM differs, trace {j} frame {f}"
                assert zlib.crc32(frames[j].tobytes()) == int(t["frame_crc"][f]), f"pixels differ, trace {j} frame {f}"
    eng.close()


def test_recorder_and_replay_agree_on_a_stand_in_env(tmp_path, monkeypatch):
    """The harness itself: tools/record_retro_trace.py run against a stand-in `retro` module backed by the oracle emulator
    writes traces that the replay above accepts (file layout, players=1 button truncation, reset frame)."""
    import importlib.util
    import sys
    import types

    import oracle

    class Env:
        def __init__(self, state, players):
            self.emu = oracle.Atari(); self.state = state; self.players = players
        def reset(self):
            self.emu.reset_to_state(oracle.STATE_START_1P if self.players == 1 else oracle.STATE_START_2P)
            return np.zeros((210, 160, 3), np.uint8)
        def step(self, a):
            full = np.zeros(16, np.uint8); full[:len(a)] = a
            rgb = oracle.fb_to_rgb(self.emu.step(full))
            ram = self.emu.ram
            return rgb, 0.0, False, {"score1": int(ram[13]), "score2": int(ram[14])}
        def get_ram(self):
            return self.emu.ram
        def close(self):
            pass

    fake = types.ModuleType("retro")
    fake.make = lambda game, state="Start", players=1: Env(state, players)
    fake.Actions = types.SimpleNamespace(FILTERED=2)
    monkeypatch.setitem(sys.modules, "retro", fake)
    spec = importlib.util.spec_from_file_location("record_retro_trace", os.path.join(ROOT, "tools", "record_retro_trace.py"))
    rec = importlib.util.module_from_spec(spec); spec.loader.exec_module(rec)
    rng = np.random.RandomState(1)
    for state, players in (("Start.2P", 2), ("Start", 1)):
        tr = rec.record(state, players, rec.random_actions(rng, 150, hold=4), keep_every=50)
        path = tmp_path / f"{players}.npz"
        np.savez_compressed(path, **tr)
        _replay(str(path))
