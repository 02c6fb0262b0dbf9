#!/usr/bin/env python3
"""Record gym-retro traces of the reference's two environments for tests/test_retro_trace.py.

Run this where gym-retro (with the Pong-Atari2600 ROM imported) is installed -- it is NOT installable in the build image
(no network, not in the wheelhouse), which is why tests/golden/retro_traces/ ships empty and the replay test skips itself.
It drives the environments exactly as the reference does (main.py:21-24, 40, 51-56, 77):

    env = retro.make('Pong-Atari2600', state='Start.2P', players=2)      # or retro.make('Pong-Atari2600') for the robot game
    env.use_restricted_actions = retro.Actions.FILTERED
    env.reset()
    observation, reward, is_done, info = env.step(action)               # action: 16 (2 players) or 8 buttons

and stores, per trace, one .npz with
    state        'Start' | 'Start.2P'
    players      1 | 2
    actions      u8[T][16]   the button vector passed to env.step (players=1: only the first 8 are passed on)
    ram          u8[T][128]  env.get_ram() after each step
    score        i32[T][2]   info['score1'], info['score2']
    frame_crc    u32[T]      zlib.crc32 of the (210,160,3) observation
    frames       u8[K][210][160][3] + frame_index i32[K]: every `--keep-every`-th observation in full
    reset_obs    u8[210][160][3]  the observation env.reset() returned

    python tools/record_retro_trace.py --out tests/golden/retro_traces --frames 2100 --traces 4
"""
import argparse
import os
import zlib

import numpy as np


def random_actions(rng, n_frames, hold):
    """BLANK_ACTION (config.py:21-23) with random paddle decisions written to a[4:6] / a[6:8] (main.py:91-92), each held `hold` frames."""
    acts = np.zeros((n_frames, 16), np.uint8)
    acts[:, 0] = 1; acts[:, 15] = 1
    cur = (0, 0)
    for f in range(n_frames):
        if f % hold == 0:
            cur = (rng.randint(0, 3), rng.randint(0, 3))
        r, l = cur
        acts[f, 4] = r == 1; acts[f, 5] = r == 2; acts[f, 6] = l == 1; acts[f, 7] = l == 2
    return acts


def record(state, players, actions, keep_every):
    import retro
    if players == 1:
        env = retro.make('Pong-Atari2600')                                   # main.py:40
    else:
        env = retro.make('Pong-Atari2600', state=state, players=players)     # main.py:21, 51
    env.use_restricted_actions = retro.Actions.FILTERED                      # main.py:23, 55
    reset_obs = env.reset()
    T = len(actions)
    ram = np.zeros((T, 128), np.uint8); score = np.zeros((T, 2), np.int32); crc = np.zeros(T, np.uint32)
    frames, index = [], []
    for f in range(T):
        a = actions[f][:8 * players]
        obs, _, done, info = env.step(a)
        ram[f] = np.asarray(env.get_ram(), np.uint8)[:128]
        score[f] = (info.get('score1', -1), info.get('score2', -1))
        crc[f] = zlib.crc32(np.ascontiguousarray(obs, np.uint8).tobytes())
        if f % keep_every == 0:
            frames.append(np.asarray(obs, np.uint8).copy()); index.append(f)
        if done:
            ram, score, crc, actions = ram[:f + 1], score[:f + 1], crc[:f + 1], actions[:f + 1]
            break
    env.close()
    return dict(state=state, players=players, actions=actions, ram=ram, score=score, frame_crc=crc,
                frames=np.stack(frames), frame_index=np.asarray(index, np.int32), reset_obs=np.asarray(reset_obs, np.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/retro_traces")
    ap.add_argument("--frames", type=int, default=2100)
    ap.add_argument("--traces", type=int, default=4, help="per start state")
    ap.add_argument("--keep-every", type=int, default=25)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    rng = np.random.RandomState(20260)
    for state, players in (("Start.2P", 2), ("Start", 1)):
        for t in range(args.traces):
            acts = random_actions(rng, args.frames, hold=1 + 3 * t)
            rec = record(state, players, acts, args.keep_every)
            path = os.path.join(args.out, f"{state.replace('.', '_')}_{t}.npz")
            np.savez_compressed(path, **rec)
            print(path, len(rec["ram"]), "frames")


if __name__ == "__main__":
    main()
