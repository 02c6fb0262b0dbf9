#!/usr/bin/env python3
"""Search for an action trace whose emulated frame equals the reference's obs.npy pixel for pixel.

obs.npy (/root/reference/obs.npy, copied as tests/golden/obs.npy) is the one real gym-retro frame the reference ships: score
0-0, ball at rows 144-147 / columns 64-65, left paddle rows 149-164, right paddle rows 154-169.  This tool looks for button
traces from the 'Start.2P' state (main.py:21,51) that make the ORACLE emulator produce exactly that frame:

  stage A  two "tracking" controllers (each paddle follows the ball with a random, occasionally re-drawn offset, so rallies
           last and the ball leaves the paddles at many angles) are played from many seeds until a frame has the ball in
           obs.npy's box with the score still 0-0;
  stage B  everything after the ball's last reversal before that frame is free for the paddles: random piecewise-constant
           controls per paddle (they do not interact before the frame) until each paddle sits in obs.npy's rows.

Found with seeds 0..1499: seed 1406 puts the ball in place at frame 352 (last reversal at frame 341); stage B seeds 805 (left)
and 36 (right) place the paddles; the resulting frame equals obs.npy in all 210 x 160 x 3 bytes.  The trace is committed as
tests/golden/obs_trace.npz and replayed by tests/test_oracle_golden.py (oracle) and tests/test_gpu_parity.py (CUDA cores).
What this pins against a real gym-retro/Stella frame: the whole rendering (object sizes, columns, colours, playfield,
score digits, the row parity of the two paddles), and that this frame is REACHABLE by the emulated game's own ball and paddle
dynamics from the emulated start state.  Facts found on the way: at column 64 the ball only ever sits on rows = 4 (mod 5);
the left paddle's top row is always odd, the right paddle's always even -- obs.npy obeys all three.

    python tools/search_obs_trace.py [--seeds 1500] [--out tests/golden/obs_trace.npz]
"""
import argparse
import concurrent.futures as cf
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TARGET_BALL, TARGET_LEFT, TARGET_RIGHT = (144, 147, 64, 65), (149, 164), (154, 169)
COLU_BALL, COLU_LEFT, COLU_RIGHT = 0x0E, 0x38, 0xC8          # TIA colour values of (236,236,236), (213,130,74), (92,186,92)


def box(fb, colu):
    ys, xs = np.nonzero(fb[34:194] == colu)
    return None if len(ys) == 0 else (int(ys.min()) + 34, int(ys.max()) + 34, int(xs.min()), int(xs.max()))


def buttons(left, right):
    a = np.zeros(16, np.uint8)
    a[0] = a[15] = 1                                           # BLANK_ACTION, config.py:21-23
    a[4] = right == 1; a[5] = right == 2; a[6] = left == 1; a[7] = left == 2
    return a


def stage_a(seed, max_frames=1500):
    """Tracking controllers with random offsets.  Returns (seed, [frames where the ball is in place], controls)."""
    import oracle
    rng = np.random.RandomState(seed)
    env = oracle.Atari(); env.reset_to_state(oracle.STATE_START_2P)
    off_l, off_r = rng.randint(-9, 10), rng.randint(-9, 10)
    ctl, hits = [], []
    l = r = 0
    b = lb = rb = None
    for f in range(max_frames):
        if b is not None and lb is not None and rb is not None:
            bc = (b[0] + b[1]) / 2
            lc, rc = (lb[0] + lb[1]) / 2 + off_l, (rb[0] + rb[1]) / 2 + off_r
            l = 1 if bc < lc - 2 else (2 if bc > lc + 2 else 0)
            r = 1 if bc < rc - 2 else (2 if bc > rc + 2 else 0)
        if rng.rand() < 0.02:
            off_l = rng.randint(-9, 10)
        if rng.rand() < 0.02:
            off_r = rng.randint(-9, 10)
        ctl.append((l, r))
        fb = env.step(buttons(l, r))
        if env.ram[13] or env.ram[14]:
            break
        b, lb, rb = box(fb, COLU_BALL), box(fb, COLU_LEFT), box(fb, COLU_RIGHT)
        if b == TARGET_BALL:
            hits.append(f)
    return seed, hits, ctl


def replay(ctl, upto):
    import oracle
    env = oracle.Atari(); env.reset_to_state(oracle.STATE_START_2P)
    xs, fb = [], None
    for f in range(upto + 1):
        fb = env.step(buttons(*ctl[f]))
        b = box(fb, COLU_BALL)
        xs.append(b[2] if b else None)
    return fb, xs


def stage_b(args):
    ctl, frame, free_from, side, seed = args
    rng = np.random.RandomState(seed)
    ctl = list(ctl[:frame + 1])
    f = free_from
    while f <= frame:
        d, k = rng.randint(0, 3), rng.randint(1, 25)
        for g in range(f, min(frame + 1, f + k)):
            ctl[g] = (d, ctl[g][1]) if side == 0 else (ctl[g][0], d)
        f += k
    fb, _ = replay(ctl, frame)
    pb = box(fb, COLU_LEFT if side == 0 else COLU_RIGHT)
    ok = box(fb, COLU_BALL) == TARGET_BALL and pb is not None and pb[:2] == (TARGET_LEFT if side == 0 else TARGET_RIGHT)
    return ok, [c[side] for c in ctl[free_from:frame + 1]]


def main():
    import oracle
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=1500)
    ap.add_argument("--tries", type=int, default=1200)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "obs_trace.npz"))
    args = ap.parse_args()
    obs = np.load(os.path.join(ROOT, "tests", "golden", "obs.npy"))
    workers = os.cpu_count() or 1
    with cf.ProcessPoolExecutor(workers) as ex:
        for seed, hits, ctl in ex.map(stage_a, range(args.seeds), chunksize=4):
            for frame in hits:
                _, xs = replay(ctl, frame)
                free_from = 0
                for f in range(2, frame + 1):
                    if None not in (xs[f], xs[f - 1], xs[f - 2]) and (xs[f] - xs[f - 1]) * (xs[f - 1] - xs[f - 2]) < 0:
                        free_from = f + 1
                print(f"stage A: seed {seed} ball in place at frame {frame}, paddles free from frame {free_from}", flush=True)
                sol = {}
                for side in (0, 1):
                    for ok, part in ex.map(stage_b, [(ctl, frame, free_from, side, s) for s in range(args.tries)], chunksize=8):
                        if ok:
                            sol[side] = part
                            break
                if len(sol) < 2:
                    continue
                final = list(ctl[:frame + 1])
                for i, g in enumerate(range(free_from, frame + 1)):
                    final[g] = (sol[0][i], sol[1][i])
                fb, _ = replay(final, frame)
                rgb = oracle.fb_to_rgb(fb)
                diff = int((rgb != obs).any(axis=-1).sum())
                print(f"stage B: paddles placed; {diff} pixels differ from obs.npy", flush=True)
                if diff == 0:
                    acts = np.stack([buttons(*c) for c in final])
                    np.savez_compressed(args.out, actions=acts, state=np.int32(oracle.STATE_START_2P), seed=np.int32(seed))
                    print("wrote", args.out, acts.shape)
                    return 0
    print("no pixel-exact trace found; closest misses are listed above")
    return 1


if __name__ == "__main__":
    sys.exit(main())
