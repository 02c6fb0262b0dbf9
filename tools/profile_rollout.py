#!/usr/bin/env python3
"""Short, fixed-length invocation of the fused rollout for ncu (one warm launch, one profiled)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import neuro_genetic_pong_self_play_b200 as ngp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--population", type=int, default=1024)
ap.add_argument("--max-frames", type=int, default=120)
ap.add_argument("--launches", type=int, default=2)
ap.add_argument("--games", type=int, default=6, help="games per genome (population 1, games 1 = a single live lane)")
a = ap.parse_args()
cfg = ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, POPULATION_SIZE=a.population, MAX_FRAMES=a.max_frames, GAMES_TO_PLAY=a.games)
eng = ngp.Engine(cfg, device=0)
g = eng.init_population(a.population, seed=1)
for i in range(a.launches - 1):          # warm launches (instruction cache, clocks); only the last one is timed
    eng.evaluate(g, seed=3, generation=i)
eng.profile_enable(True)
out = eng.evaluate(g, seed=3, generation=a.launches - 1)
ms, n = eng.profile_read()
print(f"population={a.population} envs={a.population * a.games} max_frames={a.max_frames} frames/launch={out['frames_total']} "
      f"rollout_ms/launch={ms / n:.3f} frames/s={out['frames_total'] / (ms / n * 1e-3):.4g}")
