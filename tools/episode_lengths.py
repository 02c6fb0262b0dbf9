#!/usr/bin/env python3
"""Why a generation of config 2 takes what it takes: episode-length distribution per generation against the rollout's
kernel time (population 1 024, round-robin, six games per genome), and how the long games sit in the warps."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import neuro_genetic_pong_self_play_b200 as ngp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, POPULATION_SIZE=n)
eng = ngp.Engine(cfg, device=0)
for kv in sys.argv[3:]:                      # tuning switches for geometry sweeps: rollout_block=64 rollout_nosync=1 ...
    k, v = kv.split("=")
    eng.set_option(k, int(v))
g = eng.init_population(n, seed=1)
eng.profile_enable(True)
for gen in range(gens):
    out = eng.evaluate(g, seed=3, generation=gen, want_detail=True)
    ms, k = eng.profile_read()
    f = out["frames"].cpu().numpy().astype(np.int64).ravel()
    per_warp = f.reshape(-1, 32)
    longest_warp = per_warp.max(axis=1).argmax()
    w = np.sort(per_warp[longest_warp])[::-1]
    live_at = [(per_warp[longest_warp] > t).sum() for t in (0, 250, 500, 1000, 1500, 2000, 3000)]
    print(json.dumps({"generation": gen, "kernel_ms": round(ms / max(1, k), 2), "frames": int(f.sum()), "mean": round(float(f.mean()), 1),
                      "p50": int(np.percentile(f, 50)), "p90": int(np.percentile(f, 90)), "p99": int(np.percentile(f, 99)), "max": int(f.max()),
                      "ms_per_max_frame": round(ms / max(1, k) / f.max(), 4), "longest_warp_top8": w[:8].tolist(),
                      "longest_warp_live_lanes_at_0_250_500_1000_1500_2000_3000": [int(x) for x in live_at],
                      "envs_longer_than_1000": int((f > 1000).sum()), "envs_longer_than_2000": int((f > 2000).sum())}), flush=True)
    g = eng.ga_step(g, out["fitness"], seed=1, generation=gen)["genomes"]
