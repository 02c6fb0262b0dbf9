"""A/B of the tail compaction (csrc/ngp_core.cu) on one GPU: the same evolved population evaluated with and without it.
    python tools/profile_compaction.py [population] [generations of evolution first]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neuro_genetic_pong_self_play_b200 as ngp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
evolve = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = ngp.Engine(ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, POPULATION_SIZE=n), device=0)
g = eng.init_population(n, seed=1234)
for gen in range(evolve):
    out = eng.evaluate(g, seed=99, generation=gen, sync=False)
    g = eng.ga_step(g, out["fitness"], seed=99, generation=gen)["genomes"]
torch.cuda.synchronize()
res = {}
for name, opt in (("compact", 0), ("single_launch", 1), ("compact_again", 0)):
    eng.set_option("rollout_nocompact", opt)
    eng.profile_enable(True); eng.profile_read()
    frames = 0
    for rep in range(3):
        frames = eng.evaluate(g, seed=99, generation=evolve)["frames_total"]
    ms, k = eng.profile_read()
    res[name] = {"kernel_ms": ms / k, "env_frames_per_s": frames / (ms / k * 1e-3), "frames": frames}
print(json.dumps({"population": n, "envs": n * 6, "evolved_generations": evolve, **res}))
