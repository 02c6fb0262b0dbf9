"""Worker of tests/test_gpu_multirank.py (launched under torch.distributed.run, one rank per GPU, NCCL): a sharded
run_generations with the asynchronous per-generation exchange; writes what the test compares across ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import neuro_genetic_pong_self_play_b200 as ngp  # noqa: E402
from neuro_genetic_pong_self_play_b200 import parallel  # noqa: E402
from neuro_genetic_pong_self_play_b200.reference_api import Toolbox, run_generations  # noqa: E402


def main():
    out_dir, n_total, ngen = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = ngp.Config(POPULATION_SIZE=n_total, MAX_FRAMES=300)
    eng = ngp.Engine(cfg, device=local)
    tb = Toolbox(cfg, eng, seed=17)
    lo, hi = parallel.shard_bounds(n_total, world, rank)
    # every rank builds the whole initial population from the one seed and keeps its shard: the global population does
    # not depend on the number of ranks
    genomes = tb.population(n_total)[lo:hi].contiguous()
    ex = parallel.Exchange(eng, n_total, cfg.HALL_OF_FAME_AMOUNT)
    genomes, fitness, log = run_generations(tb, genomes, ngen, verbose=False, exchange=ex)
    hg, hf = tb.hall_of_fame.tensors()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), genomes=genomes.cpu().numpy(), fitness=fitness.cpu().numpy(),
             hof_genomes=hg.cpu().numpy(), hof_fitness=hf.cpu().numpy(), lo=lo, hi=hi,
             log=json.dumps(log))
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
