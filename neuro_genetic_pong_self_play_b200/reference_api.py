"""The reference's call surface, batched underneath.

Names and argument meaning follow the reference so its call sites read the same:
  NeuralNetwork(nodes, weights=..., bias=...).run(vec6)        numpy_nn.py:35-50, 120-137
  find_stuff(observation)                                      utils.py:14-19
  evaluate(individual) -> (fitness,)                           main.py:28-66
  toolbox.select / mate / mutate / map / evaluate              ga.py:83-94
  run_generations(...)  (eaSimple-shaped loop + stats)         main.py:157-173
  save_checkpoint / load_latest_population                     utils.py:116-125, ga.py:32-53
Every call lands in libngp.so's CUDA kernels through Engine; nothing here computes on the CPU except
list/array marshalling and the hall-of-fame bookkeeping (a few hundred comparisons per generation)."""
from __future__ import annotations

import glob
import os
import time
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import Config
from .engine import Engine

_engines = {}


def get_engine(config: Config | None = None, device: int | None = None) -> Engine:
    config = config or Config()
    key = (config, torch.cuda.current_device() if device is None else device)
    if key not in _engines:
        _engines[key] = Engine(config, device)
    return _engines[key]


class NeuralNetwork:
    """numpy_nn.NeuralNetwork's inference surface (training is out of scope)."""

    def __init__(self, nodes: Sequence[int], learning_rate=0.1, bias=None, weights=None, engine: Engine | None = None):
        if weights is None:
            raise _lib.NgpError("random weight initialisation (numpy_nn.py:71-82) is outside the GA path; pass weights")
        self.nodes = list(nodes)
        self.bias = 1 if bias else 0
        cfg = Config(NETWORK_SHAPE=tuple(self.nodes), BIAS=bool(bias))
        self._eng = engine or get_engine(cfg)
        g = np.asarray(weights, np.float32)
        need = cfg.gene_size()
        if len(g) < need:
            raise ValueError(f"need {need} weights, got {len(g)}")
        # populate_weights tolerates extra genes with a warning (numpy_nn.py:66-67)
        self._genome = torch.from_numpy(np.ascontiguousarray(g[:need])).to(self._eng.device).reshape(1, need)

    def run(self, input_vector):
        if len(input_vector) != self.nodes[0]:
            raise Exception("input vector wrong shape")            # numpy_nn.py:121-122
        x = torch.tensor(np.asarray(input_vector, np.float32).reshape(1, 1, -1), device=self._eng.device)
        act, _ = self._eng.mlp_forward(self._genome, x, want_out=False)
        return [1, 0] if int(act.item()) == _lib.ACT_UP else [0, 1]


def find_stuff(observation: np.ndarray, engine: Engine | None = None):
    """(210,160,3) uint8 -> [ball, left, right], each (row, col) float array or None."""
    eng = engine or get_engine()
    f = torch.from_numpy(np.ascontiguousarray(observation, np.uint8)).to(eng.device).reshape(1, 210, 160, 3)
    loc, valid = eng.find_stuff(f)
    loc = loc.cpu().numpy()[0].astype(np.float64); valid = valid.cpu().numpy()[0]
    return [loc[t] if valid[t] else None for t in range(3)]


class HallOfFame:
    """deap.tools.HallOfFame(maxsize) semantics: best distinct individuals ever seen, best first."""

    def __init__(self, maxsize: int):
        self.maxsize = maxsize
        self.genomes: List[np.ndarray] = []
        self.fitness: List[float] = []

    def __len__(self):
        return len(self.genomes)

    def update(self, genomes: np.ndarray, fitness: np.ndarray):
        if self.maxsize == 0:
            return
        # only candidates that can enter: better than the current worst or while not full
        order = range(len(fitness))
        for i in order:
            f = float(fitness[i])
            if len(self.genomes) == 0:
                self.genomes.append(genomes[i].copy()); self.fitness.append(f)
                continue
            if f > self.fitness[-1] or len(self.genomes) < self.maxsize:
                if any(np.array_equal(genomes[i], h) for h in self.genomes):
                    continue
                if len(self.genomes) >= self.maxsize:
                    self.genomes.pop(); self.fitness.pop()
                pos = 0
                while pos < len(self.fitness) and self.fitness[pos] > f:
                    pos += 1
                self.genomes.insert(pos, genomes[i].copy()); self.fitness.insert(pos, f)

    def tensors(self, device):
        if not self.genomes:
            return None, None
        return (torch.from_numpy(np.stack(self.genomes)).to(device), torch.tensor(self.fitness, dtype=torch.float64, device=device))


class Toolbox:
    """The slice of the DEAP toolbox the reference registers (ga.py:83-94), on device tensors."""

    def __init__(self, config: Config | None = None, engine: Engine | None = None, seed: int = 0):
        self.config = config or Config()
        self.engine = engine or get_engine(self.config)
        self.seed = seed
        self.generation = 0
        self.hall_of_fame = HallOfFame(self.config.HALL_OF_FAME_AMOUNT)

    def population(self, n: int) -> torch.Tensor:
        return self.engine.init_population(n, seed=self.seed)

    def evaluate(self, individual, render=False):
        """main.evaluate for ONE individual -> 1-tuple (main.py:66)."""
        g = torch.tensor(np.asarray(individual, np.float32).reshape(1, -1), device=self.engine.device)
        return (float(self.map_evaluate(g)[0].item()),)

    def map_evaluate(self, genomes: torch.Tensor) -> torch.Tensor:
        """toolbox.map(toolbox.evaluate, population): one fused launch for the whole population."""
        hg, hf = self.hall_of_fame.tensors(self.engine.device)
        out = self.engine.evaluate(genomes, hg, hf, seed=self.seed, generation=self.generation)
        self.last_frames = out["frames_total"]
        return out["fitness"]

    def vary(self, genomes: torch.Tensor, fitness: torch.Tensor):
        """toolbox.select + varAnd(mate, mutate) as one GA step."""
        return self.engine.ga_step(genomes, fitness, seed=self.seed, generation=self.generation)


def run_generations(toolbox: Toolbox, genomes: torch.Tensor, ngen: int, fitness: Optional[torch.Tensor] = None, verbose: bool = True):
    """deap.algorithms.eaSimple as the reference drives it (main.py:165-170): evaluate, update the hall of
    fame, then ngen x (select, vary, evaluate the invalid, update hall of fame, log avg/std/min/max)."""
    log = []

    def record(gen, nevals, fit, t0):
        f = fit.double()
        row = dict(gen=gen, nevals=int(nevals), avg=f.mean().item(), std=f.std(unbiased=False).item(), min=f.min().item(),
                   max=f.max().item(), frames=toolbox.last_frames, seconds=time.time() - t0)
        log.append(row)
        if verbose:
            print("{gen}\t{nevals}\t{avg:.6g}\t{std:.6g}\t{min:.6g}\t{max:.6g}".format(**row))

    t0 = time.time()
    if fitness is None:
        fitness = toolbox.map_evaluate(genomes)
        toolbox.hall_of_fame.update(genomes.cpu().numpy(), fitness.cpu().numpy())
        record(0, genomes.shape[0], fitness, t0)
    for _ in range(ngen):
        toolbox.generation += 1
        t0 = time.time()
        nxt = toolbox.vary(genomes, fitness)
        children, invalid, parents = nxt["genomes"], nxt["invalid"].bool(), nxt["parent_idx"].long()
        evaluated = toolbox.map_evaluate(children)
        # eaSimple re-evaluates only individuals whose fitness was invalidated; clones keep their parent's
        fitness = torch.where(invalid, evaluated, fitness.index_select(0, parents))
        genomes = children
        toolbox.hall_of_fame.update(genomes.cpu().numpy(), fitness.cpu().numpy())
        record(toolbox.generation, int(invalid.sum().item()), fitness, t0)
    return genomes, fitness, log


# ---- checkpoint / resume (utils.py:116-125, ga.py:13-53) -------------------------------------------
def save_checkpoint(toolbox: Toolbox, genomes: torch.Tensor, fitness: torch.Tensor, directory: str = "checkpoints/checkpoints") -> str:
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "c_{}.npz".format(time.strftime("%H_%M_%S")))
    hof = toolbox.hall_of_fame
    np.savez(path, population=genomes.cpu().numpy(), fitness=fitness.cpu().numpy(),
             hof_genomes=np.stack(hof.genomes) if len(hof) else np.zeros((0, genomes.shape[1]), np.float32),
             hof_fitness=np.asarray(hof.fitness, np.float64), seed=np.uint64(toolbox.seed), generation=np.uint64(toolbox.generation),
             network_shape=np.asarray(toolbox.config.NETWORK_SHAPE, np.int32))
    return path


def load_latest_population(toolbox: Toolbox, directory: str = "checkpoints/checkpoints"):
    """Newest checkpoint by ctime (ga.py:32-38), sorted by fitness descending (ga.py:45), truncated or
    topped up with fresh random individuals to POPULATION_SIZE (ga.py:13-29).  Returns (genomes, fitness|None)."""
    files = glob.glob(os.path.join(directory, "*.npz"))
    n = toolbox.config.POPULATION_SIZE
    if not files:
        return toolbox.population(n), None
    cp = np.load(max(files, key=os.path.getctime))
    if tuple(int(v) for v in cp["network_shape"]) != tuple(toolbox.config.NETWORK_SHAPE):
        raise _lib.NgpError("checkpoint network_shape differs from the configured NETWORK_SHAPE; rebuild the Toolbox with it")
    order = np.argsort(-cp["fitness"], kind="stable")
    pop, fit = cp["population"][order][:n], cp["fitness"][order][:n]
    toolbox.seed = int(cp["seed"]); toolbox.generation = int(cp["generation"])
    toolbox.hall_of_fame.genomes = [g for g in cp["hof_genomes"]]; toolbox.hall_of_fame.fitness = [float(f) for f in cp["hof_fitness"]]
    dev = toolbox.engine.device
    genomes = torch.from_numpy(pop).to(dev)
    if len(pop) < n:
        fresh = toolbox.engine.init_population(n - len(pop), seed=toolbox.seed + 1 + toolbox.generation)
        return torch.cat([genomes, fresh]), None
    return genomes, torch.from_numpy(fit).to(dev)
