"""One-way importer of the reference's checkpoints (utils.save_checkpoint, /root/reference/utils.py:116-125) without DEAP.

DEAP is not installable here, so the test builds a checkpoint with stand-in modules named like DEAP's (`deap.creator.Individual`
= list subclass with a `fitness` attribute, `deap.creator.Fitness` with `wvalues`, `deap.tools.support.HallOfFame` with
maxsize/keys/items) -- the pickle stream then references exactly the class paths a real reference checkpoint does -- removes the
modules again and reads the file back through read_reference_checkpoint."""
import operator
import pickle
import random
import sys
import types

import numpy as np


def _write_reference_style_checkpoint(path, genes, fits, hof_idx):
    creator = types.ModuleType("deap.creator"); support = types.ModuleType("deap.tools.support"); deap = types.ModuleType("deap")
    tools = types.ModuleType("deap.tools")

    class Fitness(object):
        weights = (1.0,)
        def __init__(self, values=()):
            self.wvalues = tuple(values)
    class Individual(list):
        pass
    class HallOfFame(object):
        def __init__(self, maxsize):
            self.maxsize = maxsize; self.keys = []; self.items = []; self.similar = operator.eq
    for cls, mod in ((Fitness, creator), (Individual, creator), (HallOfFame, support)):
        cls.__module__ = mod.__name__; cls.__qualname__ = cls.__name__; setattr(mod, cls.__name__, cls)
    mods = {"deap": deap, "deap.creator": creator, "deap.tools": tools, "deap.tools.support": support}
    sys.modules.update(mods)
    try:
        pop = []
        for g, f in zip(genes, fits):
            ind = Individual(float(x) for x in g)
            ind.fitness = Fitness(() if np.isnan(f) else (float(f),))
            pop.append(ind)
        hof = HallOfFame(len(hof_idx))
        hof.items = [pop[i] for i in hof_idx]; hof.keys = [pop[i].fitness for i in reversed(hof_idx)]
        with open(path, "wb") as f:
            pickle.dump(dict(population=pop, hall_of_fame=hof, rndstate=random.getstate(), network_shape=[6, 2, 2]), f)
    finally:
        for k in mods:
            sys.modules.pop(k, None)


def test_read_reference_checkpoint_without_deap(tmp_path):
    from neuro_genetic_pong_self_play_b200.reference_api import read_reference_checkpoint
    rng = np.random.RandomState(0)
    genes = rng.standard_normal((7, 20)).astype(np.float32)
    fits = np.array([0.5, -1.0, np.nan, 2.25, 0.0, 1.5, np.nan])
    path = tmp_path / "c_12_00_00.pkl"
    _write_reference_style_checkpoint(path, genes, fits, hof_idx=[3, 5, 0])
    assert "deap" not in sys.modules
    cp = read_reference_checkpoint(str(path))
    assert cp["network_shape"] == (6, 2, 2)
    assert cp["population"].dtype == np.float32 and np.array_equal(cp["population"], genes)
    assert np.array_equal(np.isnan(cp["fitness"]), np.isnan(fits)) and np.array_equal(cp["fitness"][~np.isnan(fits)], fits[~np.isnan(fits)])
    assert np.array_equal(cp["hof_genomes"], genes[[3, 5, 0]]) and cp["hof_fitness"].tolist() == [2.25, 1.5, 0.5]
