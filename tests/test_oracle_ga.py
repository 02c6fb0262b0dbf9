"""The numpy restatements of the DEAP operators (oracle/__init__.py; ga.py:77-94)."""
import numpy as np

import oracle


def test_sel_tournament_first_max_wins():
    fitness = np.array([1.0, 3.0, 3.0, 2.0])
    draws = np.array([[0, 2, 1, 3], [3, 0, 0, 0], [1, 2, 2, 1]], np.int32)
    assert oracle.sel_tournament(fitness, draws).tolist() == [2, 3, 1]


def test_cx_blend_formula_and_symmetry():
    rng = np.random.RandomState(0)
    x1 = rng.random_sample(20).astype(np.float32); x2 = rng.random_sample(20).astype(np.float32); u = rng.random_sample(20).astype(np.float32)
    c1, c2 = oracle.cx_blend(x1, x2, u, 0.9)
    g = (1 + 2 * 0.9) * u.astype(np.float64) - 0.9
    np.testing.assert_allclose(c1, (1 - g) * x1 + g * x2, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(c1 + c2, x1 + x2, rtol=1e-5, atol=1e-6)          # blend preserves the pair sum


def test_var_and_flags_and_mutation_gate():
    n, G = 6, 4
    parents = np.arange(n * G, dtype=np.float32).reshape(n, G)
    cx_do = np.array([1, 0, 0], np.uint8); mut_do = np.array([0, 0, 1, 0, 0, 0], np.uint8)
    child, invalid = oracle.var_and(parents, cx_do, np.full((3, G), 0.5, np.float32), mut_do, np.array([[0.1, 0.95, 0.1, 0.95]] * n, np.float32),
                                    np.ones((n, G), np.float32), 0.9, 0.0, 0.9, 0.9)
    assert invalid.tolist() == [1, 1, 1, 0, 0, 0]
    assert np.array_equal(child[3:], parents[3:])
    np.testing.assert_allclose(child[2], parents[2] + np.array([0.9, 0, 0.9, 0], np.float32))


def test_hall_of_fame_update_semantics():
    pop = [np.array([i, i], np.float32) for i in range(5)]
    fit = [0.1, 0.5, 0.3, 0.5, 0.2]
    g, f = oracle.hall_of_fame_update([], [], pop, fit, maxsize=3)
    assert f == [0.5, 0.5, 0.3]
    assert [int(x[0]) for x in g] == [3, 1, 2]          # equal fitness: the newer individual is inserted first
    g2, f2 = oracle.hall_of_fame_update(g, f, [np.array([1, 1], np.float32)], [0.9], maxsize=3)
    assert f2 == f                                       # a duplicate genome is never inserted twice
