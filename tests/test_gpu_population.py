"""GPU parity tests of the population-level pieces: hall of fame (ngp_hof_update), toolbox.select / mate / mutate as separate
callables, the Philox (non-injected) GA path, the multi-GPU exchange records, the eaSimple-shaped driver with its logbook and
checkpoint / resume -- each against the CPU restatements in oracle/ (DEAP semantics: SURVEY Appendix C)."""
import concurrent.futures as cf
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ngp():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import neuro_genetic_pong_self_play_b200 as m
    return m


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---------------------------------------------------------------------------------------------------------------------
# hall of fame
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,G_shape,maxsize", [(64, (6, 2, 2), 16), (300, (6, 2, 2), 75), (40, (6, 16, 16, 2), 10), (1500, (6, 2, 2), 1200)])
def test_hall_of_fame_update_matches_deap_semantics(ngp, n, G_shape, maxsize):
    """Several successive updates with ties (rounded fitness), duplicate genomes inside one population, duplicates of members
    (with a different fitness: the member blocks them), and a hall larger than one CTA pass (1200 > 1024 threads)."""
    import oracle
    from neuro_genetic_pong_self_play_b200.reference_api import HallOfFame
    eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=G_shape, POPULATION_SIZE=maxsize * 4), device=0)
    G = eng.gene_size
    rng = np.random.RandomState(n + maxsize)
    hof = HallOfFame(maxsize, eng)
    ref_g, ref_f = [], []
    prev = None
    for gen in range(4):
        pop = rng.random_sample((n, G)).astype(np.float32)
        fit = np.round(rng.standard_normal(n), 1)
        pop[5] = pop[3]                                     # the same genes twice in one population
        pop[7] = -0.0 * pop[7]; pop[8] = 0.0 * pop[8]       # -0.0 == 0.0 under Python's ==
        if prev is not None:
            pop[10:14] = prev[0][20:24]                     # clones of earlier individuals, re-evaluated to another fitness
            fit[10:12] = prev[1][20:22] + 5.0
        hof.update(_cuda(pop), _cuda(fit))
        ref_g, ref_f = oracle.hall_of_fame_update(ref_g, ref_f, pop, fit, maxsize)
        hg, hf = hof.tensors()
        assert len(hof) == len(ref_f)
        assert np.array_equal(hf.cpu().numpy(), np.array(ref_f)), gen
        assert np.array_equal(hg.cpu().numpy(), np.stack(ref_g)), gen
        prev = (pop, fit)
    assert len(hof.items) == len(ref_f) and hof.keys == ref_f
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------
# toolbox.select / mate / mutate as separate callables (ga.py:89-94)
# ---------------------------------------------------------------------------------------------------------------------
def test_toolbox_select_mate_mutate(ngp):
    import oracle
    from neuro_genetic_pong_self_play_b200.reference_api import Toolbox
    cfg = ngp.Config(POPULATION_SIZE=64)
    eng = ngp.Engine(cfg, device=0)
    tb = Toolbox(cfg, eng, seed=5)
    G, T = eng.gene_size, cfg.TOURNAMENT_SIZE
    rng = np.random.RandomState(1)
    genomes = rng.random_sample((64, G)).astype(np.float32)
    fitness = np.round(rng.standard_normal(64), 1)
    # select: k != n, injected draws
    draws = rng.randint(0, 64, size=(40, T)).astype(np.int32)
    chosen, idx = tb.select(_cuda(genomes), _cuda(fitness), k=40, draws=_cuda(draws))
    ref = oracle.sel_tournament(fitness, draws)
    assert np.array_equal(idx.cpu().numpy(), ref) and np.array_equal(chosen.cpu().numpy(), genomes[ref])
    # mate in place on two rows of a population tensor
    pop = _cuda(genomes)
    u = rng.random_sample(G).astype(np.float32)
    a, b = tb.mate(pop[2], pop[9], u=_cuda(u))
    c1, c2 = oracle.cx_blend(genomes[2], genomes[9], u, cfg.CROSSOVER_BLEND_ALPHA)
    assert np.array_equal(pop[2].cpu().numpy(), c1) and np.array_equal(pop[9].cpu().numpy(), c2)
    assert a.data_ptr() == pop[2].data_ptr()
    assert np.array_equal(pop[3].cpu().numpy(), genomes[3])          # neighbours untouched
    # mutate in place, returns a 1-tuple like DEAP
    mu_u = rng.random_sample(G).astype(np.float32); z = rng.standard_normal(G).astype(np.float32)
    (m,) = tb.mutate(pop[20], u=_cuda(mu_u), z=_cuda(z))
    ref_m = oracle.mut_gaussian(genomes[20], mu_u, z, cfg.GAUSSIAN_MUTATION_MEAN, cfg.GAUSSIAN_MUTATION_SIGMA, cfg.PROBABILITY_OF_MUTATING_A_SINGLE_GENE)
    assert np.array_equal(m.cpu().numpy(), ref_m)
    # the Philox streams of the separate callables are those of the fused ngp_ga_step: replaying select + mate + mutate
    # slot by slot reproduces its children
    tb.generation = 3
    fused = tb.vary(_cuda(genomes), _cuda(fitness))
    nz = oracle.ga_noise_philox(64, G, T, seed=5, generation=3, cxpb=cfg.CROSSOVER_BLEND_PROBABILITY, mutpb=cfg.GAUSSIAN_MUTATION_PROBABILITY)
    off, idx = tb.select(_cuda(genomes), _cuda(fitness))
    assert torch.equal(idx, fused["parent_idx"])
    off = off.contiguous()
    for p in range(32):
        if nz["cx_do"][p]:
            tb.mate(off[2 * p], off[2 * p + 1], pair=p)
    for i in range(64):
        if nz["mut_do"][i]:
            tb.mutate(off[i], slot=i)
    assert torch.equal(off, fused["genomes"])
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------
# Philox (non-injected) path: the counter layouts restated on the CPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [64, 65, 1024])
def test_philox_ga_path_matches_restated_counter_layout(ngp, n):
    import oracle
    cfg = ngp.Config(POPULATION_SIZE=n)
    eng = ngp.Engine(cfg, device=0)
    G, T = eng.gene_size, cfg.TOURNAMENT_SIZE
    seed, gen = 0x1234ABCD5678, 7
    pop = eng.init_population(n, seed=seed)
    assert np.array_equal(pop.cpu().numpy(), oracle.init_population_philox(n, G, seed))
    fitness = np.round(np.random.RandomState(n).standard_normal(n), 1)
    out = eng.ga_step(pop, _cuda(fitness), seed=seed, generation=gen)
    nz = oracle.ga_noise_philox(n, G, T, seed, gen, cfg.CROSSOVER_BLEND_PROBABILITY, cfg.GAUSSIAN_MUTATION_PROBABILITY)
    parents = oracle.sel_tournament(fitness, nz["sel_draws"])
    assert np.array_equal(out["parent_idx"].cpu().numpy(), parents)                  # bit-exact draws and winners
    # injecting the restated noise reproduces the Philox run: decisions and uniforms exactly, the normals to float32 rounding
    inj = eng.ga_step(pop, _cuda(fitness), seed=seed, generation=gen, noise={k: _cuda(v) for k, v in nz.items()})
    assert torch.equal(inj["parent_idx"], out["parent_idx"]) and torch.equal(inj["invalid"], out["invalid"])
    a, b = inj["genomes"].cpu().numpy(), out["genomes"].cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=0, atol=4e-6)                               # sigma * |dz|, dz <= ~2 ulp of z
    pg = pop.cpu().numpy()[parents]
    child, invalid = oracle.var_and(pg, nz["cx_do"], nz["cx_u"], np.zeros(n, np.uint8), nz["mut_u"], nz["mut_z"], cfg.CROSSOVER_BLEND_ALPHA,
                                    0.0, 0.9, 0.9)
    untouched = nz["mut_do"] == 0
    assert np.array_equal(b[untouched], child[untouched])                           # crossover-only individuals: bit-exact
    cx_ind = np.array((np.repeat(nz["cx_do"], 2).tolist() + [0])[:n], np.uint8)
    assert np.array_equal(out["invalid"].cpu().numpy(), cx_ind | nz["mut_do"])
    eng.close()


def test_order_statistics_tournament(ngp):
    """selTournament for large populations (csrc/ngp_ops.cu select_order_stat_kernel): winners drawn from the order statistics
    of the tournament on the fitness-sorted population.  Checked: (1) the counter layout and arithmetic restated with numpy
    reproduce the device picks; (2) the winners' rank distribution is the tournament's P(R >= r) = ((N - r) / N)^T, and equals
    the draw-by-draw kernel's statistically; (3) equal fitness: uniform inside the tie group; (4) config 5 scale runs."""
    import oracle
    n, T = 4096, 1024
    cfg = ngp.Config(POPULATION_SIZE=n)                      # TOURNAMENT_SIZE = n // 4
    eng = ngp.Engine(cfg, device=0)
    rng = np.random.RandomState(12)
    fitness = np.round(rng.standard_normal(n), 2)            # many ties
    seed, gen, k = 0xABCDEF0123, 5, 200_000
    eng.set_option("select_os_min_t", 1)
    got = eng.select(_cuda(fitness), k, seed=seed, generation=gen).cpu().numpy()
    # (1) restatement: stable sort by (-fitness, index); one Philox block per slot
    order = np.lexsort((np.arange(n), -fitness))
    fs = fitness[order]
    w0, w1, w2, _ = oracle.philox4x32_np(np.arange(k, dtype=np.uint64), 0, gen, 0x53454C32, seed)
    u = ((((w0.astype(np.uint64) << np.uint64(32)) | w1.astype(np.uint64)) >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    r = np.clip(np.floor(n * -np.expm1(np.log(u) / T)).astype(np.int64), 0, n - 1)
    first = np.searchsorted(-fs, -fs[r], side="left"); last = np.searchsorted(-fs, -fs[r], side="right")
    pick = first + ((w2.astype(np.uint64) * (last - first).astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
    want = order[pick]
    assert (got == want).mean() > 0.9999                     # log/expm1 may round differently at a rank boundary once in a while
    assert np.array_equal(fitness[got], fitness[want]) or (fitness[got] != fitness[want]).mean() < 1e-4
    # (2) rank distribution against the closed form and against the draw-by-draw kernel
    rank_of = np.empty(n, np.int64); rank_of[order] = np.arange(n)
    group_first = np.searchsorted(-fs, -fs, side="left")      # first rank of each individual's tie group
    g_got = group_first[rank_of[got]]
    eng.set_option("select_os_min_t", 0)
    ref = eng.select(_cuda(fitness), 50_000, seed=seed, generation=gen).cpu().numpy()        # T < 8192: N/4 Philox draws per slot
    g_ref = group_first[rank_of[ref]]
    edges = np.unique(group_first)
    cdf = 1.0 - ((n - edges) / n) ** T                       # P(best group starts before edge)
    cdf = np.append(cdf, 1.0)
    for sample in (g_got, g_ref):
        emp = np.searchsorted(np.sort(sample), edges, side="left") / len(sample)
        assert np.abs(np.append(emp, 1.0) - cdf).max() < 4.0 / np.sqrt(len(sample))          # Kolmogorov-Smirnov style bound
    # (3) uniform inside the best tie group
    top = np.nonzero(fitness == fitness.max())[0]
    if len(top) == 1:
        fitness[rng.choice(n, 7, replace=False)] = fitness.max(); top = np.nonzero(fitness == fitness.max())[0]
        eng.set_option("select_os_min_t", 1)
        got = eng.select(_cuda(fitness), k, seed=seed, generation=gen + 1).cpu().numpy()
    winners = got[np.isin(got, top)]
    counts = np.array([(winners == t).sum() for t in top])
    assert len(winners) > 1000 and np.abs(counts / len(winners) - 1.0 / len(top)).max() < 5.0 / np.sqrt(len(winners))
    eng.close()
    # (4) BASELINE config 5's largest population: one GA step on 2^20 genomes uses the order statistics by default
    n = 1 << 20
    eng = ngp.Engine(ngp.Config(POPULATION_SIZE=n), device=0)
    pop = eng.init_population(n, seed=1)
    fit = _cuda(np.random.RandomState(1).standard_normal(n))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.ga_step(pop, fit, seed=3, generation=0)
    ev0.record(); out = eng.ga_step(pop, fit, seed=3, generation=1); ev1.record(); torch.cuda.synchronize()
    parents = out["parent_idx"].cpu().numpy()
    assert parents.min() >= 0 and parents.max() < n
    best_ranks = np.argsort(np.argsort(-fit.cpu().numpy()))[parents]
    assert np.median(best_ranks) < 8 and ev0.elapsed_time(ev1) < 50.0        # T = 262144: winners are the top few; was 945 ms draw by draw
    eng.close()


def test_philox_hof_pick_matches_restated_counter_layout(ngp):
    """Games 3..5 with the hall-of-fame opponent drawn by Philox equal the same evaluation with the restated picks injected."""
    import oracle
    cfg = ngp.Config(POPULATION_SIZE=16, MAX_FRAMES=400)
    eng = ngp.Engine(cfg, device=0)
    rng = np.random.RandomState(4)
    genomes = _cuda((rng.standard_normal((16, eng.gene_size)) * 2).astype(np.float32))
    hof = _cuda((rng.standard_normal((5, eng.gene_size)) * 2).astype(np.float32)); hof_fit = _cuda(np.array([2.0, 1.5, 1.0, 0.5, 0.25]))
    a = eng.evaluate(genomes, hof, hof_fit, seed=77, generation=9, want_detail=True)
    picks = oracle.hof_pick_philox(16, 5, 77, 9)
    b = eng.evaluate(genomes, hof, hof_fit, _cuda(picks), seed=77, generation=9, want_detail=True)
    assert torch.equal(a["rewards"], b["rewards"]) and torch.equal(a["frames"], b["frames"])
    assert len(set(picks.reshape(-1).tolist())) > 1
    # the random-action stream is keyed by the generation too (fresh draws every generation, main.py:139-140)
    c = eng.evaluate(genomes, hof, hof_fit, _cuda(picks), seed=77, generation=10, want_detail=True)
    assert a["frames_total"] > 0 and c["frames_total"] > 0
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------
# exchange records
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizes,counts", [((6, 6), (2, 2)), ((4, 3, 3), (1, 1, 1)), ((300, 299), (40, 40))])
def test_exchange_records_match_restated_layout(ngp, sizes, counts):
    import oracle
    eng = ngp.Engine(ngp.Config(), device=0)
    G = eng.gene_size
    rng = np.random.RandomState(sum(sizes))
    n_max, k_max = max(sizes), max(counts)
    assert eng.exchange_bytes(n_max, k_max) == oracle.exchange_bytes(n_max, k_max, G)
    recs, fits, gens = [], [], []
    for n, k in zip(sizes, counts):
        g = rng.random_sample((n, G)).astype(np.float32); f = np.round(rng.standard_normal(n), 1)
        rec = eng.pack_elites(_cuda(g), _cuda(f), k, n_max, k_max)
        assert np.array_equal(rec.cpu().numpy(), oracle.pack_record(g, f, k, n_max, k_max))
        recs.append(rec); fits.append(f); gens.append(g)
    gathered = torch.cat(recs)
    fit, eg, ef = eng.unpack_elites(gathered, len(sizes), n_max, k_max, sum(sizes), sum(counts))
    rf, rg, re_ = oracle.unpack_records(gathered.cpu().numpy(), len(sizes), n_max, k_max, G)
    assert np.array_equal(fit.cpu().numpy(), rf) and np.array_equal(eg.cpu().numpy(), rg) and np.array_equal(ef.cpu().numpy(), re_)
    assert np.array_equal(rf, np.concatenate(fits))
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------
# eaSimple-shaped driver + logbook, checkpoint / resume
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_eval(args):
    import oracle
    genome, hof_g, hof_f, pick, seed, gen, gid = args
    return oracle.evaluate([6, 2, 2], genome, hof_g, hof_f, pick, seed=seed, genome_id=gid, generation=gen)[0]


def test_run_generations_matches_oracle_driven_easimple(ngp):
    """3 generations of the reference's loop (main.py:165-170 -> DEAP eaSimple: evaluate, HallOfFame.update, select, varAnd,
    evaluate the invalid, update, log) with injected GA noise: logbook, final population, fitness and hall of fame against a
    loop driven entirely by the CPU oracle."""
    import oracle
    from neuro_genetic_pong_self_play_b200.reference_api import Toolbox, run_generations
    n, ngen, seed = 16, 3, 21
    cfg = ngp.Config(POPULATION_SIZE=n)
    eng = ngp.Engine(cfg, device=0)
    G, T, H = eng.gene_size, cfg.TOURNAMENT_SIZE, cfg.HALL_OF_FAME_AMOUNT
    rng = np.random.RandomState(seed)
    pop0 = (rng.standard_normal((n, G)) * 2).astype(np.float32)
    noises = {g: dict(sel_draws=rng.randint(0, n, size=(n, T)).astype(np.int32), cx_do=(rng.random_sample(n // 2) < 0.9).astype(np.uint8),
                      cx_u=rng.random_sample((n // 2, G)).astype(np.float32), mut_do=(rng.random_sample(n) < 0.9).astype(np.uint8),
                      mut_u=rng.random_sample((n, G)).astype(np.float32), mut_z=rng.standard_normal((n, G)).astype(np.float32))
              for g in range(1, ngen + 1)}
    tb = Toolbox(cfg, eng, seed=seed)
    genomes, fitness, log = run_generations(tb, _cuda(pop0), ngen, verbose=False,
                                            noise_fn=lambda g: {k: _cuda(v) for k, v in noises[g].items()})

    with cf.ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        def evaluate(pop, hof_g, hof_f, gen):
            nh = len(hof_f)
            picks = oracle.hof_pick_philox(n, nh, seed, gen) if nh else np.zeros((n, 3), np.int32)
            hg = np.stack(hof_g) if nh else None
            return np.array(list(ex.map(_oracle_eval, [(pop[i], hg, np.array(hof_f) if nh else None, picks[i], seed, gen, i) for i in range(n)])))
        pop = pop0
        hof_g, hof_f = [], []
        fit = evaluate(pop, hof_g, hof_f, 0)
        hof_g, hof_f = oracle.hall_of_fame_update(hof_g, hof_f, pop, fit, H)
        rows = [(0, n, fit.mean(), fit.std(), fit.min(), fit.max())]
        for gen in range(1, ngen + 1):
            nz = noises[gen]
            parents = oracle.sel_tournament(fit, nz["sel_draws"])
            child, invalid = oracle.var_and(pop[parents], nz["cx_do"], nz["cx_u"], nz["mut_do"], nz["mut_u"], nz["mut_z"], cfg.CROSSOVER_BLEND_ALPHA,
                                            cfg.GAUSSIAN_MUTATION_MEAN, cfg.GAUSSIAN_MUTATION_SIGMA, cfg.PROBABILITY_OF_MUTATING_A_SINGLE_GENE)
            ev = evaluate(child, hof_g, hof_f, gen)
            fit = np.where(invalid == 1, ev, fit[parents])
            pop = child
            hof_g, hof_f = oracle.hall_of_fame_update(hof_g, hof_f, pop, fit, H)
            rows.append((gen, int(invalid.sum()), fit.mean(), fit.std(), fit.min(), fit.max()))
    assert np.array_equal(genomes.cpu().numpy(), pop)
    assert np.array_equal(fitness.cpu().numpy(), fit)
    hg, hf = tb.hall_of_fame.tensors()
    assert np.array_equal(hf.cpu().numpy(), np.array(hof_f)) and np.array_equal(hg.cpu().numpy(), np.stack(hof_g))
    assert len(log) == ngen + 1
    for row, (gen, nevals, avg, std, mn, mx) in zip(log, rows):
        assert row["gen"] == gen and row["nevals"] == nevals
        np.testing.assert_allclose([row["avg"], row["std"], row["min"], row["max"]], [avg, std, mn, mx], rtol=1e-12, atol=1e-15)
    eng.close()


def test_checkpoint_round_trip_truncate_and_top_up(ngp, tmp_path):
    """utils.save_checkpoint / ga.load_or_create_pop (utils.py:116-125, ga.py:13-53): newest file wins, population sorted by
    fitness descending, truncated or topped up to POPULATION_SIZE (loaded individuals keep their fitness, new ones are
    invalid), hall of fame, RNG position (seed + generation), NETWORK_SHAPE and BIAS restored or checked."""
    from neuro_genetic_pong_self_play_b200.reference_api import Toolbox, load_latest_population, run_generations, save_checkpoint
    d = str(tmp_path / "checkpoints" / "checkpoints")
    cfg = ngp.Config(POPULATION_SIZE=12, MAX_FRAMES=150)
    eng = ngp.Engine(cfg, device=0)
    tb = Toolbox(cfg, eng, seed=3)
    g0, f0 = load_latest_population(tb, d)                      # no checkpoint yet: fresh population, nobody evaluated
    assert f0 is None and tuple(g0.shape) == (12, eng.gene_size)
    genomes, fitness, log = run_generations(tb, g0, 2, verbose=False)
    path = save_checkpoint(tb, genomes, fitness, d)
    assert os.path.exists(path) and tb.generation == 2 and len(tb.hall_of_fame) > 0
    hg, hf = (t.clone() for t in tb.hall_of_fame.tensors())

    tb2 = Toolbox(cfg, eng, seed=999)
    g2, f2 = load_latest_population(tb2, d)
    order = np.argsort(-fitness.cpu().numpy(), kind="stable")
    assert np.array_equal(g2.cpu().numpy(), genomes.cpu().numpy()[order]) and np.array_equal(f2.cpu().numpy(), fitness.cpu().numpy()[order])
    assert tb2.seed == 3 and tb2.generation == 2
    assert torch.equal(tb2.hall_of_fame.tensors()[0], hg) and torch.equal(tb2.hall_of_fame.tensors()[1], hf)

    small = cfg.replace(POPULATION_SIZE=8)                      # truncate: the best 8
    eng8 = ngp.Engine(small, device=0)
    g8, f8 = load_latest_population(Toolbox(small, eng8, seed=0), d)
    assert np.array_equal(g8.cpu().numpy(), genomes.cpu().numpy()[order][:8]) and not torch.isnan(f8).any()
    big = cfg.replace(POPULATION_SIZE=20)                       # top up: 8 fresh individuals without fitness
    eng20 = ngp.Engine(big, device=0)
    tb20 = Toolbox(big, eng20, seed=0)
    g20, f20 = load_latest_population(tb20, d)
    assert tuple(g20.shape) == (20, eng.gene_size) and torch.isnan(f20[12:]).all() and not torch.isnan(f20[:12]).any()
    assert np.array_equal(g20[:12].cpu().numpy(), genomes.cpu().numpy()[order])
    fresh = g20[12:].cpu().numpy()
    assert 0.0 <= fresh.min() and fresh.max() < 1.0
    # resuming evaluates only the invalid individuals (eaSimple), the others keep their loaded fitness
    g_res, f_res, log_res = run_generations(tb20, g20, 0, fitness=f20, verbose=False)
    assert log_res[0]["nevals"] == 8 and torch.equal(f_res[:12], f20[:12]) and not torch.isnan(f_res).any()
    # a checkpoint of another network shape is refused
    other = ngp.Config(POPULATION_SIZE=12, NETWORK_SHAPE=(6, 4, 2))
    engo = ngp.Engine(other, device=0)
    with pytest.raises(ngp.NgpError):
        load_latest_population(Toolbox(other, engo, seed=0), d)
    for e in (eng, eng8, eng20, engo):
        e.close()


def test_resume_from_a_reference_checkpoint(ngp, tmp_path):
    """load_latest_population continues from a checkpoint the REFERENCE wrote (pickled DEAP objects, utils.py:116-125): sorted by
    fitness, invalid individuals re-evaluated, hall of fame restored without its invalid members (utils.py:96-98), topped up."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_reference_checkpoint import _write_reference_style_checkpoint
    from neuro_genetic_pong_self_play_b200.reference_api import Toolbox, load_latest_population
    cfg = ngp.Config(POPULATION_SIZE=8)
    tb = Toolbox(cfg, ngp.Engine(cfg, device=0), seed=5)
    rng = np.random.RandomState(3)
    genes = rng.standard_normal((6, tb.engine.gene_size)).astype(np.float32)
    fits = np.array([0.25, np.nan, 1.5, -0.5, 0.75, 0.0])
    _write_reference_style_checkpoint(tmp_path / "c_01_02_03.pkl", genes, fits, hof_idx=[2, 4, 1])
    genomes, fitness = load_latest_population(tb, str(tmp_path))
    g, f = genomes.cpu().numpy(), fitness.cpu().numpy()
    assert g.shape == (8, tb.engine.gene_size)
    assert np.array_equal(g[:5], genes[[2, 4, 0, 5, 3]]) and f[:5].tolist() == [1.5, 0.75, 0.25, 0.0, -0.5]
    assert np.array_equal(g[5], genes[1]) and np.isnan(f[5:]).all()          # the invalid individual and the two fresh ones
    hg, hf = tb.hall_of_fame.tensors()
    assert hf.cpu().tolist() == [1.5, 0.75] and np.array_equal(hg.cpu().numpy(), genes[[2, 4]])
    tb.engine.close()


def test_replay_shows_the_game_the_batched_path_played(ngp):
    """pickle_inspector.py / evaluate(individual, render=True): the frame-by-frame replay of one individual (explicit-action API +
    ngp_mlp_forward on the host loop) ends after the same number of env.step calls with the same reward as the fused evaluation of
    the same game, and keeps one upscaled frame per step (render_game's repeat_upsample 4 x 4)."""
    from neuro_genetic_pong_self_play_b200.reference_api import Toolbox, replay, repeat_upsample
    cfg = ngp.Config(POPULATION_SIZE=4)
    eng = ngp.Engine(cfg, device=0)
    tb = Toolbox(cfg, eng, seed=3)
    rng = np.random.RandomState(11)
    genomes = (rng.standard_normal((4, eng.gene_size)) * 2).astype(np.float32)
    ref = eng.evaluate(_cuda(genomes), seed=3, generation=0, want_detail=True)
    for gi, game in ((0, 0), (1, 2), (2, 1)):
        out = replay(tb, _cuda(genomes[gi]), game=game)
        assert out["steps"] == int(ref["frames"][gi, game].item())
        assert out["reward"] == ref["rewards"][gi, game].item()
        assert tuple(out["frames"].shape) == (out["steps"], 840, 640, 3) and out["frames"].dtype == torch.uint8
    small = torch.arange(24, dtype=torch.uint8, device="cuda").reshape(2, 4, 3)
    up = repeat_upsample(small, 2, 3)
    assert torch.equal(up.cpu(), torch.from_numpy(np.repeat(np.repeat(small.cpu().numpy(), 2, axis=0), 3, axis=1)))
    assert repeat_upsample(small, 0, 3) is small
    eng.close()


def test_reference_call_surface(ngp, golden, obs_npy):
    """NeuralNetwork(nodes, weights, bias).run, find_stuff(obs), toolbox.evaluate(individual), toolbox.map(toolbox.evaluate, pop)
    read like the reference's call sites (numpy_nn.py:35-50,120-137; utils.py:14-19; main.py:28-66; ga.py:83)."""
    from neuro_genetic_pong_self_play_b200 import reference_api as api
    assert [None if v is None else v.tolist() for v in api.find_stuff(obs_npy)] == [[111.5, 64.5], [122.5, 17.5], [127.5, 141.5]]
    assert api.find_stuff(np.zeros_like(obs_npy)) == [None, None, None]
    genome = np.random.RandomState(0).random_sample(20)
    nn = api.NeuralNetwork(nodes=[6, 2, 2], weights=list(genome), bias=True)
    assert nn.run([0.403125, 0.696875, 0.403125, 0.696875, 0.796875, 0.765625]) == [0, 1]
    with pytest.raises(Exception, match="input vector wrong shape"):
        nn.run([0.0] * 5)
    with pytest.warns(UserWarning):
        api.NeuralNetwork(nodes=[6, 2, 2], weights=list(genome) + [1.0], bias=True)
    cfg = ngp.Config(POPULATION_SIZE=8, MAX_FRAMES=120)
    tb = api.Toolbox(cfg, ngp.Engine(cfg, device=0), seed=1)
    pop = tb.population(8)
    fits = tb.map(tb.evaluate, pop)
    assert len(fits) == 8 and all(isinstance(f, tuple) and len(f) == 1 for f in fits)
    assert tb.evaluate(pop[0].cpu().tolist()) == fits[0]          # an individual alone plays the games it plays in the population
    assert list(tb.map(len, [[1, 2], [3]])) == [2, 1]
    tb.engine.close()
