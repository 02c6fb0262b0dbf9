"""GPU parity tests proper: the CUDA path (through libngp.so's C ABI) against the CPU oracle on the
same seeded inputs, bit-exact for all integer/byte/index work and for the FP64 episode arithmetic."""
import concurrent.futures as cf
import os
import zlib

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


@pytest.fixture(scope="module")
def engine(ngp):
    e = ngp.Engine(ngp.Config(), device=0)
    yield e
    e.close()


def _random_actions(rng, n_frames, hold):
    acts = np.zeros((n_frames, 16), np.uint8)
    acts[:, 0] = 1; acts[:, 15] = 1
    cur = (0, 0)
    for f in range(n_frames):
        if f % hold == 0:
            cur = (rng.randint(0, 3), rng.randint(0, 3))
        r, l = cur
        acts[f, 4] = r == 1; acts[f, 5] = r == 2; acts[f, 6] = l == 1; acts[f, 7] = l == 2
    return acts


def _oracle_trace(args):
    import oracle
    state, acts = args
    env = oracle.Atari()
    env.reset_to_state(state)
    n = len(acts)
    ram = np.zeros((n, 128), np.uint8); regs = np.zeros((n, 8), np.uint8); dig = np.zeros((n, 8), np.uint32)
    crc = np.zeros(n, np.uint32); loc = np.zeros((n, 3, 2), np.float32); valid = np.zeros((n, 3), np.uint8)
    for f in range(n):
        fb = env.step(acts[f])
        rgb = oracle.fb_to_rgb(fb)
        ram[f] = env.ram; regs[f] = env.cpu_regs; dig[f] = env.tia_digest; crc[f] = zlib.crc32(rgb.tobytes())
        l, v = oracle.find_stuff(rgb)
        loc[f] = l.astype(np.float32); valid[f] = v
    return ram, regs, dig, crc, loc, valid


@pytest.mark.parametrize("core", [0, 1])
def test_env_step_trace_parity(engine, core):
    """core 0 = table-driven interpreter, core 1 = statically translated cartridge.  RAM, CPU registers, TIA digest (collision latches, positions, paddle charges), every frame
    (CRC of the RGB frame) and the fused find_stuff result, frame by frame, on random action traces
    from both start states (BASELINE config 2's bit-exact RAM/frame check: a 64-environment subset, more than 2 000 frames
    each)."""
    n_envs, n_frames = 64, 2100
    rng = np.random.RandomState(42)
    states = [i % 2 for i in range(n_envs)]
    acts = [_random_actions(rng, n_frames, hold=1 + 3 * (i % 5)) for i in range(n_envs)]
    with cf.ProcessPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        ref = list(ex.map(_oracle_trace, zip(states, acts)))
    for state in (0, 1):
        idx = [i for i in range(n_envs) if states[i] == state]
        engine.env_reset(len(idx), state)
        for f in range(n_frames):
            a = torch.from_numpy(np.stack([acts[i][f] for i in idx])).cuda()
            out = engine.env_step(a, core=core)
            ram = out["ram"].cpu().numpy(); regs = out["regs"].cpu().numpy(); frames = out["frames"].cpu().numpy()
            loc = out["loc"].cpu().numpy(); valid = out["valid"].cpu().numpy(); dig = engine.env_digest().cpu().numpy().view(np.uint32)
            for j, i in enumerate(idx):
                r_ram, r_regs, r_dig, r_crc, r_loc, r_valid = ref[i]
                assert regs[j, 7] == 0, "emulator error flag"
                assert np.array_equal(ram[j], r_ram[f]), f"RAM env {i} frame {f}"
                assert np.array_equal(regs[j, :7], r_regs[f, :7]), f"regs env {i} frame {f}"
                assert np.array_equal(dig[j], r_dig[f]), f"TIA digest env {i} frame {f}"
                assert zlib.crc32(frames[j].tobytes()) == r_crc[f], f"frame pixels env {i} frame {f}"
                assert np.array_equal(valid[j], r_valid[f]) and np.array_equal(loc[j][r_valid[f] == 1], r_loc[f][r_valid[f] == 1]), \
                    f"fused observation env {i} frame {f}"


def test_start_states_and_obs_npy_geometry(engine, obs_npy):
    import oracle
    for state in (0, 1):
        engine.env_reset(2, state)
        a = torch.zeros((2, 16), dtype=torch.uint8, device="cuda")
        out = engine.env_step(a)
        env = oracle.Atari(); env.reset_to_state(state); fb = env.step([0] * 16)
        assert np.array_equal(out["ram"][0].cpu().numpy(), env.ram) and np.array_equal(out["ram"][1].cpu().numpy(), env.ram)
        assert np.array_equal(out["frames"][0].cpu().numpy(), oracle.fb_to_rgb(fb))
    # the static rows of the frame (score digits, walls) are pixel-identical to obs.npy
    fr = out["frames"][0].cpu().numpy()
    assert np.array_equal(fr[:34], obs_npy[:34]) and np.array_equal(fr[194:], obs_npy[194:])


def _oracle_eval(args):
    import oracle
    nodes, genome, hof, hof_fit, pick, seed, gid = args
    return oracle.evaluate(nodes, genome, hof, hof_fit, pick, seed=seed, genome_id=gid)


def test_fused_evaluate_reference_schedule_parity(ngp, engine):
    """main.evaluate for a small population: per-game rewards (FP64) and episode lengths bit-exact."""
    rng = np.random.RandomState(3)
    n, G = 12, engine.gene_size
    genomes = (rng.random_sample((n, G)) * 6 - 3).astype(np.float32)
    genomes[:4] = rng.random_sample((4, G)).astype(np.float32)          # ga.py:85 style init for some
    hof = (rng.random_sample((4, G)) * 4 - 2).astype(np.float32)
    hof_fit = np.array([0.9, 0.5, 0.25, -0.1])
    pick = rng.randint(0, 4, size=(n, 3)).astype(np.int32)
    with cf.ProcessPoolExecutor(max_workers=min(12, os.cpu_count() or 1)) as ex:
        ref = list(ex.map(_oracle_eval, [([6, 2, 2], genomes[g], hof, hof_fit, pick[g], 11, g) for g in range(n)]))
    out = engine.evaluate(torch.from_numpy(genomes).cuda(), torch.from_numpy(hof).cuda(), torch.from_numpy(hof_fit).cuda(),
                          torch.from_numpy(pick).cuda(), seed=11, generation=0, want_detail=True)
    rewards = out["rewards"].cpu().numpy(); frames = out["frames"].cpu().numpy(); fitness = out["fitness"].cpu().numpy()
    for g in range(n):
        fit, r, f = ref[g]
        assert np.array_equal(frames[g], f), (g, frames[g], f)
        assert np.array_equal(rewards[g], r), (g, rewards[g], r)
        assert fitness[g] == fit
    assert out["frames_total"] == int(frames.sum())
    # the interpreter core gives the same bits as the translated core
    eng_i = ngp.Engine(ngp.Config(CORE=ngp.CORE_INTERPRETER), device=0)
    out_i = eng_i.evaluate(torch.from_numpy(genomes).cuda(), torch.from_numpy(hof).cuda(), torch.from_numpy(hof_fit).cuda(),
                           torch.from_numpy(pick).cuda(), seed=11, generation=0, want_detail=True)
    assert torch.equal(out_i["rewards"], out["rewards"]) and torch.equal(out_i["frames"], out["frames"])
    eng_i.close()
    # no hall of fame: games 3..5 fall back to HardcodedAi with multiplier 1 (main.py:43-53)
    out2 = engine.evaluate(torch.from_numpy(genomes[:2]).cuda(), seed=11, want_detail=True)
    for g in range(2):
        fit, r, f = _oracle_eval(([6, 2, 2], genomes[g], None, None, (0, 0, 0), 11, g))
        assert np.array_equal(out2["rewards"][g].cpu().numpy(), r) and np.array_equal(out2["frames"][g].cpu().numpy(), f)


def _oracle_selfplay(args):
    import oracle
    right, left, seed, env_id = args
    r = oracle.selfplay_game([6, 2, 2], right, left, seed=seed, env_id=env_id)
    return r.reward, r.frames


def test_fused_round_robin_parity(ngp):
    cfg = ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, GAMES_TO_PLAY=3)
    eng = ngp.Engine(cfg, device=0)
    rng = np.random.RandomState(5)
    n = 8
    genomes = (rng.standard_normal((n, eng.gene_size)) * 2).astype(np.float32)
    jobs = [(genomes[g], genomes[(g + k + 1) % n], 9, g * 3 + k) for g in range(n) for k in range(3)]
    with cf.ProcessPoolExecutor(max_workers=min(12, os.cpu_count() or 1)) as ex:
        ref = list(ex.map(_oracle_selfplay, jobs))
    out = eng.evaluate(torch.from_numpy(genomes).cuda(), seed=9, want_detail=True)
    rewards = out["rewards"].cpu().numpy().reshape(-1); frames = out["frames"].cpu().numpy().reshape(-1)
    assert np.array_equal(rewards, np.array([r for r, _ in ref]))
    assert np.array_equal(frames, np.array([f for _, f in ref]))
    eng.close()


def test_evaluate_population_1024_properties(ngp):
    """BASELINE config 2 at full size: size-independent properties (the oracle cannot run 6144
    episodes in a test): determinism, host-buffer path == device path, shard invariance, bounds."""
    cfg = ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, POPULATION_SIZE=1024)
    eng = ngp.Engine(cfg, device=0)
    genomes = eng.init_population(1024, seed=0)
    a = eng.evaluate(genomes, seed=1, want_detail=True)
    b = eng.evaluate(genomes, seed=1, want_detail=True)
    assert torch.equal(a["fitness"], b["fitness"]) and torch.equal(a["frames"], b["frames"])
    # CTA-synchronous launch geometry (what large populations use) gives the same bits
    for blk in ("128", "256"):
        eng.set_option("rollout_block", int(blk))
        try:
            c = eng.evaluate(genomes, seed=1, want_detail=True)
        finally:
            eng.set_option("rollout_block", 0)
        assert torch.equal(a["fitness"], c["fitness"]) and torch.equal(a["frames"], c["frames"]) and torch.equal(a["rewards"], c["rewards"])
    fit_h, total_h = eng.evaluate_host(genomes.cpu().numpy(), seed=1)
    assert np.array_equal(fit_h, a["fitness"].cpu().numpy()) and total_h == a["frames_total"]
    frames = a["frames"].cpu().numpy()
    assert frames.min() > 60 and frames.max() <= 5 * 2002 + 64
    assert np.isfinite(a["fitness"].cpu().numpy()).all()
    assert a["frames_total"] == int(frames.sum())
    # fitness is the in-order mean of the per-game rewards (main.py:65)
    r = a["rewards"].cpu().numpy()
    acc = np.zeros(1024)
    for k in range(6):
        acc = acc + r[:, k]
    assert np.array_equal(acc / 6.0, a["fitness"].cpu().numpy())
    eng.close()


def test_find_stuff_kernel(engine, golden, obs_npy):
    frames = np.concatenate([golden["fs_frames"], obs_npy[None]])
    loc, valid = engine.find_stuff(torch.from_numpy(frames).cuda())
    loc = loc.cpu().numpy(); valid = valid.cpu().numpy()
    ref_loc = np.concatenate([golden["fs_loc"], np.array([[[111.5, 64.5], [122.5, 17.5], [127.5, 141.5]]])])
    ref_valid = np.concatenate([golden["fs_valid"], np.ones((1, 3), np.uint8)])
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(loc[ref_valid == 1], ref_loc[ref_valid == 1].astype(np.float32))
    # ragged / edge: a single frame, all-zero frames (the reference's None case, tests.py:41,53)
    z = torch.zeros((3, 210, 160, 3), dtype=torch.uint8, device="cuda")
    _, v = engine.find_stuff(z)
    assert v.sum().item() == 0


def test_find_stuff_matches_fused_observation(engine):
    """K2 on frames produced by K1 equals K1's own fused observation (size-independent property)."""
    engine.env_reset(64, 1)
    rng = np.random.RandomState(0)
    for f in range(120):
        a = np.zeros((64, 16), np.uint8); a[:, 0] = 1; a[:, 15] = 1
        a[:, 4] = rng.randint(0, 2, 64); a[:, 7] = rng.randint(0, 2, 64)
        out = engine.env_step(torch.from_numpy(a).cuda())
        loc, valid = engine.find_stuff(out["frames"])
        assert torch.equal(valid, out["valid"]) and torch.equal(loc * valid[..., None], out["loc"] * out["valid"][..., None])


def test_mlp_forward_kernel(ngp, golden):
    """rtol 1e-5 against numpy_nn in FP32, argmax agreement 100 % (BASELINE north_star)."""
    for name in ("mlp_default", "mlp_default_wide_range", "mlp_mid"):
        nodes = tuple(int(v) for v in golden[name + "_nodes"])
        eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=nodes), device=0)
        g = torch.from_numpy(golden[name + "_genomes"]).cuda()
        x = torch.from_numpy(golden[name + "_x"].astype(np.float32)).cuda()
        act, out = eng.mlp_forward(g, x)
        np.testing.assert_allclose(out.cpu().numpy(), golden[name + "_out"], rtol=1e-5, atol=1e-7)
        assert np.array_equal(act.cpu().numpy(), golden[name + "_act"])
        eng.close()
    eng = ngp.Engine(ngp.Config(), device=0)
    g = torch.tensor([[0.0] * 14 + [50.0] * 3 + [60.0] * 3], dtype=torch.float32, device="cuda")
    act, _ = eng.mlp_forward(g, torch.full((1, 1, 6), 0.5, dtype=torch.float32, device="cuda"))
    assert act.item() == ngp.ACT_UP          # saturation tie rule, SURVEY Appendix A13
    eng.close()


def test_mlp_forward_wide(ngp, golden):
    nodes = (6, 512, 512, 2)
    eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=nodes), device=0)
    assert eng.gene_size == 267266
    genomes = np.stack([(np.random.RandomState(9000 + g).standard_normal(267266) * 0.05).astype(np.float32) for g in range(4)])
    act, out = eng.mlp_forward(torch.from_numpy(genomes).cuda(), torch.from_numpy(golden["mlp_wide_x"].astype(np.float32)).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), golden["mlp_wide_out"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(act.cpu().numpy(), golden["mlp_wide_act"])
    eng.close()


def test_ga_step_injected_noise_parity(ngp):
    """selection indices and children bit-exact given identical noise (BASELINE north_star)."""
    import oracle
    for n in (64, 65, 256):
        cfg = ngp.Config(POPULATION_SIZE=n)
        eng = ngp.Engine(cfg, device=0)
        G, T = eng.gene_size, cfg.TOURNAMENT_SIZE
        rng = np.random.RandomState(n)
        genomes = rng.random_sample((n, G)).astype(np.float32)
        fitness = np.round(rng.standard_normal(n), 1)                    # rounded: exercises ties
        noise = dict(sel_draws=rng.randint(0, n, size=(n, T)).astype(np.int32), cx_do=(rng.random_sample(n // 2) < 0.9).astype(np.uint8),
                     cx_u=rng.random_sample((n // 2, G)).astype(np.float32), mut_do=(rng.random_sample(n) < 0.9).astype(np.uint8),
                     mut_u=rng.random_sample((n, G)).astype(np.float32), mut_z=rng.standard_normal((n, G)).astype(np.float32))
        out = eng.ga_step(torch.from_numpy(genomes).cuda(), torch.from_numpy(fitness).cuda(), noise={k: torch.from_numpy(v).cuda() for k, v in noise.items()})
        parents = oracle.sel_tournament(fitness, noise["sel_draws"])
        assert np.array_equal(out["parent_idx"].cpu().numpy(), parents)
        child, invalid = oracle.var_and(genomes[parents], noise["cx_do"], noise["cx_u"], noise["mut_do"], noise["mut_u"], noise["mut_z"],
                                        cfg.CROSSOVER_BLEND_ALPHA, cfg.GAUSSIAN_MUTATION_MEAN, cfg.GAUSSIAN_MUTATION_SIGMA,
                                        cfg.PROBABILITY_OF_MUTATING_A_SINGLE_GENE)
        assert np.array_equal(out["genomes"].cpu().numpy(), child)
        assert np.array_equal(out["invalid"].cpu().numpy(), invalid)
        st = out["stats"].cpu().numpy()
        np.testing.assert_allclose(st, [fitness.mean(), fitness.std(), fitness.min(), fitness.max()], rtol=1e-12, atol=1e-15)
        eng.close()


def test_ga_step_philox_statistics(ngp):
    cfg = ngp.Config(POPULATION_SIZE=4096)
    eng = ngp.Engine(cfg, device=0)
    genomes = eng.init_population(4096, seed=7)
    g = genomes.cpu().numpy()
    assert 0.0 <= g.min() and g.max() < 1.0 and abs(g.mean() - 0.5) < 0.01
    fitness = torch.arange(4096, dtype=torch.float64, device="cuda")
    a = eng.ga_step(genomes, fitness, seed=3, generation=1)
    b = eng.ga_step(genomes, fitness, seed=3, generation=1)
    c = eng.ga_step(genomes, fitness, seed=3, generation=2)
    assert torch.equal(a["genomes"], b["genomes"]) and torch.equal(a["parent_idx"], b["parent_idx"])
    assert not torch.equal(a["parent_idx"], c["parent_idx"])
    p = a["parent_idx"].cpu().numpy()
    assert p.min() >= 0 and p.max() < 4096 and p.mean() > 4000          # tournament of 1024 picks near-best
    frac_invalid = a["invalid"].float().mean().item()
    assert abs(frac_invalid - 0.99) < 0.01                                # 1-(1-.9)(1-.9)
    d = (a["genomes"] - genomes[a["parent_idx"].long()]).cpu().numpy()
    assert abs(d.std() - np.sqrt(0.9 * 0.81 * 0.9 + 0.0)) < 0.35         # mutation noise present, finite
    eng.close()


def test_env_step_fast_flavour_equals_verify_flavour(ngp):
    """ngp_env_step without a frame request runs the fused no-framebuffer flavour (quick span accounting);
    RAM and observation must equal the pixel-rendering flavour frame by frame."""
    rng = np.random.RandomState(8)
    n, frames = 64, 400
    acts = np.zeros((frames, n, 16), np.uint8); acts[:, :, 0] = 1; acts[:, :, 15] = 1
    for f in range(frames):
        if f % 6 == 0:
            r = rng.randint(0, 3, n); l = rng.randint(0, 3, n)
        acts[f, :, 4] = r == 1; acts[f, :, 5] = r == 2; acts[f, :, 6] = l == 1; acts[f, :, 7] = l == 2
    outs = []
    for want_frames, core in ((True, ngp.CORE_INTERPRETER), (False, ngp.CORE_TRANSLATED), (False, ngp.CORE_INTERPRETER)):
        eng = ngp.Engine(ngp.Config(), device=0)
        eng.env_reset(n, ngp.STATE_START_2P)
        rec = []
        for f in range(frames):
            o = eng.env_step(torch.from_numpy(acts[f]).cuda(), want_frames=want_frames, core=core)
            rec.append((o["ram"].clone(), o["loc"] * o["valid"][..., None], o["valid"].clone()))
        outs.append(rec)
        eng.close()
    for f in range(frames):
        for k in (1, 2):
            assert torch.equal(outs[0][f][0], outs[k][f][0]), f"RAM frame {f}"
            assert torch.equal(outs[0][f][2], outs[k][f][2]) and torch.equal(outs[0][f][1], outs[k][f][1]), f"observation frame {f}"


@pytest.mark.parametrize("nodes,envs", [((6, 512, 512, 2), 64), ((6, 512, 512, 2), 40), ((6, 128, 192, 2), 64), ((6, 64, 64, 64, 2), 17),
                                        ((6, 512, 512, 2), 128), ((6, 200, 72, 2), 100), ((6, 512, 512, 512, 2), 130)])
def test_mlp_forward_tensor_core_path(ngp, nodes, envs):
    """Wide hidden layers with >= 16 environments per genome run on tcgen05 (3xTF32, TMEM accumulators).  Tolerance:
    rtol 1e-5 against the FP64 oracle (the north_star's FP32 bar; plain TF32 would be ~1e-3), argmax 100 %; the
    FP32 FFMA path must agree to the same tolerance.  Ragged shapes: envs not a multiple of 64, outputs not of 128."""
    import oracle
    eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=nodes), device=0)
    rng = np.random.RandomState(envs)
    n = 3
    genomes = (rng.standard_normal((n, eng.gene_size)) * 0.08).astype(np.float32)
    x = rng.random_sample((n, envs, 6)).astype(np.float32)
    act, out = eng.mlp_forward(torch.from_numpy(genomes).cuda(), torch.from_numpy(x).cuda())
    eng.set_option("mlp_no_tf32", 1)
    try:
        act_f, out_f = eng.mlp_forward(torch.from_numpy(genomes).cuda(), torch.from_numpy(x).cuda())
    finally:
        eng.set_option("mlp_no_tf32", 0)
    ref_out = np.zeros((n, envs, nodes[-1])); ref_act = np.zeros((n, envs), np.uint8)
    for g in range(n):
        for e in range(envs):
            o, a = oracle.mlp_forward(list(nodes), genomes[g], x[g, e].astype(np.float64))
            ref_out[g, e] = o; ref_act[g, e] = a
    np.testing.assert_allclose(out.cpu().numpy(), ref_out, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out_f.cpu().numpy(), ref_out, rtol=1e-5, atol=1e-7)
    assert np.array_equal(act.cpu().numpy(), ref_act) and np.array_equal(act_f.cpu().numpy(), ref_act)
    # prepared genome set (ngp_mlp_prepare + ngp_mlp_forward_prepared): weights streamed from the packed copy into tensor memory
    gd = torch.from_numpy(genomes).cuda()
    eng.mlp_prepare(gd)
    act_p, out_p = eng.mlp_forward_prepared(gd, torch.from_numpy(x).cuda())
    np.testing.assert_allclose(out_p.cpu().numpy(), ref_out, rtol=1e-5, atol=1e-7)
    assert np.array_equal(act_p.cpu().numpy(), ref_act)
    act_p2, out_p2 = eng.mlp_forward_prepared(gd, torch.from_numpy(x).cuda())
    assert torch.equal(out_p, out_p2) and torch.equal(act_p, act_p2)
    with pytest.raises(ngp.NgpError):                   # another genome set was not prepared
        eng.mlp_forward_prepared(gd.clone(), torch.from_numpy(x).cuda())
    eng.close()


def test_translated_core_refuses_another_cartridge(ngp):
    """The statically translated 6507 core and its super-blocks are generated from the bundled cartridge; a different
    2 KiB image must be refused on that path (error, not silently wrong frames) while the interpreter core still runs it."""
    from neuro_genetic_pong_self_play_b200.engine import load_rom
    rom = bytearray(load_rom())
    rom[0x7F0] ^= 0xFF                                   # a byte of padding: never executed by the cartridge's code
    eng = ngp.Engine(ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, POPULATION_SIZE=2, MAX_FRAMES=5), device=0, rom=bytes(rom))
    g = eng.init_population(2, seed=1)
    with pytest.raises(ngp.NgpError):
        eng.evaluate(g, seed=1)
    eng.close()
    eng = ngp.Engine(ngp.Config(SCHEDULE=ngp.SCHEDULE_ROUND_ROBIN, POPULATION_SIZE=2, MAX_FRAMES=5, CORE=ngp.CORE_INTERPRETER), device=0, rom=bytes(rom))
    out = eng.evaluate(eng.init_population(2, seed=1), seed=1)
    assert int(out["frames_total"]) == 2 * 6 * 5
    eng.close()


def test_find_stuff_kernel_dense_random_frames(engine):
    """K2 on frames where target bytes are everywhere (every 16-byte vector takes the exact-accounting path, several targets
    and single-channel matches inside one vector, matches in the first and last vector of a row) and on sparse ones, against
    the restated find_stuff.  Sums are integers, so the comparison is exact."""
    import oracle
    rng = np.random.RandomState(123)
    cols = np.array([[236, 236, 236], [213, 130, 74], [92, 186, 92], [144, 72, 17], [0, 0, 0]], np.uint8)
    frames = []
    for density in (0.9, 0.3, 0.02, 0.0005):
        f = np.empty((210, 160, 3), np.uint8); f[:] = cols[3]
        pick = rng.random_sample((210, 160)) < density
        f[pick] = cols[rng.randint(0, 5, pick.sum())]
        byte_noise = rng.random_sample((210, 160, 3)) < density * 0.2            # single-channel matches
        f[byte_noise] = rng.choice(cols[:3].ravel(), byte_noise.sum())
        frames.append(f)
    edge = np.zeros((210, 160, 3), np.uint8); edge[34, 0] = cols[0]; edge[193, 159] = cols[1]; edge[100, 5] = cols[2]; edge[33, 7] = cols[0]; edge[194, 9] = cols[2]
    frames.append(edge)
    frames = np.stack(frames)
    loc, valid = engine.find_stuff(torch.from_numpy(frames).cuda())
    loc = loc.cpu().numpy(); valid = valid.cpu().numpy()
    for i, f in enumerate(frames):
        rl, rv = oracle.find_stuff(f)
        assert np.array_equal(valid[i], rv), i
        assert np.array_equal(loc[i][rv == 1], rl[rv == 1].astype(np.float32)), i


@pytest.mark.parametrize("core", [0, 1])
def test_cuda_core_reproduces_obs_npy_pixel_for_pixel(engine, obs_npy, core):
    """The committed button trace (tests/golden/obs_trace.npz) through ngp_env_step: the frame equals the reference's real
    gym-retro frame obs.npy byte for byte, and the fused find_stuff result is the reference's."""
    tr = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs_trace.npz"))
    engine.env_reset(2, int(tr["state"]))
    for a in tr["actions"][:-1]:
        engine.env_step(torch.from_numpy(np.stack([a, a])).cuda(), want_frames=False, want_obs=False, core=core)
    a = tr["actions"][-1]
    out = engine.env_step(torch.from_numpy(np.stack([a, a])).cuda(), core=core)
    for j in range(2):
        assert np.array_equal(out["frames"][j].cpu().numpy(), obs_npy)
    assert out["valid"][0].cpu().tolist() == [1, 1, 1]
    assert out["loc"][0].cpu().tolist() == [[111.5, 64.5], [122.5, 17.5], [127.5, 141.5]]
    loc, valid = engine.find_stuff(out["frames"])
    assert torch.equal(loc[0], out["loc"][0]) and valid[0].cpu().tolist() == [1, 1, 1]
