"""CPU ORACLE bindings (test infrastructure, NOT product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  It wraps ``oracle/_build/liboracle.so``
(built from a2600_oracle.c + episode_oracle.c by ``oracle/Makefile``) and holds numpy
restatements of the reference's DEAP operators (ga.py:77-94) that take
their noise explicitly.

Parity status: the numpy-level pieces (find_stuff, MLP, reward, clamp, bots) are pinned
against the imported reference (tools/make_golden.py -> tests/golden/reference_vectors.npz).
The emulator and the DEAP operators live in third-party packages that are absent from
/root/reference (gym-retro/Stella, deap; requirements.txt:2-3, unpinned): PARITY UNPINNED.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
ROM_PATH = os.path.join(os.path.dirname(_HERE), "neuro_genetic_pong_self_play_b200", "data", "video_olympics.a26")

ACT_NONE, ACT_UP, ACT_DOWN = 0, 1, 2
POLICY_HARDCODED, POLICY_SCORE_HARDCODED, POLICY_MLP = 0, 1, 2
STATE_START_1P, STATE_START_2P = 0, 1
GAMES_TO_PLAY = 6
TRACE_BYTES = 144


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("a2600_oracle.c", "episode_oracle.c", "a2600_oracle.h", "episode_oracle.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class Shape(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int), ("nodes", ctypes.c_int * 8), ("bias", ctypes.c_int)]

    @classmethod
    def make(cls, nodes, bias=True):
        arr = (ctypes.c_int * 8)(*(list(nodes) + [0] * (8 - len(nodes))))
        return cls(len(nodes), arr, 1 if bias else 0)


class Policy(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("genome", ctypes.POINTER(ctypes.c_float))]


class Input(ctypes.Structure):
    _fields_ = [("swchb", ctypes.c_uint8), ("fire", ctypes.c_uint8), ("dec", ctypes.c_uint8), ("inc", ctypes.c_uint8)]


class Obs(ctypes.Structure):
    _fields_ = [("loc", (ctypes.c_double * 2) * 3), ("valid", ctypes.c_uint8 * 3)]


class EpisodeResult(ctypes.Structure):
    _fields_ = [("frames", ctypes.c_int), ("score1", ctypes.c_int), ("score2", ctypes.c_int),
                ("total_frames", ctypes.c_double), ("reward", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, u8p, fp, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
        L.a26o_new.restype = vp; L.a26o_new.argtypes = [ctypes.c_char_p]
        L.a26o_free.argtypes = [vp]
        L.a26o_power_on.argtypes = [vp]
        L.a26o_run_frame.restype = ctypes.c_int; L.a26o_run_frame.argtypes = [vp, ctypes.POINTER(Input), vp]
        L.a26o_ram.restype = u8p; L.a26o_ram.argtypes = [vp]
        L.a26o_cpu_regs.argtypes = [vp, vp]
        L.a26o_cycles.restype = ctypes.c_uint64; L.a26o_cycles.argtypes = [vp]
        L.a26o_instructions.restype = ctypes.c_uint64; L.a26o_instructions.argtypes = [vp]
        L.a26o_tia_digest.argtypes = [vp, vp]
        L.a26o_fb_to_rgb.argtypes = [vp, vp]
        L.a26o_build_paddle_table.argtypes = [vp]
        L.eo_gene_size.restype = ctypes.c_int; L.eo_gene_size.argtypes = [ctypes.POINTER(Shape)]
        L.eo_det_exp.restype = ctypes.c_double; L.eo_det_exp.argtypes = [ctypes.c_double]
        L.eo_mlp_forward.argtypes = [ctypes.POINTER(Shape), vp, vp, vp, ctypes.POINTER(ctypes.c_int)]
        L.eo_find_stuff.argtypes = [vp, ctypes.POINTER(Obs)]
        L.eo_clamp.restype = ctypes.c_int; L.eo_clamp.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.eo_reward.restype = ctypes.c_double; L.eo_reward.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.eo_philox_bit.restype = ctypes.c_uint32; L.eo_philox_bit.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.eo_philox4x32.argtypes = [vp, vp, vp]
        L.eo_bot_act.restype = ctypes.c_int; L.eo_bot_act.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_int]
        L.eo_inference_vector.argtypes = [vp, vp, ctypes.c_double, ctypes.c_double, vp]
        L.eo_action_to_input.argtypes = [vp, ctypes.c_int, ctypes.POINTER(Input)]
        L.eo_reset_to_state.argtypes = [vp, ctypes.c_int]
        L.eo_set_button_map.argtypes = [vp]; L.eo_get_button_map.argtypes = [vp]
        L.eo_episode.argtypes = [vp, ctypes.POINTER(Shape), Policy, Policy, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                 ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(EpisodeResult), vp, ctypes.c_int]
        L.eo_evaluate.restype = ctypes.c_double
        L.eo_evaluate.argtypes = [ctypes.c_char_p, ctypes.POINTER(Shape), vp, vp, vp, ctypes.c_int, vp, ctypes.c_uint64,
                                  ctypes.c_uint64, ctypes.c_uint32, vp, vp]
        L.eo_selfplay_game.argtypes = [ctypes.c_char_p, ctypes.POINTER(Shape), vp, vp, ctypes.c_uint64, ctypes.c_uint64,
                                       ctypes.c_uint32, ctypes.POINTER(EpisodeResult)]
        _lib = L
    return _lib


def load_rom() -> bytes:
    with open(ROM_PATH, "rb") as f:
        rom = f.read()
    assert len(rom) == 2048
    return rom


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class Atari:
    """Single-environment oracle emulator (gym-retro env.step/reset stand-in)."""

    def __init__(self, rom: bytes | None = None):
        self._rom = rom or load_rom()
        self._h = ctypes.c_void_p(lib().a26o_new(self._rom))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().a26o_free(self._h)
            self._h = None

    def power_on(self):
        lib().a26o_power_on(self._h)

    def reset_to_state(self, state_id: int):
        lib().eo_reset_to_state(self._h, state_id)
        self.players = 1 if state_id == STATE_START_1P else 2       # main.py:40 makes the robot game with players=1

    def run_frame(self, swchb=0x3F, fire=0, dec=0, inc=0, want_frame=True):
        fb = np.zeros((210, 160), np.uint8) if want_frame else None
        inp = Input(swchb, fire, dec, inc)
        err = lib().a26o_run_frame(self._h, ctypes.byref(inp), _ptr(fb) if want_frame else None)
        if err:
            raise RuntimeError(f"oracle emulator error {err}")
        return fb

    def step(self, action16, want_frame=True):
        a = np.ascontiguousarray(action16, dtype=np.uint8)
        inp = Input()
        lib().eo_action_to_input(_ptr(a), int(getattr(self, "players", 2)), ctypes.byref(inp))
        return self.run_frame(inp.swchb, inp.fire, inp.dec, inp.inc, want_frame)

    @property
    def ram(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().a26o_ram(self._h), shape=(128,)).copy()

    @property
    def cpu_regs(self) -> np.ndarray:
        out = np.zeros(8, np.uint8)
        lib().a26o_cpu_regs(self._h, _ptr(out))
        return out

    @property
    def tia_digest(self) -> np.ndarray:
        out = np.zeros(8, np.uint32)
        lib().a26o_tia_digest(self._h, _ptr(out))
        return out

    @property
    def cycles(self) -> int:
        return int(lib().a26o_cycles(self._h))

    @property
    def instructions(self) -> int:
        return int(lib().a26o_instructions(self._h))

    def episode(self, shape: Shape, left, right, mult=1.0, seed=0, env_id=0, max_frames=0, trace_cap=0, generation=0, players=None):
        """left/right: ("hardcoded"|"score", None) or ("mlp", float32 genome)."""
        keep = []

        def pol(p):
            kind, g = p
            if kind == "mlp":
                g = np.ascontiguousarray(g, np.float32)
                keep.append(g)
                return Policy(POLICY_MLP, g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
            return Policy(POLICY_HARDCODED if kind == "hardcoded" else POLICY_SCORE_HARDCODED, None)

        res = EpisodeResult()
        trace = np.zeros((max(trace_cap, 1), TRACE_BYTES), np.uint8)
        if players is None:
            players = int(getattr(self, "players", 2))
        lib().eo_episode(self._h, ctypes.byref(shape), pol(left), pol(right), float(mult), int(seed), int(generation), int(env_id),
                         int(players), int(max_frames), ctypes.byref(res), _ptr(trace) if trace_cap else None, trace_cap)
        return res, trace[: min(res.frames, trace_cap)]


def button_map() -> np.ndarray:
    m = np.zeros(16, np.uint8)
    lib().eo_get_button_map(_ptr(m))
    return m


def set_button_map(m):
    m = np.ascontiguousarray(m, np.uint8)
    assert m.shape == (16,)
    lib().eo_set_button_map(_ptr(m))


def fb_to_rgb(fb: np.ndarray) -> np.ndarray:
    fb = np.ascontiguousarray(fb, np.uint8)
    rgb = np.zeros((210, 160, 3), np.uint8)
    lib().a26o_fb_to_rgb(_ptr(fb), _ptr(rgb))
    return rgb


def ntsc_palette() -> np.ndarray:
    pal = (ctypes.c_uint32 * 128).in_dll(lib(), "a26o_ntsc_palette")
    return np.array(pal, dtype=np.uint32)


def paddle_table() -> np.ndarray:
    t = np.zeros(4097, np.uint32)
    lib().a26o_build_paddle_table(_ptr(t))
    return t


def find_stuff(rgb: np.ndarray):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    ob = Obs()
    lib().eo_find_stuff(_ptr(rgb), ctypes.byref(ob))
    loc = np.array([[ob.loc[t][0], ob.loc[t][1]] for t in range(3)], np.float64)
    valid = np.array([ob.valid[t] for t in range(3)], np.uint8)
    return loc, valid


def mlp_forward(nodes, genome, x, bias=True):
    sh = Shape.make(nodes, bias)
    g = np.ascontiguousarray(genome, np.float32)
    xv = np.ascontiguousarray(x, np.float64)
    out = np.zeros(nodes[-1], np.float64)
    act = ctypes.c_int(0)
    lib().eo_mlp_forward(ctypes.byref(sh), _ptr(g), _ptr(xv), _ptr(out), ctypes.byref(act))
    return out, act.value


def bot_act(kind: int, x, score1: int = 0, score2: int = 0) -> int:
    xv = np.ascontiguousarray(x, np.float64)
    return int(lib().eo_bot_act(kind, _ptr(xv), score1, score2))


def inference_vector(ball, last, me_row: float, enemy_row: float) -> np.ndarray:
    b = np.ascontiguousarray(ball, np.float64); l = np.ascontiguousarray(last, np.float64); out = np.zeros(6, np.float64)
    lib().eo_inference_vector(_ptr(b), _ptr(l), float(me_row), float(enemy_row), _ptr(out))
    return out


def det_exp(x: float) -> float:
    return lib().eo_det_exp(float(x))


def philox4x32(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().eo_philox4x32(_ptr(c), _ptr(k), _ptr(o))
    return o


def evaluate(nodes, genome, hof_genomes=None, hof_fitness=None, hof_pick=(0, 0, 0), seed=0, genome_id=0, bias=True, generation=0):
    sh = Shape.make(nodes, bias)
    g = np.ascontiguousarray(genome, np.float32)
    n_hof = 0 if hof_genomes is None else len(hof_genomes)
    hg = np.ascontiguousarray(hof_genomes if n_hof else np.zeros((1, len(g))), np.float32)
    hf = np.ascontiguousarray(hof_fitness if n_hof else np.zeros(1), np.float64)
    pick = np.ascontiguousarray(hof_pick, np.int32)
    rewards = np.zeros(GAMES_TO_PLAY, np.float64); frames = np.zeros(GAMES_TO_PLAY, np.int32)
    fit = lib().eo_evaluate(load_rom(), ctypes.byref(sh), _ptr(g), _ptr(hg), _ptr(hf), n_hof, _ptr(pick), int(seed),
                            int(generation), int(genome_id), _ptr(rewards), _ptr(frames))
    return fit, rewards, frames


def selfplay_game(nodes, right, left, seed=0, env_id=0, bias=True, generation=0):
    sh = Shape.make(nodes, bias)
    r = np.ascontiguousarray(right, np.float32); l = np.ascontiguousarray(left, np.float32)
    res = EpisodeResult()
    lib().eo_selfplay_game(load_rom(), ctypes.byref(sh), _ptr(r), _ptr(l), int(seed), int(generation), int(env_id), ctypes.byref(res))
    return res


# ----------------------------------------------------------------------------------------
# DEAP operator restatements with explicit noise (ga.py:77-94).
# Genes are float32 (the product's genome dtype); arithmetic order is fixed so the CUDA
# kernels can match bit-for-bit.
# ----------------------------------------------------------------------------------------

def sel_tournament(fitness: np.ndarray, draws: np.ndarray) -> np.ndarray:
    """deap.tools.selTournament: draws[k, T] aspirant indices; winner = max fitness,
    first maximal in draw order (ga.py:94)."""
    fitness = np.asarray(fitness, np.float64)
    out = np.empty(draws.shape[0], np.int32)
    for k in range(draws.shape[0]):
        best = draws[k, 0]
        for j in draws[k, 1:]:
            if fitness[j] > fitness[best]:
                best = j
        out[k] = best
    return out


def cx_blend(x1: np.ndarray, x2: np.ndarray, u: np.ndarray, alpha: float):
    """deap.tools.cxBlend (ga.py:89): gamma=(1+2a)u-a; c1=(1-g)x1+g x2; c2=g x1+(1-g)x2, in float32."""
    f = np.float32
    x1 = x1.astype(f); x2 = x2.astype(f); u = u.astype(f)
    g = f(1.0 + 2.0 * alpha) * u - f(alpha)
    one = f(1.0)
    c1 = (one - g) * x1 + g * x2
    c2 = g * x1 + (one - g) * x2
    return c1.astype(f), c2.astype(f)


def mut_gaussian(x: np.ndarray, gene_u: np.ndarray, z: np.ndarray, mu: float, sigma: float, indpb: float):
    """deap.tools.mutGaussian (ga.py:91-92): per gene, if u < indpb: x += mu + sigma*z (z ~ N(0,1))."""
    f = np.float32
    x = x.astype(f)
    step = f(mu) + f(sigma) * z.astype(f)
    return np.where(gene_u.astype(f) < f(indpb), x + step, x).astype(f)


def var_and(parents: np.ndarray, cx_do: np.ndarray, cx_u: np.ndarray, mut_do: np.ndarray, mut_u: np.ndarray,
            mut_z: np.ndarray, alpha: float, mu: float, sigma: float, indpb: float):
    """deap.algorithms.varAnd on already-selected (cloned) parents: pairs (0,1),(2,3).. mate when
    cx_do[pair]; then every individual mutates when mut_do[i].  Returns children, invalid flags."""
    n = parents.shape[0]
    child = parents.astype(np.float32).copy()
    invalid = np.zeros(n, np.uint8)
    for p in range(n // 2):
        if cx_do[p]:
            a, b = cx_blend(child[2 * p], child[2 * p + 1], cx_u[p], alpha)
            child[2 * p], child[2 * p + 1] = a, b
            invalid[2 * p] = invalid[2 * p + 1] = 1
    for i in range(n):
        if mut_do[i]:
            child[i] = mut_gaussian(child[i], mut_u[i], mut_z[i], mu, sigma, indpb)
            invalid[i] = 1
    return child, invalid


def hall_of_fame_update(hof_genomes, hof_fitness, pop, fitness, maxsize):
    """deap.tools.HallOfFame.update: keep the best `maxsize` distinct individuals, best first."""
    hof_g = [np.asarray(g, np.float32) for g in hof_genomes]
    hof_f = list(hof_fitness)
    for g, f in zip(pop, fitness):
        g = np.asarray(g, np.float32)
        if len(hof_g) == 0 and maxsize != 0:
            hof_g.append(g.copy()); hof_f.append(float(f))
            continue
        if f > hof_f[-1] or len(hof_g) < maxsize:
            if any(np.array_equal(g, h) for h in hof_g):
                continue
            if len(hof_g) >= maxsize:
                hof_g.pop(); hof_f.pop()
            # insert sorted, best first; DEAP's bisect_right on the ascending key list puts a new
            # individual BEFORE existing members of equal fitness
            pos = 0
            while pos < len(hof_f) and hof_f[pos] > f:
                pos += 1
            hof_g.insert(pos, g.copy()); hof_f.insert(pos, float(f))
    return hof_g, hof_f


# ----------------------------------------------------------------------------------------
# Restatement of the multi-GPU exchange record (include/ngp.h: ngp_pack_elites / ngp_unpack_elites):
#   [ n_local i32 | k_local i32 | 8 B pad | fitness f64[n_max] | elite_fitness f64[k_max] | elite_genomes f32[k_max][G] ]
# padded to a multiple of 16 bytes.  Elites = the k best of the shard, best first, ties by lower index.
# ----------------------------------------------------------------------------------------
def exchange_bytes(n_max: int, k_max: int, G: int) -> int:
    raw = 16 + n_max * 8 + k_max * 8 + k_max * G * 4
    return (raw + 15) & ~15


def pack_record(genomes: np.ndarray, fitness: np.ndarray, k: int, n_max: int, k_max: int) -> np.ndarray:
    n, G = genomes.shape
    rec = np.zeros(exchange_bytes(n_max, k_max, G), np.uint8)
    rec[:8] = np.array([n, k], np.int32).view(np.uint8)
    order = np.argsort(-np.asarray(fitness, np.float64), kind="stable")[:k]
    f = np.zeros(n_max, np.float64); f[:n] = fitness
    ef = np.zeros(k_max, np.float64); ef[:k] = np.asarray(fitness, np.float64)[order]
    eg = np.zeros((k_max, G), np.float32); eg[:k] = genomes[order]
    o = 16
    rec[o:o + n_max * 8] = f.view(np.uint8); o += n_max * 8
    rec[o:o + k_max * 8] = ef.view(np.uint8); o += k_max * 8
    rec[o:o + k_max * G * 4] = eg.reshape(-1).view(np.uint8)
    return rec


def unpack_records(gathered: np.ndarray, world: int, n_max: int, k_max: int, G: int):
    size = exchange_bytes(n_max, k_max, G)
    fit, eg, ef = [], [], []
    for r in range(world):
        rec = np.ascontiguousarray(gathered[r * size:(r + 1) * size])
        n, k = rec[:8].view(np.int32)
        o = 16
        fit.append(rec[o:o + n_max * 8].view(np.float64)[:n]); o += n_max * 8
        ef.append(rec[o:o + k_max * 8].view(np.float64)[:k]); o += k_max * 8
        eg.append(rec[o:o + k_max * G * 4].view(np.float32).reshape(k_max, G)[:k])
    return np.concatenate(fit), np.concatenate(eg), np.concatenate(ef)


# ----------------------------------------------------------------------------------------
# Restatement of the product's Philox4x32-10 counter layouts (csrc/ga_streams.cuh, rollout.cuh plan_env): the noise the
# CUDA kernels draw when none is injected.  Vectorised numpy Philox, pinned to the C one (and so to the Random123 KATs).
# ----------------------------------------------------------------------------------------
STREAM_SELECT, STREAM_CXDO, STREAM_CXU, STREAM_MUTDO = 0x53454C31, 0x43584431, 0x43585531, 0x4D544431
STREAM_MUTU, STREAM_MUTZ, STREAM_INIT, STREAM_HOF = 0x4D545531, 0x4D545A31, 0x494E4931, 0x484F4621


def philox4x32_np(c0, c1, c2, c3, seed: int):
    """Vectorised Philox4x32-10: counters are broadcastable integer arrays, key = (seed lo, seed hi); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, np.uint64) & 0xFFFFFFFF for c in (c0, c1, c2, c3)))
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0; n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1; n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask; k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _u01(words):
    return ((words >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def init_population_philox(n: int, G: int, seed: int) -> np.ndarray:
    total = n * G
    q = np.arange((total + 3) // 4, dtype=np.uint64)
    w = np.stack(philox4x32_np(q & 0xFFFFFFFF, q >> np.uint64(32), 0, STREAM_INIT, seed), axis=1).reshape(-1)[:total]
    return _u01(w).reshape(n, G)


def ga_noise_philox(n: int, G: int, T: int, seed: int, generation: int, cxpb: float, mutpb: float):
    """The noise ngp_ga_step draws for (seed, generation): same keys as the injected-noise interface (ngp_noise)."""
    gen = generation & 0xFFFFFFFF
    nb = (T + 3) // 4
    slot = np.arange(n, dtype=np.uint64)[:, None]; blk = np.arange(nb, dtype=np.uint64)[None, :]
    w = np.stack(philox4x32_np(slot, blk, gen, STREAM_SELECT, seed), axis=2).reshape(n, nb * 4)[:, :T]
    sel = ((w.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int32)
    pairs = n // 2
    cx_do = (_u01(philox4x32_np(np.arange(pairs, dtype=np.uint64), 0, gen, STREAM_CXDO, seed)[0]) < np.float32(cxpb)).astype(np.uint8)
    gb = (G + 3) // 4
    pr = np.arange(pairs, dtype=np.uint64)[:, None]; gq = np.arange(gb, dtype=np.uint64)[None, :]
    cx_u = _u01(np.stack(philox4x32_np(pr, gq, gen, STREAM_CXU, seed), axis=2).reshape(pairs, gb * 4)[:, :G])
    ind = np.arange(n, dtype=np.uint64)
    mut_do = (_u01(philox4x32_np(ind, 0, gen, STREAM_MUTDO, seed)[0]) < np.float32(mutpb)).astype(np.uint8)
    mut_u = _u01(np.stack(philox4x32_np(ind[:, None], gq, gen, STREAM_MUTU, seed), axis=2).reshape(n, gb * 4)[:, :G])
    z0, z1, _, _ = philox4x32_np(ind[:, None], np.arange(G, dtype=np.uint64)[None, :], gen, STREAM_MUTZ, seed)
    u1 = ((z0 >> np.uint32(8)).astype(np.float64) + 1.0) / 16777216.0
    u2 = _u01(z1).astype(np.float64)
    mut_z = (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)       # the kernel: sqrtf, logf, cospif
    return dict(sel_draws=sel, cx_do=cx_do, cx_u=cx_u, mut_do=mut_do, mut_u=mut_u, mut_z=mut_z)


def hof_pick_philox(n: int, n_hof: int, seed: int, generation: int) -> np.ndarray:
    """Hall-of-fame opponents ngp_evaluate draws for games 3..5 of genome g: counter (g, game, generation, 'HOF!')."""
    g = np.arange(n, dtype=np.uint64)[:, None]; k = np.arange(3, 6, dtype=np.uint64)[None, :]
    w = philox4x32_np(g, k, generation & 0xFFFFFFFF, STREAM_HOF, seed)[0]
    return (w % np.uint32(n_hof)).astype(np.int32)
