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
        L.eo_philox_bit.restype = ctypes.c_uint32; L.eo_philox_bit.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.eo_philox4x32.argtypes = [vp, vp, vp]
        L.eo_action_to_input.argtypes = [vp, ctypes.POINTER(Input)]
        L.eo_reset_to_state.argtypes = [vp, ctypes.c_int]
        L.eo_episode.argtypes = [vp, ctypes.POINTER(Shape), Policy, Policy, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint32,
                                 ctypes.c_int, ctypes.POINTER(EpisodeResult), vp, ctypes.c_int]
        L.eo_evaluate.restype = ctypes.c_double
        L.eo_evaluate.argtypes = [ctypes.c_char_p, ctypes.POINTER(Shape), vp, vp, vp, ctypes.c_int, vp, ctypes.c_uint64,
                                  ctypes.c_uint32, vp, vp]
        L.eo_selfplay_game.argtypes = [ctypes.c_char_p, ctypes.POINTER(Shape), vp, vp, ctypes.c_uint64, ctypes.c_uint32,
                                       ctypes.POINTER(EpisodeResult)]
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
        lib().eo_action_to_input(_ptr(a), ctypes.byref(inp))
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

    def episode(self, shape: Shape, left, right, mult=1.0, seed=0, env_id=0, max_frames=0, trace_cap=0):
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
        lib().eo_episode(self._h, ctypes.byref(shape), pol(left), pol(right), float(mult), int(seed), int(env_id),
                         int(max_frames), ctypes.byref(res), _ptr(trace) if trace_cap else None, trace_cap)
        return res, trace[: min(res.frames, trace_cap)]


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


def det_exp(x: float) -> float:
    return lib().eo_det_exp(float(x))


def philox4x32(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().eo_philox4x32(_ptr(c), _ptr(k), _ptr(o))
    return o


def evaluate(nodes, genome, hof_genomes=None, hof_fitness=None, hof_pick=(0, 0, 0), seed=0, genome_id=0, bias=True):
    sh = Shape.make(nodes, bias)
    g = np.ascontiguousarray(genome, np.float32)
    n_hof = 0 if hof_genomes is None else len(hof_genomes)
    hg = np.ascontiguousarray(hof_genomes if n_hof else np.zeros((1, len(g))), np.float32)
    hf = np.ascontiguousarray(hof_fitness if n_hof else np.zeros(1), np.float64)
    pick = np.ascontiguousarray(hof_pick, np.int32)
    rewards = np.zeros(GAMES_TO_PLAY, np.float64); frames = np.zeros(GAMES_TO_PLAY, np.int32)
    fit = lib().eo_evaluate(load_rom(), ctypes.byref(sh), _ptr(g), _ptr(hg), _ptr(hf), n_hof, _ptr(pick), int(seed),
                            int(genome_id), _ptr(rewards), _ptr(frames))
    return fit, rewards, frames


def selfplay_game(nodes, right, left, seed=0, env_id=0, bias=True):
    sh = Shape.make(nodes, bias)
    r = np.ascontiguousarray(right, np.float32); l = np.ascontiguousarray(left, np.float32)
    res = EpisodeResult()
    lib().eo_selfplay_game(load_rom(), ctypes.byref(sh), _ptr(r), _ptr(l), int(seed), int(env_id), ctypes.byref(res))
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
