#!/usr/bin/env python3
"""Stand-alone operator sweep (BASELINE.json configs[4]): K2 find_stuff, K3 grouped MLP forward, K4 GA step on
synthetic obs.npy-shaped inputs, population 2^10..2^20.  Prints one JSON line per measurement with the achieved
algorithmic HBM bandwidth against MEASURED_PEAKS.json."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import neuro_genetic_pong_self_play_b200 as ngp  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    SRC = "measured"
except Exception:
    PEAK, SRC = 6650.0, "fallback"


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def report(op, cfg, seconds, alg_bytes, units, unit_name):
    gbs = alg_bytes / seconds / 1e9
    print(json.dumps({"op": op, "config": cfg, "ms": seconds * 1e3, unit_name + "_per_s": units / seconds, "algorithmic_gbs": gbs,
                      "hbm_peak_gbs": PEAK, "peak_source": SRC, "frac": gbs / PEAK}), flush=True)


def synth_frames(n, seed=0):
    rng = np.random.RandomState(seed)
    f = np.zeros((n, 210, 160, 3), np.uint8)
    f[:] = (144, 72, 17); f[:, 24:34] = 236; f[:, 194:] = 236
    for i in range(n):
        r, c = rng.randint(34, 190), rng.randint(0, 158); f[i, r:r + 4, c:c + 2] = 236
        r = rng.randint(34, 178); f[i, r:r + 16, 16:20] = (213, 130, 74)
        r = rng.randint(34, 178); f[i, r:r + 16, 140:144] = (92, 186, 92)
    return f


def main():
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else ["find_stuff", "mlp_small", "mlp_wide", "ga"]
    eng = ngp.Engine(ngp.Config(), device=0)
    # K2: frames larger than L2 (126 MB): 4096 frames = 413 MB
    base = torch.from_numpy(synth_frames(64)).cuda()
    for n in (1024, 4096, 16384) if "find_stuff" in only else ():
        frames = base.repeat(n // 64, 1, 1, 1).contiguous()
        t = timed(lambda: eng.find_stuff(frames))
        report("find_stuff", {"frames": n, "bytes_per_frame": 76800 + 27}, t, n * (76800 + 27), n, "frames")
        del frames
    # K3 default net: one env per genome, HBM-bound on genomes + inputs + outputs
    for logn in (10, 14, 17, 20) if "mlp_small" in only else ():
        n = 1 << logn
        g = eng.init_population(n, seed=1)
        x = torch.rand((n, 1, 6), device="cuda")
        t = timed(lambda: eng.mlp_forward(g, x, want_out=False))
        report("mlp_forward[6,2,2]", {"genomes": n, "envs": 1}, t, n * (20 * 4 + 24 + 1), n, "inferences")
    # K3 wide net, 64 envs per genome: weights streamed once per genome
    engw = ngp.Engine(ngp.Config(NETWORK_SHAPE=(6, 512, 512, 2)), device=0)
    G = engw.gene_size
    envs_w = int(sys.argv[sys.argv.index("--envs") + 1]) if "--envs" in sys.argv else 64
    for n in (64, 256, 1024) if "mlp_wide" in only else ():
        g = (torch.randn((n, G), device="cuda") * 0.05)
        x = torch.rand((n, envs_w, 6), device="cuda")
        engw.mlp_prepare(g)
        for path in ("prepared_tmem_3xtf32", "tcgen05_3xtf32", "fp32_ffma"):
            if path == "fp32_ffma":
                engw.set_option("mlp_no_tf32", 1)
            fn = (lambda: engw.mlp_forward_prepared(g, x, want_out=False)) if path.startswith("prepared") else (lambda: engw.mlp_forward(g, x, want_out=False))
            t = timed(fn, iters=5, warm=2)
            engw.set_option("mlp_no_tf32", 0)
            flops = 2.0 * G * envs_w * n
            print(json.dumps({"op": "mlp_forward[6,512,512,2]", "path": path, "config": {"genomes": n, "envs": envs_w}, "ms": t * 1e3,
                              "inferences_per_s": n * envs_w / t, "algorithmic_gbs": n * G * 4 / t / 1e9, "hbm_peak_gbs": PEAK,
                              "frac": n * G * 4 / t / 1e9 / PEAK, "tflops_algorithmic": flops / t / 1e12}), flush=True)
    engw.close()
    # K4: GA step, ~3*N*G*4 bytes
    for logn in (10, 14, 17) if "ga" in only else ():
        n = 1 << logn
        e = ngp.Engine(ngp.Config(POPULATION_SIZE=n), device=0)
        g = e.init_population(n, seed=2)
        fit = torch.rand(n, dtype=torch.float64, device="cuda")
        t = timed(lambda: e.ga_step(g, fit, seed=5, generation=1), iters=5, warm=2)
        report("ga_step", {"population": n, "genes": 20, "tournament": n // 4}, t, 3 * n * 20 * 4 + n * 12, n, "offspring")
        e.close()
    eng.close()


if __name__ == "__main__":
    main()
