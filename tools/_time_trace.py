import sys, time, zlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import neuro_genetic_pong_self_play_b200 as ngp
eng = ngp.Engine(ngp.Config(), device=0)
eng.env_reset(32, 0)
a = torch.zeros((32, 16), dtype=torch.uint8, device="cuda")
T = {}
def tick(k, t0):
    torch.cuda.synchronize(); T[k] = T.get(k, 0) + time.perf_counter() - t0
for f in range(60):
    t0 = time.perf_counter(); out = eng.env_step(a, core=1); tick('step', t0)
    t0 = time.perf_counter(); ram = out["ram"].cpu().numpy(); regs = out["regs"].cpu().numpy(); tick('small_copies', t0)
    t0 = time.perf_counter(); frames = out["frames"].cpu().numpy(); tick('frames_copy', t0)
    t0 = time.perf_counter(); dig = eng.env_digest().cpu().numpy(); tick('digest', t0)
    t0 = time.perf_counter(); c = [zlib.crc32(frames[j].tobytes()) for j in range(32)]; tick('crc', t0)
print({k: round(v / 60 * 1e3, 2) for k, v in T.items()}, 'ms per frame')
