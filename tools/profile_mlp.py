#!/usr/bin/env python3
"""One wide grouped-MLP forward ([6,512,512,2], 64 envs/genome) for ncu."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import neuro_genetic_pong_self_play_b200 as ngp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=(6, 512, 512, 2)), device=0)
g = torch.randn((n, eng.gene_size), device="cuda") * 0.05
x = torch.rand((n, 64, 6), device="cuda")
for _ in range(3):
    act, _ = eng.mlp_forward(g, x, want_out=False)
torch.cuda.synchronize()
print("ok", act.float().mean().item())
