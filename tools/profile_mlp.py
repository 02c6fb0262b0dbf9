#!/usr/bin/env python3
"""One wide grouped-MLP forward ([6,512,512,2], `envs` envs/genome) for ncu: python tools/profile_mlp.py [genomes] [envs] [prepared]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import neuro_genetic_pong_self_play_b200 as ngp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prepared = len(sys.argv) > 3 and sys.argv[3] == "prepared"
eng = ngp.Engine(ngp.Config(NETWORK_SHAPE=(6, 512, 512, 2)), device=0)
g = torch.randn((n, eng.gene_size), device="cuda") * 0.05
x = torch.rand((n, envs, 6), device="cuda")
if prepared:
    eng.mlp_prepare(g)
for _ in range(3):
    act, _ = eng.mlp_forward_prepared(g, x, want_out=False) if prepared else eng.mlp_forward(g, x, want_out=False)
torch.cuda.synchronize()
print("ok", act.float().mean().item())
