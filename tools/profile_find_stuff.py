#!/usr/bin/env python3
"""A few find_stuff launches on 16 384 synthetic frames (1.26 GB, larger than L2) for ncu: python tools/profile_find_stuff.py [frames]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import neuro_genetic_pong_self_play_b200 as ngp
from bench_ops import synth_frames
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = ngp.Engine(ngp.Config(), device=0)
frames = torch.from_numpy(synth_frames(64)).cuda().repeat(n // 64, 1, 1, 1).contiguous()
for _ in range(3):
    loc, valid = eng.find_stuff(frames)
torch.cuda.synchronize()
print("ok", valid.float().mean().item())
