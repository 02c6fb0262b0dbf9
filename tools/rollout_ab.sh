#!/bin/bash
# A/B harness for rollout-kernel changes: lock-step latency runs (dense and sparse warps) + a short bench line.
# usage (on the GPU box): bash tools/rollout_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out/${tag}_rollout_ab.txt
{
python tools/profile_rollout.py --population 1024 --max-frames 300
python tools/profile_rollout.py --population 1 --max-frames 600
python tools/profile_rollout.py --population 1 --games 1 --max-frames 600
python tools/profile_rollout.py --population 2048 --max-frames 300
python tools/profile_rollout.py --population 16384 --max-frames 100
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras
} 2>gpurun_out/${tag}_rollout_ab.err | tee $out
