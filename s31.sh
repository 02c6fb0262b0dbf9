mkdir -p gpurun_out; rm -f gpurun_out/s31_*
timeout 300 python -m pytest tests -m gpu -x -q -k "fused_round_robin or fast_flavour" > gpurun_out/s31_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s31_pytest.log
python tools/profile_rollout.py --population 1024 --max-frames 60 > gpurun_out/s31.log 2>&1
python tools/profile_rollout.py --population 1024 --max-frames 300 >> gpurun_out/s31.log 2>&1
python tools/profile_rollout.py --population 32768 --max-frames 300 >> gpurun_out/s31.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-saturated 2>&1 | cut -c1-200 >> gpurun_out/s31.log
tail -3 gpurun_out/s31_pytest.log; cat gpurun_out/s31.log
