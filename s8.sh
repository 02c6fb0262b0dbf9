mkdir -p gpurun_out; rm -f gpurun_out/s8_*
timeout 900 python -m pytest tests -m gpu -x -q -k "trace_parity or fused or population_1024 or fast_flavour or start_states" > gpurun_out/s8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s8_pytest.log
python tools/profile_rollout.py --population 1024 --max-frames 60 > gpurun_out/s8_p60.log 2>&1
python tools/profile_rollout.py --population 1024 --max-frames 300 > gpurun_out/s8_p300.log 2>&1
python tools/profile_rollout.py --population 32768 --max-frames 300 > gpurun_out/s8_sat300.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -o gpurun_out/s8_prof_p300 python tools/profile_rollout.py --population 1024 --max-frames 300 > gpurun_out/s8_ncu.log 2>&1
tail -n 3 gpurun_out/s8_pytest.log gpurun_out/s8_p60.log gpurun_out/s8_p300.log gpurun_out/s8_sat300.log
cut -c1-200 gpurun_out/s8_bench.log
