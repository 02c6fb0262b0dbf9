mkdir -p gpurun_out; rm -f gpurun_out/s21_*
timeout 900 python -m pytest tests -m gpu -x -q -k "fused_round_robin or fast_flavour" > gpurun_out/s21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s21_pytest.log
python tools/profile_rollout.py --population 1024 --max-frames 60 > gpurun_out/s21_p60.log 2>&1
python tools/profile_rollout.py --population 1024 --max-frames 300 > gpurun_out/s21_p300.log 2>&1
python tools/profile_rollout.py --population 32768 --max-frames 300 > gpurun_out/s21_sat300.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s21_bench.log 2>&1
tail -n 3 gpurun_out/s21_pytest.log gpurun_out/s21_p60.log gpurun_out/s21_p300.log; cat gpurun_out/s21_sat300.log
cut -c1-200 gpurun_out/s21_bench.log; grep -o '"saturated".*' gpurun_out/s21_bench.log | cut -c1-300
