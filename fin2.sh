mkdir -p gpurun_out; rm -f gpurun_out/fin_*
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/fin_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/fin_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/fin_smoke.log
timeout 600 python bench.py > gpurun_out/fin_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/fin_bench_n1.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-saturated > gpurun_out/fin_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -o gpurun_out/fin_prof_cfg2 python tools/profile_rollout.py --population 1024 --max-frames 60 > gpurun_out/fin_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -o gpurun_out/fin_prof_sat python tools/profile_rollout.py --population 32768 --max-frames 100 --launches 2 > gpurun_out/fin_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -o gpurun_out/fin_prof_p300 python tools/profile_rollout.py --population 1024 --max-frames 300 > gpurun_out/fin_ncu3.log 2>&1
for pop in 1024 2048 4096 8192 16384 32768; do python tools/profile_rollout.py --population $pop --max-frames 300 >> gpurun_out/fin_sweep.log 2>&1; done
tail -3 gpurun_out/fin_pytest.log; cat gpurun_out/fin_smoke.log; cut -c1-300 gpurun_out/fin_bench_n1.log; cut -c1-200 gpurun_out/fin_bench_ref.log; cat gpurun_out/fin_sweep.log
