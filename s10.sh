mkdir -p gpurun_out; rm -f gpurun_out/s10_*
timeout 300 python -m pytest tests -m gpu -x -q -k "mlp" > gpurun_out/s10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s10_pytest.log
timeout 300 python tools/bench_ops.py --only mlp_wide > gpurun_out/s10_ops.log 2>&1
for cfg in "2048 32" "2048 64" "2048 128" "4096 32" "4096 128" "4096 256" "8192 128" "8192 256" "8192 512"; do set -- $cfg; echo "pop=$1 block=$2" >> gpurun_out/s10_geo.log; NGP_ROLLOUT_BLOCK=$2 python tools/profile_rollout.py --population $1 --max-frames 300 >> gpurun_out/s10_geo.log 2>&1; done
echo "full episodes pop 2048" >> gpurun_out/s10_geo.log
for b in 32 128; do NGP_ROLLOUT_BLOCK=$b python bench.py --population 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-saturated 2>&1 | cut -c1-180 >> gpurun_out/s10_geo.log; done
cat gpurun_out/s10_geo.log; tail -3 gpurun_out/s10_pytest.log; cut -c1-250 gpurun_out/s10_ops.log
