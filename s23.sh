mkdir -p gpurun_out; rm -f gpurun_out/s23_*
for pop in 2048 4096; do for lean in 0 1; do echo "pop=$pop lean=$lean" >> gpurun_out/s23_geo.log; NGP_ROLLOUT_LEAN=$lean python tools/profile_rollout.py --population $pop --max-frames 300 >> gpurun_out/s23_geo.log 2>&1; done; done
python bench.py --population 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-saturated 2>&1 | cut -c1-200 >> gpurun_out/s23_geo.log
NGP_ROLLOUT_LEAN=1 python bench.py --population 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-saturated 2>&1 | cut -c1-200 >> gpurun_out/s23_geo.log
cat gpurun_out/s23_geo.log
