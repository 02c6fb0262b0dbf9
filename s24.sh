mkdir -p gpurun_out; rm -f gpurun_out/s24_*
for cfg in "32 32" "64 21" "64 24" "64 32" "128 11" "32 21"; do set -- $cfg; echo "block=$1 lanes=$2" >> gpurun_out/s24_geo.log; NGP_ROLLOUT_BLOCK=$1 NGP_ROLLOUT_LANES=$2 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-saturated 2>&1 | cut -c1-190 >> gpurun_out/s24_geo.log; done
cat gpurun_out/s24_geo.log
