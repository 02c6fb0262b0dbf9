mkdir -p gpurun_out; rm -f gpurun_out/s25_*
for cfg in "512 1" "384 3" "256 3"; do set -- $cfg; echo "sat block=$1 lean=$2" >> gpurun_out/s25_geo.log; NGP_ROLLOUT_BLOCK=$1 NGP_ROLLOUT_LEAN=$2 python tools/profile_rollout.py --population 32768 --max-frames 300 >> gpurun_out/s25_geo.log 2>&1; done
echo "pop 8192" >> gpurun_out/s25_geo.log
for cfg in "512 1" "384 3"; do set -- $cfg; echo "block=$1 lean=$2" >> gpurun_out/s25_geo.log; NGP_ROLLOUT_BLOCK=$1 NGP_ROLLOUT_LEAN=$2 python tools/profile_rollout.py --population 8192 --max-frames 300 >> gpurun_out/s25_geo.log 2>&1; done
cat gpurun_out/s25_geo.log
