mkdir -p gpurun_out; rm -f gpurun_out/s4_*
timeout 600 python -m pytest tests -m gpu -x -q -k "find_stuff or mlp" > gpurun_out/s4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s4_pytest.log
python tools/bench_ops.py --only find_stuff,mlp_wide > gpurun_out/s4_ops.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:find_stuff -s 20 -c 1 -o gpurun_out/s4_prof_find_stuff python tools/bench_ops.py --only find_stuff > gpurun_out/s4_ncu_fs.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 6 -c 3 -o gpurun_out/s4_prof_mlp python tools/profile_mlp.py 1024 > gpurun_out/s4_ncu_mlp.log 2>&1
tail -n 3 gpurun_out/s4_pytest.log
cut -c1-330 gpurun_out/s4_ops.log
tail -3 gpurun_out/s4_ncu_mlp.log
