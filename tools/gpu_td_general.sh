mkdir -p gpurun_out
python tools/bench_td_general.py acrobot 296 10 | tee gpurun_out/td_general.log
python tools/bench_td_general.py cartpole 296 10 | tee -a gpurun_out/td_general.log
python tools/bench_td_general.py acrobot 148 10 | tee -a gpurun_out/td_general.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:general_td_update_kernel -s 3 -c 1 -f -o gpurun_out/prof_td python tools/bench_td_general.py acrobot 296 2 > gpurun_out/prof_td.log 2>&1
tail -2 gpurun_out/prof_td.log
