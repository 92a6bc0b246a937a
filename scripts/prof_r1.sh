#!/bin/bash
# round-1 profiling pass (run under gpurun): launch list of one bench step + full captures of the top kernels
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --table > gpurun_out/bench_a.json 2> gpurun_out/table_a.txt
# launch list: skip the 2*3 warm-up steps (resident+e2e each 26 launches... keep all, small)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_a.log 2>&1
# full capture: conv engine (launch index chosen to hit k3 cin32, cin64, cin96 and k7) and the GF kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 \
    -o gpurun_out/prof_conv_a python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_conv_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gf_ -s 2 -c 2 \
    -o gpurun_out/prof_gf_a python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_gf_a.log 2>&1
ls -la gpurun_out
