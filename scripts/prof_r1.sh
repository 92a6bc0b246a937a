#!/bin/bash
# round-1 profiling pass (run under gpurun): the default bench line, the launch list of one bench step, and full
# captures of the top kernels.  Summarise afterwards with: python scripts/summarize_profiles.py gpurun_out profiles r1
set -x
mkdir -p gpurun_out
python bench.py --table > gpurun_out/bench_r1.json 2> gpurun_out/table_r1.txt
FWD="--no-cpu-baseline --bwd-steps 0 --bf16-steps 0 --pgd-frames 0"
# launch list of the same forward step (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 $FWD > gpurun_out/ncu_launch_r1.log 2>&1
# full captures (batch 16, the bench workload): one forward's worth of engine launches (16 convs + stem_out), and the GF kernels
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 17 -c 17 \
    -o gpurun_out/prof_conv_r1 -f python bench.py --steps 1 --warmup 3 $FWD > gpurun_out/ncu_conv_r1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gf_forward_march -s 2 -c 2 \
    -o gpurun_out/prof_gf_r1 -f python bench.py --steps 1 --warmup 3 $FWD > gpurun_out/ncu_gf_r1.log 2>&1
ls -la gpurun_out
