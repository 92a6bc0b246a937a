#!/bin/bash
# final round-2 profiling pass (short form of prof_r2.sh): launch list of one bench step + full capture of every kernel of
# one forward step.  Summarise with: python scripts/summarize_profiles.py gpurun_out profiles r2f
TAG=${1:-r2f}
mkdir -p gpurun_out
FWD="--no-cpu-baseline --bwd-steps 0 --bf16-steps 0 --pgd-frames 0 --pgd-weak-frames 0 --config-d-steps 0 --config-b-steps 0 --eager-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 $FWD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off \
    -o /tmp/prof_fwd_$TAG -f python scripts/one_step.py fwd > gpurun_out/ncu_fwd_$TAG.log 2>&1
ncu -i /tmp/prof_fwd_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_fwd_$TAG.raw.csv 2> /dev/null
