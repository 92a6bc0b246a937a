#!/bin/bash
# round-2 profiling pass (run under gpurun): the launch list of one bench step and full captures of every kernel of one
# forward step and of one forward+backward-to-input step.  TAG (default r2) names the outputs.
# Summarise afterwards with: python scripts/summarize_profiles.py gpurun_out profiles $TAG
TAG=${1:-r2}
set -x
mkdir -p gpurun_out
FWD="--no-cpu-baseline --bwd-steps 0 --bf16-steps 0 --pgd-frames 0 --pgd-weak-frames 0 --config-d-steps 0 --config-b-steps 0 --eager-steps 0"
# launch list of the same forward step (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 $FWD > gpurun_out/ncu_launch_$TAG.log 2>&1
# full captures of every kernel of ONE forward step (batch 16, the bench workload; per-operator path = same launches)
timeout 1500 ncu --set full --clock-control none --profile-from-start off \
    -o /tmp/prof_fwd_$TAG -f python scripts/one_step.py fwd > gpurun_out/ncu_fwd_$TAG.log 2>&1
# ... of one forward+backward-to-input step (batch 8: same kernels and per-pixel traffic, half the replay time)
timeout 1800 ncu --set full --clock-control none --profile-from-start off \
    -o /tmp/prof_bwd_$TAG -f python scripts/one_step.py bwd > gpurun_out/ncu_bwd_$TAG.log 2>&1
# ... and of the bf16-storage forward
timeout 1500 ncu --set full --clock-control none --profile-from-start off \
    -o /tmp/prof_bf16_$TAG -f python scripts/one_step.py bf16 > gpurun_out/ncu_bf16_$TAG.log 2>&1
# the .ncu-rep files stay on the box (too large to pull back); their raw pages travel as CSV
for k in fwd bwd bf16; do
  ncu -i /tmp/prof_${k}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${k}_$TAG.raw.csv 2> /dev/null
done
ls -la gpurun_out /tmp/*.ncu-rep
