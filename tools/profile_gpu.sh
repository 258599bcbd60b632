#!/usr/bin/env bash
# Run on the GPU box (via gpurun).  Writes everything under gpurun_out/.
#   1. launch list (per-launch device time) of a short bench run
#   2. ncu --set full capture of the dominant HP1 kernel and the HP2 cast kernel
# Numbers printed by bench.py under ncu are never bench values.
set -u
mkdir -p gpurun_out
STEPS=${STEPS:-6}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --hp2-envs 2048 \
    > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp1_step_kernel -s 6 -c 2 \
    -f -o gpurun_out/hp1_step_r1 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --no-hp2 \
    > gpurun_out/ncu_hp1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp2_cast_kernel -s 2 -c 1 \
    -f -o gpurun_out/hp2_cast_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --hp2-envs 2048 \
    > gpurun_out/ncu_hp2.log 2>&1
ls -la gpurun_out
