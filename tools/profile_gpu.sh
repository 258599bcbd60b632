#!/usr/bin/env bash
# Run on the GPU box (via gpurun).  Writes everything under gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r1}
STEPS=${STEPS:-6}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --hp2-envs 2048 \
    > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp1_step_kernel -s ${SKIP:-20} -c 2 \
    -f -o gpurun_out/hp1_step_$TAG python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --no-hp2 \
    > gpurun_out/ncu_hp1.log 2>&1
if [ "${HP2:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp2_cast_kernel -s 2 -c 1 \
    -f -o gpurun_out/hp2_cast_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --hp2-envs 2048 \
    > gpurun_out/ncu_hp2.log 2>&1
fi
ls -la gpurun_out | tail -5
