#!/usr/bin/env bash
# Round-2 ncu captures (run on the GPU box via gpurun; one GPU).  Everything lands in gpurun_out/; tools/summarize_ncu.py and
# tools/summarize_launches.py turn the reports into profiles/*.md here.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2}
# 1. launch list of the bench's timed loops (per-launch durations are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs --hp2-envs 2048 \
    > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# 2. the fused HP1 step kernel (specialised instantiation), two launches of the timed loop
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp1_step_kernel -s 20 -c 2 \
    -f -o gpurun_out/hp1_step_$TAG python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-hp2 --no-configs \
    > gpurun_out/ncu_hp1_$TAG.log 2>&1
# 3. the ray caster: 64x48 tile path (scene in shared memory) and the records-only tile path of the 1024-box scene
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp2_cast_kernel -s 2 -c 1 \
    -f -o gpurun_out/hp2_cast_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --hp2-envs 2048 \
    > gpurun_out/ncu_hp2_$TAG.log 2>&1
E=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hp2_cast_kernel -s 1 -c 1 \
    -f -o gpurun_out/hp2_cast_cfg3_$TAG python tools/dbg/cfg3_once.py > gpurun_out/ncu_hp2_cfg3_$TAG.log 2>&1
# 4. the observation gather's push kernel beside the chained steps (one GPU, emulated world of 2: the stores go to local buffers)
ONLY=ncu timeout 600 ncu --set full --clock-control none --import-source on -k regex:obs_gather_push_kernel -s 40 -c 1 \
    -f -o gpurun_out/obs_push_$TAG python tools/dbg/dbg_gather_loop.py > gpurun_out/ncu_push_$TAG.log 2>&1
ls -la gpurun_out | tail -8
