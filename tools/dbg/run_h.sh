for pct in -1 25 50 100; do echo "CARVEOUT $pct"; AGX_CARVEOUT_PCT=$pct python tools/hp1_time.py --reps 2 2>&1 | cut -c1-200; done
for pct in -1 25; do echo "LOOPBACK pct=$pct"; AGX_CARVEOUT_PCT=$pct timeout 200 python -m pytest tests/test_hp1_gpu.py -q -x -k "loopback or timed_out" 2>&1 | tail -2; done
bash tools/run_variants.sh gpurun_out/r2_variants_h.jsonl trig trig2 wpb1 wpb1t wpb4 mb10t > /dev/null 2>&1; cut -c1-230 gpurun_out/r2_variants_h.jsonl
timeout 200 python -m pytest tests/test_hp2_gpu.py tests/test_hp2_reference_fixtures.py tests/test_graph_step_gpu.py tests/test_sharded_task_gpu.py -q -m gpu 2>&1 | tail -4
