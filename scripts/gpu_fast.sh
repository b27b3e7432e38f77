#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_tc.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_fused.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_fast.json; tail -3 gpurun_out/bench_fast.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_fast.csv python bench.py --steps 1 --warmup 1 --tracks 128 --dets 128 --no-cpu-baseline > gpurun_out/ncu_fast.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_p -c 4 -o gpurun_out/prof_pair python bench.py --steps 1 --warmup 0 --tracks 128 --dets 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
