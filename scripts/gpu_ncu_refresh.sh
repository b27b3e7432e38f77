#!/bin/bash
# ncu evidence of the bench command at HEAD: launch list (gpu__time_duration per launch) + one --set full capture of the fused
# pair kernels.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_head.csv python bench.py --steps 1 --warmup 1 --tracks 256 --dets 256 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list rc=$? (${SECONDS}s)"
SECONDS=0
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_p -c 6 -o gpurun_out/prof_pair_head -f python bench.py --steps 1 --warmup 0 --tracks 128 --dets 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? (${SECONDS}s)"
ls -la gpurun_out/ | tail -8
