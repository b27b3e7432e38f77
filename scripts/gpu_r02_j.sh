#!/bin/bash
# round 2, GPU call J: final ncu evidence at HEAD -- launch list of the bench command + --set full capture of the fused pair kernels
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_head.csv python bench.py --steps 1 --warmup 1 --tracks 256 --dets 256 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list rc=$? (${SECONDS}s)"
SECONDS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_p -c 6 -o gpurun_out/r02_pair_parity_tc_head -f python bench.py --steps 1 --warmup 0 --tracks 256 --dets 256 --no-cpu-baseline --no-extra > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? (${SECONDS}s)"
ls -la gpurun_out/ | grep "r02_pair_parity_tc_head\|r02_launches_head"
