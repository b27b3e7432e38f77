#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_p -c 6 -f -o gpurun_out/prof_pair_gen2 python bench.py --steps 1 --warmup 0 --tracks 128 --dets 128 --no-cpu-baseline > gpurun_out/ncu_full_gen2.log 2>&1; echo "ncu full rc=$?"
