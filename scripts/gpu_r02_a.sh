#!/bin/bash
# round 2, GPU call A: GPU test suite, default bench line, ncu launch list + --set full capture of the fused pair kernels
set -u
mkdir -p gpurun_out
SECONDS=0
python -m pytest tests -m gpu -x -q 2>&1 | tail -5; echo "pytest ${SECONDS}s"
SECONDS=0
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$? ${SECONDS}s"; tail -c 3000 gpurun_out/r02_bench_default.json
SECONDS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_parity_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu launch list rc=$? (${SECONDS}s)"
SECONDS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_p -c 6 -f -o gpurun_out/r02_pair_parity_tc python bench.py --steps 1 --warmup 0 --tracks 256 --dets 256 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$? (${SECONDS}s)"
ls -la gpurun_out | tail -8
