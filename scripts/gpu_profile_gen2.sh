#!/bin/bash
# Evidence run for the gen-2 matcher: full GPU test suite, bench (default steps), ncu launch list, ncu --set full of the three kernels.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['phase_ms'], d['objects_encoded_per_s'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --tracks 256 --dets 256 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_p -c 6 -f -o gpurun_out/prof_pair_final python bench.py --steps 1 --warmup 0 --tracks 128 --dets 128 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
