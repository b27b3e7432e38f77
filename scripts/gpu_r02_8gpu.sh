#!/bin/bash
# round 2, 8-GPU call: headline bench (weak) + its strong-scaling leg incl. the score gather, C4 at its named shape, north-star target shape
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
$TR bench.py --gpus 8 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$? ${SECONDS}s"; tail -c 1800 gpurun_out/r02_bench_8gpu.json
SECONDS=0
$TR bench.py --gpus 8 --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c4_8gpu.json 2> gpurun_out/r02_bench_c4_8gpu.err; echo "c4 rc=$? ${SECONDS}s"; tail -c 1200 gpurun_out/r02_bench_c4_8gpu.json
SECONDS=0
$TR scripts/bench_target.py > gpurun_out/r02_target_shape_8gpu.json 2> gpurun_out/r02_target_8gpu.err; echo "target rc=$? ${SECONDS}s"; tail -c 600 gpurun_out/r02_target_shape_8gpu.json
