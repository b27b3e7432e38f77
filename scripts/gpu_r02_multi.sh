#!/bin/bash
# round 2, multi-GPU call: configs[4] sweep (strong scaling incl. score gather) and, on 2 GPUs, configs[2] (DGCNN) at its named shape.
# usage: gpurun --gpus N -- bash scripts/gpu_r02_multi.sh N
set -u
N=${1:-1}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$N" = "1" ] && TR="python"
SIZES="256,1024,4096"
[ "$N" -ge 4 ] && SIZES="256,1024,4096,16384"
SECONDS=0
timeout 600 $TR scripts/bench_sweep.py --sizes $SIZES --max-seconds 12 2> gpurun_out/r02_sweep_${N}gpu.err | grep '^{' > gpurun_out/r02_sweep_${N}gpu.jsonl; echo "sweep rc=$? ${SECONDS}s"; cat gpurun_out/r02_sweep_${N}gpu.jsonl | cut -c1-330; tail -3 gpurun_out/r02_sweep_${N}gpu.err
if [ "$N" = "2" ]; then
SECONDS=0
timeout 300 $TR bench.py --gpus 2 --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_2gpu.json 2> gpurun_out/r02_bench_c3_2gpu.err; echo "c3@2 rc=$? ${SECONDS}s"; cut -c1-700 gpurun_out/r02_bench_c3_2gpu.json
fi
