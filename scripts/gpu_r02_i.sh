#!/bin/bash
# round 2, GPU call I (2 GPUs): the driver's launch of bench.py at N = 2, configs[2] (DGCNN) at N = 1 and N = 2 at HEAD
set -u
mkdir -p gpurun_out
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; }
show() { python - <<P
import json
for l in open("gpurun_out/$1.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$1", round(d["value"]), round(d["ms_per_step"], 2), d["phase_ms"], d["n_gpus"], round(d["encoder"]["objects_per_s"]), (d.get("strong") or {}).get("value"), round(d["e2e"]["value"]), d["roofline"]["frac"], d["roofline"]["traffic"])
P
}
run2 29511 r02_bench_2gpu_head --steps 5 --warmup 3; show r02_bench_2gpu_head
timeout 600 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_head.json 2> gpurun_out/r02_bench_c3_head.err; echo "c3 rc=$?"; show r02_bench_c3_head
run2 29512 r02_bench_c3_2gpu_head --config c3 --steps 3 --warmup 3 --no-cpu-baseline; show r02_bench_c3_2gpu_head
