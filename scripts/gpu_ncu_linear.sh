#!/bin/bash
# ncu --set full of cn_linear_tma_kernel on two shapes (store-bound small K, MMA-bound large K)
mkdir -p gpurun_out
for shp in "2048 256 128 128" "1024 256 1024 512"; do
  tag=$(echo $shp | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:cn_linear_tma -c 1 -s 2 -f -o gpurun_out/r02_linear_tma_$tag python scripts/ncu_linear.py $shp > gpurun_out/ncu_linear_$tag.log 2>&1
  echo "rc=$? $tag"; tail -2 gpurun_out/ncu_linear_$tag.log
done
