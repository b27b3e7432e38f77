#!/bin/bash
# One short gpurun call: the GPU tests of the newest components only (image-token matcher, F-FPS / FS samplers, avg pooling).
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_image.py tests/test_gpu_pointnet_modules.py -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_new.log
