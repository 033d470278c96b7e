#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== renderer drop-in + sampling + multi-gpu(1) =="; timeout 900 python -m pytest tests/test_renderer_dropin_gpu.py tests/test_sampling_gpu.py tests/test_smoothing_gpu.py tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -15
