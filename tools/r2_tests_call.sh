#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== new GPU tests =="; timeout 900 python -m pytest tests/test_renderer_dropin_gpu.py tests/test_sampling_gpu.py -m gpu -q -x 2>&1 | tail -15
echo "== large live-reference =="; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "baseline_sizes" --durations=5 2>&1 | tail -15
echo "== tc quick =="; bash tools/r2_bwd_tc_quick.sh 2>&1 | head -4
