#!/usr/bin/env bash
# First GPU call of round 2 (run under gpurun, one GPU, ~6-8 minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# 1. the gated tests of the variants written after round 1's GPU budget was spent (tile-sort binning, bulk-copy staging,
#    scale_modifier parity); 2. A/B timings of those variants; 3. the regular GPU suite; 4. a fresh launch list.
# Everything lands in gpurun_out/ (merged back by gpurun).  Nothing here changes a default.
set -u
mkdir -p gpurun_out
export SAGARS_TEST_EXPERIMENTAL=1
echo "== gated tests ==";          timeout 400 python -m pytest tests -m gpu -q -k "tile_sort or tma or scale_modifier or blend_wait" 2>&1 | tail -15 | tee gpurun_out/r2_gated_tests.log
unset SAGARS_TEST_EXPERIMENTAL
echo "== bench: default ==";        timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 1500 gpurun_out/r2_bench_default.json
echo "== bench: tile_sort ==";      timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --binning tile_sort > gpurun_out/r2_bench_tile_sort.json 2> gpurun_out/r2_bench_tile_sort.err; tail -c 1500 gpurun_out/r2_bench_tile_sort.json
echo "== bench c3-like: default / tile_sort (5M Gaussians: the binning-dominated case) =="
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_default.json 2>/dev/null; tail -c 700 gpurun_out/r2_bench_c3_default.json
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --binning tile_sort > gpurun_out/r2_bench_c3_tile_sort.json 2>/dev/null; tail -c 700 gpurun_out/r2_bench_c3_tile_sort.json
echo "== K=3 forward: fp32 tile kernel (default) vs warp kernel =="
timeout 200 python bench.py --workload c2_k3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_k3_default.json 2>/dev/null; tail -c 600 gpurun_out/r2_bench_k3_default.json
timeout 200 python bench.py --workload c2_k3 --steps 20 --warmup 5 --no-cpu-baseline --fwd-kernel warp_any > gpurun_out/r2_bench_k3_warp_any.json 2>/dev/null; tail -c 600 gpurun_out/r2_bench_k3_warp_any.json
echo "== staging A/B (fp32 tile forward) =="; timeout 400 python tools/gpu_check.py --staging-ab 2>&1 | grep -E "timing|ours" | tee gpurun_out/r2_staging_ab.log
echo "== build variant pack_cpos (make -C seganygaussians_b200/csrc variants, run BEFORE gpurun): parity subset + bench =="
VARIANT=$PWD/seganygaussians_b200/lib/variants/pack_cpos/libsagars.so
if [ -f "$VARIANT" ]; then
  SAGARS_LIBRARY=$VARIANT timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_pack_parity.log
  SAGARS_LIBRARY=$VARIANT timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_pack_cpos.json 2>/dev/null; tail -c 900 gpurun_out/r2_bench_pack_cpos.json
else
  echo "(not built: run make -C seganygaussians_b200/csrc variants first)"
fi
echo "== regular GPU suite ==";     timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_gpu_suite.log
