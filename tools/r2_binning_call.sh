#!/usr/bin/env bash
# GPU call: depth-first binning -- bit-identity test, then A/B bench at c2 and c3-like
set -u
mkdir -p gpurun_out
echo "== depth-first test =="; timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "depth_first" 2>&1 | tail -5
for wl in c2 c3; do for b in radix depth_first; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-c4 --binning $b > gpurun_out/r2_bin_${wl}_${b}.json 2>gpurun_out/r2_bin_${wl}_${b}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bin_${wl}_${b}.json").read().strip().splitlines()[-1])
    st=d["stage_ms_per_step"]
    print("${wl} ${b}: value ms %.3f e2e ms %.3f | pre %.3f scan %.3f dup %.3f sort %.3f ranges %.3f fwd %.3f bwd %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], st["preprocess"], st["scan_block_sums"], st["duplicate_keys"], st["radix_sort"], st["tile_ranges"], st["render_forward"], st["render_backward"]))
except Exception as e:
    print("${wl} ${b}: n/a", e); print(open("gpurun_out/r2_bin_${wl}_${b}.err").read()[-1500:])
PY
done; done
