#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -16 | tee gpurun_out/r2_gpu_suite.log
for wl in c2 c3; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/r2_cur_${wl}.json 2>gpurun_out/r2_cur_${wl}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_cur_${wl}.json").read().strip().splitlines()[-1])
    st=d["stage_ms_per_step"]
    print("${wl}: value ms %.3f e2e ms %.3f | pre %.3f scan %.3f dup %.3f sort %.3f ranges %.3f fwd %.3f bwd %.3f geom %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], st["preprocess"], st["scan_block_sums"], st["duplicate_keys"], st["radix_sort"], st["tile_ranges"], st["render_forward"], st["render_backward"], st["geom_backward"]))
except Exception as e:
    print("${wl}: n/a", e); print(open("gpurun_out/r2_cur_${wl}.err").read()[-1500:])
PY
done
