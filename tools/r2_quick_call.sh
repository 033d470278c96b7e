#!/usr/bin/env bash
# quick check of a change: parity tests (small + c2-size live reference), the c2 bench stage times, the e2e kernel timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "not baseline_sizes" 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/r2_quick_c2.json 2>gpurun_out/r2_quick_c2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_quick_c2.json").read().strip().splitlines()[-1]); st=d["stage_ms_per_step"]
    print("c2: value ms %.3f e2e ms %.3f | pre %.3f scan %.3f dup %.3f sort %.3f ranges %.3f fwd %.3f bwd %.3f geom %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], st["preprocess"], st["scan_block_sums"], st["duplicate_keys"], st["radix_sort"], st["tile_ranges"], st["render_forward"], st["render_backward"], st["geom_backward"]))
except Exception as e:
    print("c2: n/a", e); print(open("gpurun_out/r2_quick_c2.err").read()[-1500:])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 --no-cpu-baseline --no-c4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reference arm: value ms %.2f e2e ms %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
timeout 300 python tools/gpu_timeline.py --e2e > gpurun_out/r2_e2e_timeline.txt 2>&1; tail -75 gpurun_out/r2_e2e_timeline.txt
