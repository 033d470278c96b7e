#!/usr/bin/env bash
# quick GPU iteration on the tcgen05 backward: variants test, bench (stage times), timeline
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "variants_agree and cf" 2>&1 | tail -2
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-c4 --bwd-kernel tc > gpurun_out/r2_bench_bwd_tc.json 2>gpurun_out/r2_bench_bwd_tc.err; tail -2 gpurun_out/r2_bench_bwd_tc.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_bwd_tc.json").read().strip().splitlines()[-1])
print("tc: value ms %.3f e2e ms %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
SAGARS_LIBRARY=$PWD/seganygaussians_b200/lib/timeline/libsagars.so timeout 250 python tools/bt_timeline.py 2>&1 | tail -28
