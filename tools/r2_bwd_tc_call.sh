#!/usr/bin/env bash
# GPU call: the tcgen05 backward -- parity (variants test, full c2 size vs warp kernel and vs the reference), then A/B bench
set -u
mkdir -p gpurun_out
echo "== variants test =="; timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "variants_agree" 2>&1 | tail -8
echo "== full-size check =="; timeout 400 python tools/gpu_bwd_tc_check.py 2>&1 | grep -v "^  INT" | tail -40
echo "== bench default =="; timeout 200 python bench.py --steps 30 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r2_bench_bwd_default.json 2>gpurun_out/r2_bench_bwd_default.err; python - <<'PY'
import json
for n in ("default","tc"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_bwd_{n}.json").read().strip().splitlines()[-1])
        print(n, "value ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["stage_ms_per_step"])
    except Exception as e: print(n, "n/a", e)
PY
echo "== bench tc =="; timeout 200 python bench.py --steps 30 --warmup 5 --no-c4 --no-cpu-baseline --bwd-kernel tc > gpurun_out/r2_bench_bwd_tc.json 2>gpurun_out/r2_bench_bwd_tc.err; tail -3 gpurun_out/r2_bench_bwd_tc.err; python - <<'PY'
import json
for n in ("default","tc"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_bwd_{n}.json").read().strip().splitlines()[-1])
        print(n, "value ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["stage_ms_per_step"])
    except Exception as e: print(n, "n/a", e)
PY
