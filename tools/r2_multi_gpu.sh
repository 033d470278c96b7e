#!/usr/bin/env bash
# Multi-GPU call of round 2 (gpurun --gpus N -- 'bash tools/r2_multi_gpu.sh N'): the two-GPU tests, then the bench at N ranks with the
# three ways of exchanging the feature gradient, each with its c4 (BASELINE configs[3], strong scaling) leg.
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== multi-GPU tests =="; timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_multi_gpu_tests.log
run() {   # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_mg${N}_${name}.json 2> gpurun_out/r2_mg${N}_${name}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_mg${N}_${name}.json").read().splitlines() if l.startswith("{")][-1])
    c4=d.get("c4") or {}
    print("N=${N} ${name}: c2 value %.3f ms  e2e %.3f ms | c4 batch %.2f ms e2e %.2f ms (%s)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], c4.get("ms_per_batch", float("nan")), (c4.get("e2e") or {}).get("ms_per_batch", float("nan")), c4.get("error", d.get("allreduce"))))
except Exception as e:
    print("N=${N} ${name}: n/a", e); print(open("gpurun_out/r2_mg${N}_${name}.err").read()[-1200:])
PY
}
run sync --allreduce-mode sync
run overlap --allreduce-mode overlap
run multimem --allreduce multimem
run reference --impl reference --steps 10
