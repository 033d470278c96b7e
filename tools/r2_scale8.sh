#!/usr/bin/env bash
# N-GPU scaling check (gpurun --gpus N -- 'bash tools/r2_scale8.sh N [modes...]')
set -u
N=${1:-8}; shift || true
MODES=${@:-"overlap:nccl overlap:multimem"}
mkdir -p gpurun_out
for mm in $MODES; do
mode=${mm%%:*}; ar=${mm##*:}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 100 --warmup 5 --allreduce-mode $mode --allreduce $ar > gpurun_out/r2_scale${N}_${mode}_${ar}.json 2> gpurun_out/r2_scale${N}_${mode}_${ar}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_scale${N}_${mode}_${ar}.json").read().splitlines() if l.startswith("{")][-1])
    c4=d.get("c4") or {}
    print("N=${N} ${mode} ${ar}: c2 value %.3f ms  e2e %.3f ms | c4 batch %.2f ms e2e %.2f ms (%s)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], c4.get("ms_per_batch", float("nan")), (c4.get("e2e") or {}).get("ms_per_batch", float("nan")), c4.get("error", d.get("allreduce"))))
except Exception as e:
    print("N=${N} ${mode} ${ar}: n/a", e); print(open("gpurun_out/r2_scale${N}_${mode}_${ar}.err").read()[-1500:])
PY
done
