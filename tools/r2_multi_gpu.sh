#!/usr/bin/env bash
# Two-GPU call of round 2 (charged 2x; ~5 minutes):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_multi_gpu.sh'
# 1. gated two-GPU tests (own multicast all-reduce vs NCCL; blend_wait_event gating); 2. bench at N=2: NCCL serialised (default),
# NCCL overlapped with the next forward's geometry stages, own multicast all-reduce.  Nothing here changes a default.
set -u
mkdir -p gpurun_out
N=${N:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== gated multi-GPU tests =="; timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2_multi_gpu_tests.log
for variant in "" "--overlap-allreduce" "--allreduce multimem"; do
  tag=$(echo "${variant:-default}" | tr -d ' -')
  echo "== bench N=$N ${variant:-default} =="
  timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 $variant > gpurun_out/r2_bench_n${N}_${tag}.json 2> gpurun_out/r2_bench_n${N}_${tag}.err
  tail -c 900 gpurun_out/r2_bench_n${N}_${tag}.json; tail -3 gpurun_out/r2_bench_n${N}_${tag}.err
done
