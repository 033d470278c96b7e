#!/usr/bin/env bash
# A/B of the forward warp kernel's feature-row fetch: cp.async pieces (default build) vs bulk copies on the TMA unit
# (make -C seganygaussians_b200/csrc OUTD=../lib/fw_bulk OBJD=../lib/fw_bulk/obj EXTRA_NVCCFLAGS=-DSAGARS_FW_BULK=1)
set -u
mkdir -p gpurun_out
echo "== parity with the bulk-copy build =="
SAGARS_LIBRARY=$PWD/seganygaussians_b200/lib/fw_bulk/libsagars.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "oracle or golden or live_reference and not baseline_sizes" 2>&1 | tail -3
for v in default fw_bulk default fw_bulk; do
  lib=$PWD/seganygaussians_b200/lib/libsagars.so; [ $v = fw_bulk ] && lib=$PWD/seganygaussians_b200/lib/fw_bulk/libsagars.so
  SAGARS_LIBRARY=$lib timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-c4 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['stage_ms_per_step']
print('$v: value ms %.3f e2e ms %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], st['render_forward'], st['render_backward']))"
done
