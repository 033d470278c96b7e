#!/usr/bin/env bash
# End-of-round verification on one B200: full GPU suite, smoke(), both bench arms, ncu evidence, compute-sanitizer.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu =="; timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 | tee gpurun_out/r2_final_gpu_suite.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench (ours) =="; timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.json
echo "== bench (reference arm) =="; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; tail -c 400 gpurun_out/r2_final_bench_reference.json
echo "== bench c3 =="; timeout 300 python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/r2_final_bench_c3.json 2> gpurun_out/r2_final_bench_c3.err
python - <<'PY'
import json
for f in ("r2_final_bench", "r2_final_bench_c3"):
    try:
        d=json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); st=d["stage_ms_per_step"]
        print(f, "value ms %.3f e2e ms %.3f | pre %.3f scan %.3f dup %.3f sort %.3f ranges %.3f fwd %.3f bwd %.3f geom %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], st["preprocess"], st["scan_block_sums"], st["duplicate_keys"], st["radix_sort"], st["tile_ranges"], st["render_forward"], st["render_backward"], st["geom_backward"]), "roofline", d["roofline"]["frac"], "c4", (d.get("c4") or {}).get("ms_per_batch"))
    except Exception as e:
        print(f, "n/a", e)
PY
echo "== ncu =="; bash tools/r2_ncu_call.sh 2>&1 | grep -v "^{" | tail -5
echo "== sanitizer =="; bash tools/r2_sanitizer_call.sh 2>&1 | tail -20
