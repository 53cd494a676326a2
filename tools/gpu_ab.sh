# A/B of backward reduction variants
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_perf_vs_ref.py) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in smem shuffle; do
  (GSR_BWD_REDUCE=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/bench_$v.log 2>&1; echo "bench $v rc=$?"
  tail -1 gpurun_out/bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', d['value'], d['ms_per_step'], d['stages_ms'])"
done
