set -x
for v in "" 3424014; do
  (GSR_EXPERIMENT_ASSUME_M=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline) > gpurun_out/bench_exp.log 2>&1
  tail -1 gpurun_out/bench_exp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('assumeM=[$v]', d['value'], d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['value'])"
done
