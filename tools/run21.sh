set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run21
python -m pytest tests/test_gpu_fused.py tests/test_gpu_train_loop.py tests/test_gpu_densify.py -m gpu -q -x -s > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
grep "^\[" gpurun_out/${R}_pytest.log | cut -c1-300 | tail -20; tail -4 gpurun_out/${R}_pytest.log
for m in tr scan; do GSR_PACKED_BWD=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_$m.json 2> gpurun_out/${R}_bench_$m.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench_$m.json') if l.startswith('{')][-1]);print('$m', d['value'], d['fused_operator'], d['e2e']['value'])"; done
