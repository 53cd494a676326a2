set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run22
python -m pytest tests/test_gpu_fused.py tests/test_gpu_train_loop.py -m gpu -q -x > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
tail -3 gpurun_out/${R}_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench.json') if l.startswith('{')][-1]);print(d['value'], d['fused_operator'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_launches.log 2>&1
