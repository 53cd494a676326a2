set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run24
python -m pytest tests/test_gpu_fused.py -m gpu -q -x > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
tail -3 gpurun_out/${R}_pytest.log
for b in 128 256; do GSR_PACKED_TR_BATCH=$b python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_$b.json 2> gpurun_out/${R}_bench_$b.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench_$b.json') if l.startswith('{')][-1]);print('$b', d['value'], d['fused_operator'])"; done
GSR_PACKED_TR_BATCH=128 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:packed -c 40 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_launches.log 2>&1
grep "packed" gpurun_out/${R}_launches.csv | tail -4 | cut -c1-300
