set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_binning_device.py tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/r2_run6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run6_pytest.log
tail -4 gpurun_out/r2_run6_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run6_bench.json 2> gpurun_out/r2_run6_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_run6_bench.json")); print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["stages_ms"]); print(d["ref_cuda_ext"].get("speedup"), d.get("cuda_graph"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r2_run6_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_run6_ncu_launch.log 2>&1
python - <<'PY'
import csv
with open('gpurun_out/r2_run6_launches.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1000
    elif unit=='ms': v*=1000
    seq.append((row['Kernel Name'][:70],v))
idx=[i for i,(n,v) in enumerate(seq) if 'sh_forward' in n]
a,b=idx[0],idx[1]
for n,v in seq[a:b]:
    if 'elementwise' in n: continue
    print(f"{v:8.1f} us  {n}")
PY
